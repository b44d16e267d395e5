"""oracle/ref.py -- run the reference's OWN JavaScript (lib/codel.js, lib/queue.js,
lib/utils.js, verbatim) under the SpiderMonkey shell built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/gen_golden.py and by
bench.py's `--impl reference` / cpu_baseline legs.  Never by the product path.
"""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
JS = os.path.join(HERE, "_ref", "js")
BUNDLE = os.path.join(HERE, "_ref", "bundle.js")
SHIM = os.path.join(HERE, "js", "shim.js")
DRIVER = os.path.join(HERE, "js", "driver.js")


def available():
    return os.access(JS, os.X_OK) and os.path.exists(BUNDLE)


def _lit(v):
    """Python value -> JS literal (ints/floats print round-trip exact)."""
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int,)):
        return repr(int(v))
    if isinstance(v, float):
        return repr(v)
    if isinstance(v, str):
        return json.dumps(v)
    if isinstance(v, dict):
        return "{" + ",".join("%s:%s" % (json.dumps(k), _lit(x)) for k, x in v.items()) + "}"
    if hasattr(v, "tolist"):
        v = v.tolist()
    return "[" + ",".join(_lit(x) for x in v) + "]"


def _stdin_arrays(arrays):
    """Serialise arrays for __read_array() in driver.js (bulk data never goes in
    as script literals: SM 1.7 breaks beyond 65535 literals per script)."""
    out = []
    for a in arrays:
        if a is None:
            out.append("null")
            continue
        if hasattr(a, "tolist"):
            a = a.tolist()
        a = list(a)
        out.append(str(len(a)))
        for i in range(0, len(a), 1000):
            out.append(",".join(repr(x) for x in a[i:i + 1000]))
    return "\n".join(out) + "\n"


def call(expr, timeout=3600, arrays=()):
    """Evaluate a driver expression, return the JSON-decoded result.  `arrays`
    are fed over stdin; the expression pulls them with __read_array() in order."""
    if not available():
        raise RuntimeError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
    with tempfile.NamedTemporaryFile("w", suffix=".js", delete=False) as f:
        f.write('load(%s); load(%s); load(%s);\n' % (json.dumps(SHIM), json.dumps(BUNDLE), json.dumps(DRIVER)))
        f.write("print(__json(%s));\n" % expr)
        path = f.name
    try:
        out = subprocess.run([JS, path], capture_output=True, text=True, timeout=timeout,
                             input=_stdin_arrays(arrays))
    finally:
        os.unlink(path)
    if out.returncode != 0 or not out.stdout.strip():
        raise RuntimeError("reference run failed: rc=%d\n%s\n%s" % (out.returncode, out.stdout[-2000:], out.stderr[-2000:]))
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_trace(target, start_ns, now_ns, flags=None, state=None, nslots=0, cursor=0, want_slots=False):
    if len(now_ns) and max(int(max(now_ns)), int(max(start_ns))) >= 2**53:
        raise ValueError("JS harness carries total ns in a double: keep traces below 2^53 ns")
    job = {"target": target, "state": state, "nslots": nslots, "cursor": cursor, "want_slots": want_slots}
    expr = ("(function(){var job=%s; job.start_ns=__read_array(); job.now_ns=__read_array(); "
            "job.flags=__read_array(); return cb_run_trace(job);})()" % _lit(job))
    return call(expr, arrays=[start_ns, now_ns, flags])


def max_idle(target, state, now_ns):
    return call("cb_max_idle(%s,%s,%s)" % (_lit(target), _lit(state), _lit(now_ns)))


def build_slots(nbackends, nslots, draws):
    return call("cb_build_slots(%d,%d,__read_array())" % (nbackends, nslots), arrays=[draws])


def plan(conns, dead, target, mx, singleton=False):
    return call("cb_plan(%s,%s,%d,%d,%s)" % (_lit(conns), _lit(dead), target, mx, _lit(bool(singleton))))


def record_utils_tests():
    return call("cb_record_utils_tests()")


def gen_trace(seed, n, params):
    return call("cb_gen_trace(%d,%d,%s)" % (seed, n, _lit(params)))


def bench(seed, n, params, target, nslots, timeout=3600):
    return call("cb_bench(%d,%d,%s,%s,%d)" % (seed, n, _lit(params), _lit(target), nslots), timeout=timeout)


def pool_sim(target, nslots, ev_type, ev_a, ev_b, state=None, idle0=0, cursor=0, timeout=3600):
    """The event-driven pool model (cb_pool_sim in driver.js) on the reference's own
    ControlledDelay / Queue."""
    if len(ev_a) and max(int(x) for x in ev_a) >= 2**53:
        raise ValueError("JS harness carries total ns in a double: keep clocks below 2^53 ns")
    job = {"target": target, "nslots": nslots, "state": state, "idle0": idle0, "cursor": cursor}
    expr = ("(function(){var job=%s; job.ev_type=__read_array(); job.ev_a=__read_array(); "
            "job.ev_b=__read_array(); return cb_pool_sim(job);})()" % _lit(job))
    return call(expr, timeout=timeout, arrays=[ev_type, ev_a, ev_b])


def lpf_run(spares, mx, ty, a, b, c, d):
    """The pool's low-pass filter (lib/pool.js:44-100, verbatim) under the sampler and _rebalance's target
    arithmetic (cb_lpf_run in driver.js)."""
    job = {"spares": spares, "max": mx}
    expr = ("(function(){var job=%s; job.ty=__read_array(); job.a=__read_array(); job.b=__read_array(); "
            "job.c=__read_array(); job.d=__read_array(); return cb_lpf_run(job);})()" % _lit(job))
    return call(expr, arrays=[ty, a, b, c, d])


def queue_ops(op, arg):
    """A script of operations on the reference's own Queue (cb_queue_ops in driver.js): rows of
    (result, length, isEmpty)."""
    return call("cb_queue_ops(__read_array(),__read_array())", arrays=[op, arg])
