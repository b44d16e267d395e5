"""The host mirror of the reference's pending-claim FIFO (node-cueball_b200/queue.py: push -> node.remove(), shift,
forEach, peek, length, isEmpty) against operation scripts recorded from the reference's OWN lib/queue.js
(tests/golden/queue_golden.json), and against live reference runs where oracle/_ref exists."""
import pytest

import util
from oracle import ref

G = util.load_golden("queue_golden.json")


def run_mirror(op, arg):
    from cueball_b200 import Queue
    q, nodes, linked, out = Queue(), [], [], []
    for o, a in zip(op, arg):
        r = -2
        if o == 0:
            nodes.append(q.push(a)); linked.append(True); r = len(nodes) - 1   # noqa: E702
        elif o == 1:
            if q.length > 0:
                r = q.shift()
                linked[[n.n_value for n in nodes].index(r)] = False
        elif o == 2:
            if a < len(nodes) and linked[a]:
                nodes[a].remove(); linked[a] = False; r = 1                     # noqa: E702
        elif o == 3:
            seen = {"vals": 0, "removed": 0}

            def cb(v, n):
                seen["vals"] += 1
                if a > 0 and v % a == 0:
                    n.remove()
                    seen["removed"] += 1
                    linked[nodes.index(n)] = False
            q.forEach(cb)
            r = seen["vals"] + 1000 * seen["removed"]
        elif o == 4:
            p = q.peek()
            r = -1 if p is None else p
        out += [r, q.length, 1 if q.isEmpty() else 0]
    return out


@pytest.mark.parametrize("case", G["cases"], ids=[str(c["seed"]) for c in G["cases"]])
def test_queue_mirror_on_golden_scripts(case):
    assert run_mirror(case["op"], case["arg"]) == case["out"]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_queue_mirror_against_live_reference():
    op, arg = util.gen_queue_ops(5101, 800)
    assert run_mirror(op, arg) == ref.queue_ops(op, arg)


def test_shift_on_an_empty_queue_is_the_callers_bug():
    """lib/queue.js:57-62 has no empty-queue guard (callers check length first); the mirror
    raises instead of corrupting itself."""
    from cueball_b200 import Queue
    with pytest.raises(IndexError):
        Queue().shift()


# utils.assertClaimDelay (lib/utils.js:188-195): outcomes recorded from the reference's own function under the
# oracle engine (with the engine-side assert-plus stand-in of oracle/js/shim.js)
CLAIM_DELAY_CASES = [("undefined", None, "ok"), ("null", None, "ok"), ("5", 5, "ok"), ("1", 1, "ok"), ("0", 0, "throws"),
                     ("-1", -1, "throws"), ("2.5", 2.5, "throws"), ("NaN", float("nan"), "throws"),
                     ("Infinity", float("inf"), "throws"), ("'5'", "5", "throws"), ("true", True, "throws"),
                     ("1e15", 1e15, "ok"), ("300", 300, "ok")]


@pytest.mark.parametrize("js,py,want", CLAIM_DELAY_CASES, ids=[c[0] for c in CLAIM_DELAY_CASES])
def test_assert_claim_delay(js, py, want):
    from cueball_b200 import assertClaimDelay
    if ref.available():
        got = ref.call("(function(){try{utils.assertClaimDelay(%s);return 'ok';}catch(e){return 'throws';}})()" % js)
        assert got == want
    if want == "ok":
        assertClaimDelay(py)
    else:
        with pytest.raises(AssertionError):
            assertClaimDelay(py)
