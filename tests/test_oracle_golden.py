"""The C restatement (oracle/liboracle.so) against the golden vectors produced by
the reference's own JavaScript (tests/golden/*.json, see gen_golden.py), and --
when oracle/_ref is present -- against live runs of the reference on fresh
random traces.  CPU only."""
import numpy as np
import pytest

import util
from oracle import oracle as O
from oracle import ref, tracegen

G = util.load_golden("codel_golden.json")
PLAN = util.load_golden("plan_rebalance_kat.json")
SLOTS = util.load_golden("slots_golden.json")


def run_oracle_case(case, variant):
    start, now, flags = util.make_trace(case["trace"])
    cd = O.new_codel(case["target"], util.state_unhex(case["init_state"]))
    decs, slots, states = [], [], []
    lo, cursor = 0, 0
    for b in case["batches"]:
        d, s, st = O.drain(cd, start[lo:lo + b], now[lo:lo + b], None if flags is None else flags[lo:lo + b],
                           nslots=case["nslots"], cursor=cursor, variant=variant)
        cursor = st.cursor
        decs.append(d)
        slots.append(s)
        states.append(cd.as_dict())
        lo += b
    return np.concatenate(decs) if decs else np.zeros(0, np.uint8), np.concatenate(slots), states, cursor


@pytest.mark.parametrize("variant", ["flat", "list"])
@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_drain_matches_reference(case, variant):
    dec, slots, states, cursor = run_oracle_case(case, variant)
    exp = case["expect"]
    assert util.codes_to_rle(dec) == exp["dec_rle"]
    assert util.crc_i32(slots) == exp["slots_crc32"]
    assert slots[:64].tolist() == exp["slots_head"]
    assert cursor == exp["cursor"]
    assert int((dec == 1).sum()) == exp["drops"] and int((dec == 0).sum()) == exp["admits"]
    for got, want in zip(states, exp["states"]):
        util.assert_state_equal(got, util.state_unhex(want), case["name"])


def test_ms_variant_equals_ns_variant():
    case = next(c for c in G["cases"] if c["name"] == "gen_many_episodes")
    start, now, _ = util.make_trace(case["trace"])
    to_ms = np.vectorize(O.ns_to_ms, otypes=[np.float64])
    cd = O.new_codel(case["target"])
    dec, _, _ = O.drain(cd, to_ms(start), to_ms(now), nslots=case["nslots"])
    assert util.codes_to_rle(dec) == case["expect"]["dec_rle"]


@pytest.mark.parametrize("case", G["max_idle"], ids=lambda c: "t%s_%s" % (c["target"], c["now_ns"]))
def test_max_idle(case):
    cd = O.new_codel(case["target"], util.state_unhex(case["state"]))
    assert util.same_bits(O.max_idle(cd, O.ns_to_ms(case["now_ns"])), util.unhx(case["expect"]))


def _plan_via_oracle(c):
    names = [k for k, _ in c["conns"]]
    have = [len(v) for _, v in c["conns"]]
    dead = [1 if k in c["dead"] else 0 for k in names]
    want, add, rem = O.plan_rebalance(have, dead, c["target"], c["max"], c["singleton"])
    conns = dict((k, v) for k, v in c["conns"])
    return [names[i] for i in add], [conns[names[k]][o] for k, o in rem]


@pytest.mark.parametrize("c", PLAN["kat"], ids=[c["test"] for c in PLAN["kat"]])
def test_plan_rebalance_reference_kat(c):
    """The reference's own known-answer tests (test/utils.test.js:13-285)."""
    add, rem = _plan_via_oracle(c)
    assert add == c["add"]
    assert rem == c["remove"]


def test_plan_rebalance_random_vs_reference():
    assert len(PLAN["random"]) >= 400
    for c in PLAN["random"]:
        add, rem = _plan_via_oracle(c)
        assert add == c["add"] and rem == c["remove"], c


def test_plan_rebalance_asserts():
    with pytest.raises(AssertionError):
        O.plan_rebalance([0], [0], 5, 4)
    with pytest.raises(AssertionError):
        O.plan_rebalance([0], [0], -1, 4)


@pytest.mark.parametrize("c", SLOTS["build"], ids=lambda c: "B%d_S%d" % (c["B"], c["S"]))
def test_slot_table(c):
    keys, slots = O.build_slots(c["B"], c["S"], tracegen.gen_insert_draws(c["B"], c["seed"]))
    assert keys.tolist() == c["keys"]
    assert slots.tolist() == c["slots"]


def test_reshuffle_and_shuffle():
    for c in SLOTS["reshuffle"]:
        keys = np.arange(c["n"], dtype=np.int32)
        for u, want in zip(tracegen.gen_insert_draws(8, c["seed"]), c["trail"]):
            keys = O.reshuffle(keys, u)
            assert keys.tolist() == want
    for c in SLOTS["shuffle"]:
        got = O.shuffle(np.arange(c["n"]), tracegen.gen_insert_draws(c["n"], c["seed"]))
        assert got.tolist() == c["result"]


def test_current_millis_rounding():
    # secs*1000 + ns/1e6 with three roundings (lib/utils.js:198-204)
    assert O.ns_to_ms(1_000_000_000) == 1000.0
    assert O.ns_to_ms(1_308_235_026) == 1000.0 + 308235026 / 1e6
    t = 200 * 86400 * 10**9 + 123_456_789
    assert O.ns_to_ms(t) == float((t // 10**9) * 1000) + (t % 10**9) / 1e6


@pytest.mark.ref
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [101, 202, 303])
def test_live_reference_random_traces(seed):
    """Fresh seeds, reference JS vs C restatement, incl. flags and carried state."""
    rng = np.random.RandomState(seed)
    n = 12_000
    start, now = tracegen.gen_trace(n, seed=seed, de=int(rng.choice([2_000, 90_000, 600_000])),
                                    unit=int(rng.choice([64, 512, 4096])), target_ns=2_000_000)
    flags = util.make_flags(n, seed, cancel_mod=int(rng.choice([0, 3, 13])), empty_mod=int(rng.choice([0, 97])))
    if seed == 303:  # clock jitter backwards
        now = now - (rng.rand(n) < 0.01).astype(np.int64) * 120_000_000
    cut = int(rng.randint(1, n))
    cd = O.new_codel(2)
    st, cursor = None, 0
    for lo, hi in ((0, cut), (cut, n)):
        r = ref.run_trace(2, start[lo:hi], now[lo:hi], flags[lo:hi], state=st, nslots=37, cursor=cursor, want_slots=True)
        d, s, stats = O.drain(cd, start[lo:hi], now[lo:hi], flags[lo:hi], nslots=37, cursor=cursor)
        assert "".join(util.DEC_CHARS[x] for x in d) == r["dec"]
        assert s.tolist() == r["slots"]
        util.assert_state_equal(cd.as_dict(), r["state"], "seed %d" % seed)
        st, cursor = r["state"], r["cursor"]
        assert stats.cursor == cursor


@pytest.mark.ref
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_generators_agree_js_numpy():
    p = dict(t0=10**9, de=1000, target_ns=5_000_000, unit=64, start_overloaded=True, all_overloaded_from=2500)
    t = ref.gen_trace(tracegen.SEED, 4000, p)
    s, n = tracegen.gen_trace(4000, **p)
    assert t["start_ns"] == s.tolist() and t["now_ns"] == n.tolist()
