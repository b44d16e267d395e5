"""Resolver ingestion behind the C ABI: cb_srv_keys (SHA-1 on the device) against
hashlib through the host mirror's srvKey (lib/resolver.js:1157-1171), and cb_resolver_process against the
host mirror of CueBallDNSResolver#state_process (lib/resolver.js:1024-1108) over a sequence of result sets
with duplicates, re-orderings and an empty set."""
import numpy as np
import pytest

import util  # noqa: F401

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def cb():
    import cueball_b200
    return cueball_b200


def rand_backend(rng, longname=False):
    if rng.rand() < 0.4:
        addr = ":".join("%x" % rng.randint(0, 65536) for _ in range(8))
        if rng.rand() < 0.3:
            addr = "::ffff:%d.%d.%d.%d" % tuple(rng.randint(0, 256, size=4))
        elif rng.rand() < 0.3:
            addr = "fd00::%x" % rng.randint(1, 65536)
    else:
        addr = "%d.%d.%d.%d" % tuple(rng.randint(0, 256, size=4))
    n = rng.randint(120, 250) if longname else rng.randint(1, 50)
    name = "".join(chr(rng.randint(97, 123)) for _ in range(n)) + ".svc.example.com"
    return {"name": name, "port": int(rng.randint(1, 65536)), "address": addr}


def test_srv_keys_match_sha1_base64(cb):
    rng = np.random.RandomState(3)
    ctx = cb.Context(0)
    backends = [rand_backend(rng, longname=(k % 7 == 0)) for k in range(4096)]
    backends += [{"name": "", "port": 0, "address": "0.0.0.0"}, {"name": "x" * 46, "port": 80, "address": "10.0.0.1"},
                 {"name": "y" * 47, "port": 80, "address": "10.0.0.1"}]  # around the one-block / two-block boundary
    got = ctx.srv_keys(backends)
    assert got == [cb.srvKey(b) for b in backends]
    with pytest.raises(cb.CueballError):
        ctx.srv_keys([{"name": "a", "port": 1, "address": "not-an-ip"}])
    ctx.close()


def test_resolver_process_matches_state_process(cb):
    rng = np.random.RandomState(4)
    ctx = cb.Context(0)
    dev = cb.DeviceResolverState(ctx)
    pool_b = [rand_backend(rng) for _ in range(300)]
    old = {}
    ids = {}
    for step in range(12):
        if step == 5:
            srvs = []  # empty result: 'empty-set', nothing changes
        else:
            pick = [pool_b[i] for i in rng.choice(len(pool_b), size=rng.randint(1, 200), replace=True)]  # with duplicates
            srvs = [{"name": b["name"], "port": b["port"],
                     "addresses_v6": [b["address"]] if ":" in b["address"] else [],
                     "addresses_v4": [b["address"]] if ":" not in b["address"] else []} for b in pick]
        new, removed, added = cb.process_backends(old, srvs)
        g_removed, g_added, empty = dev.process(srvs)
        assert empty == (not srvs)
        assert g_removed == removed
        assert [k for k, _ in g_added] == added
        for k, be in g_added:
            assert be == new[k] or cb.srvKey(be) == k
        old = new
    ctx.close()


def test_utils_shuffle_kernel(cb):
    """utils.shuffle (lib/utils.js:207-217) through cb_shuffle: the results recorded from
    the reference engine (tests/golden/slots_golden.json), random arrays against the oracle's loop at sizes
    round the CTA (1024) and staging-tile (2048) edges and beyond one wave of CTAs, and the reference-shaped
    wrapper utils.shuffle on a list of objects."""
    from cueball_b200 import tracegen
    from cueball_b200 import utils as U
    from oracle import oracle as O

    ctx = cb.Context(0)
    slots = util.load_golden("slots_golden.json")
    for c in slots["shuffle"]:
        us = tracegen.gen_insert_draws(c["n"], c["seed"])
        assert ctx.shuffle(np.arange(c["n"]), us).tolist() == c["result"]
    rng = np.random.RandomState(2024)
    for n in (0, 1, 2, 3, 1023, 1024, 1025, 2048, 2049, 3073, 40_000, 400_000):
        us = rng.rand(n)
        if n > 4:
            us[rng.randint(n)] = 0.0
            us[rng.randint(n)] = np.nextafter(1.0, 0.0)
        arr = rng.randint(-2**31, 2**31 - 1, size=n, dtype=np.int64).astype(np.int32)
        assert np.array_equal(ctx.shuffle(arr, us), O.shuffle(arr, us)), n
    names = ["b%d" % k for k in range(257)]
    us = rng.rand(len(names))
    want = [names[i] for i in O.shuffle(np.arange(len(names)), us)]
    got = U.shuffle(list(names), draws=us.tolist(), context=ctx)
    assert got == want
    with pytest.raises(cb.CueballError):
        ctx.shuffle(np.arange(4), [0.1, 0.2, 1.0, 0.3])
