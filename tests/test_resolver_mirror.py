"""Host mirror of the resolver side: the reference's own static-resolver tests
(test/resolver_static.test.js) restated, srvKey against an independent SHA-1, and the added/removed
diffing of CueBallDNSResolver#state_process (lib/resolver.js:1024-1108)."""
import base64
import hashlib
import re

import pytest

import util  # noqa: F401
import cueball_b200 as cb


def test_static_resolver_bad_arguments():  # test/resolver_static.test.js:17-92
    R = cb.StaticIpResolver
    for args, pat in [((), r"options"), (({},), r"options.backends"), (({"backends": None},), r"options.backends"),
                      (({"backends": [None]},), r"options.backends"),
                      (({"backends": [{"address": "127.0.0.1", "port": 1234}, {}]},), r"options.backends\[1\].address"),
                      (({"backends": [{"address": "127.0.0.1", "port": 1234}, {"address": 1234, "port": "foobar"}]},),
                       r"options.backends\[1\].address"),
                      (({"backends": [{"address": "127.0.0.1", "port": 1234}, {"address": "127.0.0.1"}]},),
                       r"options.backends\[1\].port"),
                      (({"backends": [{"address": "127.0.0.1", "port": 1234}, {"address": "127.0.0.1", "port": "foobar"}]},),
                       r"options.backends\[1\].port")]:
        with pytest.raises(AssertionError) as e:
            R(*args)
        assert re.search(pat, str(e.value)), (args, str(e.value))


def test_static_resolver_no_backends():  # test/resolver_static.test.js:94-111
    r = cb.StaticIpResolver({"backends": []})
    nadded = []
    r.on("added", lambda k, b: nadded.append(k))
    r.start()
    assert r.isInState("running") and nadded == [] and r.list() == {} and r.count() == 0
    r.stop()


@pytest.mark.parametrize("options,expected", [
    ({"defaultPort": 2021, "backends": [{"address": "10.0.0.3", "port": 2022}, {"address": "10.0.0.4"}, {"address": "10.0.0.5"}]},
     [("10.0.0.3:2022", "10.0.0.3", 2022), ("10.0.0.4:2021", "10.0.0.4", 2021), ("10.0.0.5:2021", "10.0.0.5", 2021)]),
    ({"backends": [{"address": "10.0.0.3", "port": 2021}, {"address": "10.0.0.3", "port": 2020}, {"address": "10.0.0.7", "port": 2020}]},
     [("10.0.0.3:2021", "10.0.0.3", 2021), ("10.0.0.3:2020", "10.0.0.3", 2020), ("10.0.0.7:2020", "10.0.0.7", 2020)]),
], ids=["defaultPort", "several_backends"])
def test_static_resolver_emission_order_and_naming(options, expected):  # test/resolver_static.test.js:113-221
    r = cb.StaticIpResolver(options)
    found, keys = [], []
    r.on("added", lambda k, b: (keys.append(k), found.append(b)))
    r.start()
    assert r.count() == len(expected)
    assert found == [{"name": n, "address": a, "port": p} for n, a, p in expected]
    listed = r.list()
    assert list(listed.keys()) == keys and sorted(b["name"] for b in listed.values()) == sorted(n for n, _, _ in expected)
    r.stop()
    with pytest.raises(AssertionError):
        r.stop()


def test_srvkey_is_sha1_of_name_port_normalised_address():  # lib/resolver.js:1157-1171
    def indep(name, port, addr):
        return base64.b64encode(hashlib.sha1(("%s||%s||%s" % (name, port, addr)).encode()).digest()).decode()
    assert cb.srvKey({"name": "10.0.0.3:2022", "port": 2022, "address": "10.0.0.3"}) == indep("10.0.0.3:2022", 2022, "10.0.0.3")
    # ipaddr.js toNormalizedString(): no '::' compression, no leading zeros, lowercase
    k = cb.srvKey({"name": "svc.example.com", "port": 80, "address": "2001:DB8::0001"})
    assert k == indep("svc.example.com", 80, "2001:db8:0:0:0:0:0:1")
    assert k == cb.srvKey({"name": "svc.example.com", "port": 80, "address": "2001:0db8:0:0:0:0:0:1"})
    assert len(k) == 28 and k.endswith("=")  # 20 bytes in base64
    assert cb.srvKey({"name": "a", "port": 1, "address": "10.0.0.1"}) != cb.srvKey({"name": "a", "port": 2, "address": "10.0.0.1"})


def test_process_backends_diff_order():  # lib/resolver.js:1024-1108
    srvs1 = [{"name": "a.example", "port": 80, "addresses_v4": ["10.0.0.1", "10.0.0.2"], "addresses_v6": ["fd00::1"]},
             {"name": "b.example", "port": 81, "addresses_v4": ["10.0.0.3"]}]
    new1, removed, added = cb.process_backends({}, srvs1)
    assert removed == [] and added == list(new1.keys())
    # v6 addresses come first within a record (lib/resolver.js:1032-1034)
    assert [new1[k]["address"] for k in added] == ["fd00::1", "10.0.0.1", "10.0.0.2", "10.0.0.3"]
    srvs2 = [{"name": "b.example", "port": 81, "addresses_v4": ["10.0.0.3", "10.0.0.9"]},
             {"name": "a.example", "port": 80, "addresses_v4": ["10.0.0.2"]}]
    new2, removed, added = cb.process_backends(new1, srvs2)
    assert [new1[k]["address"] for k in removed] == ["fd00::1", "10.0.0.1"]  # old-map key order
    assert [new2[k]["address"] for k in added] == ["10.0.0.9"]               # new-map key order
    # nothing resolved: the reference keeps the old set and reports 'empty-set'
    same, removed, added = cb.process_backends(new2, [])
    assert same is new2 and removed == [] and added == []
