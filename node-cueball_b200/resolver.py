"""Resolver side of the path: backend identity and the 'added'/'removed'
event stream that feeds ConnectionPool.on_resolver_added/removed; names and behaviour follow the
reference.  The functions below are the host mirror; the same two steps sit behind the C ABI for a native
host (cb_srv_keys: SHA-1 of a whole result set on the device; cb_resolver_process: the diff) and are
reached from here through DeviceResolverState:

    srvKey(srv)                    lib/resolver.js:1157-1171  sha1(name || port || normalised address), base64
    StaticIpResolver               lib/resolver.js:1387-1456  fixed backend list; emission order = list order
    process_backends(old, srvs)    lib/resolver.js:1024-1108  state_process: diff of SRV/A results into
                                                              'removed' (old key order) then 'added' (new key order)
"""
import base64
import hashlib
import ipaddress


def _normalised(address):
    """ipaddr.js (>=1.1 <2, package.json:12) toNormalizedString(): IPv4 dotted quad; IPv6 eight
    groups in lowercase hex without leading zeros and WITHOUT '::' compression."""
    ip = ipaddress.ip_address(address)
    if ip.version == 4:
        return str(ip)
    v = int(ip)
    return ":".join("%x" % ((v >> (16 * (7 - k))) & 0xFFFF) for k in range(8))


def srvKey(srv):  # lib/resolver.js:1157-1171
    h = hashlib.sha1()
    h.update(str(srv["name"]).encode())
    h.update(b"||")
    h.update(str(srv["port"]).encode())
    h.update(b"||")
    h.update(_normalised(srv["address"]).encode())
    return base64.b64encode(h.digest()).decode()


def _is_ip(s):
    try:
        ipaddress.ip_address(s)
        return True
    except ValueError:
        return False


class _Emitter(object):
    def __init__(self):
        self._handlers = {}

    def on(self, event, fn):
        self._handlers.setdefault(event, []).append(fn)
        return self

    def emit(self, event, *args):
        for fn in list(self._handlers.get(event, [])):
            fn(*args)


class StaticIpResolver(_Emitter):
    """lib/resolver.js:1387-1456.  The reference wraps it in a CueBallResolver FSM whose 'running'
    state is entered after the initial 'updated'; here start() emits synchronously and
    isInState('running') turns true afterwards."""

    def __init__(self, options=None):
        _Emitter.__init__(self)
        if not isinstance(options, dict):
            raise AssertionError("options (object) is required")
        dp = options.get("defaultPort")
        if dp is not None and not isinstance(dp, (int, float)):
            raise AssertionError("options.defaultPort (number) is required")
        backends = options.get("backends")
        if not isinstance(backends, list) or any(not isinstance(b, dict) for b in backends):
            raise AssertionError("options.backends ([object]) is required")
        self.sr_backends = []
        for i, b in enumerate(backends):
            addr = b.get("address")
            if not isinstance(addr, str):
                raise AssertionError("options.backends[%d].address (string) is required" % i)
            if not _is_ip(addr):
                raise AssertionError("options.backends[%d].address must be an IP address" % i)
            port = b.get("port")
            if port is None:
                port = dp
            if isinstance(port, bool) or not isinstance(port, (int, float)):
                raise AssertionError("options.backends[%d].port (number) is required" % i)
            self.sr_backends.append({"name": "%s:%s" % (addr, port), "address": addr, "port": port})
        self.sr_state = "idle"
        self._running = False

    def start(self):
        assert self.sr_state == "idle", "cannot call start() again without calling stop()"
        self.sr_state = "started"
        for be in self.sr_backends:
            self.emit("added", srvKey(be), be)
        self.emit("updated")
        self._running = True
        self.emit("stateChanged", "running")

    def stop(self):
        assert self.sr_state == "started", "cannot call stop() again without calling start()"
        self.sr_state = "idle"
        self._running = False

    def isInState(self, st):  # the wrapping CueBallResolver FSM's states, reduced to the two this mirror visits
        return st == ("running" if self._running else "stopped")

    def count(self):
        return len(self.sr_backends)

    def list(self):
        return dict((srvKey(be), be) for be in self.sr_backends)

    def getLastError(self):
        return None

    def attach(self, pool):
        """Feed a ConnectionPool mirror: lib/pool.js:319-320 (listeners), :345-355 (initial list)."""
        self.on("added", lambda k, be: pool.on_resolver_added(k, be))
        self.on("removed", lambda k: pool.on_resolver_removed(k))
        return self


def process_backends(old_backends, srvs):
    """CueBallDNSResolver#state_process, lib/resolver.js:1024-1108: `srvs` is the list of
    {name, port, addresses_v6, addresses_v4}; v6 addresses come first.  Returns
    (new_backends, removed_keys, added_keys); the reference emits every 'removed' (in the OLD map's
    key order) before every 'added' (in the NEW map's key order).  An empty result keeps the old map
    (the reference reports 'empty-set' and changes nothing)."""
    new_backends = {}
    for srv in srvs:
        addrs = list(srv.get("addresses_v6") or []) + list(srv.get("addresses_v4") or [])
        for addr in addrs:
            final = {"name": srv["name"], "port": srv["port"], "address": addr}
            new_backends[srvKey(final)] = final
    if not new_backends:
        return old_backends, [], []
    removed = [k for k in old_backends if k not in new_backends]
    added = [k for k in new_backends if k not in old_backends]
    return new_backends, removed, added


class DeviceResolverState(object):
    """state_process through the C ABI: the srvKeys of a result set are computed on the device
    (cb_srv_keys) and diffed against the context's backend map (cb_resolver_process); the events go to a
    ConnectionPool mirror in the reference's order (every 'removed', then every 'added')."""

    def __init__(self, ctx):
        self._ctx = ctx

    def process(self, srvs):
        """`srvs` as for process_backends.  Returns (removed_keys, added [(key, backend)], empty_set)."""
        rows = []
        for srv in srvs:
            for addr in list(srv.get("addresses_v6") or []) + list(srv.get("addresses_v4") or []):
                rows.append({"name": srv["name"], "port": srv["port"], "address": addr})
        keys = self._ctx.srv_keys(rows)
        removed, added, empty = self._ctx.resolver_process(keys)
        return [k for k, _ in removed], [(k, rows[row]) for k, _, row in added], empty

    def feed(self, pool, srvs):
        removed, added, empty = self.process(srvs)
        for k in removed:
            pool.on_resolver_removed(k)
        for k, be in added:
            pool.on_resolver_added(k, be)
        return removed, [k for k, _ in added], empty
