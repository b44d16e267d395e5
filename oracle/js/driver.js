/*
 * oracle/js/driver.js -- TEST INFRASTRUCTURE (parity oracle), ours, not reference code.
 *
 * Harness around the reference's own lib/codel.js, lib/queue.js, lib/utils.js
 * (loaded verbatim from oracle/_ref/bundle.js).  What lives in FSM closures in
 * the reference and therefore cannot be imported is restated here, each piece
 * citing the lines it follows:
 *
 *   cb_run_trace    the slot-idle drain loop           lib/pool.js:733-753
 *                   handle creation (ch_started)       lib/connection-fsm.js:523
 *                   enqueue                            lib/pool.js:960
 *   cb_pool_sim     claim() enqueue side + idle-first   lib/pool.js:874-885,929-966
 *                   the slot-idle handler, k slots      lib/pool.js:733-757
 *                   the handle's hard timeout            lib/connection-fsm.js:626-642
 *                   counters                             lib/pool.js:269-283,887,962-963
 *   cb_lpf_run      the pool's 5 Hz sampler body        lib/pool.js:251-258 (setInterval closure)
 *                   and _rebalance's target sizing     lib/pool.js:556-587, around the reference's
 *                   own FIRFilter / LP_TAPS            lib/pool.js:44-100 (verbatim, module ./pool_lpf)
 *   cb_queue_ops    a script of push / shift / node.remove / forEach / peek on the reference's own
 *                   Queue (lib/queue.js, verbatim): pins the host mirror node-cueball_b200/queue.py
 *   cb_build_slots  preference-list random insertion   lib/pool.js:285-291
 *                   (Math.random replaced by injected draws) then the
 *                   reference's own planRebalance      lib/utils.js:239-393
 *
 * Batch model (SURVEY.md A.2/A.4): claim i is dequeued at clock now_ns[i];
 * flags bit0 = handle still 'waiting', bit1 = the waiter queue ran empty with a
 * slot idle right after this claim (pool.js:751-753 -> codel.empty()).
 */
var ControlledDelay = __mods['./codel'].ControlledDelay;
var Queue = __mods['./queue'];
var utils = __mods['./utils'];
var pool_lpf = __mods['./pool_lpf'];

function __json(v) {
	var t = typeof (v), i, r;
	if (v === null || v === undefined) return 'null';
	if (t === 'number') return isFinite(v) ? String(v) : ('"' + String(v) + '"');
	if (t === 'boolean') return v ? 'true' : 'false';
	if (t === 'string') return '"' + v.replace(/\\/g, '\\\\').replace(/"/g, '\\"').replace(/\n/g, '\\n') + '"';
	if (v instanceof Array) {
		r = [];
		for (i = 0; i < v.length; i++) r.push(__json(v[i]));
		return '[' + r.join(',') + ']';
	}
	r = [];
	for (i in v) if (v.hasOwnProperty(i)) r.push(__json(i) + ':' + __json(v[i]));
	return '{' + r.join(',') + '}';
}

/*
 * Bulk numeric input comes over stdin (readline), never as script literals:
 * SM 1.7 mis-indexes atoms once a single script holds > 65535 literals.
 * Format: one line with the element count (or "null"), then lines of <= 1000
 * comma-separated decimal numbers.
 */
function __read_array() {
	var hdr = readline();
	if (hdr === 'null') return null;
	var n = parseInt(hdr, 10), out = new Array(n), k = 0;
	while (k < n) {
		var parts = readline().split(',');
		for (var j = 0; j < parts.length; j++) out[k++] = Number(parts[j]);
	}
	return out;
}

function cb_state(cd) {
	return {
		count: cd.cd_count, drop_next: cd.cd_drop_next,
		first_above_time: cd.cd_first_above_time, dropping: cd.cd_dropping,
		last_empty: (cd.cd_last_empty === undefined) ? null : cd.cd_last_empty
	};
}
function cb_set_state(cd, st) {
	cd.cd_count = st.count; cd.cd_drop_next = st.drop_next;
	cd.cd_first_above_time = st.first_above_time; cd.cd_dropping = st.dropping;
	if (st.last_empty !== null && st.last_empty !== undefined)
		cd.cd_last_empty = st.last_empty;
}

/*
 * job: { target, start_ns:[], now_ns:[], flags:[]|null, state:{}|null,
 *        nslots, cursor, want_slots:bool }
 * Clocks are total nanoseconds (exact integers < 2^53) fed through the fake
 * process.hrtime(), so the reference's currentMillis() does the ms conversion.
 */
function cb_run_trace(job) {
	var cd = new ControlledDelay(job.target);
	if (job.state) cb_set_state(cd, job.state);
	var waiters = new Queue();
	var n = job.start_ns.length, i;
	/* enqueue side: pool.js:914 (handle ctor stamps ch_started) + pool.js:960 */
	for (i = 0; i < n; i++) {
		__set_clock_ns(job.start_ns[i]);
		waiters.push({
			ch_started: utils.currentMillis(),		/* connection-fsm.js:523 */
			waiting: job.flags ? ((job.flags[i] & 1) !== 0) : true,
			idx: i
		});
	}
	var dec = [], slots = [], cursor = job.cursor || 0;
	var drops = 0, admits = 0, skipped = 0;
	/* dequeue side: pool.js:733-753, one clock value per dequeued claim */
	while (waiters.length > 0) {
		var hdl = waiters.shift();				/* pool.js:734 */
		i = hdl.idx;
		__set_clock_ns(job.now_ns[i]);
		var drop = cd.overloaded(hdl.ch_started);		/* pool.js:735-736 */
		if (!hdl.waiting) {					/* pool.js:738-740 */
			dec.push('S'); slots.push(-1); skipped++;
		} else if (drop) {					/* pool.js:742-745 */
			dec.push('D'); slots.push(-1); drops++;
		} else {						/* pool.js:747 */
			dec.push('A');
			slots.push(job.nslots > 0 ? (cursor % job.nslots) : -1);
			cursor++; admits++;
		}
		if (job.flags && (job.flags[i] & 2))
			cd.empty();					/* pool.js:751-753 */
	}
	var out = { dec: dec.join(''), state: cb_state(cd), drops: drops,
	    admits: admits, skipped: skipped, cursor: cursor };
	if (job.want_slots) out.slots = slots;
	return out;
}

/* getMaxIdle at clock now_ns for a given state (lib/codel.js:109-118, verbatim call) */
function cb_max_idle(target, state, now_ns) {
	var cd = new ControlledDelay(target);
	if (state) cb_set_state(cd, state);
	__set_clock_ns(now_ns);
	return cd.getMaxIdle();
}

/*
 * Preference list + slot table.  us[k] are the injected uniform draws standing
 * in for Math.random() in on_resolver_added (lib/pool.js:287).
 * Returns { keys: preference order (backend indices), slots: backend index per slot }.
 */
function cb_build_slots(B, S, us) {
	var keys = [], k;
	for (k = 0; k < B; k++) {
		var idx = Math.floor(us[k] * (keys.length + 1));	/* pool.js:287 */
		keys.splice(idx, 0, 'b' + k);				/* pool.js:288 */
	}
	var conns = {};
	for (k = 0; k < keys.length; k++) conns[keys[k]] = [];		/* pool.js:562-565 */
	var plan = utils.planRebalance(conns, {}, S, S);		/* pool.js:590 */
	var strip = function (s) { return parseInt(s.substring(1), 10); };
	var ko = [], so = [];
	for (k = 0; k < keys.length; k++) ko.push(strip(keys[k]));
	for (k = 0; k < plan.add.length; k++) so.push(strip(plan.add[k]));	/* pool.js:632-654 order */
	return { keys: ko, slots: so };
}

/* reshuffle (lib/pool.js:501-519): pop last, random re-insert; u = injected draw */
function cb_reshuffle(keys, u) {
	if (keys.length <= 1) return keys;
	var taken = keys.pop();
	var idx = Math.floor(u * (keys.length + 1));
	keys.splice(idx, 0, taken);
	return keys;
}

/* generic planRebalance call: conns = [[key, [conn...]]...] in key order */
function cb_plan(conns, dead, target, max, singleton) {
	var o = {}, d = {}, i;
	for (i = 0; i < conns.length; i++) o[conns[i][0]] = conns[i][1];
	for (i = 0; i < dead.length; i++) d[dead[i]] = true;
	return utils.planRebalance(o, d, target, max, singleton);
}

/* utils.shuffle (lib/utils.js:207-217) with injected draws */
function cb_shuffle(arr, us) {
	var k = 0, saved = Math.random;
	Math.random = function () { return us[k++]; };
	var r = utils.shuffle(arr);
	Math.random = saved;
	return r;
}

/*
 * Run the reference's own test/utils.test.js under a recording tape stand-in:
 * every planRebalance call is logged with its arguments and result, and every
 * t.deepEqual / t.equal / t.ok outcome is counted.  The log becomes
 * tests/golden/plan_rebalance_kat.json.
 */
function cb_record_utils_tests() {
	var log = [], pass = 0, fail = 0, cur = null;
	var real = utils.planRebalance;
	function deq(a, b) { return __json(a) === __json(b); }
	utils.planRebalance = function (c, d, t, m, s) {
		var conns = [], k, dead = [];
		for (k in c) if (c.hasOwnProperty(k)) conns.push([k, c[k].slice()]);
		for (k in d) if (d.hasOwnProperty(k) && d[k] === true) dead.push(k);
		var plan = real(c, d, t, m, s);
		log.push({ test: cur, conns: conns, dead: dead, target: t, max: m,
		    singleton: !!s, add: plan.add.slice(), remove: plan.remove.slice() });
		return plan;
	};
	var T = {
		deepEqual: function (a, b) { if (deq(a, b)) pass++; else { fail++; print('FAIL ' + cur + ': ' + __json(a) + ' != ' + __json(b)); } },
		equal: function (a, b) { if (a == b) pass++; else fail++; },
		strictEqual: function (a, b) { if (a === b) pass++; else fail++; },
		ok: function (a) { if (a) pass++; else fail++; },
		notOk: function (a) { if (!a) pass++; else fail++; },
		throws: function (f) { try { f(); fail++; } catch (e) { pass++; } },
		end: function () {}
	};
	__mods['tape'] = { test: function (name, f) { cur = name; f(T); } };
	__tests['utils.test']();
	utils.planRebalance = real;
	return { pass: pass, fail: fail, calls: log };
}

/* ---- seeded synthetic trace generator (identical to oracle/tracegen.py) ---- */
function imul32(a, b) {
	var ah = (a >>> 16) & 0xffff, al = a & 0xffff;
	var bh = (b >>> 16) & 0xffff, bl = b & 0xffff;
	return ((al * bl) + (((ah * bl + al * bh) << 16) >>> 0)) >>> 0;
}
function hash32(x) {	/* lowbias32 */
	x = x >>> 0;
	x ^= x >>> 16; x = imul32(x, 0x7feb352d);
	x ^= x >>> 15; x = imul32(x, 0x846ca68b);
	x ^= x >>> 16;
	return x >>> 0;
}
/*
 * G(seed, n, p): p = { t0, de, target_ns, unit (claims per phase unit),
 *                      start_overloaded (bool), all_overloaded_from (index|-1) }
 * see oracle/tracegen.py for the specification.
 */
function cb_gen_trace(seed, n, p) {
	var start = new Array(n), now = new Array(n);
	var i = 0, ph = 0, overloaded = !!p.start_overloaded, prev = 0;
	var healthy_span = Math.floor(p.target_ns * 9 / 10), over_span = p.target_ns * 60;
	while (i < n) {
		var r = hash32((seed ^ imul32(ph + 1, 0x85ebca6b)) >>> 0);
		var len = overloaded ? (2 + (r % 8)) * p.unit : (1 + (r % 4)) * p.unit;
		var end = i + len;
		if (p.all_overloaded_from >= 0 && i >= p.all_overloaded_from) { overloaded = true; end = n; }
		else if (p.all_overloaded_from >= 0 && end > p.all_overloaded_from) end = p.all_overloaded_from;
		if (end > n) end = n;
		for (; i < end; i++) {
			var h = hash32((seed ^ imul32(i + 1, 0x9e3779b9)) >>> 0);
			var s = p.t0 + i * p.de;
			var soj = overloaded ? (p.target_ns + (h % over_span)) : (h % healthy_span);
			var t = s + soj;
			if (t < prev) t = prev;
			start[i] = s; now[i] = t; prev = t;
		}
		overloaded = !overloaded; ph++;
	}
	return { start_ns: start, now_ns: now };
}
function cb_trace_checksum(tr) {
	/* order-sensitive 32-bit checksum over both arrays (lo and hi words) */
	var c = 0x811c9dc5, i, n = tr.start_ns.length;
	for (i = 0; i < n; i++) {
		var a = tr.start_ns[i], b = tr.now_ns[i];
		var alo = a % 4294967296, ahi = Math.floor(a / 4294967296);
		var blo = b % 4294967296, bhi = Math.floor(b / 4294967296);
		c = hash32((c ^ alo) >>> 0); c = hash32((c ^ ahi) >>> 0);
		c = hash32((c ^ blo) >>> 0); c = hash32((c ^ bhi) >>> 0);
	}
	return c >>> 0;
}

/*
 * Timed run for bench.py --impl reference: the reference's Queue + ControlledDelay
 * on a generated trace; returns wall ms of the enqueue and of the drain loop.
 */
function cb_bench(seed, n, p, target, nslots) {
	var tr = cb_gen_trace(seed, n, p);
	var cd = new ControlledDelay(target), waiters = new Queue(), i;
	var t0 = Date.now();
	for (i = 0; i < n; i++) {
		__set_clock_ns(tr.start_ns[i]);
		waiters.push({ ch_started: utils.currentMillis(), waiting: true, idx: i });
	}
	var t1 = Date.now();
	var drops = 0, admits = 0, cursor = 0, lastslot = -1;
	while (waiters.length > 0) {
		var hdl = waiters.shift();
		__set_clock_ns(tr.now_ns[hdl.idx]);
		var drop = cd.overloaded(hdl.ch_started);
		if (!hdl.waiting) continue;
		if (drop) { drops++; continue; }
		lastslot = cursor % nslots; cursor++; admits++;
	}
	var t2 = Date.now();
	return { n: n, push_ms: t1 - t0, drain_ms: t2 - t1, drops: drops, admits: admits,
	    lastslot: lastslot, state: cb_state(cd), checksum: cb_trace_checksum(tr) };
}

/*
 * Event-driven pool model (SURVEY.md 8(f) row 1): the reference's ControlledDelay and Queue,
 * verbatim, driven by the restated closures of ConnectionPool#claim / the slot-idle handler
 * and the claim handle's hard timeout.
 *
 * job: { target, nslots, state|null, idle0, ev_type:[], ev_a:[], ev_b:[] }
 *   ev_type 0  CLAIM at clock ev_a (ns)             -> ticket = number of CLAIMs before it
 *           1  IDLE  at clock ev_a (ns), ev_b slots turn idle, one handler run per slot
 *           2  CANCEL ticket ev_a                    (handle.cancel() while waiting)
 * Before an event that carries a clock, every waiting handle whose hard timeout has come due
 * (clock >= started + claimTimeout) fails with ClaimTimeoutError, oldest ticket first, as its
 * S.timeout() would have fired.  Slots are anonymous; an admitted claim gets slot cursor % nslots
 * (SURVEY A.4), whether it took an idle slot at claim() time or was handed one by the drain loop.
 * Output: res = flat [ticket, code, slot]* in the order handles were resolved
 *   code 0 admitted, 1 dropped by CoDel (hdl.timeout()), 2 skipped (not waiting when dequeued),
 *        3 hard timeout; timeouts = claimTimeout per ticket; final state and counters.
 */
function cb_pool_sim(job) {
	var cd = new ControlledDelay(job.target);
	if (job.state) cb_set_state(cd, job.state);
	var waiters = new Queue();
	var handles = [], timeouts = [], res = [];
	var idle = job.idle0 || 0, cursor = job.cursor || 0, S = job.nslots;
	var counters = { 'claim': 0, 'queued-claim': 0, 'max-claim-queue': 0, 'claim-timeout': 0 };
	var first_live = 0;	/* tickets below this are resolved (speeds up the timer scan) */
	function admit(h) {
		h.state = 'claimed';
		res.push(h.ticket, 0, S > 0 ? (cursor % S) : -1);
		cursor++;
	}
	function fire_timers(clock_ns) {
		/* connection-fsm.js:626-642: S.timeout(ch_claimTimeout, onTimeout) in state 'waiting' */
		for (var t = first_live; t < handles.length; t++) {
			var h = handles[t];
			if (h.state !== 'waiting') {
				if (t === first_live) first_live++;
				continue;
			}
			if (clock_ns >= h.deadline_ns) {
				h.state = 'failed';
				counters['claim-timeout']++;		/* connection-fsm.js:629-630 */
				res.push(h.ticket, 3, -1);
			}
		}
	}
	var n = job.ev_type.length, e;
	for (e = 0; e < n; e++) {
		var ty = job.ev_type[e], a = job.ev_a[e], b = job.ev_b[e];
		if (ty === 0) {
			fire_timers(a);
			__set_clock_ns(a);
			var timeout = cd.getMaxIdle();			/* pool.js:880 */
			counters['claim']++;				/* pool.js:887 */
			var h = { ticket: handles.length, ch_started: utils.currentMillis(),	/* connection-fsm.js:523 */
			    state: 'waiting', deadline_ns: a + timeout * 1000000 };
			handles.push(h);
			timeouts.push(timeout);
			if (idle > 0) {					/* pool.js:934-951 */
				idle--;
				admit(h);
			} else {
				waiters.push(h);			/* pool.js:960 */
				if (waiters.length > counters['max-claim-queue'])	/* pool.js:962 */
					counters['max-claim-queue'] = waiters.length;
				counters['queued-claim']++;		/* pool.js:963 */
			}
		} else if (ty === 1) {
			fire_timers(a);
			__set_clock_ns(a);
			for (var s = 0; s < b; s++) {			/* one handler run per slot */
				var handed = false;
				while (waiters.length > 0) {		/* pool.js:733 */
					var hdl = waiters.shift();
					var drop = cd.overloaded(hdl.ch_started);	/* pool.js:735-736 */
					if (hdl.state !== 'waiting') {		/* pool.js:738-740 */
						res.push(hdl.ticket, 2, -1);
						continue;
					}
					if (drop) {				/* pool.js:742-745 */
						hdl.state = 'failed';
						counters['claim-timeout']++;	/* connection-fsm.js:633-638 */
						res.push(hdl.ticket, 1, -1);
						continue;
					}
					admit(hdl);				/* pool.js:747-748 */
					handed = true;
					break;
				}
				if (!handed) {
					cd.empty();				/* pool.js:751-753 */
					idle++;					/* pool.js:756 */
				}
			}
		} else {
			var hc = handles[a];
			if (hc && hc.state === 'waiting') hc.state = 'cancelled';	/* connection-fsm.js:580-587 */
		}
	}
	return { res: res, timeouts: timeouts, state: cb_state(cd), counters: counters,
	    idle: idle, cursor: cursor, queued: waiters.length };
}

/*
 * The pool's sizing arithmetic around the reference's own FIRFilter (lib/pool.js:44-100, verbatim).
 * One event per sampler tick or rebalance, flat arrays:
 *   ty 0: the 5 Hz sampler (lib/pool.js:251-258): conns=a, idleq+initq=b        -> lpf.put(busy + spares)
 *   ty 1: _rebalance's target (lib/pool.js:556-587): total=a, idleq=b, initq=c, waiters=d
 *         -> out: target, clamped (0/1), lpf.get()
 * job: { spares, max }.  Returns { taps, out: flat [target, clamped, lpf]* for the ty-1 events }.
 */
function cb_lpf_run(job) {
	var lpf = new pool_lpf.FIRFilter(pool_lpf.LP_TAPS);
	var out = [], n = job.ty.length;
	for (var e = 0; e < n; e++) {
		if (job.ty[e] === 0) {
			var conns = job.a[e];
			var sp = job.b[e];				/* p_idleq.length + p_initq.length, pool.js:256 */
			var bz = conns - sp;				/* pool.js:257 */
			lpf.put(bz + job.spares);			/* pool.js:258 */
			continue;
		}
		var total = job.a[e];
		var spares = job.b[e] + job.c[e] - job.d[e];		/* pool.js:563-564 */
		if (spares < 0) spares = 0;
		var busy = total - spares;				/* pool.js:567 */
		if (busy < 0) busy = 0;
		var extras = job.d[e] - job.c[e];			/* pool.js:570 */
		if (extras < 0) extras = 0;
		var target = busy + extras + job.spares;		/* pool.js:574 */
		var g = lpf.get();
		var min = Math.ceil(g);					/* pool.js:576 */
		var clamped = 0;
		if (target < min * 1.05) {				/* pool.js:577-583 */
			target = min;
			clamped = 1;
		}
		if (target > job.max) target = job.max;			/* pool.js:585-586 */
		out.push(target, clamped, g);
	}
	var taps = [];
	for (var i = 0; i < pool_lpf.LP_TAPS.length; i++) taps.push(pool_lpf.LP_TAPS[i]);
	return { taps: taps, out: out, lp_int: pool_lpf.LP_INT };
}

/*
 * A script of operations on the reference's own Queue (lib/queue.js, verbatim).  op[i], arg[i]:
 *   0 push(arg)            1 shift() (skipped by the harness when the queue is empty: the reference has no guard,
 *                            SURVEY A.3 quirk 7)
 *   2 node.remove() of the arg-th pushed node, if it is still linked
 *   3 forEach: collect the values in order, removing (from inside the callback) every node whose value is
 *     divisible by arg
 *   4 peek()
 * Returns per op [result, length, isEmpty]: result = the shifted value / 1 when a remove happened / the number of
 * values visited plus 1000 x the number removed / -1 for a null peek / -2 for a skipped op.
 */
function cb_queue_ops(op, arg) {
	var q = new Queue(), nodes = [], linked = [], out = [], vals;
	for (var i = 0; i < op.length; i++) {
		var r = -2, a = arg[i];
		if (op[i] === 0) {
			nodes.push(q.push(a)); linked.push(true); r = nodes.length - 1;
		} else if (op[i] === 1) {
			if (q.length > 0) {
				r = q.shift();
				for (var k = 0; k < nodes.length; k++)
					if (linked[k] && nodes[k].n_value === r && nodes[k].n_next === null) linked[k] = false;
			}
		} else if (op[i] === 2) {
			if (a < nodes.length && linked[a]) { nodes[a].remove(); linked[a] = false; r = 1; }
		} else if (op[i] === 3) {
			vals = 0; var removed = 0;
			q.forEach(function (v, n) {
				vals++;
				if (a > 0 && v % a === 0) {
					n.remove(); removed++;
					for (var k = 0; k < nodes.length; k++) if (nodes[k] === n) linked[k] = false;
				}
			});
			r = vals + 1000 * removed;
		} else if (op[i] === 4) {
			var p = q.peek();
			r = (p === null) ? -1 : p;
		}
		out.push(r, q.length, q.isEmpty() ? 1 : 0);
	}
	return out;
}
