/*
 * shim/js/pool_queue.js -- the pending-claim side of lib/pool.js on the device queue (cb_queue_*), through
 * the N-API addon over libcueball_b200.so.  Not executed in this repository (no Node.js here); the same
 * protocol is what node-cueball_b200/pool.py (the Python mirror) runs, and that IS tested against the
 * reference's pool model on a B200 and on the CPU (tests/test_gpu_pool_mirror.py,
 * tests/test_pool_mirror_emulated.py).  See INTEGRATION.md.
 *
 * What it replaces in lib/pool.js (the slot FSMs, sockets and p_idleq's FSM objects stay where they are):
 *
 *   this.p_waiters = new Queue(); this.p_codel = new ControlledDelay(...)       :191-200
 *       -> this.p_claims = new PendingClaims(options.targetClaimDelay, nIdleSlots)
 *   claim(): timeout = p_codel.getMaxIdle(); idle slot first, else p_waiters.push + counters   :880, :929-966
 *       -> var r = this.p_claims.claim(handle);   // r.timeout = claimTimeout; r.admitted => take p_idleq.shift()
 *   the slot-idle handler's while loop + codel.empty() + p_idleq.push             :733-757
 *       -> if (!this.p_claims.slotIdle(fsm)) this.p_idleq.push(fsm);
 *   handle.cancel() while waiting                                                 connection-fsm.js:580-587
 *       -> this.p_claims.cancel(handle)
 *   the handles' S.timeout(claimTimeout)                                          connection-fsm.js:626-642
 *       -> one pool timer calling this.p_claims.expire()
 *   counters claim / queued-claim / max-claim-queue / claim-timeout, waiterCount    :269-283, :834-857
 *       -> this.p_claims.stats()
 */
'use strict';
const native = require('../napi/build/Release/cueball_b200.node');

const R_ADMIT = 0n, R_CODEL_DROP = 1n, R_SKIPPED = 2n, R_HARD_TIMEOUT = 3n;

function hrns() { return process.hrtime.bigint(); }

function PendingClaims(targetClaimDelay, idleSlots, options) {
	options = options || {};
	this.pc_ctx = options.ctx || native.create(options.device || 0);
	native.codelInit(this.pc_ctx, targetClaimDelay);
	var cap = options.queueCapacity || 65536;
	native.queueCreate(this.pc_ctx, cap, 4 * cap, idleSlots || 0);
	this.pc_handles = new Map();			/* ticket (BigInt) -> claim handle, until resolved */
	this.pc_start = new BigInt64Array(1);
	this.pc_tmo = new Float64Array(1);
	this.pc_ticket = new BigInt64Array(1);
	this.pc_res = new BigInt64Array(3 * 4 * cap);
	this.pc_admitted = [];				/* handles admitted by the call in progress */
}

/* turn the device's (ticket, code, slot) records into handle transitions, in the device's order */
PendingClaims.prototype._dispatch = function () {
	var n = native.queueResults(this.pc_ctx, this.pc_res);
	for (var i = 0; i < n; i++) {
		var ticket = this.pc_res[3 * i], code = this.pc_res[3 * i + 1];
		var hdl = this.pc_handles.get(ticket);
		if (hdl === undefined)
			continue;
		this.pc_handles.delete(ticket);
		if (code === R_ADMIT)
			this.pc_admitted.push(hdl);		/* the caller hands it a slot FSM: hdl.try(fsm), lib/pool.js:747 */
		else if (code === R_CODEL_DROP || code === R_HARD_TIMEOUT)
			hdl.timeout();				/* ClaimTimeoutError, lib/pool.js:742-745 / connection-fsm.js:633-638 */
		/* R_SKIPPED: the handle left 'waiting' before it was dequeued, lib/pool.js:738-740 */
	}
};

/* ConnectionPool#claim's queue side.  Returns { timeout: ms, admitted: bool }. */
PendingClaims.prototype.claim = function (hdl) {
	this.pc_start[0] = hrns();
	var ticket = native.queuePush(this.pc_ctx, this.pc_start, this.pc_tmo);
	hdl.ch_ticket = ticket;
	this.pc_handles.set(ticket, hdl);
	this.pc_admitted.length = 0;
	this._dispatch();
	return ({ timeout: this.pc_tmo[0], admitted: this.pc_admitted.indexOf(hdl) !== -1 });
};

/* A slot FSM turned idle.  true: a waiting handle took it (hdl.try(fsm) done); false: the queue ran dry
 * (codel.empty() has been called on the device) and the caller keeps the slot on p_idleq. */
PendingClaims.prototype.slotIdle = function (fsm) {
	this.pc_admitted.length = 0;
	native.queueIdle(this.pc_ctx, hrns(), 1);
	this._dispatch();
	if (this.pc_admitted.length === 0)
		return (false);
	this.pc_admitted[0].try(fsm);
	return (true);
};

PendingClaims.prototype.cancel = function (hdl) {
	if (hdl.ch_ticket === undefined || !this.pc_handles.has(hdl.ch_ticket))
		return;
	this.pc_ticket[0] = hdl.ch_ticket;
	native.queueCancel(this.pc_ctx, this.pc_ticket);
	this.pc_handles.delete(hdl.ch_ticket);
};

PendingClaims.prototype.expire = function () {
	native.queueExpire(this.pc_ctx, hrns());
	this._dispatch();
};

/* an idle slot FSM was removed by a rebalance (setUnwanted, lib/pool.js:601-631) */
PendingClaims.prototype.idleSlotsRemoved = function (k) {
	return (native.queueTrimIdle(this.pc_ctx, k));
};

/* the pieces of getStats() this object owns (lib/pool.js:834-857) */
PendingClaims.prototype.stats = function () {
	var qi = native.queueInfo(this.pc_ctx);
	var counters = {};
	if (qi.claim) counters['claim'] = qi.claim;
	if (qi.queuedClaim) counters['queued-claim'] = qi.queuedClaim;
	if (qi.maxClaimQueue) counters['max-claim-queue'] = qi.maxClaimQueue;
	if (qi.claimTimeout) counters['claim-timeout'] = qi.claimTimeout;
	return ({ counters: counters, idleConnections: qi.idle, waiterCount: qi.queued });
};

module.exports = { PendingClaims: PendingClaims };
