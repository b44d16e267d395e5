/*
 * shim/js/codel.js -- drop-in for the reference's lib/codel.js (same exports, same
 * method names) backed by the N-API addon over libcueball_b200.so.  Not executed in
 * this repository (no Node.js here); see INTEGRATION.md.
 *
 *   const mod_codel = require('cueball-b200/shim/js/codel');   // instead of './codel'
 *   this.p_codel = new mod_codel.ControlledDelay(options.targetClaimDelay);
 *   ... p_codel.overloaded(hdl.ch_started) / .empty() / .getMaxIdle()   (lib/pool.js:735,752,880)
 */
'use strict';
const native = require('../napi/build/Release/cueball_b200.node');

function hrns() { return process.hrtime.bigint(); }

function ControlledDelay(targetClaimDelay, ctx) {
	this.cd_ctx = ctx || native.create(0);
	native.codelInit(this.cd_ctx, targetClaimDelay);	/* throws like mod_assert.finite */
	this.cd_targdelay = targetClaimDelay;
}
/*
 * `start` is hdl.ch_started.  The reference stamps it with utils.currentMillis() (a Number, double
 * milliseconds, lib/connection-fsm.js:523): that call shape works unchanged -- the clock is then read
 * the same way (lib/utils.js:198-204) and the comparison runs on the doubles, as in lib/codel.js:56-86.
 * A host that keeps BigInt nanoseconds on the handle (ch_started_ns) passes those instead and the
 * device does the ms conversion itself.  Anything else throws, like mod_assert.number(start).
 */
function currentMillis() {	/* lib/utils.js:198-204 */
	var t = process.hrtime();
	return (t[0] * 1000 + t[1] / 1000000);
}
ControlledDelay.prototype.overloaded = function (start) {
	if (typeof (start) === 'bigint')
		return (native.overloaded(this.cd_ctx, start, hrns()));
	if (typeof (start) !== 'number')
		throw (new TypeError('start (number) is required'));
	return (native.overloadedMs(this.cd_ctx, start, currentMillis()));
};
ControlledDelay.prototype.empty = function () { native.empty(this.cd_ctx, hrns()); };
ControlledDelay.prototype.getMaxIdle = function () {
	return (native.getMaxIdle(this.cd_ctx, hrns()));
};
/* batch form: the whole drain loop lib/pool.js:733-753 in one device call */
ControlledDelay.prototype.admitBatch = function (start, now, flags, decision, slot) {
	return (native.admitBatch(this.cd_ctx, start, now, flags, decision, slot));
};

module.exports = { ControlledDelay: ControlledDelay };
