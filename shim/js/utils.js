/*
 * shim/js/utils.js -- planRebalance with the reference's signature (lib/utils.js:239-393)
 * computed by cb_plan_rebalance, and shuffle (lib/utils.js:207-217) by cb_shuffle.  Not executed in this repository (no Node.js here).
 */
'use strict';
const native = require('../napi/build/Release/cueball_b200.node');
var ctx = null;

function planRebalance(inSpares, dead, target, max, singleton) {
	if (ctx === null) ctx = native.create(0);
	var keys = Object.keys(inSpares);
	var have = new Int32Array(keys.length), dd = new Uint8Array(keys.length);
	keys.forEach(function (k, i) {
		have[i] = (inSpares[k] || []).length;
		dd[i] = dead[k] === true ? 1 : 0;
	});
	var r = native.planRebalance(ctx, have, dd, target, max, !!singleton);
	var plan = { add: [], remove: [] }, i;
	for (i = 0; i < r.add.length; i++) plan.add.push(keys[r.add[i]]);
	for (i = 0; i < r.rmKey.length; i++) plan.remove.push(inSpares[keys[r.rmKey[i]]][r.rmOrd[i]]);
	return (plan);
}
/* lib/utils.js:207-217: same draws from Math.random(), same result, in place */
function shuffle(array) {
	if (ctx === null) ctx = native.create(0);
	var draws = new Float64Array(array.length), i;
	for (i = 0; i < draws.length; i++) draws[i] = Math.random();
	var perm = native.shuffle(ctx, draws), copy = array.slice();
	for (i = 0; i < perm.length; i++) array[i] = copy[perm[i]];
	return (array);
}
module.exports = { planRebalance: planRebalance, shuffle: shuffle };
