/*
 * oracle/js/shim.js -- TEST INFRASTRUCTURE (parity oracle), ours, not reference code.
 *
 * Minimal host environment so the SpiderMonkey 1.7 shell (ES3 + JS1.7) can load
 * the reference's lib/utils.js, lib/codel.js, lib/queue.js unmodified:
 *   - three ES5 polyfills the files use,
 *   - a require() over a module table with inert stand-ins for the npm
 *     dependencies that are not vendored in the reference tree (assert-plus is
 *     reduced to the checks the hot path relies on; artedi/os are never called
 *     on the path),
 *   - a fake `process.hrtime()` driven by the harness clock, so that the
 *     reference's own utils.currentMillis() (lib/utils.js:198-204) runs verbatim.
 */
if (!Object.keys) Object.keys = function (o) {
	var r = []; for (var k in o) if (o.hasOwnProperty(k)) r.push(k); return r; };
if (!Object.defineProperty) Object.defineProperty = function (o, n, d) {
	if (d.get) o.__defineGetter__(n, d.get);
	if (d.set) o.__defineSetter__(n, d.set);
	if ('value' in d) o[n] = d.value;
	return o; };
if (!Number.isFinite) Number.isFinite = function (v) {
	return typeof (v) === 'number' && isFinite(v); };

/* lib/pool.js:65,79 keep the filter taps and the sample ring in Float64Arrays: an Array of Numbers holds
 * the same doubles (SM 1.7 has no typed arrays) */
if (typeof (Float64Array) === 'undefined') {
	var Float64Array = function (n) { var a = new Array(n); for (var i = 0; i < n; i++) a[i] = 0.0; return a; };
}

var __mods = {}, __tests = {};
/* harness clock: integer seconds + integer nanoseconds, like process.hrtime() */
var __hr_secs = 0, __hr_ns = 0;
var process = { hrtime: function () { return [__hr_secs, __hr_ns]; } };
/* set the clock from a total-nanoseconds value (exact integer < 2^53) */
function __set_clock_ns(t) {
	__hr_secs = Math.floor(t / 1e9);
	__hr_ns = t - __hr_secs * 1e9;
}

function __fail(m) { throw new Error('assertion failed: ' + m); }
__mods['assert-plus'] = {
	ok: function (c, m) { if (!c) __fail(m || 'ok'); },
	number: function (v, m) { if (typeof (v) !== 'number') __fail(m + ' (number)'); },
	finite: function (v, m) { if (typeof (v) !== 'number' || !isFinite(v)) __fail(m + ' (finite)'); },
	object: function (v, m) { if (typeof (v) !== 'object' || v === null) __fail(m + ' (object)'); },
	equal: function (a, b, m) { if (a != b) __fail(m || 'equal'); },
	optionalFinite: function (v, m) { if (v !== undefined && v !== null && (typeof (v) !== 'number' || !isFinite(v))) __fail(m); },
	optionalObject: function () {}, optionalNumber: function () {}, optionalBool: function () {},
	optionalString: function () {}, string: function () {}, bool: function () {}, func: function () {},
	arrayOfString: function () {}, uuid: function () {}, notStrictEqual: function (a, b, m) { if (a === b) __fail(m); },
	strictEqual: function (a, b, m) { if (a !== b) __fail(m); }, deepEqual: function () {}
};
__mods['artedi'] = {}; __mods['os'] = {};
function require(n) {
	if (n === '../lib/utils') n = './utils';
	if (!(n in __mods)) throw new Error('no module ' + n);
	return __mods[n];
}
