/*
 * shim/napi/cueball_b200_napi.c -- N-API addon over the C ABI of include/cueball_b200.h.
 *
 * NOT BUILT IN THIS REPOSITORY'S ENVIRONMENT: there is no Node.js toolchain
 * (no `node`, no node_api.h).  It is the binding a maintainer would add
 * (see INTEGRATION.md); build with node-gyp against libcueball_b200.so:
 *
 *   { "targets": [{ "target_name": "cueball_b200", "sources": ["cueball_b200_napi.c"],
 *       "include_dirs": ["../../include"], "libraries": ["-lcueball_b200"] }] }
 *
 * JS surface (consumed by the modules under shim/js):
 *   create(device) -> ctx (external)                      destroy(ctx)
 *   codelInit(ctx, targetMs)                              new ControlledDelay(), lib/codel.js:24
 *   overloaded(ctx, startNs: BigInt, nowNs: BigInt) -> bool   lib/codel.js:56
 *   overloadedMs(ctx, startMs: Number, nowMs: Number) -> bool  the same on the reference's own double-ms clock
 *                                                          (hdl.ch_started as lib/connection-fsm.js:523 stamps it)
 *   empty(ctx, nowNs)                                     lib/codel.js:91
 *   getMaxIdle(ctx, nowNs) -> ms                          lib/codel.js:109
 *   admitBatch(ctx, start: BigInt64Array, now: BigInt64Array, flags: Uint8Array|null,
 *              decision: Uint8Array, slot: Int32Array|null) -> {drops, admits, skipped}
 *                                                          the drain loop lib/pool.js:733-753
 *   backendAdded(ctx, key, u) / backendRemoved(ctx, key) / reshuffle(ctx, u)
 *                                                          lib/pool.js:285-313, 501-519
 *   planRebalance(ctx, have: Int32Array, dead: Uint8Array, target, max, singleton)
 *   shuffle(ctx, draws: Float64Array) -> Int32Array (the permutation of utils.shuffle)
 *              -> {want: Int32Array, add: Int32Array, rmKey: Int32Array, rmOrd: Int32Array}
 *                                                          lib/utils.js:239-393
 *   the pending-claim queue resident in HBM (cb_queue_*; what shim/js/pool_queue.js drives):
 *   queueCreate(ctx, capacity, resultCapacity, idleSlots)
 *   queuePush(ctx, startNs: BigInt64Array, timeouts: Float64Array|null) -> first ticket (BigInt)
 *                                                          n x ConnectionPool#claim, lib/pool.js:874-966
 *   queueCancel(ctx, tickets: BigInt64Array)              handle.cancel(), lib/connection-fsm.js:580-587
 *   queueExpire(ctx, nowNs)                               the handles' claim timers, lib/connection-fsm.js:626-642
 *   queueIdle(ctx, nowNs, k)                              k slots turned idle: the handler lib/pool.js:733-757
 *   queueTrimIdle(ctx, k) -> removed                      idle slots that went away, lib/pool.js:601-631
 *   queueResults(ctx, out: BigInt64Array) -> records written: (ticket, code, slot) triples, resolution order
 *   queueInfo(ctx) -> {claim, queuedClaim, maxClaimQueue, claimTimeout, idle, queued, nResults}
 *                                                          lib/pool.js:269-283 counters, p_idleq/p_waiters lengths
 * Errors: CB_EINVAL -> throws (the reference throws AssertionError for programmer errors);
 * everything else -> throws Error(cb_last_error()).  Arguments are validated before the C ABI sees them:
 * a clock must be a BigInt that fits int64 (a Number is NOT silently read as 0), typed arrays must have the
 * element type named above and at least as many elements as the batch.
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>

#include "cueball_b200.h"

#define NAPI_CALL(env, call)                                          \
	do {                                                              \
		if ((call) != napi_ok) {                                      \
			napi_throw_error((env), NULL, "N-API call failed: " #call); \
			return NULL;                                              \
		}                                                             \
	} while (0)

/* NULL (with a TypeError pending) unless `v` is a context made by create(): callers return at once */
static cb_ctx *
get_ctx(napi_env env, napi_value v)
{
	void *p = NULL;
	if (napi_get_value_external(env, v, &p) != napi_ok || p == NULL) {
		napi_throw_type_error(env, NULL, "ctx expected");
		return NULL;
	}
	return (cb_ctx *)p;
}
#define GET_CTX(ctx, env, v)                   \
	cb_ctx *ctx = get_ctx((env), (v));         \
	if (ctx == NULL)                           \
		return NULL

static napi_value
throw_cb(napi_env env, cb_ctx *ctx, int rc)
{
	if (rc == CB_EINVAL)
		napi_throw_type_error(env, "AssertionError", cb_last_error(ctx));
	else
		napi_throw_error(env, NULL, cb_last_error(ctx));
	return NULL;
}

static void
ctx_finalize(napi_env env, void *data, void *hint)
{
	(void)env; (void)hint;
	cb_destroy((cb_ctx *)data);
}

static napi_value
Create(napi_env env, napi_callback_info info)
{
	size_t argc = 1; napi_value argv[1], out; int32_t dev = 0; cb_ctx *ctx = NULL;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc > 0) napi_get_value_int32(env, argv[0], &dev);
	int rc = cb_create(&ctx, dev);
	if (rc != CB_OK) return throw_cb(env, NULL, rc);
	NAPI_CALL(env, napi_create_external(env, ctx, ctx_finalize, NULL, &out));
	return out;
}

static napi_value
CodelInit(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2]; double target;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	NAPI_CALL(env, napi_get_value_double(env, argv[1], &target));
	int rc = cb_codel_init(ctx, target);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

/* a clock in nanoseconds: a BigInt that fits int64; anything else throws (false is returned) */
static bool
big(napi_env env, napi_value v, int64_t *out)
{
	bool lossless = false;
	*out = 0;
	if (napi_get_value_bigint_int64(env, v, out, &lossless) != napi_ok || !lossless) {
		napi_throw_type_error(env, NULL, "clock must be a BigInt of nanoseconds that fits int64 "
		    "(use overloadedMs for the reference's double milliseconds)");
		return false;
	}
	return true;
}

static napi_value
Overloaded(napi_env env, napi_callback_info info)
{
	size_t argc = 3; napi_value argv[3], out; int drop = 0; int64_t st, nw;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	if (!big(env, argv[1], &st) || !big(env, argv[2], &nw)) return NULL;
	int rc = cb_codel_overloaded(ctx, st, nw, &drop);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_get_boolean(env, drop != 0, &out));
	return out;
}

/* the reference's own call shape: hdl.ch_started and the clock as double milliseconds (lib/codel.js:56-58) */
static napi_value
OverloadedMs(napi_env env, napi_callback_info info)
{
	size_t argc = 3; napi_value argv[3], out; int drop = 0; double st, nw; napi_valuetype t1, t2;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	if (napi_typeof(env, argv[1], &t1) != napi_ok || napi_typeof(env, argv[2], &t2) != napi_ok ||
	    t1 != napi_number || t2 != napi_number) {
		napi_throw_type_error(env, "AssertionError", "start (number) is required");	/* mod_assert.number, lib/codel.js:57 */
		return NULL;
	}
	NAPI_CALL(env, napi_get_value_double(env, argv[1], &st));
	NAPI_CALL(env, napi_get_value_double(env, argv[2], &nw));
	int rc = cb_codel_overloaded_ms(ctx, st, nw, &drop);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_get_boolean(env, drop != 0, &out));
	return out;
}

static napi_value
Empty(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2]; int64_t nw;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	if (!big(env, argv[1], &nw)) return NULL;
	int rc = cb_codel_empty(ctx, nw);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
GetMaxIdle(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2], out; double ms = 0; int64_t nw;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	if (!big(env, argv[1], &nw)) return NULL;
	int rc = cb_codel_max_idle(ctx, nw, &ms);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_create_double(env, ms, &out));
	return out;
}

/*
 * A typed array argument of element type `want` with at least `need` elements.  *ok = false (with a
 * TypeError pending) when it is something else or too short; null / undefined is allowed only when
 * `optional` and yields NULL with *ok = true.
 */
static void *
typed(napi_env env, napi_value v, napi_typedarray_type want, size_t need, bool optional, const char *what,
    size_t *len, bool *ok)
{
	napi_valuetype t; void *data = NULL; napi_typedarray_type tt; napi_value ab; size_t off, n = 0;
	*ok = false;
	if (len) *len = 0;
	if (napi_typeof(env, v, &t) != napi_ok) {
		napi_throw_type_error(env, NULL, what);
		return NULL;
	}
	if (t == napi_null || t == napi_undefined) {
		if (optional) *ok = true;
		else napi_throw_type_error(env, NULL, what);
		return NULL;
	}
	if (napi_get_typedarray_info(env, v, &tt, &n, &data, &ab, &off) != napi_ok || tt != want || n < need) {
		napi_throw_type_error(env, NULL, what);
		return NULL;
	}
	if (len) *len = n;
	*ok = true;
	return data;
}

static napi_value
AdmitBatch(napi_env env, napi_callback_info info)
{
	size_t argc = 6, n = 0; napi_value argv[6], out, v; bool ok;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 6) {
		napi_throw_type_error(env, NULL, "admitBatch(ctx, start, now, flags, decision, slot)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	/* the batch is as long as `start`; every other array must hold at least that many elements */
	const int64_t *start = typed(env, argv[1], napi_bigint64_array, 0, false, "start: BigInt64Array expected", &n, &ok);
	if (!ok) return NULL;
	const int64_t *now = typed(env, argv[2], napi_bigint64_array, n, false, "now: BigInt64Array of start's length expected", NULL, &ok);
	if (!ok) return NULL;
	const uint8_t *flags = typed(env, argv[3], napi_uint8_array, n, true, "flags: Uint8Array of start's length (or null) expected", NULL, &ok);
	if (!ok) return NULL;
	uint8_t *decision = typed(env, argv[4], napi_uint8_array, n, false, "decision: Uint8Array of start's length expected", NULL, &ok);
	if (!ok) return NULL;
	int32_t *slot = typed(env, argv[5], napi_int32_array, n, true, "slot: Int32Array of start's length (or null) expected", NULL, &ok);
	if (!ok) return NULL;
	cb_batch_stats st;
	int rc = cb_admit_batch(ctx, start, now, flags, n, decision, slot, NULL, &st);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_create_object(env, &out));
	napi_create_double(env, (double)st.drops, &v);   napi_set_named_property(env, out, "drops", v);
	napi_create_double(env, (double)st.admits, &v);  napi_set_named_property(env, out, "admits", v);
	napi_create_double(env, (double)st.skipped, &v); napi_set_named_property(env, out, "skipped", v);
	return out;
}

static napi_value
BackendAdded(napi_env env, napi_callback_info info)
{
	size_t argc = 3; napi_value argv[3]; int32_t key; double u;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	NAPI_CALL(env, napi_get_value_int32(env, argv[1], &key));
	NAPI_CALL(env, napi_get_value_double(env, argv[2], &u));
	int rc = cb_backend_added(ctx, key, NULL, u);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
BackendRemoved(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2]; int32_t key;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	NAPI_CALL(env, napi_get_value_int32(env, argv[1], &key));
	int rc = cb_backend_removed(ctx, key);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
Reshuffle(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2]; double u;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	GET_CTX(ctx, env, argv[0]);
	NAPI_CALL(env, napi_get_value_double(env, argv[1], &u));
	int rc = cb_reshuffle(ctx, u);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
i32array(napi_env env, const int32_t *src, size_t n)
{
	napi_value ab, ta; void *data;
	napi_create_arraybuffer(env, n * 4, &data, &ab);
	for (size_t i = 0; i < n; i++) ((int32_t *)data)[i] = src[i];
	napi_create_typedarray(env, napi_int32_array, n, ab, 0, &ta);
	return ta;
}

static napi_value
PlanRebalance(napi_env env, napi_callback_info info)
{
	size_t argc = 6, B = 0; napi_value argv[6], out; int32_t target, max; bool single = false, ok;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 5) {
		napi_throw_type_error(env, NULL, "planRebalance(ctx, have, dead, target, max[, singleton])");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	const int32_t *have = typed(env, argv[1], napi_int32_array, 0, false, "have: Int32Array expected", &B, &ok);
	if (!ok) return NULL;
	const uint8_t *dead = typed(env, argv[2], napi_uint8_array, B, false, "dead: Uint8Array of have's length expected", NULL, &ok);
	if (!ok) return NULL;
	NAPI_CALL(env, napi_get_value_int32(env, argv[3], &target));
	NAPI_CALL(env, napi_get_value_int32(env, argv[4], &max));
	if (argc > 5) napi_get_value_bool(env, argv[5], &single);
	size_t tot = 0;
	for (size_t k = 0; k < B; k++) {
		if (have[k] < 0) {
			napi_throw_type_error(env, "AssertionError", "have[] must be >= 0");
			return NULL;
		}
		tot += (size_t)have[k];
	}
	int32_t *want = calloc(B + 1, 4), *add = calloc((size_t)(max > 0 ? max : 0) + 1, 4);
	int32_t *rk = calloc(tot + 1, 4), *ro = calloc(tot + 1, 4), na = 0, nr = 0;
	int rc = cb_plan_rebalance(ctx, have, dead, (int32_t)B, target, max, single, want, add, &na, rk, ro, &nr);
	if (rc == CB_OK) {
		napi_create_object(env, &out);
		napi_set_named_property(env, out, "want", i32array(env, want, B));
		napi_set_named_property(env, out, "add", i32array(env, add, (size_t)na));
		napi_set_named_property(env, out, "rmKey", i32array(env, rk, (size_t)nr));
		napi_set_named_property(env, out, "rmOrd", i32array(env, ro, (size_t)nr));
	}
	free(want); free(add); free(rk); free(ro);
	return rc == CB_OK ? out : throw_cb(env, ctx, rc);
}

/* shuffle(ctx, draws: Float64Array) -> Int32Array: the permutation utils.shuffle (lib/utils.js:207-217) applies
 * to an array of draws.length elements when Math.random() returns draws[0], draws[1], ... */
static napi_value
Shuffle(napi_env env, napi_callback_info info)
{
	size_t argc = 2, n = 0; napi_value argv[2]; bool ok;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 2) {
		napi_throw_type_error(env, NULL, "shuffle(ctx, draws)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	const double *us = typed(env, argv[1], napi_float64_array, 0, false, "draws: Float64Array expected", &n, &ok);
	if (!ok) return NULL;
	int32_t *perm = calloc(n + 1, 4);
	for (size_t i = 0; i < n; i++) perm[i] = (int32_t)i;
	int rc = cb_shuffle(ctx, perm, n, us);
	napi_value out = rc == CB_OK ? i32array(env, perm, n) : NULL;
	free(perm);
	return rc == CB_OK ? out : throw_cb(env, ctx, rc);
}

/* a non-negative count that fits the C ABI's size_t / uint64_t: a Number (integers below 2^53) */
static bool
count_arg(napi_env env, napi_value v, const char *what, uint64_t *out)
{
	double d = -1.0;
	*out = 0;
	if (napi_get_value_double(env, v, &d) != napi_ok || !(d >= 0.0) || d > 9007199254740992.0 || d != (double)(uint64_t)d) {
		napi_throw_type_error(env, NULL, what);
		return false;
	}
	*out = (uint64_t)d;
	return true;
}

static napi_value
QueueCreate(napi_env env, napi_callback_info info)
{
	size_t argc = 4; napi_value argv[4]; uint64_t cap, rcap, idle;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 4) {
		napi_throw_type_error(env, NULL, "queueCreate(ctx, capacity, resultCapacity, idleSlots)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	if (!count_arg(env, argv[1], "capacity: a positive integer expected", &cap) ||
	    !count_arg(env, argv[2], "resultCapacity: a positive integer expected", &rcap) ||
	    !count_arg(env, argv[3], "idleSlots: a non-negative integer expected", &idle))
		return NULL;
	int rc = cb_queue_create(ctx, (size_t)cap, (size_t)rcap, idle);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
QueuePush(napi_env env, napi_callback_info info)
{
	size_t argc = 3, n = 0; napi_value argv[3], out; bool ok; int64_t first = 0;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 3) {
		napi_throw_type_error(env, NULL, "queuePush(ctx, startNs, timeouts)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	const int64_t *start = typed(env, argv[1], napi_bigint64_array, 0, false, "startNs: BigInt64Array expected", &n, &ok);
	if (!ok) return NULL;
	double *tmo = typed(env, argv[2], napi_float64_array, n, true, "timeouts: Float64Array of startNs's length (or null) expected", NULL, &ok);
	if (!ok) return NULL;
	int rc = cb_queue_push(ctx, start, n, &first, tmo);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_create_bigint_int64(env, first, &out));
	return out;
}

static napi_value
QueueCancel(napi_env env, napi_callback_info info)
{
	size_t argc = 2, n = 0; napi_value argv[2]; bool ok;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 2) {
		napi_throw_type_error(env, NULL, "queueCancel(ctx, tickets)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	const int64_t *tickets = typed(env, argv[1], napi_bigint64_array, 0, false, "tickets: BigInt64Array expected", &n, &ok);
	if (!ok) return NULL;
	int rc = cb_queue_cancel(ctx, tickets, n);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
QueueExpire(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2]; int64_t nw;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 2) {
		napi_throw_type_error(env, NULL, "queueExpire(ctx, nowNs)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	if (!big(env, argv[1], &nw)) return NULL;
	int rc = cb_queue_expire(ctx, nw);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
QueueIdle(napi_env env, napi_callback_info info)
{
	size_t argc = 3; napi_value argv[3]; int64_t nw; uint64_t k;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 3) {
		napi_throw_type_error(env, NULL, "queueIdle(ctx, nowNs, k)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	if (!big(env, argv[1], &nw) || !count_arg(env, argv[2], "k: a non-negative integer expected", &k)) return NULL;
	int rc = cb_queue_idle(ctx, nw, k);
	return rc == CB_OK ? NULL : throw_cb(env, ctx, rc);
}

static napi_value
QueueTrimIdle(napi_env env, napi_callback_info info)
{
	size_t argc = 2; napi_value argv[2], out; uint64_t k, removed = 0;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 2) {
		napi_throw_type_error(env, NULL, "queueTrimIdle(ctx, k)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	if (!count_arg(env, argv[1], "k: a non-negative integer expected", &k)) return NULL;
	int rc = cb_queue_trim_idle(ctx, k, &removed);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_create_double(env, (double)removed, &out));
	return out;
}

/* queueResults(ctx, out: BigInt64Array) -> number of (ticket, code, slot) records written into `out`;
 * `out` must have room for every waiting record (3 elements each; queueInfo().nResults), else CB_EINVAL */
static napi_value
QueueResults(napi_env env, napi_callback_info info)
{
	size_t argc = 2, len = 0, n = 0; napi_value argv[2], out; bool ok;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 2) {
		napi_throw_type_error(env, NULL, "queueResults(ctx, out)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	int64_t *triples = typed(env, argv[1], napi_bigint64_array, 3, false, "out: BigInt64Array of 3 elements per record expected", &len, &ok);
	if (!ok) return NULL;
	int rc = cb_queue_results(ctx, triples, len / 3, &n);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_create_double(env, (double)n, &out));
	return out;
}

static napi_value
QueueInfo(napi_env env, napi_callback_info info)
{
	size_t argc = 1; napi_value argv[1], out, v; cb_queue_info qi;
	NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
	if (argc < 1) {
		napi_throw_type_error(env, NULL, "queueInfo(ctx)");
		return NULL;
	}
	GET_CTX(ctx, env, argv[0]);
	int rc = cb_queue_info_get(ctx, &qi);
	if (rc != CB_OK) return throw_cb(env, ctx, rc);
	NAPI_CALL(env, napi_create_object(env, &out));
	/* counters are exact as Numbers below 2^53 (the reference's are Numbers, lib/pool.js:269-283) */
	napi_create_double(env, (double)qi.claim, &v);            napi_set_named_property(env, out, "claim", v);
	napi_create_double(env, (double)qi.queued_claim, &v);     napi_set_named_property(env, out, "queuedClaim", v);
	napi_create_double(env, (double)qi.max_claim_queue, &v);  napi_set_named_property(env, out, "maxClaimQueue", v);
	napi_create_double(env, (double)qi.claim_timeout, &v);    napi_set_named_property(env, out, "claimTimeout", v);
	napi_create_double(env, (double)qi.idle, &v);             napi_set_named_property(env, out, "idle", v);
	napi_create_double(env, (double)(qi.tail - qi.head), &v); napi_set_named_property(env, out, "queued", v);
	napi_create_double(env, (double)qi.n_results, &v);        napi_set_named_property(env, out, "nResults", v);
	return out;
}

static napi_value
Init(napi_env env, napi_value exports)
{
	napi_property_descriptor d[] = {
		{ "create", NULL, Create, NULL, NULL, NULL, napi_default, NULL },
		{ "codelInit", NULL, CodelInit, NULL, NULL, NULL, napi_default, NULL },
		{ "overloaded", NULL, Overloaded, NULL, NULL, NULL, napi_default, NULL },
		{ "overloadedMs", NULL, OverloadedMs, NULL, NULL, NULL, napi_default, NULL },
		{ "empty", NULL, Empty, NULL, NULL, NULL, napi_default, NULL },
		{ "getMaxIdle", NULL, GetMaxIdle, NULL, NULL, NULL, napi_default, NULL },
		{ "admitBatch", NULL, AdmitBatch, NULL, NULL, NULL, napi_default, NULL },
		{ "backendAdded", NULL, BackendAdded, NULL, NULL, NULL, napi_default, NULL },
		{ "backendRemoved", NULL, BackendRemoved, NULL, NULL, NULL, napi_default, NULL },
		{ "reshuffle", NULL, Reshuffle, NULL, NULL, NULL, napi_default, NULL },
		{ "planRebalance", NULL, PlanRebalance, NULL, NULL, NULL, napi_default, NULL },
		{ "shuffle", NULL, Shuffle, NULL, NULL, NULL, napi_default, NULL },
		{ "queueCreate", NULL, QueueCreate, NULL, NULL, NULL, napi_default, NULL },
		{ "queuePush", NULL, QueuePush, NULL, NULL, NULL, napi_default, NULL },
		{ "queueCancel", NULL, QueueCancel, NULL, NULL, NULL, napi_default, NULL },
		{ "queueExpire", NULL, QueueExpire, NULL, NULL, NULL, napi_default, NULL },
		{ "queueIdle", NULL, QueueIdle, NULL, NULL, NULL, napi_default, NULL },
		{ "queueTrimIdle", NULL, QueueTrimIdle, NULL, NULL, NULL, napi_default, NULL },
		{ "queueResults", NULL, QueueResults, NULL, NULL, NULL, napi_default, NULL },
		{ "queueInfo", NULL, QueueInfo, NULL, NULL, NULL, napi_default, NULL },
	};
	napi_define_properties(env, exports, sizeof (d) / sizeof (d[0]), d);
	return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
