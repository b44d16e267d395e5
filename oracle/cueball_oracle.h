/*
 * oracle/cueball_oracle.h -- CPU restatement of node-cueball's claim-queue hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker (or as the timed CPU baseline) -- never on the product path.
 *
 * Parity status: PINNED.  This restatement is checked bit-for-bit against the
 * reference's own JavaScript executed under the SpiderMonkey engine vendored in
 * the reference tree (oracle/_ref, see oracle/Makefile) and against the golden
 * vectors generated from it (tests/golden/, tests/test_oracle_golden.py), and
 * its planRebalance against the reference's own known-answer tests
 * (test/utils.test.js, recorded in tests/golden/plan_rebalance_kat.json).
 *
 * All citations are file:line in the reference tree (TritonDataCenter/node-cueball
 * v2.10.3).  Build with -ffp-contract=off: every double operation below must
 * round exactly once, like the JavaScript it restates.
 */
#ifndef CUEBALL_ORACLE_H
#define CUEBALL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBO_CODEL_INTERVAL 100.0 /* lib/codel.js:16 */

/* decision codes shared with include/cueball_b200.h */
#define CBO_ADMIT 0   /* hdl.try(fsm)      lib/pool.js:747 */
#define CBO_DROP 1    /* hdl.timeout()     lib/pool.js:742-745 */
#define CBO_SKIPPED 2 /* not 'waiting'     lib/pool.js:738-740 */

/* per-claim flag bits */
#define CBO_F_WAITING 1     /* handle still in state 'waiting' */
#define CBO_F_EMPTY_AFTER 2 /* queue drained with a slot idle: codel.empty(), lib/pool.js:751-753 */

/* lib/codel.js:24-32 (+ cd_last_empty, which the ctor leaves undefined) */
typedef struct cbo_codel {
	double cd_targdelay;
	double cd_first_above_time;
	double cd_drop_next;
	double cd_count; /* a JS Number */
	int cd_dropping;
	int cd_has_last_empty; /* 0 = undefined */
	double cd_last_empty;
} cbo_codel;

typedef struct cbo_batch_stats {
	uint64_t drops, admits, skipped;
	uint64_t cursor; /* admit cursor after the batch (slot = cursor mod nslots) */
} cbo_batch_stats;

/* lib/utils.js:198-204 */
double cbo_current_millis(int64_t hr_secs, int64_t hr_ns);
double cbo_ns_to_ms(int64_t total_ns); /* hrtime split + currentMillis */

void cbo_codel_init(cbo_codel *cd, double target);                 /* codel.js:24-32 */
int cbo_codel_can_drop(cbo_codel *cd, double now, double start);   /* codel.js:34-46 */
double cbo_codel_get_drop_next(const cbo_codel *cd, double now);   /* codel.js:48-50 */
int cbo_codel_overloaded(cbo_codel *cd, double now, double start); /* codel.js:56-86 */
void cbo_codel_empty(cbo_codel *cd, double now);                   /* codel.js:91-94 */
double cbo_codel_get_max_idle(const cbo_codel *cd, double now);    /* codel.js:109-118 */

/* lib/queue.js: sentinel doubly-linked FIFO */
typedef struct cbo_queue cbo_queue;
typedef struct cbo_qnode cbo_qnode;
cbo_queue *cbo_queue_new(void);
void cbo_queue_free(cbo_queue *q);
cbo_qnode *cbo_queue_push(cbo_queue *q, void *v); /* queue.js:52-56 */
void *cbo_queue_shift(cbo_queue *q);              /* queue.js:57-62 */
int cbo_queue_is_empty(const cbo_queue *q);       /* queue.js:46-48 */
size_t cbo_queue_length(const cbo_queue *q);      /* queue.js:71-75 */
void cbo_qnode_remove(cbo_qnode *n);              /* queue.js:28-37 */

/*
 * The drain loop lib/pool.js:733-753 in batch form over a linked-list queue of
 * handles (faithful variant) or straight over the arrays (flat variant, same
 * results; the fair "CPU speed" baseline).  flags may be NULL (all waiting).
 * decision/slot may be NULL (count only).
 */
int cbo_drain_batch(cbo_codel *cd, const int64_t *start_ns, const int64_t *now_ns,
    const uint8_t *flags, size_t n, int64_t nslots, uint64_t cursor,
    uint8_t *decision, int32_t *slot, cbo_batch_stats *st);
int cbo_drain_batch_flat(cbo_codel *cd, const int64_t *start_ns, const int64_t *now_ns,
    const uint8_t *flags, size_t n, int64_t nslots, uint64_t cursor,
    uint8_t *decision, int32_t *slot, cbo_batch_stats *st);
/* same, timestamps already in double milliseconds */
int cbo_drain_batch_ms(cbo_codel *cd, const double *start_ms, const double *now_ms,
    const uint8_t *flags, size_t n, int64_t nslots, uint64_t cursor,
    uint8_t *decision, int32_t *slot, cbo_batch_stats *st);

/*
 * utils.planRebalance (lib/utils.js:239-393).  Keys are 0..B-1 in the caller's
 * key order (JS object insertion order).  have[k] = number of open connections
 * of key k, dead[k] != 0 = declared dead.  Outputs:
 *   want[k]           wantedSpares (0 where the JS map has no entry)
 *   add[]/n_add       plan.add   : key indices, forward key order, grouped
 *   rm_key/rm_ord     plan.remove: (key, ordinal of the connection within
 *                     inSpares[key]) in REVERSE key order, oldest first
 * add must hold max(0,sum(want)) entries at most `max`; rm_* sum(have).
 * Returns 0, or -1 on the reference's assertion failures.
 */
int cbo_plan_rebalance(const int32_t *have, const uint8_t *dead, int32_t B,
    int32_t target, int32_t max, int singleton, int32_t *want,
    int32_t *add, int32_t *n_add, int32_t *rm_key, int32_t *rm_ord, int32_t *n_rm);

/* preference list: lib/pool.js:285-291 (insert), :293-296 (remove), :501-519 (reshuffle) */
void cbo_pref_insert(int32_t *keys, int32_t *len, int32_t key, double u);
int cbo_pref_remove(int32_t *keys, int32_t *len, int32_t key);
void cbo_pref_reshuffle(int32_t *keys, int32_t len, double u);
/* B insertions with draws us[], then planRebalance({}, {}, S, S): slot table */
int cbo_build_slots(int32_t B, int32_t S, const double *us, int32_t *keys_out,
    int32_t *slot_backend_out);
/* utils.shuffle, lib/utils.js:207-217, draws injected */
void cbo_shuffle(int32_t *arr, int32_t n, const double *us);

/*
 * Event-driven pool model, same semantics as cb_pool_sim in
 * oracle/js/driver.js, which drives the reference's own ControlledDelay / Queue:
 *   ev_type 0  CLAIM at clock ev_a (ns): ConnectionPool#claim, lib/pool.js:874-885 (claimTimeout =
 *              getMaxIdle()), :887 (counter), :929-966 (idle-first fast path, else enqueue + counters)
 *           1  IDLE  at clock ev_a (ns): ev_b slots turn idle, the handler lib/pool.js:733-757 runs
 *              once per slot, every clock read returning ev_a
 *           2  CANCEL ticket ev_a (handle.cancel() while waiting, lib/connection-fsm.js:580-587)
 * Before a clocked event, waiting handles whose hard timeout is due (clock >= started + claimTimeout,
 * lib/connection-fsm.js:626-642) fail in ticket order.  res receives [ticket, code, slot] triples in
 * resolution order (code 0 admitted, 1 CoDel drop, 2 skipped, 3 hard timeout; at most two per CLAIM: a handle
 * that timed out is still dequeued, and skipped, later);
 * timeouts_ms[ticket] = the claim's claimTimeout.
 */
typedef struct cbo_pool_counters {
	uint64_t claim, queued_claim, max_claim_queue, claim_timeout; /* lib/pool.js:269-283 */
} cbo_pool_counters;
typedef struct cbo_pool_result {
	uint64_t n_res;   /* triples written to res */
	uint64_t idle;    /* slots on the idle queue at the end */
	uint64_t cursor;  /* admits so far (slot = cursor mod nslots) */
	uint64_t queued;  /* handles still on the waiter queue */
	cbo_pool_counters counters;
} cbo_pool_result;
int cbo_pool_sim(cbo_codel *cd, int64_t nslots, uint64_t idle0, uint64_t cursor0, const uint8_t *ev_type,
    const int64_t *ev_a, const int64_t *ev_b, size_t n_events, int64_t *res /* 6 per CLAIM */,
    double *timeouts_ms /* 1 per CLAIM */, cbo_pool_result *out);

/*
 * The pool's low-pass filter and sizing arithmetic ("LPF clamp input"), same event
 * format as cb_lpf_run in oracle/js/driver.js, which drives the reference's own FIRFilter:
 *   cbo_lpf_taps     genTaps(128, -0.2), lib/pool.js:64-76
 *   ty 0  the 5 Hz sampler, lib/pool.js:251-258: a = connections, b = idleq + initq   -> put(busy + spares)
 *   ty 1  _rebalance's target, lib/pool.js:556-587: a = total, b = idleq, c = initq, d = waiters
 *         -> out[3k..] = target, clamped (0/1), FIRFilter.get() (lib/pool.js:90-100)
 * Returns the number of ty-1 events.
 */
#define CBO_LPF_TAPS 128
void cbo_lpf_taps(double *taps /* CBO_LPF_TAPS */);
size_t cbo_lpf_run(int64_t spares, int64_t max, const uint8_t *ty, const int64_t *a, const int64_t *b,
    const int64_t *c, const int64_t *d, size_t n_events, double *out /* 3 per ty-1 event */);

#ifdef __cplusplus
}
#endif
#endif
