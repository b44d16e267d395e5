/*
 * cueball_b200.h -- C ABI of the B200-native claim-queue admission path.
 *
 * The reference (TritonDataCenter/node-cueball v2.10.3) is pure JavaScript and has
 * NO FFI/plugin interface of its own; its seams on this path
 * are plain JS objects.  Each entry point below names the reference interface it
 * replaces (file:line in the reference tree); the N-API stub that binds them to
 * ControlledDelay / Queue / planRebalance-shaped JS classes is in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; every function returns
 *     0 (CB_OK) or a negative CB_E* code; cb_last_error() gives the message.
 *     (Reference convention: programmer errors throw synchronously, e.g.
 *     lib/codel.js:25,57, lib/pool.js:875-878; the JS shim turns CB_E* into throws.)
 *   - "host" pointers may be pageable or pinned (cb_host_alloc); "dev" pointers are
 *     CUDA device pointers on the context's device.
 *   - timestamps are int64 nanoseconds on a monotone clock, i.e. process.hrtime()
 *     as secs*1e9+ns; the device converts them to double milliseconds exactly as
 *     utils.currentMillis() does (lib/utils.js:198-204).  The *_ms variants take the
 *     double milliseconds a JS caller already holds (hdl.ch_started).
 *   - decisions and slot indices are bit-exact against the reference path on the
 *     same inputs; there is NO CPU fallback: without a CUDA device cb_create fails.
 *   - one cb_ctx == one ConnectionPool's claim queue: one ControlledDelay state
 *     (lib/pool.js:195-200), one preference list / slot table, one CUDA stream.
 *     A context is not re-entrant (Node's event loop is single threaded too).
 */
#ifndef CUEBALL_B200_H
#define CUEBALL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 1

/* return codes */
#define CB_OK 0
#define CB_EINVAL (-1)  /* bad argument (the reference would throw an AssertionError) */
#define CB_ECUDA (-2)   /* CUDA runtime / driver failure (incl. "no device") */
#define CB_ENOMEM (-3)
#define CB_ESTATE (-4)  /* call out of order (e.g. batch before cb_codel_init) */
#define CB_ENCCL (-5)   /* NCCL unavailable or failed */

/* per-claim decision codes written to `decision` */
#define CB_ADMIT 0   /* hdl.try(fsm)                      lib/pool.js:747 */
#define CB_DROP 1    /* hdl.timeout() -> ClaimTimeoutError lib/pool.js:742-745, lib/errors.js:35 */
#define CB_SKIPPED 2 /* handle no longer 'waiting'         lib/pool.js:738-740 */

/* per-claim input flag bits (`flags` arrays; NULL means every claim is CB_F_WAITING) */
#define CB_F_WAITING 1     /* hdl.isInState('waiting')        lib/pool.js:738 */
#define CB_F_EMPTY_AFTER 2 /* waiter queue ran empty with a slot idle right after this
                              claim: codel.empty()            lib/pool.js:751-753 */

/* cb_batch_stats.path */
#define CB_PATH_PARALLEL 0   /* monotone clock: three-kernel decomposition */
#define CB_PATH_SEQUENTIAL 1 /* clock went backwards inside the batch (or adversarial
                                episode density): exact one-warp state machine */

typedef struct cb_ctx cb_ctx;

/* ControlledDelay's fields, lib/codel.js:24-32 (+ cd_last_empty, :92) */
typedef struct cb_codel_state {
	double target_ms;        /* cd_targdelay */
	double first_above_time; /* cd_first_above_time (0 = unset) */
	double drop_next;        /* cd_drop_next */
	uint64_t count;          /* cd_count (a JS Number; exact below 2^53) */
	int32_t dropping;        /* cd_dropping */
	int32_t has_last_empty;  /* 0 = cd_last_empty is undefined */
	double last_empty;       /* cd_last_empty */
} cb_codel_state;

typedef struct cb_batch_stats {
	uint64_t n;              /* claims examined in the last batch */
	uint64_t drops;          /* CB_DROP  -> counter 'claim-timeout', lib/pool.js:269-283 */
	uint64_t admits;         /* CB_ADMIT */
	uint64_t skipped;        /* CB_SKIPPED */
	uint64_t cursor;         /* admit cursor after the batch (slot = cursor mod nslots) */
	uint64_t total_drops;    /* cumulative over the context's life */
	uint64_t total_admits;
	uint64_t global_drops;   /* summed over ranks by cb_allreduce_counts (else = drops) */
	uint64_t global_admits;
	uint64_t global_n;
	cb_codel_state state;    /* ControlledDelay after the batch */
	int32_t path;            /* CB_PATH_* */
	uint32_t n_events;       /* class-change events found by the scan kernel */
	uint32_t n_intervals;    /* drop intervals produced by the resolve kernel */
	uint32_t reserved;
} cb_batch_stats;

/* one SRV answer as the resolver hands it over in 'added'(key, backend),
 * docs/api.adoc:354-453.  The reference drops priority/weight (lib/resolver.js:1358);
 * they are honoured only in CB_BALANCE_SRV_WEIGHTED mode (an extension). */
typedef struct cb_backend {
	int32_t priority; /* RFC 2782; lower = preferred */
	int32_t weight;   /* RFC 2782 */
} cb_backend;

#define CB_BALANCE_UNIFORM 0      /* the reference's behaviour: round-robin planRebalance */
#define CB_BALANCE_SRV_WEIGHTED 1 /* extension, no reference oracle ("parity unpinned") */

/* ---------------------------------------------------------------- lifecycle */

/* Create a context on CUDA device `device`.  Fails (CB_ECUDA) when no device or the
 * device is not sm_100.  Replaces: new ConnectionPool's p_codel/p_waiters/p_keys
 * construction, lib/pool.js:191-200. */
int cb_create(cb_ctx **out, int device);
void cb_destroy(cb_ctx *ctx);
const char *cb_last_error(const cb_ctx *ctx); /* ctx may be NULL: last create error */
int cb_abi_version(void);
/* Order the context's work after a caller-owned CUDA stream (cudaStream_t); NULL = own stream.
 * Device-buffer batches are PIPELINED: up to `depth` batches are in flight on internal streams
 * (scan of batch k+1 overlaps resolve of batch k and emit of batch k-1).  Every batch starts
 * after the work already queued on the caller's stream; its outputs are complete after
 * cb_join() (stream-ordered) or cb_synchronize()/cb_batch_stats_get() (host-blocking). */
int cb_set_stream(cb_ctx *ctx, void *cuda_stream);
int cb_set_pipeline_depth(cb_ctx *ctx, int depth); /* 1 = strictly in order on one stream; default 8, at most 12 */
int cb_join(cb_ctx *ctx);        /* the caller's stream waits for every batch enqueued so far */
int cb_synchronize(cb_ctx *ctx); /* the host waits for every batch enqueued so far */
/* pinned host memory for the *_batch host-buffer calls */
void *cb_host_alloc(cb_ctx *ctx, size_t bytes);
void cb_host_free(cb_ctx *ctx, void *p);

/* -------------------------------------------------------------------- CoDel */

/* new ControlledDelay(targetClaimDelay): lib/codel.js:24-32, called lib/pool.js:197-200.
 * target_ms must be a positive integer number of ms (utils.assertClaimDelay,
 * lib/utils.js:188-195) else CB_EINVAL. */
int cb_codel_init(cb_ctx *ctx, double target_ms);
int cb_codel_get_state(cb_ctx *ctx, cb_codel_state *out);
int cb_codel_set_state(cb_ctx *ctx, const cb_codel_state *in);
/* ControlledDelay#empty(): lib/codel.js:91-94 at clock now_ns */
int cb_codel_empty(cb_ctx *ctx, int64_t now_ns);
/* ControlledDelay#getMaxIdle(): lib/codel.js:109-118, called per claim() lib/pool.js:880 */
int cb_codel_max_idle(cb_ctx *ctx, int64_t now_ns, double *out_ms);

/* ------------------------------------------------------- the admission batch */

/*
 * The slot-idle drain loop, lib/pool.js:733-753, over a batch of n pending claims in
 * FIFO order (lib/queue.js): claim i was enqueued at start_ns[i] (hdl.ch_started,
 * lib/connection-fsm.js:523) and is dequeued at clock now_ns[i] (the clock read of
 * ControlledDelay#overloaded, lib/codel.js:58).  For every claim, in order:
 * overloaded() advances the CoDel state (also for non-waiting handles, pool.js:735-740),
 * then the claim is skipped / dropped / admitted; the j-th admitted claim of the
 * context's life takes slot (j mod nslots) of the slot table.
 *
 * Outputs (any may be NULL): decision[i] in CB_*, slot[i] (-1 unless admitted),
 * backend[i] = slot table entry (-1 unless admitted).  stats may be NULL.
 * Host-buffer form: copies in, computes, copies out, returns when done.
 */
int cb_admit_batch(cb_ctx *ctx, const int64_t *start_ns, const int64_t *now_ns,
    const uint8_t *flags, size_t n, uint8_t *decision, int32_t *slot,
    int32_t *backend, cb_batch_stats *stats);
int cb_admit_batch_ms(cb_ctx *ctx, const double *start_ms, const double *now_ms,
    const uint8_t *flags, size_t n, uint8_t *decision, int32_t *slot,
    int32_t *backend, cb_batch_stats *stats);
/* Device-buffer form: asynchronous on the context's stream; the queue segment stays
 * resident in HBM.  Read results with cb_batch_stats_get (which synchronises). */
int cb_admit_batch_dev(cb_ctx *ctx, const int64_t *d_start_ns, const int64_t *d_now_ns,
    const uint8_t *d_flags, size_t n, uint8_t *d_decision, int32_t *d_slot,
    int32_t *d_backend);
int cb_admit_batch_ms_dev(cb_ctx *ctx, const double *d_start_ms, const double *d_now_ms,
    const uint8_t *d_flags, size_t n, uint8_t *d_decision, int32_t *d_slot,
    int32_t *d_backend);
int cb_batch_stats_get(cb_ctx *ctx, cb_batch_stats *out);
/* ControlledDelay#overloaded(start) for ONE claim (lib/codel.js:56-86): the scalar
 * form the unmodified pool.js loop calls; *drop = 0/1. */
int cb_codel_overloaded(cb_ctx *ctx, int64_t start_ns, int64_t now_ns, int *drop);
/* the same with both clocks as the reference's double milliseconds (utils.currentMillis(),
 * lib/utils.js:198-204): what an unmodified pool.js hands over, p_codel.overloaded(hdl.ch_started),
 * lib/pool.js:735-736 with lib/connection-fsm.js:523 */
int cb_codel_overloaded_ms(cb_ctx *ctx, double start_ms, double now_ms, int *drop);

/* ----------------------------------------- backends, preference list, slots */

/*
 * Resolver 'added'(key, backend) for backends 0..B-1 in order (lib/pool.js:285-291):
 * key k is inserted at index floor(insert_u[k] * (len+1)) of the preference list
 * (insert_u[k] in [0,1) stands in for Math.random()), then the slot table is rebuilt
 * as planRebalance({}, {}, nslots, nslots) (lib/pool.js:590, lib/utils.js:239-393):
 * slot s belongs to backend slot_backend[s], add-order of plan.add (pool.js:632-654).
 * tbl may be NULL in CB_BALANCE_UNIFORM mode.
 */
int cb_backends_set(cb_ctx *ctx, const cb_backend *tbl, int32_t B, const double *insert_u,
    int32_t nslots, int mode);
/* single resolver events + decoherence reshuffle (lib/pool.js:285-313, :501-519);
 * each rebuilds the slot table for the current nslots */
int cb_backend_added(cb_ctx *ctx, int32_t key, const cb_backend *be, double u);
int cb_backend_removed(cb_ctx *ctx, int32_t key);
int cb_reshuffle(cb_ctx *ctx, double u);
/* read back: preference order (keys_out[B]), slot table (slot_backend_out[nslots]) */
int cb_backends_get(cb_ctx *ctx, int32_t *n_keys, int32_t *keys_out, int32_t *n_slots,
    int32_t *slot_backend_out);

/* utils.shuffle(array), lib/utils.js:207-217 (Fisher-Yates from the top; exported by the reference, no caller
 * in lib/): arr[n] is permuted in place, us[t] in [0,1) stands in for the t-th Math.random().  One kernel,
 * every output position traced independently (cb_topology.cuh). */
int cb_shuffle(cb_ctx *ctx, int32_t *arr, size_t n, const double *us);

/*
 * utils.planRebalance(connections, dead, target, max, singleton): lib/utils.js:239-393,
 * callers lib/pool.js:590 and lib/set.js:408.  Keys are positions 0..B-1 in the caller's
 * key order; have[k] = open connections of key k, dead[k] != 0 = declared dead.
 * want_out[B] = wantedSpares; add_out (<= max entries) = plan.add as key positions,
 * forward order; rm_key_out/rm_ord_out (<= sum(have)) = plan.remove as (key, ordinal of
 * the connection within its key's list), reverse key order, oldest first.
 * CB_EINVAL where the reference asserts (target < 0, max < target).
 */
int cb_plan_rebalance(cb_ctx *ctx, const int32_t *have, const uint8_t *dead, int32_t B,
    int32_t target, int32_t max, int singleton, int32_t *want_out, int32_t *add_out,
    int32_t *n_add, int32_t *rm_key_out, int32_t *rm_ord_out, int32_t *n_rm);

/* ------------------------------------------------------- resolver ingestion */

/* the identity and the event stream either side of the preference list.
 *   cb_srv_keys          srvKey() of n backends at once (lib/resolver.js:1157-1171): base64 of
 *                        sha1(name || "||" || String(port) || "||" || ipaddr.js-normalised address), SHA-1 on
 *                        the device.  keys_out: n x CB_SRVKEY_LEN characters, no terminator.  CB_EINVAL when
 *                        an address is not an IP address (ipaddr.parse throws).
 *   cb_resolver_process  CueBallDNSResolver#state_process (lib/resolver.js:1024-1108) on the context's
 *                        backend map: `keys` are the srvKeys of the new SRV/A result in record order (a key that
 *                        appears twice keeps its first position, the later record wins); outputs the 'removed'
 *                        events (old map order) and then the 'added' events (new map order) the reference
 *                        emits, as key ids for cb_backend_removed / cb_backend_added (ids are handed out here,
 *                        0, 1, 2, ... in order of first appearance) plus, for 'added', the input row of the
 *                        backend record.  An empty result changes nothing and sets *empty_set ('empty-set'
 *                        counter, :1052-1061).  Output arrays may be NULL; room for n entries ('added') and for
 *                        the current number of backends ('removed'). */
#define CB_SRVKEY_LEN 28
int cb_srv_keys(cb_ctx *ctx, const char *const *names, const int32_t *ports, const char *const *addresses, size_t n,
    char *keys_out);
int cb_resolver_process(cb_ctx *ctx, const char *keys, size_t n, int32_t *removed_ids, char *removed_keys,
    int32_t *n_removed, int32_t *added_ids, int32_t *added_rows, int32_t *n_added, int32_t *empty_set);

/* ----------------------------------------------------------------- multi-GPU */

/* One process per GPU; the queue is sharded by contiguous segment, each rank's
 * segment has its own ControlledDelay; one ncclAllReduce(sum) per batch gives every
 * rank the global drop/admit counts (cb_batch_stats.global_*). */
int cb_comm_unique_id(void *out, size_t *bytes); /* ncclGetUniqueId; bytes in/out (>=128) */
int cb_comm_init(cb_ctx *ctx, int nranks, int rank, const void *unique_id, size_t bytes);
int cb_allreduce_counts(cb_ctx *ctx); /* async on the context's stream */

/* The same global counts after EVERY batch without a collective: each rank owns a board in HBM and stores
 * the batch's {drops, admits, n} into it (self-tagged 16-byte words); one thread of an extra CTA of the
 * batch's emit kernel then reads that batch's row of every rank's board over NVLink peer memory and keeps
 * running totals.  No rank waits for another on the data path (flow control only holds a rank that is 64
 * batches ahead of the slowest reader).  Every rank must submit the same number of batches.
 *   cb_peers_board   : this context's board as a CUDA IPC handle (64 bytes) and/or a device pointer.
 *   cb_peers_connect : join `nranks` (<= 16) as `rank`; boards by IPC handle (nranks x 64 bytes, entry
 *                      `rank` ignored) or by pointer (contexts of this process); nranks == 0 disconnects.
 *   cb_peers_counts  : running global totals after this rank's latest batch (waits, bounded, until every
 *                      rank has posted it).  Counter 'claim-timeout' of the whole sharded pool,
 *                      lib/pool.js:269-283. */
typedef struct cb_global_counts {
	uint64_t steps;  /* batches every rank has completed */
	uint64_t drops, admits, n;
} cb_global_counts;
int cb_peers_board(cb_ctx *ctx, void *ipc_handle_out, size_t *len, void **local_ptr_out);
int cb_peers_connect(cb_ctx *ctx, int nranks, int rank, const void *ipc_handles, void *const *local_ptrs);
int cb_peers_counts(cb_ctx *ctx, cb_global_counts *out);

/* One logical queue over R GPUs: batch k of rank r is the queue segment that
 * follows batch k of rank r-1 (and batch k of rank 0 follows batch k-1 of rank R-1), exactly as if
 * one ConnectionPool had drained the concatenation.  The ControlledDelay state and the admit cursor
 * travel round a ring of device mailboxes: the kernel that commits a batch waits for its
 * predecessor's state in this rank's mailbox and stores its own result into the next rank's mailbox
 * through NVLink peer memory -- no host round trip, no collective on the data path.  Every rank must
 * submit the same number of batches; rank r's cb_codel_get_state is the state after its latest
 * segment (the last rank's is the queue's).
 *   cb_ring_mailbox : this context's mailbox as a CUDA IPC handle (64 bytes; for other processes)
 *                     and/or as a device pointer (for contexts of the same process).
 *   cb_ring_connect : join a ring of `nranks` as `rank`, given the NEXT rank's ((rank+1)%nranks)
 *                     mailbox by IPC handle or by pointer; nranks == 0 leaves the ring.  Call it on
 *                     every rank before the first batch, after cb_codel_init / cb_codel_set_state
 *                     (rank 0's state at that moment is the queue's initial state). */
int cb_ring_mailbox(cb_ctx *ctx, void *ipc_handle_out, size_t *len, void **local_ptr_out);
int cb_ring_connect(cb_ctx *ctx, int nranks, int rank, const void *next_ipc_handle, void *next_local_ptr);

/* ------------------------------------------------- pending-claim queue in HBM */

/* The enqueue side and the event form of the drain loop.  Specified by
 * tools/proto_queue.py, checked against the pinned model of the reference on CPU (tests/test_proto_queue.py)
 * and on a B200 (tests/test_gpu_queue.py, tests/test_gpu_pool_mirror.py).
 *
 * cb_queue_push    n x ConnectionPool#claim (lib/pool.js:874-966): claimTimeout = getMaxIdle()
 *                  (lib/codel.js:109-118) per claim, idle slots first, the rest appended to the device
 *                  queue; tickets are first_ticket, first_ticket+1, ...  Enqueue clocks (start_ns) must
 *                  be non-decreasing across pushes.  Timers due at start_ns[0] fire first.
 * cb_queue_cancel  handle.cancel() of waiting handles (lib/connection-fsm.js:580-587).
 * cb_queue_expire  hard timeouts due at now_ns (lib/connection-fsm.js:626-642) -> ClaimTimeoutError.
 * cb_queue_idle    k slots turn idle at clock now_ns: the handler lib/pool.js:733-757 once per slot with
 *                  every clock read returning now_ns.  One kernel: the claims at or above the target
 *                  delay are a prefix of the FIFO (bisection), ControlledDelay#overloaded has a closed
 *                  form along it, and the number of handles the k slots consume is a prefix-sum
 *                  selection over the waiting bits.  Expires due timers first; calls codel.empty() when
 *                  the queue runs dry with slots left, which then sit on the idle queue.
 * cb_queue_results resolution records since the last call, in the order the reference would resolve
 *                  the handles: (ticket, code, slot) with code 0 admitted (hdl.try), 1 dropped by CoDel
 *                  (hdl.timeout()), 2 skipped (no longer waiting when dequeued), 3 hard timeout.
 * All calls are synchronous.  The ControlledDelay state, admit cursor and counters are the context's
 * (cb_codel_get_state, cb_batch_stats_get) and are shared with the batch entry points. */
typedef struct cb_queue_info {
	uint64_t head, tail;   /* absolute FIFO positions; tail - head handles are queued */
	uint64_t next_ticket;
	uint64_t idle;         /* slots on the idle queue (p_idleq.length) */
	uint64_t claim, queued_claim, max_claim_queue, claim_timeout; /* lib/pool.js:269-283 counters */
	uint64_t n_results;    /* records waiting for cb_queue_results */
	uint32_t error, reserved;
} cb_queue_info;
int cb_queue_create(cb_ctx *ctx, size_t capacity, size_t result_capacity, uint64_t idle_slots);
int cb_queue_push(cb_ctx *ctx, const int64_t *start_ns, size_t n, int64_t *first_ticket, double *timeouts_ms);
int cb_queue_cancel(cb_ctx *ctx, const int64_t *tickets, size_t n);
int cb_queue_expire(cb_ctx *ctx, int64_t now_ns);
int cb_queue_idle(cb_ctx *ctx, int64_t now_ns, uint64_t k_slots);
/* up to k idle slots leave the idle queue (slots that went away with a topology change: setUnwanted on
 * idle slots, lib/pool.js:601-631); *removed (may be NULL) = how many did */
int cb_queue_trim_idle(cb_ctx *ctx, uint64_t k, uint64_t *removed);
int cb_queue_results(cb_ctx *ctx, int64_t *triples, size_t cap_records, size_t *n_records);
int cb_queue_info_get(cb_ctx *ctx, cb_queue_info *out);

/* ------------------------------------------------------------- measurement */

/* Per-kernel CUDA-event timing on the launching stream (bench.py's roofline leg).
 * on = 1: every batch records events around each kernel; cb_profile_get synchronises and returns
 *         accumulated milliseconds and launch counts.
 * on = 2: a device-buffer batch call launches ONLY its scan kernel, on the context's stream and without
 *         events: the caller brackets a burst of such calls over different inputs with its own two events
 *         (one kernel between two events also measures the launch latency; a burst averages it out).
 * on = 0: off. */
#define CB_K_SCAN 0    /* pass 1: timestamps -> CoDel classes / events (reads 16 B/claim) */
#define CB_K_RESOLVE 1 /* resolve kernel (prepare + commit halves): episode resolution (serial over episodes, tiny) */
#define CB_K_EMIT 2    /* pass 3: decisions + slots (writes 5 B/claim) */
#define CB_K_COUNT 3
int cb_profile_enable(cb_ctx *ctx, int on);
int cb_profile_get(cb_ctx *ctx, double ms_out[CB_K_COUNT], uint64_t launches_out[CB_K_COUNT]);
uint64_t cb_launch_count(const cb_ctx *ctx); /* kernels launched so far by this context */
/* %globaltimer stamps (ns) of the last batch's prepare / commit kernels at their phase boundaries:
 * prepare [6] entry, [0] pairing, [1] run searches, [2] merge, [3] speculation, [4] end;
 * commit [7] entry, [5] done */
int cb_debug_timestamps(cb_ctx *ctx, uint64_t out[16]);
/* Same for the batch that last used pipeline lane `lane` (0 .. depth-1): out[8]/[9] scan first-CTA
 * start / last-tile end, [6]/[4] prepare, [7]/[5] commit, [10]/[11] emit (tools/pipeline_timeline.py). */
int cb_debug_lane_timestamps(cb_ctx *ctx, int lane, uint64_t out[16]);

#ifdef __cplusplus
}
#endif
#endif /* CUEBALL_B200_H */
