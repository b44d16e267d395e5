// cueball_b200.cu -- the C ABI (include/cueball_b200.h) over the sm_100a kernels.
// Host-side plumbing only: buffers, streams, launches, the batch pipeline (up to 8 lanes, one
// stream each: a batch's scan -> prepare -> commit -> emit are stream-ordered, the commits of
// consecutive batches are ordered on the device by the chain word / ring mailbox), the chunked
// H2D/compute/D2H pipeline of the host-buffer entry points, the pending-claim queue entry points
// and a dlopen()ed NCCL for the one collective.
#include "../../include/cueball_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <new>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "cb_device.cuh"
#include "cb_emit.cuh"
#include "cb_queue.cuh"
#include "cb_resolve.cuh"
#include "cb_scan.cuh"
#include "cb_topology.cuh"

using namespace cb;

namespace {

char g_create_err[512] = "";

// ---- NCCL through dlopen: no link-time dependency; multi-GPU is optional --------------
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm_t;
struct NcclApi {
	void *h = nullptr;
	int (*GetUniqueId)(nccl_uid *) = nullptr;
	int (*CommInitRank)(nccl_comm_t *, int, nccl_uid, int) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
	int (*CommDestroy)(nccl_comm_t) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	bool load() {
		if (h)
			return true;
		h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
		if (!h)
			h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h)
			h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h)
			return false;
		GetUniqueId = (int (*)(nccl_uid *))dlsym(h, "ncclGetUniqueId");
		CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid, int))dlsym(h, "ncclCommInitRank");
		AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
		CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
		GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
		return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
	}
};
NcclApi g_nccl;
constexpr int NCCL_UINT64 = 5, NCCL_SUM = 0;

constexpr int CB_LANES = 8;                   // most batches in flight (lanes)
constexpr int CB_DEFAULT_DEPTH = 5;           // lanes used unless CB_PIPELINE_DEPTH / cb_set_pipeline_depth says otherwise
constexpr int NBUF = 3;                       // staging buffers of the host-buffer pipeline
constexpr size_t DEFAULT_CHUNK = 4u << 20;    // claims per pipeline chunk

// per-batch scratch: everything one batch's four kernels hand to each other
struct Lane {
	size_t cap_tiles = 0;
	BatchCtl *ctl = nullptr;
	SumDesc *sdesc = nullptr;
	unsigned int *ev_count = nullptr, *rec_fc = nullptr, *rec_cend = nullptr, *xcnt = nullptr, *hj_info = nullptr;
	unsigned int *ticket = nullptr;
	unsigned long long *ev = nullptr, *evc = nullptr, *ev_q = nullptr, *ev_d = nullptr;
	unsigned long long *rec_code = nullptr, *xev = nullptr, *hj_lo = nullptr, *hj_b = nullptr, *tile_now0 = nullptr;
	unsigned long long *iv_lo = nullptr, *iv_hi = nullptr, *iv_cum = nullptr;
	double *ev_now = nullptr, *hj_F = nullptr, *idx_ms = nullptr;
	cudaStream_t st = nullptr; // all four kernels of the lane's batch, in order (highest stream priority)
	cudaEvent_t res_done = nullptr, emit_done = nullptr, coll_done = nullptr, join_ev = nullptr;
	bool used = false, coll_pending = false;
	unsigned long long epoch = 0;
	// CB_GRAPH=1 (not yet validated on hardware): the batch's kernels as one CUDA graph per
	// (clock type, flags) variant, its node parameters rewritten per batch: one launch call per batch
	struct Graph {
		cudaGraph_t g = nullptr;
		cudaGraphExec_t ex = nullptr;
		cudaGraphNode_t n_scan = nullptr, n_res = nullptr, n_commit = nullptr, n_emit = nullptr;
	} gr[2][2];
};

struct Staging {
	void *start = nullptr, *now = nullptr;
	uint8_t *flags = nullptr, *decision = nullptr;
	int32_t *slot = nullptr, *backend = nullptr;
	cudaEvent_t in_done = nullptr, out_done = nullptr;
};

} // namespace

struct cb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;     // "main" stream: the caller's, or own_stream
	cudaStream_t own_stream = nullptr;
	bool user_stream = false;
	cudaStream_t st_coll = nullptr;
	cudaStream_t s_in = nullptr, s_out = nullptr;
	cudaEvent_t ev_user = nullptr;
	DevState *g = nullptr;
	bool codel_ready = false;
	double target_ms = 0;
	char err[512] = "";
	uint64_t launches = 0;
	// batch pipeline
	int depth = CB_DEFAULT_DEPTH;
	Lane lane[CB_LANES];
	unsigned long long seq = 0;        // batches enqueued so far
	ChainMem *d_chain = nullptr;           // the chain word; in ring mode this rank's mailbox
	unsigned long long chain_seq = 0;      // batches enqueued (since the ring was connected, in ring mode)
	// one logical queue over several GPUs: the state goes round a ring of mailboxes (peer memory)
	// totals as of the last cb_batch_stats_get-style read, valid while no batch has been enqueued since
	bool tot_valid = false;
	unsigned long long tot_seq = 0, tot_drops_h = 0, tot_admits_h = 0, tot_epoch = 0, epoch = 0;
	// the pending-claim queue resident in HBM (cb_queue_*)
	QueueDev *d_q = nullptr;
	long long *q_start = nullptr, *q_deadline = nullptr, *q_ticket = nullptr, *q_res = nullptr, *q_stage = nullptr;
	unsigned char *q_waiting = nullptr;
	double *q_tmo = nullptr;
	size_t q_cap = 0, q_res_cap = 0, q_stage_cap = 0;
	// F-chain variant of the commit kernel (CB_FCHAIN=1; not yet validated on hardware)
	bool fchain = false, fseed = true;
	bool fused = false; // CB_FUSED=1
	bool graph = false; // CB_GRAPH=1
	int prio_least = 0, prio_greatest = 0;
	int ring_n = 0, ring_rank = 0;
	ChainMem *ring_next = nullptr;         // the next rank's mailbox
	bool ring_next_ipc = false;            // ... opened from an IPC handle (to be closed)
	int last_lane = -1;
	bool last_flags = false;
	int emit_grid = 148 * 4;           // persistent emit CTAs (from the device's SM count)
	// topology
	int B = 0, nslots = 0, target_slots = 0, mode = CB_BALANCE_UNIFORM, key_cap = 0, slot_cap = 0;
	int32_t *d_keys = nullptr, *d_want = nullptr, *d_add = nullptr, *d_slot_backend = nullptr;
	int32_t *d_prio = nullptr, *d_weight = nullptr;
	int *d_ints = nullptr; // [0]=len [1]=n_add [2]=nslots [3]=status [4]=n_rm [5]=scalar out
	double *d_dbl = nullptr;
	// host-buffer pipeline
	size_t chunk = DEFAULT_CHUNK, stage_cap = 0;
	Staging st[NBUF];
	// profiling
	bool prof = false;
	cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
	double prof_ms[CB_K_COUNT] = {0, 0, 0};
	uint64_t prof_n[CB_K_COUNT] = {0, 0, 0};
	// nccl
	nccl_comm_t comm = nullptr;
	int nranks = 1, rank = 0;
};

namespace {

int fail(cb_ctx *c, int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(c ? c->err : g_create_err, 512, fmt, ap);
	va_end(ap);
	return code;
}
#define CU(c, call)                                                                              \
	do {                                                                                         \
		cudaError_t e_ = (call);                                                                 \
		if (e_ != cudaSuccess)                                                                   \
			return fail((c), CB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

template <typename T> int dev_alloc(cb_ctx *c, T **p, size_t count) {
	if (*p) {
		cudaFree(*p);
		*p = nullptr;
	}
	cudaError_t e = cudaMalloc((void **)p, count * sizeof(T) + 64);
	if (e != cudaSuccess)
		return fail(c, CB_ENOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
	return CB_OK;
}

static int ring_close(cb_ctx *c) {
	if (c->ring_next && c->ring_next_ipc)
		cudaIpcCloseMemHandle(c->ring_next);
	c->ring_next = nullptr;
	c->ring_next_ipc = false;
	c->ring_n = 0;
	return CB_OK;
}

// wait for every batch in flight (and the main stream)
int drain(cb_ctx *c) {
	for (int l = 0; l < CB_LANES; l++) {
		if (c->lane[l].st)
			CU(c, cudaStreamSynchronize(c->lane[l].st));
	}
	CU(c, cudaStreamSynchronize(c->st_coll));
	CU(c, cudaStreamSynchronize(c->stream));
	return CB_OK;
}

// A lane's two streams are created when the lane is first used: the device multiplexes streams
// onto a few hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default), and every stream that
// exists makes it likelier that two busy ones share a queue and serialise.
int lane_streams(cb_ctx *c, Lane &L) {
	if (L.st)
		return CB_OK;
	int least = 0, greatest = 0;
	CU(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));
	c->prio_least = least;
	c->prio_greatest = greatest;
	CU(c, cudaStreamCreateWithPriority(&L.st, cudaStreamNonBlocking, greatest));
	return CB_OK;
}

int ensure_scratch(cb_ctx *c, Lane &L, size_t tiles) {
	if (tiles <= L.cap_tiles)
		return CB_OK;
	// growing a lane's scratch: nothing of this lane may be in flight
	if (L.used) {
		if (L.st)
			CU(c, cudaStreamSynchronize(L.st));
		if (L.coll_pending)
			CU(c, cudaEventSynchronize(L.coll_done));
	}
	size_t cap = tiles + tiles / 4 + 64;
	const size_t evcap = (cap + 1) * (EVK + 3) + EV_FRONT;
	int rc;
	if ((rc = dev_alloc(c, &L.sdesc, cap)) || (rc = dev_alloc(c, &L.ev_count, cap)) ||
	    (rc = dev_alloc(c, &L.ev, cap * EVK)) || (rc = dev_alloc(c, &L.rec_fc, cap)) ||
	    (rc = dev_alloc(c, &L.rec_code, cap)) || (rc = dev_alloc(c, &L.rec_cend, cap)) ||
	    (rc = dev_alloc(c, &L.tile_now0, cap)) || (rc = dev_alloc(c, &L.xev, (cap + 1) * 3)) ||
	    (rc = dev_alloc(c, &L.xcnt, cap + 1)) || (rc = dev_alloc(c, &L.hj_lo, cap + 1)) ||
	    (rc = dev_alloc(c, &L.hj_b, cap + 1)) || (rc = dev_alloc(c, &L.hj_F, cap + 1)) ||
	    (rc = dev_alloc(c, &L.hj_info, cap + 1)) || (rc = dev_alloc(c, &L.evc, evcap)) ||
	    (rc = dev_alloc(c, &L.ev_q, evcap)) || (rc = dev_alloc(c, &L.ev_d, evcap)) ||
	    (rc = dev_alloc(c, &L.ev_now, evcap)) || (rc = dev_alloc(c, &L.iv_lo, evcap + 8)) ||
	    (rc = dev_alloc(c, &L.iv_hi, evcap + 8)) || (rc = dev_alloc(c, &L.iv_cum, evcap + 8)))
		return rc;
	if (!L.idx_ms && (rc = dev_alloc(c, &L.idx_ms, NIDX_MAX)))
		return rc;
	// the emit kernel's sum descriptors carry an epoch tag, so they only need zeroing once
	CU(c, cudaMemset(L.sdesc, 0, cap * sizeof(SumDesc)));
	L.cap_tiles = cap;
	return CB_OK;
}

// CB_GRAPH=1: scan -> (prepare -> commit | fused resolve) -> emit as the nodes of one CUDA graph per lane and
// variant.  The first batch of a variant builds and instantiates it; every later one rewrites the nodes'
// parameters (pointers, sizes, grids) in the instantiated graph and launches it: one launch call per batch
// instead of three or four.  The scan node keeps the lowest priority, the others the highest.
template <bool NS>
int graph_batch(cb_ctx *c, Lane &L, bool flags, bool use_f, bool fused, ScanParams &sp, ResolveParams &rp,
    EmitParams &ep, size_t tiles, unsigned egrid, size_t cidx_bytes, cudaStream_t s) {
	Lane::Graph &G = L.gr[NS ? 1 : 0][flags ? 1 : 0];
	void *a_scan[] = {(void *)&sp}, *a_res[] = {(void *)&rp}, *a_emit[] = {(void *)&ep};
	cudaKernelNodeParams k_scan = {}, k_res = {}, k_commit = {}, k_emit = {};
	k_scan.func = flags ? (void *)scan_kernel<NS, true> : (void *)scan_kernel<NS, false>;
	k_scan.gridDim = dim3((unsigned)tiles);
	k_scan.blockDim = dim3(TPB);
	k_scan.kernelParams = a_scan;
	if (fused) {
		k_res.func = use_f ? (flags ? (void *)resolve_fused_kernel<NS, true, true> : (void *)resolve_fused_kernel<NS, false, true>)
		                   : (flags ? (void *)resolve_fused_kernel<NS, true, false> : (void *)resolve_fused_kernel<NS, false, false>);
	} else {
		k_res.func = flags ? (void *)prepare_kernel<NS, true> : (void *)prepare_kernel<NS, false>;
		k_commit.func = use_f ? (flags ? (void *)commit_fchain_kernel<NS, true> : (void *)commit_fchain_kernel<NS, false>)
		                      : (flags ? (void *)commit_kernel<NS, true> : (void *)commit_kernel<NS, false>);
		k_commit.gridDim = dim3(1);
		k_commit.blockDim = dim3(CTPB);
		k_commit.sharedMemBytes = (unsigned)cidx_bytes;
		k_commit.kernelParams = a_res;
	}
	k_res.gridDim = dim3(1);
	k_res.blockDim = dim3(RTPB);
	k_res.sharedMemBytes = (unsigned)(NIDX_MAX * sizeof(double));
	k_res.kernelParams = a_res;
	k_emit.func = flags ? (void *)emit_kernel<true> : (void *)emit_kernel<false>;
	k_emit.gridDim = dim3(egrid);
	k_emit.blockDim = dim3(ETPB);
	k_emit.sharedMemBytes = (unsigned)ep.table_smem_bytes;
	k_emit.kernelParams = a_emit;
	if (!G.ex) {
		CU(c, cudaGraphCreate(&G.g, 0));
		CU(c, cudaGraphAddKernelNode(&G.n_scan, G.g, nullptr, 0, &k_scan));
		CU(c, cudaGraphAddKernelNode(&G.n_res, G.g, &G.n_scan, 1, &k_res));
		cudaGraphNode_t last = G.n_res;
		if (!fused) {
			CU(c, cudaGraphAddKernelNode(&G.n_commit, G.g, &G.n_res, 1, &k_commit));
			last = G.n_commit;
		}
		CU(c, cudaGraphAddKernelNode(&G.n_emit, G.g, &last, 1, &k_emit));
		cudaKernelNodeAttrValue lo, hi;
		memset(&lo, 0, sizeof(lo));
		memset(&hi, 0, sizeof(hi));
		lo.priority = c->prio_least;
		hi.priority = c->prio_greatest;
		CU(c, cudaGraphKernelNodeSetAttribute(G.n_scan, cudaKernelNodeAttributePriority, &lo));
		CU(c, cudaGraphKernelNodeSetAttribute(G.n_res, cudaKernelNodeAttributePriority, &hi));
		if (G.n_commit)
			CU(c, cudaGraphKernelNodeSetAttribute(G.n_commit, cudaKernelNodeAttributePriority, &hi));
		CU(c, cudaGraphKernelNodeSetAttribute(G.n_emit, cudaKernelNodeAttributePriority, &hi));
		CU(c, cudaGraphInstantiate(&G.ex, G.g, 0));
	} else {
		CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_scan, &k_scan));
		CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_res, &k_res));
		if (G.n_commit)
			CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_commit, &k_commit));
		CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_emit, &k_emit));
	}
	CU(c, cudaGraphLaunch(G.ex, s));
	return CB_OK;
}

// Enqueue the four kernels of one batch.  depth 1 (or profiling): all on the main stream, in
// order.  depth > 1: the batch goes to its lane's stream (lanes rotate), the commit kernels are
// chained across lanes by one event, so consecutive batches overlap; `in_ready` (optional) gates
// the scan on the batch's inputs, and the batch's emit_done event is returned through `out_done`.
template <bool NS>
int enqueue_batch(cb_ctx *c, const void *d_start, const void *d_now, const uint8_t *d_flags, size_t n,
    uint8_t *d_decision, int32_t *d_slot, int32_t *d_backend, cudaEvent_t in_ready, cudaEvent_t *out_done) {
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	if (out_done)
		*out_done = nullptr;
	if (n == 0)
		return CB_OK;
	if (n > (1ull << 40))
		return fail(c, CB_EINVAL, "batch too large");
	const bool piped = (c->depth > 1) && !c->prof;
	const int li = piped ? (int)(c->seq % (unsigned long long)c->depth) : 0;
	Lane &L = c->lane[li];
	const size_t tiles = (n + TILE - 1) / TILE;      // scan / prepare tiles
	const size_t etiles = (n + ETILE - 1) / ETILE;   // emit tiles
	int rc = ensure_scratch(c, L, tiles);
	if (rc)
		return rc;
	if (piped) {
		// size every active lane on the first batch of a new size: the allocations (which synchronise)
		// happen once, up front, not one lane at a time over the next `depth` batches
		for (int l = 0; l < c->depth; l++)
			if ((rc = ensure_scratch(c, c->lane[l], tiles)) || (rc = lane_streams(c, c->lane[l])))
				return rc;
	}
	L.epoch++;
	const bool flags = d_flags != nullptr;

	ScanParams sp;
	sp.start = d_start;
	sp.now = d_now;
	sp.flags = d_flags;
	sp.n = n;
	sp.b = L.ctl;
	sp.target = c->target_ms;
	sp.rec_fc = L.rec_fc;
	sp.rec_code = L.rec_code;
	sp.rec_cend = L.rec_cend;
	sp.tile_now0 = L.tile_now0;
	sp.ev_count = L.ev_count;
	sp.ev = L.ev;
	sp.vec_ok = (((uintptr_t)d_start | (uintptr_t)d_now) & 15) == 0;
	sp.target_ns = (long long)(c->target_ms * 1000000.0);

	ResolveParams rp;
	rp.start = d_start;
	rp.now = d_now;
	rp.flags = d_flags;
	rp.n = n;
	rp.tiles = tiles;
	rp.g = c->g;
	rp.bc = L.ctl;
	rp.rec_fc = L.rec_fc;
	rp.rec_code = L.rec_code;
	rp.rec_cend = L.rec_cend;
	rp.tile_now0 = L.tile_now0;
	rp.idx_ms = L.idx_ms;
	{
		ChainLink &ln = rp.link;
		ln.err = &c->d_chain->err;
		ln.wait_word = &c->d_chain->seq;
		if (c->ring_n > 0) {
			ln.wait_val = c->chain_seq * (unsigned long long)c->ring_n + (unsigned long long)c->ring_rank;
			ln.in_state = &c->d_chain->st;
			ln.post_word = &c->ring_next->seq;
			ln.post_state = &c->ring_next->st;
			ln.timeout_ns = 30000000000ull; // the other ranks are other processes: allow for start-up skew
		} else {
			ln.wait_val = c->chain_seq;
			ln.in_state = nullptr;
			ln.post_word = &c->d_chain->seq;
			ln.post_state = nullptr;
			ln.timeout_ns = 4000000000ull;
		}
		ln.post_val = ln.wait_val + 1;
	}
	rp.xev = L.xev;
	rp.xcnt = L.xcnt;
	rp.hj_lo = L.hj_lo;
	rp.hj_b = L.hj_b;
	rp.hj_F = L.hj_F;
	rp.hj_info = L.hj_info;
	rp.ev_count = L.ev_count;
	rp.ev = L.ev;
	rp.evc = L.evc;
	rp.ev_now = L.ev_now;
	rp.ev_q = L.ev_q;
	rp.ev_d = L.ev_d;
	rp.ev_cap = (tiles + 1) * (EVK + 3) + EV_FRONT;
	rp.iv_lo = L.iv_lo;
	rp.iv_hi = L.iv_hi;
	rp.iv_cum = L.iv_cum;
	rp.decision = d_decision;
	rp.slot = d_slot;
	rp.backend = d_backend;
	rp.slot_backend = c->d_slot_backend;
	rp.nslots = c->nslots;
	const bool use_f = c->fchain;
	rp.f_in = use_f ? c->d_chain->fs : nullptr;
	rp.f_out = use_f ? (c->ring_n > 0 ? c->ring_next->fs : c->d_chain->fs) : nullptr;
	rp.f_ring = c->ring_n > 0 ? (unsigned long long)c->ring_n : 1ull;

	EmitParams ep;
	ep.flags = d_flags;
	ep.n = n;
	ep.tiles = etiles;
	ep.g = c->g;
	ep.bc = L.ctl;
	ep.iv_lo = L.iv_lo;
	ep.iv_hi = L.iv_hi;
	ep.iv_cum = L.iv_cum;
	ep.decision = d_decision;
	ep.slot = d_slot;
	ep.backend = d_backend;
	ep.slot_backend = c->d_slot_backend;
	ep.nslots = c->nslots;
	ep.slot_magic = c->nslots > 0 ? (~0ull) / (unsigned long long)c->nslots : 0ull;
	size_t tbytes = ((size_t)c->nslots * 4 + 15) & ~(size_t)15;
	ep.table_smem_bytes = (d_backend && c->nslots > 0 && tbytes <= 48 * 1024) ? (int)tbytes : 0;
	ep.sdesc = L.sdesc;
	ep.epoch = L.epoch;
	ep.ticket = L.ticket;
	ep.link = rp.link;
	ep.vec_ok = (((uintptr_t)d_decision & 7) == 0) && (((uintptr_t)d_slot & 15) == 0) &&
	    (((uintptr_t)d_backend & 15) == 0);
	// persistent CTAs stride over the tiles; with flags the admit-count look-back needs every
	// earlier tile to be resident or earlier in the same CTA's sequence: both hold for a grid
	// within the resident capacity (a CTA visits its tiles in increasing order)
	const unsigned egrid = (unsigned)(etiles < (size_t)c->emit_grid ? etiles : (size_t)c->emit_grid);

	// One stream per lane: scan -> prepare -> commit -> emit of a batch are stream-ordered, batches
	// of different lanes overlap freely.  The only cross-lane edge is the ControlledDelay state:
	// commit(k) follows commit(k-1) (emit(k-1) when that batch had waiting-flags, which finishes
	// the slot cursor in its emit kernel); that edge is a device-side chain word, not an event.
	cudaStream_t s = piped ? L.st : c->stream;
	if (piped) {
		if (c->user_stream) { // the caller's earlier work (e.g. filling the inputs) comes first
			CU(c, cudaEventRecord(c->ev_user, c->stream));
			CU(c, cudaStreamWaitEvent(s, c->ev_user, 0));
		}
		if (L.used && L.coll_pending) // the collective of the lane's previous batch owns its control block
			CU(c, cudaStreamWaitEvent(s, L.coll_done, 0));
	}
	if (in_ready)
		CU(c, cudaStreamWaitEvent(s, in_ready, 0));
	if (use_f && c->fseed) {
		// the state was set from the host: seed the slot this batch reads.  In a ring only the queue's
		// first batch (rank 0) is seeded; every other batch gets its value from its predecessor.
		if (c->ring_n == 0 || (c->ring_rank == 0 && c->chain_seq == 0))
			fchain_seed_kernel<<<1, 1, 0, s>>>(c->g, c->d_chain->fs, rp.link.wait_val, rp.f_ring);
		c->fseed = false;
	}
	if (c->prof)
		CU(c, cudaEventRecord(c->pe[0], s));
	if (piped && c->graph) {
		const size_t stride0 = (tiles + NIDX_MAX - 1) / NIDX_MAX;
		rc = graph_batch<NS>(c, L, flags, use_f, c->fused, sp, rp, ep, tiles, egrid, ((tiles + stride0 - 1) / stride0) * sizeof(double), s);
		if (rc)
			return rc;
		c->chain_seq++;
		if (c->comm) // the graph has no interior point to record: the collective starts after the emit
			CU(c, cudaEventRecord(L.res_done, s));
		c->launches += c->fused ? 3 : 4;
		CU(c, cudaGetLastError());
		if (out_done || (flags && c->comm))
			CU(c, cudaEventRecord(L.emit_done, s));
		L.used = true;
		L.coll_pending = false;
		c->last_lane = li;
		c->last_flags = flags;
		c->seq++;
		if (out_done)
			*out_done = L.emit_done;
		return CB_OK;
	}
	if (piped) {
		// the lane's stream has the highest priority (prepare / commit / emit take the first CTA slot
		// that frees up); the scan, which fills every SM, is launched into it at the lowest
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3((unsigned)tiles);
		cfg.blockDim = dim3(TPB);
		cfg.stream = s;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributePriority;
		at[0].val.priority = c->prio_least;
		cfg.attrs = at;
		cfg.numAttrs = 1;
		if (flags)
			CU(c, cudaLaunchKernelEx(&cfg, scan_kernel<NS, true>, sp));
		else
			CU(c, cudaLaunchKernelEx(&cfg, scan_kernel<NS, false>, sp));
	} else if (flags)
		scan_kernel<NS, true><<<(unsigned)tiles, TPB, 0, s>>>(sp);
	else
		scan_kernel<NS, false><<<(unsigned)tiles, TPB, 0, s>>>(sp);
	if (c->prof)
		CU(c, cudaEventRecord(c->pe[1], s));
	const bool fused = c->fused; // CB_FUSED=1: prepare + commit in one launch (not yet validated on hardware)
	if (fused) {
		if (use_f) {
			if (flags)
				resolve_fused_kernel<NS, true, true><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
			else
				resolve_fused_kernel<NS, false, true><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
		} else if (flags)
			resolve_fused_kernel<NS, true, false><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
		else
			resolve_fused_kernel<NS, false, false><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
	}
	// prepare: independent of the ControlledDelay state, so it overlaps the other lanes' commits
	else if (flags)
		prepare_kernel<NS, true><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
	else
		prepare_kernel<NS, false><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
	// commit: the only stage that is serial across batches.  The order is kept inside the kernels:
	// commit(k) spins (one thread) on the chain word until batch k-1 has published its state, so
	// the hand-over costs an L2 round trip instead of a cross-stream event (~5 us measured).
	const size_t cidx_bytes = ((tiles + ((tiles + NIDX_MAX - 1) / NIDX_MAX) - 1) / ((tiles + NIDX_MAX - 1) / NIDX_MAX)) * sizeof(double);
	if (fused) {
		// the commit ran inside resolve_fused_kernel
	} else if (use_f) {
		if (flags)
			commit_fchain_kernel<NS, true><<<1, CTPB, cidx_bytes, s>>>(rp);
		else
			commit_fchain_kernel<NS, false><<<1, CTPB, cidx_bytes, s>>>(rp);
	} else if (flags)
		commit_kernel<NS, true><<<1, CTPB, cidx_bytes, s>>>(rp);
	else
		commit_kernel<NS, false><<<1, CTPB, cidx_bytes, s>>>(rp);
	c->chain_seq++;
	if (c->prof)
		CU(c, cudaEventRecord(c->pe[2], s));
	if (piped && c->comm) // the drop-count collective starts from here
		CU(c, cudaEventRecord(L.res_done, s));
	if (flags)
		emit_kernel<true><<<egrid, ETPB, ep.table_smem_bytes, s>>>(ep);
	else
		emit_kernel<false><<<egrid, ETPB, ep.table_smem_bytes, s>>>(ep);
	c->launches += fused ? 3 : 4;
	CU(c, cudaGetLastError());
	if (out_done || (piped && flags && c->comm))
		CU(c, cudaEventRecord(L.emit_done, s));
	L.used = true;
	L.coll_pending = false;
	c->last_lane = li;
	c->last_flags = flags;
	c->seq++;
	if (out_done)
		*out_done = L.emit_done;
	if (c->prof) {
		CU(c, cudaEventRecord(c->pe[3], s));
		CU(c, cudaEventSynchronize(c->pe[3]));
		for (int k = 0; k < CB_K_COUNT; k++) {
			float ms = 0;
			CU(c, cudaEventElapsedTime(&ms, c->pe[k], c->pe[k + 1]));
			c->prof_ms[k] += ms;
			c->prof_n[k]++;
		}
	}
	return CB_OK;
}

int fetch_stats(cb_ctx *c, cb_batch_stats *out) {
	int rc = drain(c);
	if (rc)
		return rc;
	DevState h;
	BatchCtl b;
	memset(&b, 0, sizeof(b));
	unsigned long long chain[2] = {0, 0};
	CU(c, cudaMemcpy(chain, c->d_chain, sizeof(chain), cudaMemcpyDeviceToHost));
	if (chain[1] != 0 || (c->ring_n == 0 && chain[0] != c->chain_seq))
		return fail(c, CB_ESTATE, "commit chain broken: %llu of %llu batches committed, %llu hand-over timeouts", chain[0],
		    c->chain_seq, chain[1]);
	CU(c, cudaMemcpy(&h, c->g, sizeof(h), cudaMemcpyDeviceToHost));
	if (c->last_lane >= 0)
		CU(c, cudaMemcpy(&b, c->lane[c->last_lane].ctl, sizeof(b), cudaMemcpyDeviceToHost));
	c->tot_valid = true;
	c->tot_seq = c->chain_seq;
	c->tot_epoch = c->epoch;
	c->tot_drops_h = h.tot_drops;
	c->tot_admits_h = h.tot_admits;
	memset(out, 0, sizeof(*out));
	out->n = b.n;
	out->drops = b.drops;
	out->admits = b.admits;
	out->skipped = b.skipped;
	out->cursor = h.cursor;
	out->total_drops = h.tot_drops;
	out->total_admits = h.tot_admits;
	out->global_drops = c->comm ? b.gcounts[0] : b.drops;
	out->global_admits = c->comm ? b.gcounts[1] : b.admits;
	out->global_n = c->comm ? b.gcounts[2] : b.n;
	out->state.target_ms = h.target;
	out->state.first_above_time = h.F;
	out->state.drop_next = h.D;
	out->state.count = h.C;
	out->state.dropping = h.B;
	out->state.has_last_empty = h.hasE;
	out->state.last_empty = h.E;
	out->path = b.path;
	out->n_events = b.n_events;
	out->n_intervals = b.n_intervals;
	return CB_OK;
}

int ensure_staging(cb_ctx *c, size_t n) {
	size_t need = n < c->chunk ? n : c->chunk;
	need = (need + ETILE - 1) / ETILE * ETILE;
	if (need <= c->stage_cap)
		return CB_OK;
	int rc = drain(c);
	if (rc)
		return rc;
	for (int b = 0; b < NBUF; b++) {
		Staging &s = c->st[b];
		if ((rc = dev_alloc(c, (unsigned long long **)&s.start, need)) ||
		    (rc = dev_alloc(c, (unsigned long long **)&s.now, need)) || (rc = dev_alloc(c, &s.flags, need)) ||
		    (rc = dev_alloc(c, &s.decision, need)) || (rc = dev_alloc(c, &s.slot, need)) ||
		    (rc = dev_alloc(c, &s.backend, need)))
			return rc;
		if (!s.in_done) {
			CU(c, cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
			CU(c, cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
		}
	}
	c->stage_cap = need;
	return CB_OK;
}

// host-buffer batch: chunks flow H2D -> kernels -> D2H; the CoDel state and the admit cursor
// stay on the device between chunks (a chunk is just a batch of the pipeline).
template <bool NS>
int host_batch(cb_ctx *c, const void *start, const void *now, const uint8_t *flags, size_t n,
    uint8_t *decision, int32_t *slot, int32_t *backend, cb_batch_stats *stats) {
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	if (n && (!start || !now))
		return fail(c, CB_EINVAL, "start/now must not be NULL");
	static const bool dbg_t = getenv("CB_DEBUG_E2E") != nullptr;
	auto tnow = []() {
		struct timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
	};
	const double T0 = dbg_t ? tnow() : 0;
	int rc;
	DevState h0;
	if (c->tot_valid && c->tot_seq == c->chain_seq && c->tot_epoch == c->epoch) {
		// nothing ran since the totals were last read (which drained the pipeline): no round trip
		h0.tot_drops = c->tot_drops_h;
		h0.tot_admits = c->tot_admits_h;
	} else {
		rc = drain(c);
		if (rc)
			return rc;
		CU(c, cudaMemcpy(&h0, c->g, sizeof(h0), cudaMemcpyDeviceToHost));
	}
	const double T1 = dbg_t ? tnow() : 0;
	rc = ensure_staging(c, n);
	if (rc)
		return rc;
	// Staging chunk: a batch that would fit one chunk is still cut in ~3 so that the copy-in of
	// one piece overlaps the kernels and the copy-out of the previous ones (measured with
	// tools/e2e_breakdown.py: 1M claims 1.04 ms as one chunk, 0.50-0.54 ms as 2-4 chunks).
	size_t chunk = c->chunk;
	{
		size_t third = ((n + 2) / 3 + ETILE - 1) / ETILE * ETILE;
		const size_t floor_chunk = 64 * (size_t)ETILE; // 256K claims: below this the per-chunk cost wins
		if (third < floor_chunk)
			third = floor_chunk;
		if (third < chunk)
			chunk = third;
	}
	const size_t nchunks = (n + chunk - 1) / chunk;
	for (size_t ci = 0; ci < nchunks; ci++) {
		Staging &s = c->st[ci % NBUF];
		const size_t lo = ci * chunk, m = (n - lo < chunk) ? n - lo : chunk;
		if (ci >= NBUF)
			CU(c, cudaStreamWaitEvent(c->s_in, s.out_done, 0));
		CU(c, cudaMemcpyAsync(s.start, (const char *)start + lo * 8, m * 8, cudaMemcpyHostToDevice, c->s_in));
		CU(c, cudaMemcpyAsync(s.now, (const char *)now + lo * 8, m * 8, cudaMemcpyHostToDevice, c->s_in));
		if (flags)
			CU(c, cudaMemcpyAsync(s.flags, flags + lo, m, cudaMemcpyHostToDevice, c->s_in));
		CU(c, cudaEventRecord(s.in_done, c->s_in));
		cudaEvent_t done = nullptr;
		rc = enqueue_batch<NS>(c, s.start, s.now, flags ? s.flags : nullptr, m, decision ? s.decision : nullptr,
		    slot ? s.slot : nullptr, backend ? s.backend : nullptr, s.in_done, &done);
		if (rc)
			return rc;
		if (done)
			CU(c, cudaStreamWaitEvent(c->s_out, done, 0));
		if (decision)
			CU(c, cudaMemcpyAsync(decision + lo, s.decision, m, cudaMemcpyDeviceToHost, c->s_out));
		if (slot)
			CU(c, cudaMemcpyAsync(slot + lo, s.slot, m * 4, cudaMemcpyDeviceToHost, c->s_out));
		if (backend)
			CU(c, cudaMemcpyAsync(backend + lo, s.backend, m * 4, cudaMemcpyDeviceToHost, c->s_out));
		CU(c, cudaEventRecord(s.out_done, c->s_out));
	}
	const double T2 = dbg_t ? tnow() : 0;
	CU(c, cudaStreamSynchronize(c->s_out));
	const double T3 = dbg_t ? tnow() : 0;
	if (dbg_t)
		fprintf(stderr, "[cb e2e] n=%zu chunks=%zu  entry sync+state %.0f us | enqueue %.0f us | wait %.0f us\n", n, nchunks,
		    T1 - T0, T2 - T1, T3 - T2);
	if (stats) {
		rc = fetch_stats(c, stats);
		if (rc)
			return rc;
		// totals of the whole call, not of its last chunk
		stats->n = n;
		stats->drops = stats->total_drops - h0.tot_drops;
		stats->admits = stats->total_admits - h0.tot_admits;
		stats->skipped = n - stats->drops - stats->admits;
		if (!c->comm) {
			stats->global_drops = stats->drops;
			stats->global_admits = stats->admits;
			stats->global_n = n;
		}
	} else {
		rc = drain(c);
		if (rc)
			return rc;
	}
	return CB_OK;
}

int rebuild_slots(cb_ctx *c) {
	// planRebalance({}, {}, nslots, nslots) over the preference list, then the slot table
	if (c->B <= 0 || c->target_slots <= 0) {
		c->nslots = 0;
		return CB_OK;
	}
	c->nslots = c->target_slots;
	cudaStream_t s = c->stream;
	int *d_nadd = c->d_ints + 1, *d_nslots = c->d_ints + 2;
	if (c->mode == CB_BALANCE_SRV_WEIGHTED) {
		weighted_want_kernel<<<1, 1, 0, s>>>(c->d_keys, c->B, c->d_prio, c->d_weight, c->nslots, c->d_want,
		    c->d_add, d_nadd);
	} else {
		// have = 0, dead = 0: reuse d_want's neighbour buffers as zeroed inputs
		int32_t *d_have = c->d_add + c->slot_cap; // zeroed tail, see cb_backends_set
		uint8_t *d_dead = (uint8_t *)(d_have + c->key_cap);
		plan_rebalance_kernel<<<1, 1, 0, s>>>(d_have, d_dead, c->B, c->nslots, c->nslots, 0, c->d_want, c->d_add,
		    d_nadd, nullptr, nullptr, c->d_ints + 4);
	}
	slot_table_kernel<<<8, 256, 0, s>>>(c->d_keys, c->d_add, d_nadd, c->d_slot_backend, c->slot_cap, d_nslots);
	c->launches += 2;
	CU(c, cudaGetLastError());
	int got = 0;
	CU(c, cudaMemcpyAsync(&got, d_nslots, sizeof(int), cudaMemcpyDeviceToHost, s));
	CU(c, cudaStreamSynchronize(s));
	c->nslots = got; // planRebalance hands out exactly `target` slots when B > 0
	return CB_OK;
}

int ensure_topology(cb_ctx *c, int keys, int slots) {
	if (keys <= c->key_cap && slots <= c->slot_cap)
		return CB_OK;
	int kc = keys > c->key_cap ? keys + 64 : c->key_cap;
	int sc = slots > c->slot_cap ? ((slots + 3) & ~3) + 64 : c->slot_cap;
	int32_t *ok = nullptr, *opr = nullptr, *ow = nullptr;
	int oldB = c->B;
	// keep the current preference list / SRV table across a growth
	if (c->d_keys && oldB > 0) {
		ok = (int32_t *)malloc(sizeof(int32_t) * oldB * 3);
		opr = ok + oldB;
		ow = opr + oldB;
		cudaMemcpy(ok, c->d_keys, sizeof(int32_t) * oldB, cudaMemcpyDeviceToHost);
		cudaMemcpy(opr, c->d_prio, sizeof(int32_t) * oldB, cudaMemcpyDeviceToHost);
		cudaMemcpy(ow, c->d_weight, sizeof(int32_t) * oldB, cudaMemcpyDeviceToHost);
	}
	int rc;
	if ((rc = dev_alloc(c, &c->d_keys, kc)) || (rc = dev_alloc(c, &c->d_want, kc)) ||
	    (rc = dev_alloc(c, &c->d_prio, kc)) || (rc = dev_alloc(c, &c->d_weight, kc)) ||
	    (rc = dev_alloc(c, &c->d_add, (size_t)sc + kc + kc)) || (rc = dev_alloc(c, &c->d_slot_backend, sc)))
		return rc;
	CU(c, cudaMemset(c->d_add, 0, ((size_t)sc + kc + kc) * sizeof(int32_t)));
	CU(c, cudaMemset(c->d_prio, 0, sizeof(int32_t) * kc));
	CU(c, cudaMemset(c->d_weight, 0, sizeof(int32_t) * kc));
	if (ok) {
		cudaMemcpy(c->d_keys, ok, sizeof(int32_t) * oldB, cudaMemcpyHostToDevice);
		cudaMemcpy(c->d_prio, opr, sizeof(int32_t) * oldB, cudaMemcpyHostToDevice);
		cudaMemcpy(c->d_weight, ow, sizeof(int32_t) * oldB, cudaMemcpyHostToDevice);
		free(ok);
	}
	c->key_cap = kc;
	c->slot_cap = sc;
	return CB_OK;
}

} // namespace

// =========================================================================== lifecycle
extern "C" int cb_abi_version(void) { return CB_ABI_VERSION; }

extern "C" const char *cb_last_error(const cb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int cb_create(cb_ctx **out, int device) {
	if (!out)
		return fail(nullptr, CB_EINVAL, "out is NULL");
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(nullptr, CB_ECUDA, "no CUDA device (%s): this library has no CPU fallback",
		    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
	if (device < 0 || device >= ndev)
		return fail(nullptr, CB_EINVAL, "device %d out of range (have %d)", device, ndev);
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess)
		return fail(nullptr, CB_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
	if (prop.major != 10)
		return fail(nullptr, CB_ECUDA, "device %d is sm_%d%d; this build is sm_100a only", device, prop.major,
		    prop.minor);
	cb_ctx *c = new (std::nothrow) cb_ctx();
	if (!c)
		return fail(nullptr, CB_ENOMEM, "out of host memory");
	c->device = device;
	c->emit_grid = prop.multiProcessorCount * 4;
	const char *ck = getenv("CB_CHUNK");
	if (ck && atoll(ck) >= ETILE)
		c->chunk = ((size_t)atoll(ck) + ETILE - 1) / ETILE * ETILE;
	const char *dp = getenv("CB_PIPELINE_DEPTH");
	if (dp && atoi(dp) >= 1 && atoi(dp) <= CB_LANES)
		c->depth = atoi(dp);
#define CU0(call)                                                                          \
	do {                                                                                   \
		cudaError_t e_ = (call);                                                           \
		if (e_ != cudaSuccess) {                                                           \
			fail(nullptr, CB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));              \
			cb_destroy(c);                                                                 \
			return CB_ECUDA;                                                               \
		}                                                                                  \
	} while (0)
	CU0(cudaSetDevice(device));
	CU0(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	{
		// the drop-count collective is tiny and latency-bound: highest priority, or its kernel queues
		// behind every scan CTA of the batches in flight
		int least = 0, greatest = 0;
		CU0(cudaDeviceGetStreamPriorityRange(&least, &greatest));
		CU0(cudaStreamCreateWithPriority(&c->st_coll, cudaStreamNonBlocking, greatest));
	}
	CU0(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
	CU0(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
	CU0(cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming));
	c->stream = c->own_stream;
	CU0(cudaMalloc((void **)&c->g, sizeof(DevState)));
	CU0(cudaMemset(c->g, 0, sizeof(DevState)));
	if (getenv("CB_FCHAIN") && atoi(getenv("CB_FCHAIN")) != 0) {
		c->fchain = true;
		CU0(cudaFuncSetAttribute(commit_fchain_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
		CU0(cudaFuncSetAttribute(commit_fchain_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
		CU0(cudaFuncSetAttribute(commit_fchain_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
		CU0(cudaFuncSetAttribute(commit_fchain_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	}
	c->graph = getenv("CB_GRAPH") && atoi(getenv("CB_GRAPH")) != 0;
	if (getenv("CB_FUSED") && atoi(getenv("CB_FUSED")) != 0) {
		c->fused = true;
#define FUSED_ATTR(a, b, f)                                                                                 \
	CU0(cudaFuncSetAttribute(resolve_fused_kernel<a, b, f>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8))
		FUSED_ATTR(true, true, true);
		FUSED_ATTR(true, true, false);
		FUSED_ATTR(true, false, true);
		FUSED_ATTR(true, false, false);
		FUSED_ATTR(false, true, true);
		FUSED_ATTR(false, true, false);
		FUSED_ATTR(false, false, true);
		FUSED_ATTR(false, false, false);
#undef FUSED_ATTR
	}
	CU0(cudaMalloc((void **)&c->d_chain, sizeof(ChainMem)));
	CU0(cudaMemset(c->d_chain, 0, sizeof(ChainMem)));
	for (int l = 0; l < CB_LANES; l++) {
		Lane &L = c->lane[l];
		CU0(cudaMalloc((void **)&L.ctl, sizeof(BatchCtl)));
		CU0(cudaMemset(L.ctl, 0, sizeof(BatchCtl)));
		CU0(cudaMalloc((void **)&L.ticket, 64));
		CU0(cudaMemset(L.ticket, 0, 64));
		CU0(cudaEventCreateWithFlags(&L.res_done, cudaEventDisableTiming));
		CU0(cudaEventCreateWithFlags(&L.join_ev, cudaEventDisableTiming));
		CU0(cudaEventCreateWithFlags(&L.emit_done, cudaEventDisableTiming));
		CU0(cudaEventCreateWithFlags(&L.coll_done, cudaEventDisableTiming));
	}
	CU0(cudaMalloc((void **)&c->d_ints, 64 * sizeof(int)));
	CU0(cudaMemset(c->d_ints, 0, 64 * sizeof(int)));
	CU0(cudaMalloc((void **)&c->d_dbl, 64 * sizeof(double)));
	for (int k = 0; k < 4; k++)
		CU0(cudaEventCreate(&c->pe[k]));
	CU0(cudaFuncSetAttribute(prepare_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(prepare_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(prepare_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(prepare_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(commit_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(commit_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(commit_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(commit_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
	CU0(cudaFuncSetAttribute(emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
#undef CU0
	*out = c;
	return CB_OK;
}

extern "C" void cb_destroy(cb_ctx *c) {
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaStream_t all[] = {c->st_coll, c->s_in, c->s_out, c->own_stream};
	for (cudaStream_t s : all)
		if (s)
			cudaStreamSynchronize(s);
	if (c->comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy(c->comm);
	for (int l = 0; l < CB_LANES; l++) {
		Lane &L = c->lane[l];
		if (L.st) {
			cudaStreamSynchronize(L.st);
			cudaStreamDestroy(L.st);
		}
		for (int a = 0; a < 2; a++)
			for (int b = 0; b < 2; b++) {
				if (L.gr[a][b].ex)
					cudaGraphExecDestroy(L.gr[a][b].ex);
				if (L.gr[a][b].g)
					cudaGraphDestroy(L.gr[a][b].g);
			}
		void *fr[] = {L.idx_ms, L.ctl, L.sdesc, L.ev_count, L.rec_fc, L.rec_cend, L.xcnt, L.hj_info, L.ticket, L.ev, L.evc,
		    L.ev_q, L.ev_d, L.rec_code, L.xev, L.hj_lo, L.hj_b, L.tile_now0, L.iv_lo, L.iv_hi, L.iv_cum, L.ev_now, L.hj_F};
		for (void *p : fr)
			if (p)
				cudaFree(p);
		cudaEvent_t evs[] = {L.res_done, L.emit_done, L.coll_done, L.join_ev};
		for (cudaEvent_t e : evs)
			if (e)
				cudaEventDestroy(e);
	}
	ring_close(c);
	void *qfrees[] = {c->d_q, c->q_start, c->q_deadline, c->q_ticket, c->q_res, c->q_stage, c->q_waiting, c->q_tmo};
	for (void *p : qfrees)
		if (p)
			cudaFree(p);
	void *frees[] = {c->g, c->d_chain, c->d_keys, c->d_want, c->d_add, c->d_slot_backend, c->d_prio, c->d_weight, c->d_ints,
	    c->d_dbl};
	for (void *p : frees)
		if (p)
			cudaFree(p);
	for (int b = 0; b < NBUF; b++) {
		Staging &s = c->st[b];
		void *fs[] = {s.start, s.now, s.flags, s.decision, s.slot, s.backend};
		for (void *p : fs)
			if (p)
				cudaFree(p);
		if (s.in_done)
			cudaEventDestroy(s.in_done);
		if (s.out_done)
			cudaEventDestroy(s.out_done);
	}
	for (int k = 0; k < 4; k++)
		if (c->pe[k])
			cudaEventDestroy(c->pe[k]);
	if (c->ev_user)
		cudaEventDestroy(c->ev_user);
	for (cudaStream_t s : all)
		if (s)
			cudaStreamDestroy(s);
	delete c;
}

extern "C" int cb_set_stream(cb_ctx *c, void *stream) {
	if (!c)
		return CB_EINVAL;
	int rc = drain(c);
	if (rc)
		return rc;
	c->stream = stream ? (cudaStream_t)stream : c->own_stream;
	c->user_stream = stream != nullptr;
	return CB_OK;
}

extern "C" int cb_set_pipeline_depth(cb_ctx *c, int depth) {
	if (!c || depth < 1 || depth > CB_LANES)
		return c ? fail(c, CB_EINVAL, "pipeline depth must be 1..%d", CB_LANES) : CB_EINVAL;
	int rc = drain(c);
	if (rc)
		return rc;
	c->depth = depth;
	return CB_OK;
}

// make the main (caller's) stream wait for every batch enqueued so far
extern "C" int cb_join(cb_ctx *c) {
	if (!c)
		return CB_EINVAL;
	if (c->last_lane >= 0 && c->depth > 1) {
		for (int l = 0; l < CB_LANES; l++) {
			Lane &L = c->lane[l];
			if (!L.used || !L.st)
				continue;
			CU(c, cudaEventRecord(L.join_ev, L.st));
			CU(c, cudaStreamWaitEvent(c->stream, L.join_ev, 0));
			if (L.coll_pending)
				CU(c, cudaStreamWaitEvent(c->stream, L.coll_done, 0));
		}
	}
	return CB_OK;
}

extern "C" int cb_synchronize(cb_ctx *c) {
	if (!c)
		return CB_EINVAL;
	return drain(c);
}

extern "C" void *cb_host_alloc(cb_ctx *c, size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		if (c)
			fail(c, CB_ENOMEM, "cudaMallocHost(%zu) failed", bytes);
		return nullptr;
	}
	return p;
}
extern "C" void cb_host_free(cb_ctx *, void *p) {
	if (p)
		cudaFreeHost(p);
}

// =============================================================================== CoDel
extern "C" int cb_codel_init(cb_ctx *c, double target_ms) {
	if (!c)
		return CB_EINVAL;
	c->tot_valid = false;
	c->fseed = true;
	// utils.assertClaimDelay, lib/utils.js:188-195: finite, > 0, integral
	if (!isfinite(target_ms) || !(target_ms > 0) || target_ms != floor(target_ms))
		return fail(c, CB_EINVAL, "targetClaimDelay must be a positive integer number of ms");
	int rc = drain(c);
	if (rc)
		return rc;
	DevState h;
	memset(&h, 0, sizeof(h));
	h.target = target_ms;
	c->target_ms = target_ms;
	CU(c, cudaMemcpy(c->g, &h, sizeof(h), cudaMemcpyHostToDevice));
	c->codel_ready = true;
	return CB_OK;
}

extern "C" int cb_codel_get_state(cb_ctx *c, cb_codel_state *out) {
	if (!c || !out)
		return CB_EINVAL;
	cb_batch_stats st;
	int rc = fetch_stats(c, &st);
	if (rc)
		return rc;
	*out = st.state;
	return CB_OK;
}

extern "C" int cb_codel_set_state(cb_ctx *c, const cb_codel_state *in) {
	if (!c || !in)
		return CB_EINVAL;
	c->tot_valid = false;
	c->fseed = true;
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	int rc = drain(c);
	if (rc)
		return rc;
	DevState h;
	CU(c, cudaMemcpy(&h, c->g, sizeof(h), cudaMemcpyDeviceToHost));
	h.F = in->first_above_time;
	h.D = in->drop_next;
	h.C = in->count;
	h.B = in->dropping ? 1 : 0;
	h.hasE = in->has_last_empty ? 1 : 0;
	h.E = in->last_empty;
	CU(c, cudaMemcpy(c->g, &h, sizeof(h), cudaMemcpyHostToDevice));
	return CB_OK;
}

extern "C" int cb_codel_empty(cb_ctx *c, int64_t now_ns) {
	if (!c)
		return CB_EINVAL;
	c->tot_valid = false;
	c->fseed = true;
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	int rc = drain(c);
	if (rc)
		return rc;
	empty_kernel<<<1, 1, 0, c->stream>>>(c->g, now_ns);
	c->launches++;
	CU(c, cudaGetLastError());
	CU(c, cudaStreamSynchronize(c->stream));
	return CB_OK;
}

extern "C" int cb_codel_max_idle(cb_ctx *c, int64_t now_ns, double *out_ms) {
	if (!c || !out_ms)
		return CB_EINVAL;
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	int rc = drain(c);
	if (rc)
		return rc;
	max_idle_kernel<<<1, 1, 0, c->stream>>>(c->g, now_ns, c->d_dbl);
	c->launches++;
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(out_ms, c->d_dbl, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return CB_OK;
}

extern "C" int cb_codel_overloaded(cb_ctx *c, int64_t start_ns, int64_t now_ns, int *drop) {
	if (!c || !drop)
		return CB_EINVAL;
	c->tot_valid = false;
	c->fseed = true;
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	int rc = drain(c);
	if (rc)
		return rc;
	overloaded_one_kernel<<<1, 1, 0, c->stream>>>(c->g, start_ns, now_ns, c->d_ints + 5);
	c->launches++;
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(drop, c->d_ints + 5, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return CB_OK;
}

// =============================================================================== batches
extern "C" int cb_admit_batch(cb_ctx *c, const int64_t *start_ns, const int64_t *now_ns, const uint8_t *flags,
    size_t n, uint8_t *decision, int32_t *slot, int32_t *backend, cb_batch_stats *stats) {
	if (!c)
		return CB_EINVAL;
	return host_batch<true>(c, start_ns, now_ns, flags, n, decision, slot, backend, stats);
}
extern "C" int cb_admit_batch_ms(cb_ctx *c, const double *start_ms, const double *now_ms, const uint8_t *flags,
    size_t n, uint8_t *decision, int32_t *slot, int32_t *backend, cb_batch_stats *stats) {
	if (!c)
		return CB_EINVAL;
	return host_batch<false>(c, start_ms, now_ms, flags, n, decision, slot, backend, stats);
}
extern "C" int cb_admit_batch_dev(cb_ctx *c, const int64_t *d_start, const int64_t *d_now, const uint8_t *d_flags,
    size_t n, uint8_t *d_decision, int32_t *d_slot, int32_t *d_backend) {
	if (!c)
		return CB_EINVAL;
	return enqueue_batch<true>(c, d_start, d_now, d_flags, n, d_decision, d_slot, d_backend, nullptr, nullptr);
}
extern "C" int cb_admit_batch_ms_dev(cb_ctx *c, const double *d_start, const double *d_now, const uint8_t *d_flags,
    size_t n, uint8_t *d_decision, int32_t *d_slot, int32_t *d_backend) {
	if (!c)
		return CB_EINVAL;
	return enqueue_batch<false>(c, d_start, d_now, d_flags, n, d_decision, d_slot, d_backend, nullptr, nullptr);
}
extern "C" int cb_batch_stats_get(cb_ctx *c, cb_batch_stats *out) {
	if (!c || !out)
		return CB_EINVAL;
	return fetch_stats(c, out);
}

// ============================================================================== topology
extern "C" int cb_backends_set(cb_ctx *c, const cb_backend *tbl, int32_t B, const double *insert_u, int32_t nslots,
    int mode) {
	if (!c || B < 0 || nslots < 0 || (B > 0 && !insert_u))
		return c ? fail(c, CB_EINVAL, "bad backend table arguments") : CB_EINVAL;
	if (mode != CB_BALANCE_UNIFORM && mode != CB_BALANCE_SRV_WEIGHTED)
		return fail(c, CB_EINVAL, "unknown balance mode %d", mode);
	if (mode == CB_BALANCE_SRV_WEIGHTED && !tbl)
		return fail(c, CB_EINVAL, "weighted mode needs the SRV table");
	for (int k = 0; k < B; k++)
		if (!(insert_u[k] >= 0.0 && insert_u[k] < 1.0))
			return fail(c, CB_EINVAL, "insert_u[%d] not in [0,1)", k);
	{
		int rc0 = drain(c); // the emit kernels of batches in flight read the slot table
		if (rc0)
			return rc0;
	}
	c->B = 0; // the old list is replaced
	int rc = ensure_topology(c, B, nslots);
	if (rc)
		return rc;
	c->B = B;
	c->target_slots = nslots;
	c->nslots = 0;
	c->mode = mode;
	if (B == 0) {
		int z[3] = {0, 0, 0};
		CU(c, cudaMemcpy(c->d_ints, z, sizeof(z), cudaMemcpyHostToDevice));
		return CB_OK;
	}
	double *d_us = nullptr;
	CU(c, cudaMalloc((void **)&d_us, sizeof(double) * (B > 0 ? B : 1)));
	CU(c, cudaMemcpyAsync(d_us, insert_u, sizeof(double) * B, cudaMemcpyHostToDevice, c->stream));
	if (tbl) {
		int32_t *hp = (int32_t *)malloc(sizeof(int32_t) * 2 * B);
		for (int k = 0; k < B; k++) {
			hp[k] = tbl[k].priority;
			hp[B + k] = tbl[k].weight;
		}
		cudaMemcpyAsync(c->d_prio, hp, sizeof(int32_t) * B, cudaMemcpyHostToDevice, c->stream);
		cudaMemcpyAsync(c->d_weight, hp + B, sizeof(int32_t) * B, cudaMemcpyHostToDevice, c->stream);
		cudaStreamSynchronize(c->stream);
		free(hp);
	}
	pref_build_kernel<<<1, TOPO_TPB, 0, c->stream>>>(c->d_keys, B, d_us);
	c->launches++;
	CU(c, cudaMemcpyAsync(c->d_ints, &B, sizeof(int), cudaMemcpyHostToDevice, c->stream));
	CU(c, cudaGetLastError());
	rc = rebuild_slots(c);
	cudaStreamSynchronize(c->stream);
	cudaFree(d_us);
	return rc;
}

extern "C" int cb_backend_added(cb_ctx *c, int32_t key, const cb_backend *be, double u) {
	if (!c || key < 0 || !(u >= 0.0 && u < 1.0))
		return c ? fail(c, CB_EINVAL, "bad arguments") : CB_EINVAL;
	{
		int rc0 = drain(c);
		if (rc0)
			return rc0;
	}
	int need_keys = (key + 1 > c->B + 1) ? key + 1 : c->B + 1;
	int rc = ensure_topology(c, need_keys, c->target_slots);
	if (rc)
		return rc;
	if (be) {
		CU(c, cudaMemcpyAsync(c->d_prio + key, &be->priority, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
		CU(c, cudaMemcpyAsync(c->d_weight + key, &be->weight, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	}
	pref_insert_kernel<<<1, TOPO_TPB, 0, c->stream>>>(c->d_keys, c->d_ints, key, u);
	c->launches++;
	CU(c, cudaGetLastError());
	c->B++;
	return rebuild_slots(c);
}

extern "C" int cb_backend_removed(cb_ctx *c, int32_t key) {
	if (!c)
		return CB_EINVAL;
	{
		int rc0 = drain(c);
		if (rc0)
			return rc0;
	}
	pref_remove_kernel<<<1, TOPO_TPB, 0, c->stream>>>(c->d_keys, c->d_ints, key, c->d_ints + 3);
	c->launches++;
	CU(c, cudaGetLastError());
	int status = 0;
	CU(c, cudaMemcpyAsync(&status, c->d_ints + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	if (status != 0) // mod_assert.notStrictEqual(idx, -1, ...), lib/pool.js:295
		return fail(c, CB_EINVAL, "resolver key %d not found", key);
	c->B--;
	return rebuild_slots(c);
}

extern "C" int cb_reshuffle(cb_ctx *c, double u) {
	if (!c || !(u >= 0.0 && u < 1.0))
		return c ? fail(c, CB_EINVAL, "bad arguments") : CB_EINVAL;
	if (c->B <= 1)
		return CB_OK; // lib/pool.js:502-503
	{
		int rc0 = drain(c);
		if (rc0)
			return rc0;
	}
	pref_reshuffle_kernel<<<1, TOPO_TPB, 0, c->stream>>>(c->d_keys, c->d_ints, u);
	c->launches++;
	CU(c, cudaGetLastError());
	return rebuild_slots(c);
}

extern "C" int cb_backends_get(cb_ctx *c, int32_t *n_keys, int32_t *keys_out, int32_t *n_slots,
    int32_t *slot_backend_out) {
	if (!c)
		return CB_EINVAL;
	if (n_keys)
		*n_keys = c->B;
	if (n_slots)
		*n_slots = (c->B > 0) ? c->nslots : 0;
	if (keys_out && c->B > 0)
		CU(c, cudaMemcpyAsync(keys_out, c->d_keys, sizeof(int32_t) * c->B, cudaMemcpyDeviceToHost, c->stream));
	if (slot_backend_out && c->B > 0 && c->nslots > 0)
		CU(c, cudaMemcpyAsync(slot_backend_out, c->d_slot_backend, sizeof(int32_t) * c->nslots,
		    cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return CB_OK;
}

extern "C" int cb_plan_rebalance(cb_ctx *c, const int32_t *have, const uint8_t *dead, int32_t B, int32_t target,
    int32_t max, int singleton, int32_t *want_out, int32_t *add_out, int32_t *n_add, int32_t *rm_key_out,
    int32_t *rm_ord_out, int32_t *n_rm) {
	if (!c || B < 0 || (B > 0 && (!have || !dead)))
		return c ? fail(c, CB_EINVAL, "bad arguments") : CB_EINVAL;
	if (target < 0) // mod_assert.ok(target >= 0), lib/utils.js:248
		return fail(c, CB_EINVAL, "target must be >= 0");
	if (max < target) // mod_assert.ok(max >= target), lib/utils.js:249
		return fail(c, CB_EINVAL, "max must be >= target");
	long long total_have = 0;
	for (int k = 0; k < B; k++) {
		if (have[k] < 0)
			return fail(c, CB_EINVAL, "have[%d] < 0", k);
		total_have += have[k];
	}
	size_t nb = (size_t)(B > 0 ? B : 1), na = (size_t)max + 1, nr = (size_t)total_have + 1;
	int32_t *d_have = nullptr, *d_want = nullptr, *d_add = nullptr, *d_rk = nullptr, *d_ro = nullptr;
	uint8_t *d_dead = nullptr;
	int *d_cnt = nullptr;
	CU(c, cudaMalloc((void **)&d_have, nb * 4));
	CU(c, cudaMalloc((void **)&d_want, nb * 4));
	CU(c, cudaMalloc((void **)&d_dead, nb));
	CU(c, cudaMalloc((void **)&d_add, na * 4));
	CU(c, cudaMalloc((void **)&d_rk, nr * 4));
	CU(c, cudaMalloc((void **)&d_ro, nr * 4));
	CU(c, cudaMalloc((void **)&d_cnt, 8));
	cudaStream_t s = c->stream;
	if (B > 0) {
		CU(c, cudaMemcpyAsync(d_have, have, nb * 4, cudaMemcpyHostToDevice, s));
		CU(c, cudaMemcpyAsync(d_dead, dead, nb, cudaMemcpyHostToDevice, s));
	}
	plan_rebalance_kernel<<<1, 1, 0, s>>>(d_have, d_dead, B, target, max, singleton, d_want, d_add, d_cnt, d_rk,
	    d_ro, d_cnt + 1);
	c->launches++;
	CU(c, cudaGetLastError());
	int cnt[2] = {0, 0};
	CU(c, cudaMemcpyAsync(cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s));
	CU(c, cudaStreamSynchronize(s));
	if (want_out && B > 0)
		CU(c, cudaMemcpy(want_out, d_want, nb * 4, cudaMemcpyDeviceToHost));
	if (add_out && cnt[0] > 0)
		CU(c, cudaMemcpy(add_out, d_add, (size_t)cnt[0] * 4, cudaMemcpyDeviceToHost));
	if (rm_key_out && cnt[1] > 0)
		CU(c, cudaMemcpy(rm_key_out, d_rk, (size_t)cnt[1] * 4, cudaMemcpyDeviceToHost));
	if (rm_ord_out && cnt[1] > 0)
		CU(c, cudaMemcpy(rm_ord_out, d_ro, (size_t)cnt[1] * 4, cudaMemcpyDeviceToHost));
	if (n_add)
		*n_add = cnt[0];
	if (n_rm)
		*n_rm = cnt[1];
	cudaFree(d_have);
	cudaFree(d_want);
	cudaFree(d_dead);
	cudaFree(d_add);
	cudaFree(d_rk);
	cudaFree(d_ro);
	cudaFree(d_cnt);
	return CB_OK;
}

// ============================================================================= multi-GPU
extern "C" int cb_comm_unique_id(void *out, size_t *bytes) {
	if (!out || !bytes || *bytes < sizeof(nccl_uid))
		return fail(nullptr, CB_EINVAL, "need a %zu-byte buffer", sizeof(nccl_uid));
	if (!g_nccl.load())
		return fail(nullptr, CB_ENCCL, "libnccl.so.2 not found: %s", dlerror());
	nccl_uid id;
	int r = g_nccl.GetUniqueId(&id);
	if (r != 0)
		return fail(nullptr, CB_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
	memcpy(out, &id, sizeof(id));
	*bytes = sizeof(id);
	return CB_OK;
}

extern "C" int cb_comm_init(cb_ctx *c, int nranks, int rank, const void *unique_id, size_t bytes) {
	if (!c || nranks < 1 || rank < 0 || rank >= nranks || !unique_id || bytes < sizeof(nccl_uid))
		return c ? fail(c, CB_EINVAL, "bad communicator arguments") : CB_EINVAL;
	if (!g_nccl.load())
		return fail(c, CB_ENCCL, "libnccl.so.2 not found");
	CU(c, cudaSetDevice(c->device));
	// The only collective is a 24-byte sum, latency-bound: keep its kernel to one small CTA so that
	// it fits the slot one retiring scan CTA frees (a 512+-thread CTA needs a whole SM's registers and
	// waits for the scans of every batch in flight to drain).  The host application's settings win.
	setenv("NCCL_NTHREADS", "128", 0);
	setenv("NCCL_MAX_CTAS", "1", 0);
	nccl_uid id;
	memcpy(&id, unique_id, sizeof(id));
	int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
	if (r != 0) {
		c->comm = nullptr;
		return fail(c, CB_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
	}
	c->nranks = nranks;
	c->rank = rank;
	return CB_OK;
}

// ---- one logical queue over several GPUs ------------------------------------------------------
extern "C" int cb_ring_mailbox(cb_ctx *c, void *ipc_handle_out, size_t *len, void **local_ptr_out) {
	if (!c)
		return CB_EINVAL;
	if (ipc_handle_out) {
		if (!len || *len < sizeof(cudaIpcMemHandle_t))
			return fail(c, CB_EINVAL, "the IPC handle needs %zu bytes", sizeof(cudaIpcMemHandle_t));
		cudaIpcMemHandle_t h;
		CU(c, cudaIpcGetMemHandle(&h, c->d_chain));
		memcpy(ipc_handle_out, &h, sizeof(h));
		*len = sizeof(h);
	}
	if (local_ptr_out)
		*local_ptr_out = c->d_chain;
	return CB_OK;
}

extern "C" int cb_ring_connect(cb_ctx *c, int nranks, int rank, const void *next_ipc_handle, void *next_local_ptr) {
	if (!c)
		return CB_EINVAL;
	c->tot_valid = false;
	c->fseed = true;
	int rc = drain(c);
	if (rc)
		return rc;
	ring_close(c);
	if (nranks == 0) { // back to an independent queue: the chain word restarts at the batch count
		c->chain_seq = 0;
		CU(c, cudaMemset(c->d_chain, 0, sizeof(ChainMem))); // chain word, error count and the F-chain tags
		return CB_OK;
	}
	if (nranks < 1 || rank < 0 || rank >= nranks)
		return fail(c, CB_EINVAL, "bad ring geometry");
	if ((next_ipc_handle != nullptr) == (next_local_ptr != nullptr) && nranks > 1)
		return fail(c, CB_EINVAL, "give the next rank's mailbox either as an IPC handle or as a pointer of this process");
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	if (nranks == 1 && !next_ipc_handle && !next_local_ptr) {
		c->ring_next = c->d_chain;
	} else if (next_ipc_handle) {
		cudaIpcMemHandle_t h;
		memcpy(&h, next_ipc_handle, sizeof(h));
		void *ptr = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess)
			return fail(c, CB_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
		c->ring_next = (ChainMem *)ptr;
		c->ring_next_ipc = true;
	} else {
		c->ring_next = (ChainMem *)next_local_ptr;
	}
	// this rank's mailbox: sequence 0, holding this rank's current state (rank 0 starts the queue
	// from it; the others overwrite it with what they receive)
	ChainMem m;
	memset(&m, 0, sizeof(m));
	CU(c, cudaMemcpy(&m.st, c->g, sizeof(DevState), cudaMemcpyDeviceToHost));
	CU(c, cudaMemcpy(c->d_chain, &m, sizeof(m), cudaMemcpyHostToDevice));
	c->ring_n = nranks;
	c->ring_rank = rank;
	c->chain_seq = 0;
	return CB_OK;
}

extern "C" int cb_allreduce_counts(cb_ctx *c) {
	if (!c)
		return CB_EINVAL;
	if (!c->comm)
		return fail(c, CB_ESTATE, "cb_comm_init has not been called");
	if (c->last_lane < 0)
		return CB_OK;
	// {drops, admits, n} of the last batch, summed over ranks, in place in its control block.
	// The counts are final after the resolve kernel (after the emit kernel for a flags batch); the
	// collective runs on its own stream so it never stalls the next batch's kernels.
	Lane &L = c->lane[c->last_lane];
	const bool piped = (c->depth > 1) && !c->prof;
	cudaStream_t s = piped ? c->st_coll : c->stream;
	if (piped)
		CU(c, cudaStreamWaitEvent(s, c->last_flags ? L.emit_done : L.res_done, 0));
	unsigned long long *buf = (unsigned long long *)((char *)L.ctl + offsetof(BatchCtl, gcounts));
	int r = g_nccl.AllReduce(buf, buf, 3, NCCL_UINT64, NCCL_SUM, c->comm, s);
	if (r != 0)
		return fail(c, CB_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
	if (piped) {
		CU(c, cudaEventRecord(L.coll_done, s));
		L.coll_pending = true;
	}
	return CB_OK;
}

// =========================================================================== pending-claim queue
// (kernels in cb_queue.cuh; synchronous calls: one small kernel each)
static QueueParams queue_params(cb_ctx *c) {
	QueueParams p;
	p.q = c->d_q;
	p.g = c->g;
	p.start_ns = c->q_start;
	p.deadline_ns = c->q_deadline;
	p.ticket = c->q_ticket;
	p.waiting = c->q_waiting;
	p.mask = c->q_cap - 1;
	p.res = c->q_res;
	p.res_cap = c->q_res_cap;
	p.nslots = c->nslots;
	return p;
}
static int queue_ready(cb_ctx *c) {
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	if (!c->d_q)
		return fail(c, CB_ESTATE, "cb_queue_create has not been called");
	c->tot_valid = false;
	c->fseed = true;
	return drain(c);
}
static int queue_check(cb_ctx *c) {
	CU(c, cudaGetLastError());
	QueueDev h;
	CU(c, cudaMemcpyAsync(&h, c->d_q, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	if (h.err) { // report once: clear the device flag
		const unsigned int zero = 0;
		CU(c, cudaMemcpy((char *)c->d_q + offsetof(QueueDev, err), &zero, sizeof(zero), cudaMemcpyHostToDevice));
	}
	if (h.err & Q_ERR_FULL)
		return fail(c, CB_ENOMEM, "pending-claim queue is full (%llu handles queued)", h.tail - h.head);
	if (h.err & Q_ERR_RESULTS)
		return fail(c, CB_ENOMEM, "result buffer overflow: call cb_queue_results more often");
	if (h.err & Q_ERR_NONMONO)
		return fail(c, CB_EINVAL, "enqueue clocks must be non-decreasing");
	return CB_OK;
}

extern "C" int cb_queue_create(cb_ctx *c, size_t capacity, size_t result_capacity, uint64_t idle_slots) {
	if (!c || capacity == 0 || result_capacity == 0)
		return c ? fail(c, CB_EINVAL, "queue and result capacities must be positive") : CB_EINVAL;
	int rc = drain(c);
	if (rc)
		return rc;
	size_t cap = 1024;
	while (cap < capacity)
		cap <<= 1;
	void *old[] = {c->d_q, c->q_start, c->q_deadline, c->q_ticket, c->q_res, c->q_waiting};
	for (void *p : old)
		if (p)
			cudaFree(p);
	c->d_q = nullptr;
	c->q_start = c->q_deadline = c->q_ticket = c->q_res = nullptr;
	c->q_waiting = nullptr;
	if ((rc = dev_alloc(c, &c->d_q, 1)) || (rc = dev_alloc(c, &c->q_start, cap)) ||
	    (rc = dev_alloc(c, &c->q_deadline, cap)) || (rc = dev_alloc(c, &c->q_ticket, cap)) ||
	    (rc = dev_alloc(c, &c->q_waiting, cap)) || (rc = dev_alloc(c, &c->q_res, result_capacity * 3)))
		return rc;
	QueueDev h;
	memset(&h, 0, sizeof(h));
	h.idle = idle_slots;
	CU(c, cudaMemcpy(c->d_q, &h, sizeof(h), cudaMemcpyHostToDevice));
	c->q_cap = cap;
	c->q_res_cap = result_capacity;
	return CB_OK;
}

static int queue_stage(cb_ctx *c, const int64_t *host, size_t n) {
	if (n > c->q_stage_cap) {
		if (c->q_stage)
			cudaFree(c->q_stage);
		if (c->q_tmo)
			cudaFree(c->q_tmo);
		c->q_stage = nullptr;
		c->q_tmo = nullptr;
		c->q_stage_cap = 0;
		int rc;
		if ((rc = dev_alloc(c, &c->q_stage, n + n / 2 + 64)) || (rc = dev_alloc(c, &c->q_tmo, n + n / 2 + 64)))
			return rc;
		c->q_stage_cap = n + n / 2 + 64;
	}
	CU(c, cudaMemcpyAsync(c->q_stage, host, n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
	return CB_OK;
}

extern "C" int cb_queue_push(cb_ctx *c, const int64_t *start_ns, size_t n, int64_t *first_ticket, double *timeouts_ms) {
	if (!c || (n && !start_ns))
		return c ? fail(c, CB_EINVAL, "start_ns must not be NULL") : CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	QueueDev h;
	CU(c, cudaMemcpy(&h, c->d_q, sizeof(h), cudaMemcpyDeviceToHost));
	if (first_ticket)
		*first_ticket = (int64_t)h.next_ticket;
	if (n == 0)
		return CB_OK;
	if ((rc = queue_stage(c, start_ns, n)))
		return rc;
	queue_push_kernel<<<1, QTPB, 0, c->stream>>>(queue_params(c), c->q_stage, n, c->q_tmo);
	c->launches += 1;
	if (timeouts_ms)
		CU(c, cudaMemcpyAsync(timeouts_ms, c->q_tmo, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	return queue_check(c);
}

extern "C" int cb_queue_cancel(cb_ctx *c, const int64_t *tickets, size_t n) {
	if (!c || (n && !tickets))
		return c ? fail(c, CB_EINVAL, "tickets must not be NULL") : CB_EINVAL;
	int rc = queue_ready(c);
	if (rc || n == 0)
		return rc;
	if ((rc = queue_stage(c, tickets, n)))
		return rc;
	queue_cancel_kernel<<<1, QTPB, 0, c->stream>>>(queue_params(c), c->q_stage, n);
	c->launches += 1;
	return queue_check(c);
}

extern "C" int cb_queue_expire(cb_ctx *c, int64_t now_ns) {
	if (!c)
		return CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	queue_expire_kernel<<<1, QTPB, 0, c->stream>>>(queue_params(c), now_ns);
	c->launches += 1;
	return queue_check(c);
}

extern "C" int cb_queue_idle(cb_ctx *c, int64_t now_ns, uint64_t k_slots) {
	if (!c)
		return CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	queue_idle_kernel<<<1, QTPB, 0, c->stream>>>(queue_params(c), now_ns, k_slots);
	c->launches += 1;
	return queue_check(c);
}

extern "C" int cb_queue_results(cb_ctx *c, int64_t *triples, size_t cap_records, size_t *n_records) {
	if (!c || !n_records)
		return CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	QueueDev h;
	CU(c, cudaMemcpy(&h, c->d_q, sizeof(h), cudaMemcpyDeviceToHost));
	size_t n = (size_t)h.n_res;
	if (n > c->q_res_cap)
		n = c->q_res_cap;
	if (n > cap_records)
		return fail(c, CB_EINVAL, "%zu records are waiting, room for %zu", n, cap_records);
	if (n && triples)
		CU(c, cudaMemcpy(triples, c->q_res, n * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost));
	*n_records = n;
	const unsigned long long zero = 0;
	CU(c, cudaMemcpy((char *)c->d_q + offsetof(QueueDev, n_res), &zero, sizeof(zero), cudaMemcpyHostToDevice));
	return CB_OK;
}

extern "C" int cb_queue_info_get(cb_ctx *c, cb_queue_info *out) {
	if (!c || !out)
		return CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	static_assert(sizeof(cb_queue_info) == sizeof(QueueDev), "cb_queue_info mirrors QueueDev");
	CU(c, cudaMemcpy(out, c->d_q, sizeof(*out), cudaMemcpyDeviceToHost));
	return CB_OK;
}

// =========================================================================== measurement
extern "C" int cb_profile_enable(cb_ctx *c, int on) {
	if (!c)
		return CB_EINVAL;
	int rc = drain(c); // batches in flight were enqueued under the other mode
	if (rc)
		return rc;
	c->prof = on != 0;
	for (int k = 0; k < CB_K_COUNT; k++) {
		c->prof_ms[k] = 0;
		c->prof_n[k] = 0;
	}
	return CB_OK;
}
extern "C" int cb_profile_get(cb_ctx *c, double ms_out[CB_K_COUNT], uint64_t launches_out[CB_K_COUNT]) {
	if (!c)
		return CB_EINVAL;
	for (int k = 0; k < CB_K_COUNT; k++) {
		if (ms_out)
			ms_out[k] = c->prof_ms[k];
		if (launches_out)
			launches_out[k] = c->prof_n[k];
	}
	return CB_OK;
}
extern "C" uint64_t cb_launch_count(const cb_ctx *c) { return c ? c->launches : 0; }

extern "C" int cb_debug_lane_timestamps(cb_ctx *c, int lane, uint64_t out[16]) {
	if (!c || !out || lane < 0 || lane >= CB_LANES)
		return CB_EINVAL;
	int rc = drain(c);
	if (rc)
		return rc;
	BatchCtl b;
	memset(&b, 0, sizeof(b));
	if (c->lane[lane].ctl)
		CU(c, cudaMemcpy(&b, c->lane[lane].ctl, sizeof(b), cudaMemcpyDeviceToHost));
	for (int k = 0; k < 16; k++)
		out[k] = b.dbg[k];
	return CB_OK;
}

extern "C" int cb_debug_timestamps(cb_ctx *c, uint64_t out[16]) {
	if (!c || !out)
		return CB_EINVAL;
	int rc = drain(c);
	if (rc)
		return rc;
	BatchCtl b;
	memset(&b, 0, sizeof(b));
	if (c->last_lane >= 0)
		CU(c, cudaMemcpy(&b, c->lane[c->last_lane].ctl, sizeof(b), cudaMemcpyDeviceToHost));
	for (int k = 0; k < 16; k++)
		out[k] = b.dbg[k];
	return CB_OK;
}
