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
#include <arpa/inet.h>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "cb_device.cuh"
#include "cb_emit.cuh"
#include "cb_peers.cuh"
#include "cb_queue.cuh"
#include "cb_resolver.cuh"
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

constexpr int CB_LANES = 12;                  // most batches in flight (lanes)
// Lanes used unless CB_PIPELINE_DEPTH / cb_set_pipeline_depth says otherwise.  A batch's emit kernel only gets
// the SM slots retiring scan CTAs of later batches free, so a batch stays in flight for several steps: with 5
// lanes the step was lane-bound at every size (C3: 59.4 us, 63.3 on two GPUs with count boards; 8 lanes: 56.6 / 56.0;
// C2: 11.6 -> 9.1 us).
constexpr int CB_DEFAULT_DEPTH = 8;
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
	double *ev_now = nullptr, *hj_F = nullptr;
	cudaStream_t st = nullptr; // all four kernels of the lane's batch, in order (highest stream priority)
	cudaEvent_t res_done = nullptr, emit_done = nullptr, coll_done = nullptr, join_ev = nullptr;
	bool used = false, coll_pending = false;
	unsigned long long epoch = 0;
	// graph launch mode: the batch's kernels as one CUDA graph per (clock type, flags) variant, its
	// node parameters rewritten per batch: one launch call per batch
	struct Graph {
		cudaGraph_t g = nullptr;
		cudaGraphExec_t ex = nullptr;
		cudaGraphNode_t n_scan = nullptr, n_res = nullptr, n_emit = nullptr;
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
	bool fseed = true;  // the state was set from the host: the next batch's F slot needs seeding
	int graph_mode = -1; // -1: graph launch for batches of at most graph_max_tiles tiles; 0 never; 1 always (CB_GRAPH)
	size_t graph_max_tiles = 1024;
	int prio_least = 0, prio_greatest = 0;
	// per-batch global counts over peer memory (cb_peers_*)
	PeerBoard *d_board = nullptr;          // this rank's board
	PeerBoard **d_peer_ptrs = nullptr;     // device array of the ranks' boards (peer memory)
	void *peer_ipc[PB_MAXR] = {};          // boards opened from IPC handles (to be closed)
	int peers_n = 0, peers_rank = 0;
	unsigned int peers_dbg = 0;            // CB_PEERS_DEBUG (measurement only)
	unsigned long long peer_timeout_ns = 30000000000ull; // how long a kernel waits for another rank (CB_PEER_TIMEOUT_MS)
	unsigned long long peer_step = 0;      // batches posted since cb_peers_connect
	unsigned long long *d_peer_out = nullptr;
	int ring_n = 0, ring_rank = 0;
	ChainMem *ring_next = nullptr;         // the next rank's mailbox
	bool ring_next_ipc = false;            // ... opened from an IPC handle (to be closed)
	int last_lane = -1;
	bool last_flags = false;
	int emit_grid = 148 * 4;           // persistent emit CTAs (from the device's SM count)
	int scan_ctas = 148 * CB_SCAN_MINBLOCKS; // scan CTAs resident at once
	int scan_tpw = 2;                  // tiles a scan warp streams before its CTA retires (measured: 1, 2, 3, persistent)
	// topology
	int B = 0, nslots = 0, target_slots = 0, mode = CB_BALANCE_UNIFORM, key_cap = 0, slot_cap = 0;
	int32_t *d_keys = nullptr, *d_want = nullptr, *d_add = nullptr, *d_slot_backend = nullptr;
	int32_t *d_prio = nullptr, *d_weight = nullptr, *d_plan_scratch = nullptr;
	int *d_ints = nullptr; // [0]=len [1]=n_add [2]=nslots [3]=status [4]=n_rm [5]=scalar out
	double *d_dbl = nullptr;
	// host-buffer pipeline
	size_t chunk = DEFAULT_CHUNK, stage_cap = 0;
	size_t chunk_floor = 32 * (size_t)ETILE; // host-buffer path: smallest staging chunk (CB_CHUNK_FLOOR)
	int chunk_pct = 40;                      // ... and the share of the remainder a chunk takes (CB_CHUNK_PCT)
	cudaStream_t s_in2 = nullptr;            // CB_H2D_SPLIT=1: the `now` array travels on a second copy stream
	cudaEvent_t ev_in2 = nullptr;
	Staging st[NBUF];
	// pinned landing zone of fetch_stats: {chain word, error count}, DevState, BatchCtl in one synchronisation
	struct StatsStage {
		unsigned long long chain[2];
		DevState st;
		BatchCtl ctl;
	} *h_stats = nullptr;
	// profiling
	bool prof = false;
	bool prof_scan_only = false; // cb_profile_enable(ctx, 2): launch only the scan kernel (burst timing by the caller)
	cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
	double prof_ms[CB_K_COUNT] = {0, 0, 0};
	uint64_t prof_n[CB_K_COUNT] = {0, 0, 0};
	// resolver ingestion (cb_resolver_process): r_backends of lib/resolver.js:1024-1108 as srvKey -> key id,
	// in insertion order
	std::vector<std::string> r_keys;
	std::unordered_map<std::string, int32_t> r_ids;
	int32_t r_next_id = 0;
	// nccl
	nccl_comm_t comm = nullptr;
	int nranks = 1, rank = 0;
};

namespace {

// Every entry point runs with the context's device current and leaves the caller's device as it found it
// (a host with contexts on several GPUs, or one that switches devices between calls).
struct DevGuard {
	int prev = -1;
	explicit DevGuard(const cb_ctx *c) {
		if (!c)
			return;
		int cur = -1;
		if (cudaGetDevice(&cur) == cudaSuccess && cur != c->device) {
			prev = cur;
			cudaSetDevice(c->device);
		}
	}
	~DevGuard() {
		if (prev >= 0)
			cudaSetDevice(prev);
	}
	DevGuard(const DevGuard &) = delete;
	DevGuard &operator=(const DevGuard &) = delete;
};

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

// a device allocation that lives as long as the call that made it (freed on every return path)
struct ScopedDev {
	void *p = nullptr;
	~ScopedDev() {
		if (p)
			cudaFree(p);
	}
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
	template <typename T> T *as() const { return static_cast<T *>(p); }
	ScopedDev() = default;
	ScopedDev(const ScopedDev &) = delete;
	ScopedDev &operator=(const ScopedDev &) = delete;
};

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

static void peers_close(cb_ctx *c) {
	for (int r = 0; r < PB_MAXR; r++) {
		if (c->peer_ipc[r])
			cudaIpcCloseMemHandle(c->peer_ipc[r]);
		c->peer_ipc[r] = nullptr;
	}
	c->peers_n = 0;
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

// The scan's grid: every warp streams up to `scan_tpw` tiles of its own (consecutive tiles on consecutive
// CTAs).  A CTA's life is bounded on purpose: the single-CTA resolve kernel and the emit kernel of the
// batches ahead in the pipeline run at a higher priority and take the SM slots retiring scan CTAs free;
// fully persistent scan CTAs would hold every SM until the whole batch is read.  Small batches still get
// one CTA per SM slot.
unsigned scan_grid(const cb_ctx *c, size_t tiles) {
	const size_t per_cta = (size_t)SC_WARPS * (size_t)c->scan_tpw;
	size_t g = (tiles + per_cta - 1) / per_cta;
	const size_t spread = tiles < (size_t)c->scan_ctas ? tiles : (size_t)c->scan_ctas;
	if (g < spread)
		g = spread;
	return (unsigned)g;
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
	// the emit kernel's sum descriptors carry an epoch tag, so they only need zeroing once
	CU(c, cudaMemset(L.sdesc, 0, cap * sizeof(SumDesc)));
	L.cap_tiles = cap;
	return CB_OK;
}

// Graph launch mode: scan -> resolve -> emit as the nodes of one CUDA graph per lane and variant.  The
// first batch of a variant builds and instantiates it; every later one rewrites the nodes' parameters
// (pointers, sizes, grids) in the instantiated graph and launches it: one launch call per batch instead
// of three.  The scan node keeps the lowest priority, the others the highest.
PeerLink peer_link(const cb_ctx *c, unsigned long long step) {
	PeerLink pl;
	pl.mine = c->peers_n > 0 ? c->d_board : nullptr;
	pl.peers = c->d_peer_ptrs;
	pl.R = c->peers_n;
	pl.me = c->peers_rank;
	pl.step = step;
	pl.timeout_ns = c->peer_timeout_ns; // the other ranks are other processes: allow for start-up skew
	pl.dbg = c->peers_dbg;
	if (c->peers_dbg & 1u) // measurement: boards connected, nothing posted
		pl.mine = nullptr;
	return pl;
}
template <bool NS>
int graph_batch(cb_ctx *c, Lane &L, bool flags, ScanParams &sp, ResolveParams &rp, EmitParams &ep, size_t tiles,
    unsigned egrid, cudaStream_t s) {
	Lane::Graph &G = L.gr[NS ? 1 : 0][flags ? 1 : 0];
	void *a_scan[] = {(void *)&sp}, *a_res[] = {(void *)&rp}, *a_emit[] = {(void *)&ep};
	cudaKernelNodeParams k_scan = {}, k_res = {}, k_emit = {};
	k_scan.func = flags ? (void *)scan_kernel<NS, true> : (void *)scan_kernel<NS, false>;
	k_scan.gridDim = dim3(scan_grid(c, tiles));
	k_scan.blockDim = dim3(SC_TPB);
	k_scan.sharedMemBytes = (unsigned)SC_SMEM;
	k_scan.kernelParams = a_scan;
	k_res.func = flags ? (void *)resolve_kernel<NS, true> : (void *)resolve_kernel<NS, false>;
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
		CU(c, cudaGraphAddKernelNode(&G.n_emit, G.g, &G.n_res, 1, &k_emit));
		cudaKernelNodeAttrValue lo, hi;
		memset(&lo, 0, sizeof(lo));
		memset(&hi, 0, sizeof(hi));
		lo.priority = c->prio_least;
		hi.priority = c->prio_greatest;
		CU(c, cudaGraphKernelNodeSetAttribute(G.n_scan, cudaKernelNodeAttributePriority, &lo));
		CU(c, cudaGraphKernelNodeSetAttribute(G.n_res, cudaKernelNodeAttributePriority, &hi));
		CU(c, cudaGraphKernelNodeSetAttribute(G.n_emit, cudaKernelNodeAttributePriority, &hi));
		CU(c, cudaGraphInstantiate(&G.ex, G.g, 0));
	} else {
		CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_scan, &k_scan));
		CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_res, &k_res));
		CU(c, cudaGraphExecKernelNodeSetParams(G.ex, G.n_emit, &k_emit));
	}
	CU(c, cudaGraphLaunch(G.ex, s));
	return CB_OK;
}

// Enqueue the three kernels (scan, resolve, emit) of one batch.  depth 1 (or profiling): all on the main stream, in
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
	if (n >= (1ull << 39))
		return fail(c, CB_EINVAL, "batch too large");
	if (c->peers_n > 0 && n >= (1ull << 32))
		return fail(c, CB_EINVAL, "count boards are connected: a batch must have fewer than 2^32 claims");
	const size_t tiles = (n + TILE - 1) / TILE;      // scan / prepare tiles
	const int depth = c->depth;
	const bool piped = (depth > 1) && !c->prof;
	const int li = piped ? (int)(c->seq % (unsigned long long)depth) : 0;
	Lane &L = c->lane[li];
	const size_t etiles = (n + ETILE - 1) / ETILE;   // emit tiles
	int rc = ensure_scratch(c, L, tiles);
	if (rc)
		return rc;
	if (piped) {
		// size every active lane on the first batch of a new size: the allocations (which synchronise)
		// happen once, up front, not one lane at a time over the next `depth` batches
		for (int l = 0; l < depth; l++)
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
	{
		ChainLink &ln = rp.link;
		ln.err = &c->d_chain->err;
		ln.wait_word = &c->d_chain->seq;
		if (c->ring_n > 0) {
			ln.wait_val = c->chain_seq * (unsigned long long)c->ring_n + (unsigned long long)c->ring_rank;
			ln.in_state = &c->d_chain->st;
			ln.post_word = &c->ring_next->seq;
			ln.post_state = &c->ring_next->st;
			ln.timeout_ns = c->peer_timeout_ns; // the other ranks are other processes: allow for start-up skew
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
	rp.f_in = c->d_chain->fs;
	rp.f_out = c->ring_n > 0 ? c->ring_next->fs : c->d_chain->fs;
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
	// counts for the other ranks' boards, posted by the emit kernel's last CTA; a batch takes a board step only
	// when its emit kernel really runs (the scan-only measurement mode launches none)
	ep.peer = peer_link(c, c->prof_scan_only ? ~0ull : c->peer_step);
	if (c->prof_scan_only)
		ep.peer.mine = nullptr;
	else if (c->peers_n > 0)
		c->peer_step++;

	ep.vec_ok = (((uintptr_t)d_decision & 7) == 0) && (((uintptr_t)d_slot & 15) == 0) &&
	    (((uintptr_t)d_backend & 15) == 0);
	// persistent CTAs stride over the tiles; with flags the admit-count look-back needs every
	// earlier tile to be resident or earlier in the same CTA's sequence: both hold for a grid
	// within the resident capacity (a CTA visits its tiles in increasing order)
	// (+1 CTA that handles no tile when count boards are connected: see emit_kernel)
	const unsigned egrid = (unsigned)(etiles < (size_t)c->emit_grid ? etiles : (size_t)c->emit_grid) + (ep.peer.mine ? 1u : 0u);

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
	if (c->fseed) {
		// the state was set from the host: seed the slot this batch reads.  In a ring only the queue's
		// first batch (rank 0) is seeded; every other batch gets its value from its predecessor.
		if (c->ring_n == 0 || (c->ring_rank == 0 && c->chain_seq == 0))
			fchain_seed_kernel<<<1, 1, 0, s>>>(c->g, c->d_chain->fs, rp.link.wait_val, rp.f_ring);
		c->fseed = false;
	}
	if (c->prof && !c->prof_scan_only)
		CU(c, cudaEventRecord(c->pe[0], s));
	const bool use_graph = piped && (c->graph_mode == 1 || (c->graph_mode < 0 && tiles <= c->graph_max_tiles));
	if (use_graph) {
		rc = graph_batch<NS>(c, L, flags, sp, rp, ep, tiles, egrid, s);
		if (rc)
			return rc;
		c->chain_seq++;
		if (c->comm) // the graph has no interior point to record: the collective starts after the emit
			CU(c, cudaEventRecord(L.res_done, s));
		c->launches += 3;
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
		cfg.gridDim = dim3(scan_grid(c, tiles));
		cfg.blockDim = dim3(SC_TPB);
		cfg.dynamicSmemBytes = SC_SMEM;
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
		scan_kernel<NS, true><<<scan_grid(c, tiles), SC_TPB, SC_SMEM, s>>>(sp);
	else
		scan_kernel<NS, false><<<scan_grid(c, tiles), SC_TPB, SC_SMEM, s>>>(sp);
	if (c->prof_scan_only) { // measurement: back-to-back scan launches, timed by the caller's events on this stream
		c->launches += 1;
		CU(c, cudaGetLastError());
		return CB_OK;
	}
	if (c->prof)
		CU(c, cudaEventRecord(c->pe[1], s));
	// resolve = prepare_part (independent of the ControlledDelay state: overlaps the other lanes' commits)
	// + commit_part, the only stage that is serial across batches.  The order is kept inside the kernel:
	// commit(k) spins (one thread) on the chain word until batch k-1 has published its state, so the
	// hand-over costs an L2 round trip instead of a cross-stream event (~5 us measured).
	if (flags)
		resolve_kernel<NS, true><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
	else
		resolve_kernel<NS, false><<<1, RTPB, NIDX_MAX * sizeof(double), s>>>(rp);
	c->chain_seq++;
	if (c->prof)
		CU(c, cudaEventRecord(c->pe[2], s));
	if (piped && c->comm) // the drop-count collective starts from here
		CU(c, cudaEventRecord(L.res_done, s));
	if (flags)
		emit_kernel<true><<<egrid, ETPB, ep.table_smem_bytes, s>>>(ep);
	else
		emit_kernel<false><<<egrid, ETPB, ep.table_smem_bytes, s>>>(ep);
	c->launches += 3;
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

// three small device-to-host copies into pinned memory behind whatever s_out already holds (a blocking
// cudaMemcpy each costs ~10 us); valid once s_out has been waited for
int enqueue_stats_copies(cb_ctx *c) {
	if (!c->h_stats)
		CU(c, cudaMallocHost((void **)&c->h_stats, sizeof(*c->h_stats)));
	memset(&c->h_stats->ctl, 0, sizeof(BatchCtl));
	CU(c, cudaMemcpyAsync(c->h_stats->chain, c->d_chain, sizeof(c->h_stats->chain), cudaMemcpyDeviceToHost, c->s_out));
	CU(c, cudaMemcpyAsync(&c->h_stats->st, c->g, sizeof(DevState), cudaMemcpyDeviceToHost, c->s_out));
	if (c->last_lane >= 0)
		CU(c, cudaMemcpyAsync(&c->h_stats->ctl, c->lane[c->last_lane].ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, c->s_out));
	return CB_OK;
}

int fetch_stats(cb_ctx *c, cb_batch_stats *out, bool copied = false) {
	// `copied`: the caller (host_batch) queued the copies behind its last copy-out and has waited for s_out
	if (!copied) {
		int rc = drain(c);
		if (rc)
			return rc;
		rc = enqueue_stats_copies(c);
		if (rc)
			return rc;
		CU(c, cudaStreamSynchronize(c->s_out));
	}
	const unsigned long long *chain = c->h_stats->chain;
	const DevState &h = c->h_stats->st;
	const BatchCtl &b = c->h_stats->ctl;
	if (chain[1] != 0 || (c->ring_n == 0 && chain[0] != c->chain_seq))
		return fail(c, CB_ESTATE, "commit chain broken: %llu of %llu batches committed, %llu hand-over timeouts", chain[0],
		    c->chain_seq, chain[1]);
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
	// Staging chunks.  The H2D copies run back to back on the copy engine; a chunk's kernels and its D2H
	// overlap the copy-in of the chunks behind it, so what the call pays beyond the PCIe time of its inputs
	// is the LAST chunk's kernels and copy-out.  Chunks therefore shrink towards the end (40 % of what is left,
	// at most c->chunk, at least 128K claims -- every chunk costs two copies whose fixed cost on the copy engine
	// is ~10 us each, and a chunk's kernels are latency-bound below that): the exposed tail is ~50 us at any
	// size (measured with CB_DEBUG_E2E=1: the last chunk's outputs land 45-50 us after its inputs).
	std::vector<size_t> plan;
	{
		size_t floor_chunk = c->chunk_floor; // 128K claims: below this the per-chunk cost wins
		if (floor_chunk > c->chunk)
			floor_chunk = c->chunk; // (a caller-set staging chunk smaller than that: CB_CHUNK)
		size_t left = n;
		while (left > 0) {
			size_t m = ((left / 100 * c->chunk_pct) + ETILE - 1) / ETILE * ETILE;
			if (m > c->chunk)
				m = c->chunk;
			if (m < floor_chunk)
				m = floor_chunk;
			if (m > left || (left - m < floor_chunk / 2 && left <= c->stage_cap))
				m = left; // a sliver would cost more than it saves (the staging buffers hold stage_cap claims)
			plan.push_back(m);
			left -= m;
		}
	}
	const size_t nchunks = plan.size();
	size_t lo = 0;
	bool all_waited = (decision || slot || backend) && n > 0;
	std::vector<cudaEvent_t> dbg_ev; // CB_DEBUG_E2E: when each chunk's inputs and outputs landed (device clock)
	cudaEvent_t dbg_ev0 = nullptr;
	if (dbg_t) {
		CU(c, cudaEventCreate(&dbg_ev0));
		CU(c, cudaEventRecord(dbg_ev0, c->s_in));
	}
	for (size_t ci = 0; ci < nchunks; ci++) {
		Staging &s = c->st[ci % NBUF];
		const size_t m = plan[ci];
		if (ci >= NBUF)
			CU(c, cudaStreamWaitEvent(c->s_in, s.out_done, 0));
		CU(c, cudaMemcpyAsync(s.start, (const char *)start + lo * 8, m * 8, cudaMemcpyHostToDevice, c->s_in));
		if (c->s_in2) {
			if (ci >= NBUF)
				CU(c, cudaStreamWaitEvent(c->s_in2, s.out_done, 0));
			CU(c, cudaMemcpyAsync(s.now, (const char *)now + lo * 8, m * 8, cudaMemcpyHostToDevice, c->s_in2));
			CU(c, cudaEventRecord(c->ev_in2, c->s_in2));
			CU(c, cudaStreamWaitEvent(c->s_in, c->ev_in2, 0));
		} else {
			CU(c, cudaMemcpyAsync(s.now, (const char *)now + lo * 8, m * 8, cudaMemcpyHostToDevice, c->s_in));
		}
		if (flags)
			CU(c, cudaMemcpyAsync(s.flags, flags + lo, m, cudaMemcpyHostToDevice, c->s_in));
		CU(c, cudaEventRecord(s.in_done, c->s_in));
		if (dbg_t) {
			cudaEvent_t e;
			CU(c, cudaEventCreate(&e));
			CU(c, cudaEventRecord(e, c->s_in));
			dbg_ev.push_back(e);
		}
		cudaEvent_t done = nullptr;
		rc = enqueue_batch<NS>(c, s.start, s.now, flags ? s.flags : nullptr, m, decision ? s.decision : nullptr,
		    slot ? s.slot : nullptr, backend ? s.backend : nullptr, s.in_done, &done);
		if (rc)
			return rc;
		if (done)
			CU(c, cudaStreamWaitEvent(c->s_out, done, 0));
		else
			all_waited = false;
		if (decision)
			CU(c, cudaMemcpyAsync(decision + lo, s.decision, m, cudaMemcpyDeviceToHost, c->s_out));
		if (slot)
			CU(c, cudaMemcpyAsync(slot + lo, s.slot, m * 4, cudaMemcpyDeviceToHost, c->s_out));
		if (backend)
			CU(c, cudaMemcpyAsync(backend + lo, s.backend, m * 4, cudaMemcpyDeviceToHost, c->s_out));
		CU(c, cudaEventRecord(s.out_done, c->s_out));
		if (dbg_t) {
			cudaEvent_t e;
			CU(c, cudaEventCreate(&e));
			CU(c, cudaEventRecord(e, c->s_out));
			dbg_ev.push_back(e);
		}
		lo += m;
	}
	// with outputs copied out, s_out runs behind every kernel of the call: the statistics ride the same wait
	const bool stats_inline = stats && all_waited;
	if (stats_inline && (rc = enqueue_stats_copies(c)))
		return rc;
	const double T2 = dbg_t ? tnow() : 0;
	CU(c, cudaStreamSynchronize(c->s_out));
	const double T3 = dbg_t ? tnow() : 0;
	if (dbg_t) {
		fprintf(stderr, "[cb e2e] n=%zu chunks=%zu  entry sync+state %.0f us | enqueue %.0f us | wait %.0f us\n", n, nchunks,
		    T1 - T0, T2 - T1, T3 - T2);
		for (size_t k = 0; k + 1 < dbg_ev.size(); k += 2) {
			float a = 0, b = 0;
			cudaEventElapsedTime(&a, dbg_ev0, dbg_ev[k]);
			cudaEventElapsedTime(&b, dbg_ev0, dbg_ev[k + 1]);
			fprintf(stderr, "[cb e2e]   chunk %zu (%zu claims): in by %.0f us, out by %.0f us\n", k / 2, plan[k / 2], a * 1e3,
			    b * 1e3);
			cudaEventDestroy(dbg_ev[k]);
			cudaEventDestroy(dbg_ev[k + 1]);
		}
		cudaEventDestroy(dbg_ev0);
	}
	if (stats) {
		rc = fetch_stats(c, stats, stats_inline);
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
		plan_rebalance_kernel<<<1, TOPO_TPB, 0, s>>>(d_have, d_dead, c->B, c->nslots, c->nslots, 0, c->d_want, c->d_add,
		    d_nadd, nullptr, nullptr, c->d_ints + 4, c->d_plan_scratch);
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
	    (rc = dev_alloc(c, &c->d_add, (size_t)sc + kc + kc)) || (rc = dev_alloc(c, &c->d_slot_backend, sc)) ||
	    (rc = dev_alloc(c, &c->d_plan_scratch, 2 * ((size_t)kc + 1))))
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
	int prev_dev = -1;
	if (cudaGetDevice(&prev_dev) != cudaSuccess)
		prev_dev = -1;
	c->device = device;
	c->emit_grid = prop.multiProcessorCount * 4;
	c->scan_ctas = prop.multiProcessorCount * CB_SCAN_MINBLOCKS;
	if (getenv("CB_SCAN_CTAS") && atoi(getenv("CB_SCAN_CTAS")) > 0) // tuning knobs (measurement only)
		c->scan_ctas = atoi(getenv("CB_SCAN_CTAS"));
	if (getenv("CB_SCAN_TPW") && atoi(getenv("CB_SCAN_TPW")) > 0)
		c->scan_tpw = atoi(getenv("CB_SCAN_TPW"));
	if (const char *e = getenv("CB_CHUNK_FLOOR"))
		if (atoll(e) >= (long long)ETILE)
			c->chunk_floor = (size_t)atoll(e) / ETILE * ETILE;
	if (const char *e = getenv("CB_CHUNK_PCT"))
		if (atoi(e) >= 5 && atoi(e) <= 100)
			c->chunk_pct = atoi(e);
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
			if (prev_dev >= 0)                                                             \
				cudaSetDevice(prev_dev);                                                   \
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
	if (const char *e = getenv("CB_H2D_SPLIT"))
		if (atoi(e) == 1) {
			CU0(cudaStreamCreateWithFlags(&c->s_in2, cudaStreamNonBlocking));
			CU0(cudaEventCreateWithFlags(&c->ev_in2, cudaEventDisableTiming));
		}
	CU0(cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming));
	c->stream = c->own_stream;
	CU0(cudaMalloc((void **)&c->g, sizeof(DevState)));
	CU0(cudaMemset(c->g, 0, sizeof(DevState)));
	if (getenv("CB_PEER_TIMEOUT_MS") && atoll(getenv("CB_PEER_TIMEOUT_MS")) > 0)
		c->peer_timeout_ns = (unsigned long long)atoll(getenv("CB_PEER_TIMEOUT_MS")) * 1000000ull;
	if (getenv("CB_PEERS_DEBUG"))
		c->peers_dbg = (unsigned int)atoi(getenv("CB_PEERS_DEBUG"));
	if (getenv("CB_GRAPH")) // measurement override of the launch mode: 0 = never a graph, 1 = always
		c->graph_mode = atoi(getenv("CB_GRAPH")) != 0 ? 1 : 0;
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
	CU0(cudaFuncSetAttribute(scan_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
	CU0(cudaFuncSetAttribute(scan_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
	CU0(cudaFuncSetAttribute(scan_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
	CU0(cudaFuncSetAttribute(scan_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
	CU0(cudaFuncSetAttribute(resolve_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(resolve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(resolve_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(resolve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NIDX_MAX * 8));
	CU0(cudaFuncSetAttribute(emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
	CU0(cudaFuncSetAttribute(emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
	{
		// one shared-memory carveout for every kernel of the pipeline: the scan needs nearly all of an SM's
		// shared memory, and an SM whose carveout differs between consecutive kernels is reconfigured (drained)
		// at every switch
		const void *fns[] = {(const void *)scan_kernel<true, true>, (const void *)scan_kernel<true, false>,
		    (const void *)scan_kernel<false, true>, (const void *)scan_kernel<false, false>,
		    (const void *)resolve_kernel<true, true>, (const void *)resolve_kernel<true, false>,
		    (const void *)resolve_kernel<false, true>, (const void *)resolve_kernel<false, false>,
		    (const void *)emit_kernel<true>, (const void *)emit_kernel<false>,
		    (const void *)fchain_seed_kernel};
		for (const void *f : fns)
			CU0(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
	}
#undef CU0
	if (prev_dev >= 0 && prev_dev != device) // the caller's current device is left as it was
		cudaSetDevice(prev_dev);
	*out = c;
	return CB_OK;
}

extern "C" void cb_destroy(cb_ctx *c) {
	DevGuard dev_guard_(c);
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaStream_t all[] = {c->st_coll, c->s_in, c->s_in2, c->s_out, c->own_stream};
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
		void *fr[] = {L.ctl, L.sdesc, L.ev_count, L.rec_fc, L.rec_cend, L.xcnt, L.hj_info, L.ticket, L.ev, L.evc,
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
	peers_close(c);
	if (c->d_board)
		cudaFree(c->d_board);
	if (c->d_peer_ptrs)
		cudaFree(c->d_peer_ptrs);
	if (c->d_peer_out)
		cudaFree(c->d_peer_out);
	void *qfrees[] = {c->d_q, c->q_start, c->q_deadline, c->q_ticket, c->q_res, c->q_stage, c->q_waiting, c->q_tmo};
	for (void *p : qfrees)
		if (p)
			cudaFree(p);
	void *frees[] = {c->g, c->d_chain, c->d_keys, c->d_want, c->d_add, c->d_slot_backend, c->d_prio, c->d_weight, c->d_ints,
	    c->d_dbl, c->d_plan_scratch};
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
	if (c->h_stats)
		cudaFreeHost(c->h_stats);
	for (int k = 0; k < 4; k++)
		if (c->pe[k])
			cudaEventDestroy(c->pe[k]);
	if (c->ev_user)
		cudaEventDestroy(c->ev_user);
	if (c->ev_in2)
		cudaEventDestroy(c->ev_in2);
	for (cudaStream_t s : all)
		if (s)
			cudaStreamDestroy(s);
	delete c;
}

extern "C" int cb_set_stream(cb_ctx *c, void *stream) {
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	return drain(c);
}

extern "C" void *cb_host_alloc(cb_ctx *c, size_t bytes) {
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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

extern "C" int cb_codel_overloaded_ms(cb_ctx *c, double start_ms, double now_ms, int *drop) {
	DevGuard dev_guard_(c);
	if (!c || !drop)
		return CB_EINVAL;
	c->tot_valid = false;
	c->fseed = true;
	if (!c->codel_ready)
		return fail(c, CB_ESTATE, "cb_codel_init has not been called");
	int rc = drain(c);
	if (rc)
		return rc;
	overloaded_one_ms_kernel<<<1, 1, 0, c->stream>>>(c->g, start_ms, now_ms, c->d_ints + 5);
	c->launches++;
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(drop, c->d_ints + 5, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return CB_OK;
}

// =============================================================================== batches
extern "C" int cb_admit_batch(cb_ctx *c, const int64_t *start_ns, const int64_t *now_ns, const uint8_t *flags,
    size_t n, uint8_t *decision, int32_t *slot, int32_t *backend, cb_batch_stats *stats) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	return host_batch<true>(c, start_ns, now_ns, flags, n, decision, slot, backend, stats);
}
extern "C" int cb_admit_batch_ms(cb_ctx *c, const double *start_ms, const double *now_ms, const uint8_t *flags,
    size_t n, uint8_t *decision, int32_t *slot, int32_t *backend, cb_batch_stats *stats) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	return host_batch<false>(c, start_ms, now_ms, flags, n, decision, slot, backend, stats);
}
extern "C" int cb_admit_batch_dev(cb_ctx *c, const int64_t *d_start, const int64_t *d_now, const uint8_t *d_flags,
    size_t n, uint8_t *d_decision, int32_t *d_slot, int32_t *d_backend) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	return enqueue_batch<true>(c, d_start, d_now, d_flags, n, d_decision, d_slot, d_backend, nullptr, nullptr);
}
extern "C" int cb_admit_batch_ms_dev(cb_ctx *c, const double *d_start, const double *d_now, const uint8_t *d_flags,
    size_t n, uint8_t *d_decision, int32_t *d_slot, int32_t *d_backend) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	return enqueue_batch<false>(c, d_start, d_now, d_flags, n, d_decision, d_slot, d_backend, nullptr, nullptr);
}
extern "C" int cb_batch_stats_get(cb_ctx *c, cb_batch_stats *out) {
	DevGuard dev_guard_(c);
	if (!c || !out)
		return CB_EINVAL;
	return fetch_stats(c, out);
}

// ============================================================================== topology
extern "C" int cb_backends_set(cb_ctx *c, const cb_backend *tbl, int32_t B, const double *insert_u, int32_t nslots,
    int mode) {
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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

extern "C" int cb_shuffle(cb_ctx *c, int32_t *arr, size_t n, const double *us) {
	DevGuard dev_guard_(c);
	if (!c || (n && (!arr || !us)))
		return c ? fail(c, CB_EINVAL, "bad arguments") : CB_EINVAL;
	if (n >= (1ull << 31))
		return fail(c, CB_EINVAL, "array too long");
	for (size_t k = 0; k < n; k++)
		if (!(us[k] >= 0.0 && us[k] < 1.0))
			return fail(c, CB_EINVAL, "us[%zu] is not in [0,1)", k);
	if (n < 2)
		return CB_OK; // one transposition of an element with itself, or nothing
	ScopedDev d_in, d_out, d_us;
	CU(c, d_in.alloc(n * 4));
	CU(c, d_out.alloc(n * 4));
	CU(c, d_us.alloc(n * 8));
	cudaStream_t s = c->stream;
	CU(c, cudaMemcpyAsync(d_in.p, arr, n * 4, cudaMemcpyHostToDevice, s));
	CU(c, cudaMemcpyAsync(d_us.p, us, n * 8, cudaMemcpyHostToDevice, s));
	size_t ctas = (n + TOPO_TPB - 1) / TOPO_TPB;
	if (ctas > (size_t)c->emit_grid / 2)
		ctas = (size_t)c->emit_grid / 2; // two CTAs of 1024 threads per SM
	shuffle_kernel<<<(unsigned)ctas, TOPO_TPB, 0, s>>>(d_in.as<int32_t>(), d_out.as<int32_t>(), d_us.as<double>(), (int)n);
	c->launches++;
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(arr, d_out.p, n * 4, cudaMemcpyDeviceToHost, s));
	CU(c, cudaStreamSynchronize(s));
	return CB_OK;
}

extern "C" int cb_backends_get(cb_ctx *c, int32_t *n_keys, int32_t *keys_out, int32_t *n_slots,
    int32_t *slot_backend_out) {
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	ScopedDev b_have, b_want, b_add, b_rk, b_ro, b_dead, b_cnt, b_scr;
	CU(c, b_have.alloc(nb * 4));
	CU(c, b_want.alloc(nb * 4));
	CU(c, b_dead.alloc(nb));
	CU(c, b_add.alloc(na * 4));
	CU(c, b_rk.alloc(nr * 4));
	CU(c, b_ro.alloc(nr * 4));
	CU(c, b_cnt.alloc(8));
	CU(c, b_scr.alloc(2 * (nb + 1) * 4));
	int32_t *d_have = b_have.as<int32_t>(), *d_want = b_want.as<int32_t>(), *d_add = b_add.as<int32_t>();
	int32_t *d_rk = b_rk.as<int32_t>(), *d_ro = b_ro.as<int32_t>(), *d_scr = b_scr.as<int32_t>();
	uint8_t *d_dead = b_dead.as<uint8_t>();
	int *d_cnt = b_cnt.as<int>();
	cudaStream_t s = c->stream;
	if (B > 0) {
		CU(c, cudaMemcpyAsync(d_have, have, nb * 4, cudaMemcpyHostToDevice, s));
		CU(c, cudaMemcpyAsync(d_dead, dead, nb, cudaMemcpyHostToDevice, s));
	}
	plan_rebalance_kernel<<<1, TOPO_TPB, 0, s>>>(d_have, d_dead, B, target, max, singleton, d_want, d_add, d_cnt, d_rk,
	    d_ro, d_cnt + 1, d_scr);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
		// a mailbox of another context of this process: if it lives on another GPU this device needs
		// peer access to it (the IPC path enables it lazily)
		cudaPointerAttributes pa;
		if (cudaPointerGetAttributes(&pa, next_local_ptr) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
		    pa.device != c->device) {
			cudaError_t e = cudaDeviceEnablePeerAccess(pa.device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				return fail(c, CB_ECUDA, "no peer access from device %d to device %d: %s", c->device, pa.device,
				    cudaGetErrorString(e));
			(void)cudaGetLastError();
		}
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

// ---- per-batch global counts over peer memory ---------------------------------------------------
static int peers_board(cb_ctx *c) {
	if (c->d_board)
		return CB_OK;
	CU(c, cudaMalloc((void **)&c->d_board, sizeof(PeerBoard)));
	CU(c, cudaMemset(c->d_board, 0, sizeof(PeerBoard)));
	CU(c, cudaMalloc((void **)&c->d_peer_ptrs, sizeof(PeerBoard *) * PB_MAXR));
	CU(c, cudaMalloc((void **)&c->d_peer_out, 4 * sizeof(unsigned long long)));
	return CB_OK;
}

extern "C" int cb_peers_board(cb_ctx *c, void *ipc_handle_out, size_t *len, void **local_ptr_out) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	int rc = peers_board(c);
	if (rc)
		return rc;
	if (ipc_handle_out) {
		if (!len || *len < sizeof(cudaIpcMemHandle_t))
			return fail(c, CB_EINVAL, "the IPC handle needs %zu bytes", sizeof(cudaIpcMemHandle_t));
		cudaIpcMemHandle_t h;
		CU(c, cudaIpcGetMemHandle(&h, c->d_board));
		memcpy(ipc_handle_out, &h, sizeof(h));
		*len = sizeof(h);
	}
	if (local_ptr_out)
		*local_ptr_out = c->d_board;
	return CB_OK;
}

extern "C" int cb_peers_connect(cb_ctx *c, int nranks, int rank, const void *ipc_handles, void *const *local_ptrs) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	int rc = drain(c);
	if (rc)
		return rc;
	peers_close(c);
	if (nranks == 0)
		return CB_OK;
	if (nranks < 1 || nranks > PB_MAXR || rank < 0 || rank >= nranks)
		return fail(c, CB_EINVAL, "bad peer geometry (at most %d ranks)", PB_MAXR);
	if (nranks > 1 && !ipc_handles && !local_ptrs)
		return fail(c, CB_EINVAL, "give the ranks' boards as IPC handles or as pointers of this process");
	if ((rc = peers_board(c)))
		return rc;
	CU(c, cudaMemset(c->d_board, 0, sizeof(PeerBoard)));
	PeerBoard *ptrs[PB_MAXR] = {};
	for (int r = 0; r < nranks; r++) {
		if (r == rank) {
			ptrs[r] = c->d_board;
		} else if (local_ptrs && local_ptrs[r]) {
			cudaPointerAttributes pa;
			if (cudaPointerGetAttributes(&pa, local_ptrs[r]) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
			    pa.device != c->device) {
				cudaError_t e = cudaDeviceEnablePeerAccess(pa.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
					return fail(c, CB_ECUDA, "no peer access from device %d to device %d: %s", c->device, pa.device,
					    cudaGetErrorString(e));
				(void)cudaGetLastError();
			}
			ptrs[r] = (PeerBoard *)local_ptrs[r];
		} else if (ipc_handles) {
			cudaIpcMemHandle_t h;
			memcpy(&h, (const char *)ipc_handles + (size_t)r * sizeof(h), sizeof(h));
			void *ptr = nullptr;
			cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {
				peers_close(c);
				return fail(c, CB_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
			}
			c->peer_ipc[r] = ptr;
			ptrs[r] = (PeerBoard *)ptr;
		} else {
			return fail(c, CB_EINVAL, "no board given for rank %d", r);
		}
	}
	CU(c, cudaMemcpy(c->d_peer_ptrs, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice));
	c->peers_n = nranks;
	c->peers_rank = rank;
	c->peer_step = 0;
	return CB_OK;
}

extern "C" int cb_peers_counts(cb_ctx *c, cb_global_counts *out) {
	DevGuard dev_guard_(c);
	if (!c || !out)
		return CB_EINVAL;
	if (c->peers_n <= 0)
		return fail(c, CB_ESTATE, "cb_peers_connect has not been called");
	// after every batch enqueued so far (their post kernels are on the lanes' streams)
	int rc = cb_join(c);
	if (rc)
		return rc;
	peer_counts_kernel<<<1, 32, 0, c->stream>>>(peer_link(c, c->peer_step), c->peer_step, c->d_peer_out);
	c->launches++;
	CU(c, cudaGetLastError());
	unsigned long long h[4] = {0, 0, 0, 0};
	CU(c, cudaMemcpyAsync(h, c->d_peer_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	unsigned long long err = 0;
	CU(c, cudaMemcpy(&err, (char *)c->d_board + offsetof(PeerBoard, err), sizeof(err), cudaMemcpyDeviceToHost));
	if (err != 0 || h[0] < c->peer_step)
		return fail(c, CB_ESTATE, "peer boards: %llu of %llu steps complete on every rank, %llu time-outs", h[0],
		    c->peer_step, err);
	out->steps = h[0];
	out->drops = h[1];
	out->admits = h[2];
	out->n = h[3];
	return CB_OK;
}

extern "C" int cb_allreduce_counts(cb_ctx *c) {
	DevGuard dev_guard_(c);
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

// =========================================================================== resolver ingestion
// ipaddr.js parse(address).toNormalizedString() (package.json: ipaddr.js >=1.1 <2): IPv4 dotted quad; IPv6 eight
// groups in lowercase hex without leading zeros and without '::' compression.  false: not an IP address.
static bool normalise_address(const char *addr, std::string *out) {
	unsigned char b[16];
	char buf[64];
	if (inet_pton(AF_INET, addr, b) == 1) {
		snprintf(buf, sizeof(buf), "%u.%u.%u.%u", b[0], b[1], b[2], b[3]);
		*out = buf;
		return true;
	}
	if (inet_pton(AF_INET6, addr, b) == 1) {
		std::string r;
		for (int g = 0; g < 8; g++) {
			snprintf(buf, sizeof(buf), g ? ":%x" : "%x", (unsigned)((b[2 * g] << 8) | b[2 * g + 1]));
			r += buf;
		}
		*out = r;
		return true;
	}
	return false;
}

extern "C" int cb_srv_keys(cb_ctx *c, const char *const *names, const int32_t *ports, const char *const *addresses,
    size_t n, char *keys_out) {
	DevGuard dev_guard_(c);
	if (!c || (n && (!names || !ports || !addresses || !keys_out)))
		return c ? fail(c, CB_EINVAL, "names, ports, addresses and keys_out are required") : CB_EINVAL;
	if (n == 0)
		return CB_OK;
	if (n > (1u << 24))
		return fail(c, CB_EINVAL, "too many backends");
	// the hashed bytes: name || "||" || String(port) || "||" || normalised address (lib/resolver.js:1157-1171)
	std::string blob;
	std::vector<unsigned int> offs(n + 1);
	for (size_t k = 0; k < n; k++) {
		offs[k] = (unsigned int)blob.size();
		if (!names[k] || !addresses[k])
			return fail(c, CB_EINVAL, "backend %zu: name and address are required", k);
		std::string norm;
		if (!normalise_address(addresses[k], &norm))
			return fail(c, CB_EINVAL, "backend %zu: '%s' is not an IP address", k, addresses[k]); // ipaddr.parse throws
		blob += names[k];
		blob += "||";
		blob += std::to_string(ports[k]);
		blob += "||";
		blob += norm;
	}
	offs[n] = (unsigned int)blob.size();
	ScopedDev d_blob, d_offs, d_keys;
	CU(c, d_blob.alloc(blob.size() + 16));
	CU(c, d_offs.alloc((n + 1) * sizeof(unsigned int)));
	CU(c, d_keys.alloc(n * SRVKEY_LEN));
	cudaStream_t s = c->stream;
	CU(c, cudaMemcpyAsync(d_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, s));
	CU(c, cudaMemcpyAsync(d_offs.p, offs.data(), (n + 1) * sizeof(unsigned int), cudaMemcpyHostToDevice, s));
	srv_key_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_blob.as<unsigned char>(), d_offs.as<unsigned int>(), (int)n,
	    d_keys.as<char>());
	c->launches++;
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(keys_out, d_keys.p, n * SRVKEY_LEN, cudaMemcpyDeviceToHost, s));
	CU(c, cudaStreamSynchronize(s));
	return CB_OK;
}

extern "C" int cb_resolver_process(cb_ctx *c, const char *keys, size_t n, int32_t *removed_ids, char *removed_keys,
    int32_t *n_removed, int32_t *added_ids, int32_t *added_rows, int32_t *n_added, int32_t *empty_set) {
	DevGuard dev_guard_(c);
	if (!c || (n && !keys) || !n_removed || !n_added)
		return c ? fail(c, CB_EINVAL, "keys, n_removed and n_added are required") : CB_EINVAL;
	*n_removed = 0;
	*n_added = 0;
	if (empty_set)
		*empty_set = 0;
	// newBackends: the keys in first-appearance order (a later duplicate overwrites the value, not the position)
	std::vector<std::string> nk;
	std::vector<int32_t> nrow;
	std::unordered_map<std::string, int32_t> seen;
	for (size_t k = 0; k < n; k++) {
		std::string key(keys + k * SRVKEY_LEN, SRVKEY_LEN);
		auto it = seen.find(key);
		if (it == seen.end()) {
			seen.emplace(key, (int32_t)nk.size());
			nk.push_back(key);
			nrow.push_back((int32_t)k);
		} else {
			nrow[it->second] = (int32_t)k; // the later record wins, lib/resolver.js:1043
		}
	}
	if (nk.empty()) { // 'empty-set': report it and change nothing, lib/resolver.js:1052-1061
		if (empty_set)
			*empty_set = 1;
		return CB_OK;
	}
	int32_t nr = 0, na = 0;
	for (const std::string &k : c->r_keys) { // removed: old key order, :1063-1066
		if (seen.find(k) == seen.end()) {
			if (removed_ids)
				removed_ids[nr] = c->r_ids[k];
			if (removed_keys)
				memcpy(removed_keys + (size_t)nr * SRVKEY_LEN, k.data(), SRVKEY_LEN);
			nr++;
		}
	}
	std::unordered_map<std::string, int32_t> ids;
	for (size_t j = 0; j < nk.size(); j++) { // added: new key order, :1067-1070
		auto it = c->r_ids.find(nk[j]);
		if (it == c->r_ids.end()) {
			const int32_t id = c->r_next_id++;
			ids.emplace(nk[j], id);
			if (added_ids)
				added_ids[na] = id;
			if (added_rows)
				added_rows[na] = nrow[j];
			na++;
		} else {
			ids.emplace(nk[j], it->second);
		}
	}
	c->r_keys.swap(nk); // this.r_backends = newBackends, :1072
	c->r_ids.swap(ids);
	*n_removed = nr;
	*n_added = na;
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	queue_idle_kernel<<<1, QTPB, 0, c->stream>>>(queue_params(c), now_ns, k_slots);
	c->launches += 1;
	return queue_check(c);
}

extern "C" int cb_queue_trim_idle(cb_ctx *c, uint64_t k, uint64_t *removed) {
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	int rc = queue_ready(c);
	if (rc)
		return rc;
	// slots that went away with a topology change leave the idle queue (setUnwanted on idle slots,
	// lib/pool.js:601-631); a bookkeeping word, no kernel
	QueueDev h;
	CU(c, cudaMemcpy(&h, c->d_q, sizeof(h), cudaMemcpyDeviceToHost));
	const unsigned long long take = k < h.idle ? k : h.idle;
	h.idle -= take;
	CU(c, cudaMemcpy((char *)c->d_q + offsetof(QueueDev, idle), &h.idle, sizeof(h.idle), cudaMemcpyHostToDevice));
	if (removed)
		*removed = take;
	return CB_OK;
}

extern "C" int cb_queue_results(cb_ctx *c, int64_t *triples, size_t cap_records, size_t *n_records) {
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
	if (!c)
		return CB_EINVAL;
	int rc = drain(c); // batches in flight were enqueued under the other mode
	if (rc)
		return rc;
	c->prof = on != 0;
	c->prof_scan_only = on == 2;
	for (int k = 0; k < CB_K_COUNT; k++) {
		c->prof_ms[k] = 0;
		c->prof_n[k] = 0;
	}
	return CB_OK;
}
extern "C" int cb_profile_get(cb_ctx *c, double ms_out[CB_K_COUNT], uint64_t launches_out[CB_K_COUNT]) {
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
	DevGuard dev_guard_(c);
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
