// cb_resolve.cuh -- pass 2: resolve the dropping episodes.  Two single-CTA kernels:
//
// prepare_kernel  (independent of the ControlledDelay state -> overlaps freely with the
//                 neighbouring batches of the pipeline)
//   A  pairs every closing tile with the tile that opened the run (a block-wide max-scan over
//      "last tile with a CONST claim") and evaluates the run's unclassified stretch: the class
//      there is monotone in the clock, so its (at most two) crossings are lower_bound searches,
//      one warp each.  This replaces a tile-to-tile look-back chain: a run spanning thousands of
//      tiles costs two searches.  The stretch inherited from the PREVIOUS batch (it depends on
//      the carried first_above_time) is only described, for the commit kernel.
//   B  merges run events and local events into one ordered list (block scan).
//   S  per event, speculates "first claim with now >= now_event + 100" (the drop_next a fresh
//      episode would get, lib/codel.js:48-50,82) with a two-level search.
//
// commit_kernel   (the only part serial across batches: a few microseconds)
//   resolves the carried stretch (if the batch begins inside a run), then runs
//   ControlledDelay#overloaded (lib/codel.js:56-86) over EVENTS, not claims: between two events
//   every claim has the same canDrop class, so a whole stretch is handled in closed form
//       class 0 stretch        : cd_dropping := false, nothing dropped
//       class>0, dropping      : claims with now >= cd_drop_next are dropped (a suffix: the clock
//                                is monotone) and counted                     codel.js:65-68
//       class>0, not dropping  : the state is entered at the stretch's FIRST claim iff
//                                class==2 or now - drop_next < 100             codel.js:70-83
//   and writes the drop intervals plus the new ControlledDelay state.
//
// If the clock went backwards inside the batch (or a tile overflowed its event slots) the
// decomposition does not apply and one warp of the commit kernel runs the exact per-claim state
// machine instead (CB_PATH_SEQUENTIAL), writing the outputs itself.
#pragma once
#include "cb_device.cuh"
#include "cb_scan.cuh"

namespace cb {

constexpr int RTPB = 512; // half an SM's threads: the CTA fits the slot one retiring scan CTA frees

struct ResolveParams {
	const void *start;
	const void *now;
	const uint8_t *flags;
	unsigned long long n;
	unsigned long long tiles;
	DevState *g;
	BatchCtl *bc;
	const unsigned int *rec_fc;         // per-tile records from the scan kernel
	const unsigned long long *rec_code;
	const unsigned int *rec_cend;
	const void *tile_now0;              // raw clock of every tile's first claim (coarse search index)
	double *idx_ms;                     // [<= NIDX_MAX] the coarse index in ms (prepare -> commit)
	ChainLink link;                     // this batch's place in the commit chain
	// cross-tile run events (<= 3 per closing tile; slot `tiles` = the run open at batch end)
	unsigned long long *xev;            // [(tiles+1)*3]
	unsigned int *xcnt;                 // [tiles+1]
	// heavy jobs: runs whose unclassified stretch really crosses a class boundary
	unsigned long long *hj_lo, *hj_b;   // [tiles+1]
	double *hj_F;
	unsigned int *hj_info;              // tile slot | cend0 << 28
	const unsigned int *ev_count;
	const unsigned long long *ev;
	// compacted, ordered events + per-event speculation
	unsigned long long *evc;   // (index<<2)|class
	double *ev_now;            // now[index]
	unsigned long long *ev_q;  // last index of the run the event belongs to
	unsigned long long *ev_d;  // speculated first dropped index after entry at this event
	unsigned long long ev_cap;
	// drop intervals out: [lo,hi] inclusive, cum = drops before the interval
	unsigned long long *iv_lo, *iv_hi, *iv_cum;
	// sequential path outputs
	uint8_t *decision;
	int32_t *slot;
	int32_t *backend;
	const int32_t *slot_backend;
	int nslots;
	// first_above_time hand-over of the F-chain variant of the commit kernel (see FSlot; else null):
	// slots this batch reads / writes (the next rank's mailbox in a ring) and the ring size (>= 1)
	FSlot *f_in, *f_out;
	unsigned long long f_ring;
};

template <bool NS>
__device__ __forceinline__ double now_at(const void *now, unsigned long long i) {
	using raw_t = typename TimeIn<NS>::raw_t;
	return TimeIn<NS>::to_ms(reinterpret_cast<const raw_t *>(now)[i]);
}

// Monotone predicates on the clock: level 0: now >= x (drop_next test, codel.js:65);
// level 1/2: canDrop class w.r.t. first_above_time x is >= level.
struct Pred {
	double x;
	int level;
	__device__ __forceinline__ bool operator()(double now) const {
		return level == 0 ? (now >= x) : (claim_class(true, now, x) >= level);
	}
};

// First index in [lo,hi) whose clock satisfies the (monotone) predicate, else hi.
// Two levels: a coarse index of the clock at every G-th claim lives in shared memory (no DRAM
// round trip), then a 64-ary warp-cooperative search over at most G claims: 2 dependent DRAM
// accesses for G = 2048 instead of 5 for a plain 32-ary search over 2^24 claims.
template <bool NS>
__device__ unsigned long long warp_search(const void *now, const double *s_idx, unsigned long long G,
    unsigned long long lo, unsigned long long hi, const Pred pr, int lane) {
	if (hi <= lo)
		return hi;
	{
		const unsigned long long ka = lo / G + 1, kb = (hi - 1) / G; // coarse boundaries inside (lo, hi-1]
		if (ka <= kb) {
			unsigned long long a = ka, b = kb + 1;
			while (a < b) { // first boundary whose clock satisfies the predicate (smem, all lanes alike)
				const unsigned long long mid = (a + b) >> 1;
				if (pr(s_idx[mid]))
					b = mid;
				else
					a = mid + 1;
			}
			if (a > ka)
				lo = (a - 1) * G + 1; // boundary a-1 fails: the answer lies beyond it
			if (a <= kb)
				hi = a * G; // claim a*G satisfies: the answer is at most a*G (= the fallback `hi`)
		}
	}
	while (true) {
		const unsigned long long s = hi - lo;
		if (s == 0)
			return hi;
		if (s <= 64) {
			const bool p0 = (lane < s) && pr(now_at<NS>(now, lo + lane));
			const bool p1 = (lane + 32 < s) && pr(now_at<NS>(now, lo + 32 + lane));
			const unsigned int m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
			if (m0)
				return lo + (__ffs(m0) - 1);
			if (m1)
				return lo + 32 + (__ffs(m1) - 1);
			return hi;
		}
		const unsigned long long chunk = (s + 63) / 64;
		unsigned long long i0 = lo + (unsigned long long)(lane + 1) * chunk - 1;
		unsigned long long i1 = lo + (unsigned long long)(lane + 33) * chunk - 1;
		if (i0 > hi - 1)
			i0 = hi - 1;
		if (i1 > hi - 1)
			i1 = hi - 1;
		const bool p0 = pr(now_at<NS>(now, i0)), p1 = pr(now_at<NS>(now, i1));
		const unsigned int m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
		if (!(m0 | m1))
			return hi;
		// first satisfying probe j in 0..63 and the probe before it
		const int j = m0 ? (__ffs(m0) - 1) : (32 + __ffs(m1) - 1);
		const unsigned long long pj = __shfl_sync(0xffffffffu, (j < 32) ? i0 : i1, j & 31);
		if (j > 0) {
			const int jp = j - 1;
			const unsigned long long pp = __shfl_sync(0xffffffffu, (jp < 32) ? i0 : i1, jp & 31);
			lo = pp + 1;
		}
		hi = pj; // pj itself satisfies the predicate
		if (lo >= hi)
			return hi;
	}
}
// single-thread variant (rare slow path inside the serial chain)
template <bool NS>
__device__ unsigned long long thread_lower_bound(const void *now, unsigned long long lo,
    unsigned long long hi, double D) {
	while (lo < hi) {
		unsigned long long mid = lo + (hi - lo) / 2;
		if (now_at<NS>(now, mid) >= D)
			hi = mid;
		else
			lo = mid + 1;
	}
	return lo;
}

// ---- exact per-claim state machine (one warp) ------------------------------------------
template <bool NS>
__device__ void sequential_path(const ResolveParams &p, int lane) {
	using raw_t = typename TimeIn<NS>::raw_t;
	DevState *g = p.g;
	BatchCtl *bc = p.bc;
	CodelRegs s;
	s.F = g->F;
	s.D = g->D;
	s.C = g->C;
	s.B = g->B;
	const double target = g->target;
	double E = g->E;
	int hasE = g->hasE;
	unsigned long long cursor = g->cursor, drops = 0, admits = 0, skipped = 0;
	const raw_t *gs = reinterpret_cast<const raw_t *>(p.start);
	const raw_t *gn = reinterpret_cast<const raw_t *>(p.now);
	for (unsigned long long base = 0; base < p.n; base += 32) {
		unsigned long long i = base + lane;
		bool valid = i < p.n;
		double st = valid ? TimeIn<NS>::to_ms(gs[i]) : 0.0;
		double nw = valid ? TimeIn<NS>::to_ms(gn[i]) : 0.0;
		unsigned int fl = valid ? (p.flags ? p.flags[i] : F_WAITING) : 0;
		int cnt = (p.n - base) < 32 ? (int)(p.n - base) : 32;
		uint8_t mydec = 0;
		int32_t myslot = -1;
		for (int j = 0; j < cnt; j++) {
			double nj = __shfl_sync(0xffffffffu, nw, j);
			double sj = __shfl_sync(0xffffffffu, st, j);
			unsigned int fj = __shfl_sync(0xffffffffu, fl, j);
			bool drop = codel_step(s, target, nj, sj); // lib/pool.js:735-736
			uint8_t d;
			int32_t sl = -1;
			if (!(fj & F_WAITING)) { // lib/pool.js:738-740
				d = 2;
				skipped++;
			} else if (drop) { // lib/pool.js:742-745
				d = 1;
				drops++;
			} else { // lib/pool.js:747
				d = 0;
				if (p.nslots > 0)
					sl = (int32_t)(cursor % (unsigned long long)p.nslots);
				cursor++;
				admits++;
			}
			if (fj & F_EMPTY_AFTER) { // lib/pool.js:751-753 -> lib/codel.js:91-94
				E = nj;
				hasE = 1;
				s.F = 0.0;
			}
			if (lane == j) {
				mydec = d;
				myslot = sl;
			}
		}
		if (valid) {
			if (p.decision)
				p.decision[i] = mydec;
			if (p.slot)
				p.slot[i] = myslot;
			if (p.backend)
				p.backend[i] = (myslot >= 0) ? p.slot_backend[myslot] : -1;
		}
	}
	if (lane == 0) {
		g->F = s.F;
		g->D = s.D;
		g->C = s.C;
		g->B = s.B;
		g->E = E;
		g->hasE = hasE;
		bc->cursor0 = g->cursor;
		g->cursor = cursor;
		bc->n = p.n;
		bc->drops = drops;
		bc->admits = admits;
		bc->skipped = skipped;
		g->tot_drops += drops;
		g->tot_admits += admits;
		bc->gcounts[0] = drops;
		bc->gcounts[1] = admits;
		bc->gcounts[2] = p.n;
		bc->gcounts[3] = 0;
		bc->path = 1;
		bc->n_events = 0;
		bc->n_intervals = 0;
		bc->nonmono = 0;
		bc->overflow = 0;
		bc->last_marker_p1 = 0;
	}
}

constexpr int RT = 8;          // tile slots per thread per pass (8192 per pass)
constexpr int NIDX_MAX = 8192; // coarse clock index entries kept in shared memory (64 KB)
constexpr int SEV_MAX = 1024;  // events kept in shared memory for the serial chain (32 KB)

constexpr int EV_FRONT = 3;    // slots kept free in front of the event list for the carried run's events
constexpr int CTPB = 128;      // threads of the commit kernel

#define CB_PREPARE_HEAD                                                                                    \
	template <bool NS, bool FLAGS> __global__ void __launch_bounds__(RTPB, 1) prepare_kernel(const ResolveParams p)
#include "cb_prepare_body.inc"
#undef CB_PREPARE_HEAD
#define CB_PREPARE_HEAD                                                                                    \
	template <bool NS, bool FLAGS> __device__ __forceinline__ void prepare_part(const ResolveParams &p)
#include "cb_prepare_body.inc"
#undef CB_PREPARE_HEAD

#define CB_COMMIT_HEAD                                                                                     \
	template <bool NS, bool FLAGS> __global__ void __launch_bounds__(CTPB, 1) commit_kernel(const ResolveParams p)
#define CB_COMMIT_FCHAIN false
#include "cb_commit_body.inc"
#undef CB_COMMIT_HEAD
#undef CB_COMMIT_FCHAIN
#define CB_COMMIT_HEAD                                                                                     \
	template <bool NS, bool FLAGS> __global__ void __launch_bounds__(CTPB, 1) commit_fchain_kernel(const ResolveParams p)
#define CB_COMMIT_FCHAIN true
#include "cb_commit_body.inc"
#undef CB_COMMIT_HEAD
#undef CB_COMMIT_FCHAIN
#define CB_COMMIT_HEAD                                                                                     \
	template <bool NS, bool FLAGS, bool FCHAIN_T> __device__ __forceinline__ void commit_part(const ResolveParams &p)
#define CB_COMMIT_FCHAIN FCHAIN_T
#include "cb_commit_body.inc"
#undef CB_COMMIT_HEAD
#undef CB_COMMIT_FCHAIN

// Prepare and commit in ONE launch (512 threads): the state-independent half runs, then the same CTA
// waits on the commit chain and finishes.  With the commits ordered on the device (chain word) a
// kernel boundary between the two halves buys nothing, and it costs a launch on the host (the 1M-claim
// step is launch-bound) and a launch latency on the device.  NOT yet validated on hardware: selected
// only by CB_FUSED=1.
template <bool NS, bool FLAGS, bool FCHAIN_T>
__global__ void __launch_bounds__(RTPB, 1) resolve_fused_kernel(const ResolveParams p) {
	prepare_part<NS, FLAGS>(p);
	__syncthreads(); // prepare's global writes (events, control words) are visible to the whole CTA
	commit_part<NS, FLAGS, FCHAIN_T>(p);
}

// one thread: seed the slot the first batch after a host-side state change reads
__global__ void fchain_seed_kernel(const DevState *g, FSlot *f_in, unsigned long long seq, unsigned long long ring) {
	FSlot *prev = &f_in[(seq / ring) % FSLOTS];
	prev->F = g->F;
	__threadfence();
	st_release(&prev->tag, seq);
}

// ---- scalar entry points (one claim): ControlledDelay#overloaded / #empty ---------------
__global__ void overloaded_one_kernel(DevState *g, long long start_ns, long long now_ns, int *out) {
	CodelRegs s;
	s.F = g->F;
	s.D = g->D;
	s.C = g->C;
	s.B = g->B;
	bool drop = codel_step(s, g->target, ns_to_ms(now_ns), ns_to_ms(start_ns));
	g->F = s.F;
	g->D = s.D;
	g->C = s.C;
	g->B = s.B;
	*out = drop ? 1 : 0;
}
__global__ void empty_kernel(DevState *g, long long now_ns) { // lib/codel.js:91-94
	g->E = ns_to_ms(now_ns);
	g->hasE = 1;
	g->F = 0.0;
}
__global__ void max_idle_kernel(const DevState *g, long long now_ns, double *out) { // lib/codel.js:109-118
	double bound = __dmul_rn(g->target, 10.0);
	double now = ns_to_ms(now_ns);
	if (g->hasE && g->E < __dsub_rn(now, bound))
		*out = __dmul_rn(g->target, 3.0);
	else
		*out = bound;
}

} // namespace cb
