// cb_resolve.cuh -- pass 2: resolve the dropping episodes.  One single-CTA kernel, two halves:
//
// prepare_part  (independent of the ControlledDelay state -> overlaps freely with the
//               neighbouring batches of the pipeline)
//   A  pairs every closing tile with the tile that opened the run (a block-wide max-scan over
//      "last tile with a CONST claim") and evaluates the run's unclassified stretch: the class
//      there is monotone in the clock, so its (at most two) crossings are lower_bound searches,
//      one warp each.  This replaces a tile-to-tile look-back chain: a run spanning thousands of
//      tiles costs two searches.  The stretch inherited from the PREVIOUS batch (it depends on
//      the carried first_above_time) is only described, for commit_part.
//   B  merges run events and local events into one ordered list (block scan).
//   S  per event, speculates "first claim with now >= now_event + 100" (the drop_next a fresh
//      episode would get, lib/codel.js:48-50,82) with a two-level search.
//
// commit_part   (the only part serial across batches: a few microseconds)
//   F hand-over: first_above_time evolves independently of (drop_next, count, dropping), and a batch
//   that contains any below-target claim fixes its outgoing value without knowing the incoming one.
//   The batch publishes its outgoing value first (FSlot), takes the incoming one, and resolves the
//   stretch inherited from the previous batch BEFORE the rest of the state arrives.  Then it waits
//   for the previous batch's commit (chain word / ring mailbox), reads the state once (thread 0,
//   broadcast through shared memory) and runs ControlledDelay#overloaded (lib/codel.js:56-86) over
//   EVENTS, not claims: between two events every claim has the same canDrop class, so a whole
//   stretch is handled in closed form
//       class 0 stretch        : cd_dropping := false, nothing dropped
//       class>0, dropping      : claims with now >= cd_drop_next are dropped (a suffix: the clock
//                                is monotone) and counted                     codel.js:65-68
//       class>0, not dropping  : the state is entered at the stretch's FIRST claim iff
//                                class==2 or now - drop_next < 100             codel.js:70-83
//   and writes the drop intervals plus the new ControlledDelay state.
//
// If the clock went backwards inside the batch (or a tile overflowed its event slots) the
// decomposition does not apply and one warp runs the exact per-claim state machine instead
// (CB_PATH_SEQUENTIAL), writing the outputs itself.
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
		bc->lcounts[0] = drops;
		bc->lcounts[1] = admits;
		bc->lcounts[2] = p.n;
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

// ---- first half: everything that does not depend on the ControlledDelay state ---------------------
template <bool NS, bool FLAGS> __device__ __forceinline__ void prepare_part(const ResolveParams &p) {
	CB_DYN_SHARED(double, s_idx); // clock (ms) at claim k*G, k < nidx
	__shared__ unsigned int s_warp[32];
	__shared__ int s_wmax[32];
	__shared__ unsigned int s_base;
	__shared__ unsigned int s_nev;
	__shared__ int s_fallback;
	__shared__ long long s_prevconst; // last tile with a CONST claim seen so far (-1: none)
	__shared__ unsigned int s_nheavy;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	BatchCtl *bc = p.bc;
	using raw_t = typename TimeIn<NS>::raw_t;
	if (tid == 0)
		bc->dbg[6] = gtimer();

	// coarse clock index: every `stride`-th tile's first claim (loaded while thread 0 checks
	// whether the scan kernel asked for the exact sequential path)
	const unsigned long long stride = (p.tiles + NIDX_MAX - 1) / NIDX_MAX;
	const unsigned long long G = stride * TILE;
	const unsigned int nidx = (unsigned int)((p.tiles + stride - 1) / stride);
	if (tid == 0) {
		s_fallback = (bc->nonmono != 0u) || (bc->overflow != 0u);
		s_base = 0;
		s_prevconst = -1;
		s_nheavy = 0;
	}
	for (unsigned int k = tid; k < nidx; k += RTPB) {
		const double v = TimeIn<NS>::to_ms(reinterpret_cast<const raw_t *>(p.tile_now0)[(unsigned long long)k * stride]);
		s_idx[k] = v; // commit_part searches the same index
	}
	__syncthreads();
	if (s_fallback)
		return; // the commit kernel runs the exact sequential path
	const unsigned long long carry_code = CODE_CARRY; // the carried first_above_time is the commit kernel's business
	if (tid == 0) {
		bc->cj_exists = 0;
		bc->cj_close = 0;
		bc->cj_closes = 0;
	}

	if (tid == 0) bc->dbg[0] = gtimer();
	// ---- A. pair closing tiles with the run they close; evaluate the inherited stretches ------
	// One pass handles RT consecutive tiles per thread with every load of a stage issued
	// together, so the pass costs a handful of dependent memory round trips, not RT of them.
	// (slot `tiles` is a virtual closing tile at index n: the run still open at the end of the batch)
	for (unsigned long long base = 0; base < p.tiles + 1; base += (unsigned long long)RTPB * RT) {
		const unsigned long long t0 = base + (unsigned long long)tid * RT;
		unsigned int fc[RT];
		int T0l[RT]; // block-local slot of the opener, -1 = before this thread, resolved below
		int lastc = -1;
#pragma unroll
		for (int k = 0; k < RT; k++)
			fc[k] = (t0 + k < p.tiles) ? p.rec_fc[t0 + k] : ((t0 + k == p.tiles) ? 0u : FC_NONE);
#pragma unroll
		for (int k = 0; k < RT; k++) {
			T0l[k] = lastc;
			if (fc[k] != FC_NONE && t0 + k < p.tiles)
				lastc = tid * RT + k;
		}
		int x = lastc;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o)
				x = max(x, y);
		}
		if (lane == 31)
			s_wmax[warp] = x;
		__syncthreads();
		int wprev = -1;
		for (int w = 0; w < warp; w++)
			wprev = max(wprev, s_wmax[w]);
		int ex = __shfl_up_sync(0xffffffffu, x, 1);
		if (lane == 0)
			ex = -1;
		ex = max(ex, wprev);
		const long long Tbefore = (ex >= 0) ? (long long)(base + ex) : s_prevconst;
		// stage 1: the opener's record
		long long T0[RT];
		unsigned long long code0[RT];
		unsigned int cend0[RT];
#pragma unroll
		for (int k = 0; k < RT; k++) {
			T0[k] = (T0l[k] >= 0) ? (long long)(base + T0l[k]) : Tbefore;
			code0[k] = carry_code;
			cend0[k] = 0;
			if (fc[k] != FC_NONE && T0[k] >= 0) {
				code0[k] = p.rec_code[T0[k]];
				cend0[k] = p.rec_cend[T0[k]];
			}
		}
		// stage 2: clock at the run's defining claim and at the stretch's last claim
		double Fv[RT], nb[RT];
		bool stretch[RT];
#pragma unroll
		for (int k = 0; k < RT; k++) {
			const unsigned long long lo = (unsigned long long)(T0[k] + 1) * TILE;
			const unsigned long long close_idx = (t0 + k == p.tiles) ? p.n : (t0 + k) * TILE + fc[k];
			stretch[k] = (fc[k] != FC_NONE) && (lo < close_idx);
			Fv[k] = 0.0;
			nb[k] = 0.0;
			if (stretch[k] && code0[k] == CODE_CARRY) { // inherited from the previous batch: commit's job
				stretch[k] = false;
				bc->cj_exists = 1;
				bc->cj_close = close_idx;
				bc->cj_closes = (t0 + k == p.tiles) ? 0u : 1u;
				code0[k] = CODE_ZERO;
				cend0[k] = 4u; // marks "carried": no close event from here either
			}
			if (stretch[k]) {
				Fv[k] = __dadd_rn(now_at<NS>(p.now, code0[k] == CODE_ZERO ? lo : code0[k] - 2), CODEL_INTERVAL);
				nb[k] = now_at<NS>(p.now, close_idx - 1);
			}
		}
		// stage 3: does the class rise inside the stretch?  close event?
#pragma unroll
		for (int k = 0; k < RT; k++) {
			if (t0 + k > p.tiles)
				continue;
			const bool is_tail = (t0 + k == p.tiles);
			unsigned int bits = 0;
			if (fc[k] != FC_NONE) {
				const unsigned long long close_idx = is_tail ? p.n : (t0 + k) * TILE + fc[k];
				int final_c = (int)cend0[k];
				if (stretch[k]) {
					const int cb = claim_class(true, nb[k], Fv[k]);
					if (cb > final_c) { // find where (two warps per such run, below)
						const unsigned int h = atomicAdd(&s_nheavy, 1u);
						p.hj_lo[h] = (unsigned long long)(T0[k] + 1) * TILE;
						p.hj_b[h] = close_idx - 1;
						p.hj_F[h] = Fv[k];
						p.hj_info[h] = (unsigned int)(t0 + k) | ((cend0[k] & 3u) << 28) | ((unsigned int)cb << 30);
						final_c = cb;
					}
				}
				if (final_c >= 1 && final_c < 4 && !is_tail) { // the run ends at a CONST claim (class 0)
					p.xev[(t0 + k) * 3 + 2] = (close_idx << 2);
					bits = 4u;
				}
			}
			p.xcnt[t0 + k] = bits;
		}
		__syncthreads();
		if (tid == RTPB - 1) {
			const int blockmax = max(x, wprev);
			if (blockmax >= 0)
				s_prevconst = (long long)(base + blockmax);
		}
		__syncthreads();
	}
	if (tid == 0) bc->dbg[1] = gtimer();
	// long runs: the two possible class crossings are independent searches -> one warp each
	const unsigned int nheavy = s_nheavy;
	for (unsigned int sub = warp; sub < 2 * nheavy; sub += RTPB / 32) {
		const unsigned int h = sub >> 1, level = 1 + (sub & 1u);
		const unsigned int info = p.hj_info[h];
		const unsigned int slot = info & 0x0fffffffu;
		const int cend0 = (int)((info >> 28) & 3u), cb = (int)(info >> 30);
		if ((level == 1 && cend0 != 0) || (level == 2 && cb != 2))
			continue; // that crossing does not exist in this stretch
		const unsigned long long lo = p.hj_lo[h], b = p.hj_b[h];
		const double F = p.hj_F[h];
		Pred pr;
		pr.x = F;
		pr.level = (int)level;
		const unsigned long long pos = warp_search<NS>(p.now, s_idx, G, lo, b + 1, pr, lane); // exists: class(b) >= level
		if (level == 1) {
			const int c = claim_class(true, now_at<NS>(p.now, pos), F); // 1, or 2 when both crossings coincide
			if (lane == 0) {
				p.xev[slot * 3 + 0] = (pos << 2) | (unsigned long long)c;
				atomicOr(&p.xcnt[slot], 1u);
			}
		} else {
			// a separate event only if the class was already 1 just before it
			bool separate = (cend0 == 1);
			if (cend0 == 0 && pos > lo)
				separate = claim_class(true, now_at<NS>(p.now, pos - 1), F) >= 1;
			if (lane == 0 && separate) {
				p.xev[slot * 3 + 1] = (pos << 2) | 2ull;
				atomicOr(&p.xcnt[slot], 2u);
			}
		}
	}
	__syncthreads();

	if (tid == 0) bc->dbg[2] = gtimer();
	// ---- B. merge run events and local events into one ordered list -----------------------------
	for (unsigned long long base = 0; base < p.tiles + 1; base += (unsigned long long)RTPB * RT) {
		const unsigned long long t0 = base + (unsigned long long)tid * RT;
		unsigned int xb[RT], lc[RT], cnt = 0;
#pragma unroll
		for (int k = 0; k < RT; k++) {
			const bool in = t0 + k < p.tiles;
			xb[k] = (t0 + k <= p.tiles) ? p.xcnt[t0 + k] : 0u;
			lc[k] = in ? p.ev_count[t0 + k] : 0u;
			cnt += __popc(xb[k]) + lc[k];
		}
		unsigned int x = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o)
				x += y;
		}
		if (lane == 31)
			s_warp[warp] = x;
		__syncthreads();
		if (warp == 0) {
			const unsigned int w = lane < RTPB / 32 ? s_warp[lane] : 0u;
			unsigned int ws = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned int y = __shfl_up_sync(0xffffffffu, ws, o);
				if (lane >= o)
					ws += y;
			}
			s_warp[lane] = ws - w; // exclusive
			if (lane == 31)
				s_nev = ws;        // pass total
		}
		__syncthreads();
		unsigned int off = s_base + s_warp[warp] + (x - cnt);
		if (cnt) {
			for (int k = 0; k < RT; k++) {
				const unsigned long long t = t0 + k;
				// run events of the tile's close come first (crossings <= close index <= local events)
				if ((xb[k] & 1u) && off + EV_FRONT < p.ev_cap)
					p.evc[EV_FRONT + off++] = p.xev[t * 3 + 0];
				if ((xb[k] & 2u) && off + EV_FRONT < p.ev_cap)
					p.evc[EV_FRONT + off++] = p.xev[t * 3 + 1];
				if ((xb[k] & 4u) && off + EV_FRONT < p.ev_cap)
					p.evc[EV_FRONT + off++] = p.xev[t * 3 + 2];
				if (lc[k]) {
					// a tile's events were appended unordered: sort its <= EVK entries by index
					unsigned long long e[EVK];
#pragma unroll
					for (int j = 0; j < EVK; j++)
						e[j] = (j < (int)lc[k]) ? p.ev[t * EVK + j] : ~0ull;
#pragma unroll
					for (int a = 1; a < EVK; a++) {
#pragma unroll
						for (int b = EVK - 1; b >= 1; b--) {
							if (b >= a && e[b] < e[b - 1]) {
								const unsigned long long tmp = e[b];
								e[b] = e[b - 1];
								e[b - 1] = tmp;
							}
						}
					}
#pragma unroll
					for (int j = 0; j < EVK; j++)
						if (j < (int)lc[k] && off + EV_FRONT < p.ev_cap)
							p.evc[EV_FRONT + off++] = e[j];
				}
			}
		}
		__syncthreads();
		if (tid == 0)
			s_base += s_nev;
		__syncthreads();
	}
	const unsigned int nev = s_base;
	if (nev + EV_FRONT > p.ev_cap) // cannot happen with ev_cap = (tiles+1)*(EVK+3)+EV_FRONT
		return;

	if (tid == 0) bc->dbg[3] = gtimer();
	// ---- 2. per event (one warp each): run end q, now at the event, speculated drop start
	unsigned long long *evc = p.evc + EV_FRONT;
	unsigned long long *ev_q = p.ev_q + EV_FRONT;
	unsigned long long *ev_d = p.ev_d + EV_FRONT;
	double *ev_now = p.ev_now + EV_FRONT;
	for (unsigned int k = warp; k < nev; k += RTPB / 32) {
		const unsigned long long w = evc[k];
		const unsigned long long w1 = (k + 1 < nev) ? evc[k + 1] : 3ull, w2 = (k + 2 < nev) ? evc[k + 2] : 3ull;
		const int c = (int)(w & 3ull);
		const unsigned long long idx = w >> 2;
		if (c == 0)
			continue;
		// the run ends right before the next class-0 event (at most two events ahead)
		unsigned long long q = p.n - 1;
		if ((w1 & 3ull) == 0)
			q = (w1 >> 2) - 1;
		else if ((w1 & 3ull) != 3ull && (w2 & 3ull) == 0)
			q = (w2 >> 2) - 1;
		const double nowk = now_at<NS>(p.now, idx);
		unsigned long long d;
		Pred pr;
		pr.level = 0;
		pr.x = __dadd_rn(nowk, CODEL_INTERVAL);
		d = warp_search<NS>(p.now, s_idx, G, idx + 1, q + 1, pr, lane);
		if (lane == 0) {
			ev_now[k] = nowk;
			ev_q[k] = q;
			ev_d[k] = d;
		}
	}
	__syncthreads();

	if (tid == 0) bc->dbg[4] = gtimer();
	// ---- hand-over to the commit kernel ---------------------------------------------------------
	if (tid == 0) {
		bc->nev_prep = nev;
		// F after the batch's last claim when some tile fixes it inside the batch: the last tile with
		// a CONST claim decides, later (all PASS) tiles inherit or, if it is 0, the first of them
		// opens the run
		const long long Tl = s_prevconst;
		bc->anchored = (Tl >= 0) ? 1u : 0u;
		double F = 0.0;
		if (Tl >= 0) {
			const unsigned long long code0 = p.rec_code[Tl];
			if (code0 != CODE_ZERO)
				F = __dadd_rn(now_at<NS>(p.now, code0 - 2), CODEL_INTERVAL);
			else if ((unsigned long long)(Tl + 1) < p.tiles)
				F = __dadd_rn(now_at<NS>(p.now, (unsigned long long)(Tl + 1) * TILE), CODEL_INTERVAL);
		}
		bc->F_anchor = F;
		bc->now0p100 = __dadd_rn(now_at<NS>(p.now, 0), CODEL_INTERVAL);
		bc->hasE_new = 0;
		bc->trailing_marker = 0;
		if (FLAGS) {
			const unsigned long long lm = bc->last_marker_p1;
			if (lm) {
				bc->E_val = now_at<NS>(p.now, lm - 1); // lib/codel.js:92
				bc->hasE_new = 1;
				bc->trailing_marker = (lm == p.n) ? 1u : 0u;
			}
			bc->last_marker_p1 = 0;
		}
		bc->dbg[4] = gtimer();
	}
}

// ---- second half: everything that depends on the ControlledDelay state carried from the previous
// batch; the commits of consecutive batches are ordered on the device (chain word / ring mailbox)
template <bool NS, bool FLAGS> __device__ __forceinline__ void commit_part(const ResolveParams &p) {
	CB_DYN_SHARED(double, s_cidx); // the coarse clock index prepare_part left in shared memory
	__shared__ unsigned long long s_evc[SEV_MAX], s_evq[SEV_MAX], s_evd[SEV_MAX];
	__shared__ double s_evnow[SEV_MAX];
	__shared__ unsigned int s_ncar;
	__shared__ double s_st_d[2];             // the carried state, read once by thread 0 and broadcast
	__shared__ unsigned long long s_st_u[3];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	DevState *g = p.g;
	BatchCtl *bc = p.bc;
	if (tid == 0)
		bc->dbg[7] = gtimer();
	const bool fallback = (bc->nonmono != 0u) || (bc->overflow != 0u);

	// ---- everything that does not depend on the carried state: done while the previous batch's
	// commit may still be running --------------------------------------------------------------
	const unsigned int nevp = fallback ? 0u : bc->nev_prep;
	const bool ev_smem = nevp + EV_FRONT <= SEV_MAX;
	unsigned long long *evc = ev_smem ? s_evc : p.evc;
	unsigned long long *ev_q = ev_smem ? s_evq : p.ev_q;
	unsigned long long *ev_d = ev_smem ? s_evd : p.ev_d;
	double *ev_now = ev_smem ? s_evnow : p.ev_now;
	const unsigned long long stride = (p.tiles + NIDX_MAX - 1) / NIDX_MAX;
	const unsigned long long G = stride * TILE;
	if (!fallback) {
		if (ev_smem) {
			for (unsigned int k = tid; k < nevp; k += RTPB) {
				s_evc[EV_FRONT + k] = p.evc[EV_FRONT + k];
				s_evq[EV_FRONT + k] = p.ev_q[EV_FRONT + k];
				s_evd[EV_FRONT + k] = p.ev_d[EV_FRONT + k];
				s_evnow[EV_FRONT + k] = p.ev_now[EV_FRONT + k];
			}
		}
	}
	// every state-independent word the serial part reads, fetched now (one round trip, overlapped
	// with the wait below) instead of one dependent access at a time afterwards
	const unsigned int cj_exists = fallback ? 0u : bc->cj_exists, cj_closes = bc->cj_closes, anchored = bc->anchored;
	const unsigned long long cj_close = bc->cj_close;
	const double now0p100 = bc->now0p100, F_anchor = bc->F_anchor;
	const unsigned int hasE_new = bc->hasE_new, trailing_marker = bc->trailing_marker;
	const double E_val = bc->E_val;
	double now_before_close = 0.0; // clock of the last claim of the inherited stretch
	if (cj_exists && cj_close > 0)
		now_before_close = now_at<NS>(p.now, cj_close - 1);
	{
		// ---- F hand-over (see FSlot): publish this batch's outgoing first_above_time as soon as it is
		// known, take the incoming one, and resolve the inherited stretch with it -- all before the
		// ControlledDelay state arrives
		const unsigned long long seq = p.link.wait_val;
		FSlot *mine = &p.f_out[((seq + 1) / p.f_ring) % FSLOTS]; // read by the batch at queue position seq + 1
		const FSlot *prev = &p.f_in[(seq / p.f_ring) % FSLOTS];
		const bool reset_after = FLAGS && hasE_new && trailing_marker; // trailing empty(): F = 0
		__shared__ double s_Fin;
		if (tid == 0) {
			s_ncar = 0;
			int pub = 0;
			if (!fallback && (anchored || reset_after)) {
				mine->F = reset_after ? 0.0 : F_anchor;
				__threadfence_system(); // the slot may live in the next rank's mailbox (peer memory)
				st_release_sys(&mine->tag, seq + 1);
				pub = 1;
			}
			const unsigned long long t0 = gtimer();
			while (ld_acquire_sys(&prev->tag) != seq) {
				if (gtimer() - t0 > p.link.timeout_ns) {
					atomicAdd(p.link.err, 1ull);
					break;
				}
			}
			const double Fi = ld_relaxed_f64(&prev->F);
			if (!fallback && !pub) { // one above-target run end to end: F passes through (or claim 0 opens it)
				mine->F = (Fi != 0.0) ? Fi : now0p100;
				__threadfence_system();
				st_release_sys(&mine->tag, seq + 1);
				pub = 1;
			}
			s_Fin = Fi;
		}
		__syncthreads();
		if (cj_exists) { // same steps as the serial version below, with F known early
			const double Fe = s_Fin;
			const unsigned long long close = cj_close;
			const double F = (Fe != 0.0) ? Fe : now0p100;
			const int cb = claim_class(true, now_before_close, F);
			if (cb >= 1 && warp < 2 && (warp == 0 || cb == 2)) {
				Pred pr;
				pr.x = F;
				pr.level = warp + 1;
				const unsigned long long pos = warp_search<NS>(p.now, s_cidx, G, 0, close, pr, lane);
				if (warp == 0) {
					const int c = claim_class(true, now_at<NS>(p.now, pos), F);
					if (lane == 0)
						evc[0] = (pos << 2) | (unsigned long long)c;
				} else {
					bool separate = pos > 0 && claim_class(true, now_at<NS>(p.now, pos - 1), F) >= 1;
					if (lane == 0)
						evc[1] = separate ? ((pos << 2) | 2ull) : ~0ull;
				}
			}
			__syncthreads();
			if (tid == 0) {
				unsigned long long tmp[3];
				unsigned int nc = 0;
				if (cb >= 1) {
					tmp[nc++] = evc[0];
					if (cb == 2 && evc[1] != ~0ull)
						tmp[nc++] = evc[1];
					if (cj_closes)
						tmp[nc++] = (close << 2);
				}
				for (unsigned int k = 0; k < nc; k++)
					evc[EV_FRONT - nc + k] = tmp[k];
				s_ncar = nc;
			}
			__syncthreads();
			const unsigned int nc = s_ncar;
			if ((unsigned int)warp < nc) { // speculation as if not dropping; the dropping case is redone below
				const unsigned int k = EV_FRONT - nc + warp;
				const unsigned long long w = evc[k];
				if ((w & 3ull) != 0) {
					const unsigned long long idx = w >> 2, q = close - 1;
					const double nowk = now_at<NS>(p.now, idx);
					Pred pr;
					pr.level = 0;
					pr.x = __dadd_rn(nowk, CODEL_INTERVAL);
					const unsigned long long d = warp_search<NS>(p.now, s_cidx, G, idx + 1, q + 1, pr, lane);
					if (lane == 0) {
						ev_now[k] = nowk;
						ev_q[k] = q;
						ev_d[k] = d;
					}
				}
			}
			__syncthreads();
		}
	}
	if (tid == 0) {
		// ---- hand-over: wait until the previous batch has committed its state.  It was launched
		// before this one (possibly on another stream), so it is running or about to; the bound only
		// keeps a broken chain from hanging the device.
		chain_wait(p.link, g);
		bc->dbg[12] = gtimer();
		// hand the state to the other threads through shared memory: nobody but this thread reads
		// DevState, which lane 0 overwrites at the end of the chain
		s_st_d[0] = g->F;
		s_st_d[1] = g->D;
		s_st_u[0] = g->C;
		s_st_u[1] = g->cursor;
		s_st_u[2] = (unsigned long long)g->B;
	}
	__syncthreads();
	if (fallback) {
		if (warp == 0) {
			sequential_path<NS>(p, lane);
			__syncwarp();
			if (lane == 0) {
				// the exact path fixes first_above_time only now
				FSlot *mine = &p.f_out[((p.link.wait_val + 1) / p.f_ring) % FSLOTS];
				mine->F = g->F;
				__threadfence_system();
				st_release_sys(&mine->tag, p.link.wait_val + 1);
				chain_post(p.link, g);
			}
		}
		return;
	}
	const double F_in = s_st_d[0], D_in = s_st_d[1];
	const int B_in = (int)s_st_u[2];
	const unsigned long long C_in = s_st_u[0], cursor_in = s_st_u[1];

	{
		// the inherited stretch was resolved before the wait; only a batch that opens inside a dropping
		// episode needs its first event's drop start again (drop_next comes with the state)
		if (cj_exists && B_in && s_ncar > 0 && warp == 0) {
			const unsigned int k = EV_FRONT - s_ncar;
			const unsigned long long w = evc[k];
			if ((w & 3ull) != 0 && (w >> 2) == 0) {
				Pred pr;
				pr.level = 0;
				pr.x = D_in;
				const unsigned long long d = warp_search<NS>(p.now, s_cidx, G, 0, cj_close, pr, lane);
				if (lane == 0)
					ev_d[k] = d;
			}
		}
		__syncthreads();
	}
	// ---- 3. the serial chain over events: warp 0, every lane carrying the same scalars so that
	// the searches a failed speculation needs are warp-wide (3-4 dependent accesses, not 20+) ----
	if (warp == 0) {
		const bool wr = (lane == 0);
		double D = D_in;
		unsigned long long C = C_in;
		int B = B_in;
		unsigned long long drops = 0;
		unsigned int niv = 0;
		const unsigned int k0 = EV_FRONT - s_ncar, nev = EV_FRONT + nevp;
		if (k0 == nev || (evc[k0] >> 2) > 0)
			B = 0; // the batch starts with class-0 claims
		unsigned int k = k0;
		while (k < nev) {
			const unsigned long long w = evc[k];
			const int c = (int)(w & 3ull);
			if (c == 0) {
				B = 0;
				k++;
				continue;
			}
			const unsigned long long pidx = w >> 2;
			const bool has2 = (c == 1) && (k + 1 < nev) && ((evc[k + 1] & 3ull) == 2);
			const unsigned long long q = ev_q[k];
			if (B) { // lib/codel.js:62-68, episode carried in from the previous batch
				const unsigned long long d = ev_d[k];
				if (d <= q) {
					if (wr) {
						p.iv_lo[niv] = d;
						p.iv_hi[niv] = q;
						p.iv_cum[niv] = drops;
					}
					niv++;
					drops += q - d + 1;
					C += q - d + 1;
				}
			} else { // lib/codel.js:70-83
				long long e = -1;
				double now_e = ev_now[k];
				unsigned long long dspec = ev_d[k];
				bool cond1 = __dsub_rn(now_e, D) < CODEL_INTERVAL;
				if (c == 2 || cond1) {
					e = (long long)pidx;
				} else if (has2) {
					e = (long long)(evc[k + 1] >> 2);
					now_e = ev_now[k + 1];
					dspec = ev_d[k + 1];
					cond1 = __dsub_rn(now_e, D) < CODEL_INTERVAL;
				}
				if (e >= 0) {
					C = cond1 ? (C > 2 ? C - 2 : 1) : 1;
					D = drop_next_after(now_e, C);
					B = 1;
					unsigned long long d = dspec;
					if (C != 1) { // speculation assumed count == 1 (drop_next = now + 100)
						Pred pr;
						pr.level = 0;
						pr.x = D;
						d = warp_search<NS>(p.now, s_cidx, G, (unsigned long long)e + 1, q + 1, pr, lane);
					}
					if (wr) {
						p.iv_lo[niv] = (unsigned long long)e;
						p.iv_hi[niv] = (unsigned long long)e;
						p.iv_cum[niv] = drops;
					}
					niv++;
					drops += 1;
					if (d <= q) {
						if (wr) {
							p.iv_lo[niv] = d;
							p.iv_hi[niv] = q;
							p.iv_cum[niv] = drops;
						}
						niv++;
						drops += q - d + 1;
						C += q - d + 1;
					}
				}
			}
			k += has2 ? 2 : 1;
		}
		// ---- final ControlledDelay state -------------------------------------------------
		double F;
		if (anchored)
			F = F_anchor;
		else // the whole batch is one above-target run without a reset: F is inherited, or opened by claim 0
			F = (F_in != 0.0) ? F_in : now0p100;
		if (FLAGS) {
			if (hasE_new) {
				if (wr) {
					g->E = E_val;
					g->hasE = 1;
				}
				if (trailing_marker)
					F = 0.0; // trailing empty(): lib/codel.js:93
			}
		}
		if (wr) {
			bc->dbg[5] = gtimer();
			g->F = F;
			g->D = D;
			g->C = C;
			g->B = B;
			bc->n = p.n;
			bc->path = 0;
			bc->n_events = nev - k0;
			bc->n_intervals = niv;
			bc->gcounts[2] = p.n;
			bc->gcounts[3] = drops; // CoDel-level drops (incl. non-waiting handles)
			bc->lcounts[2] = p.n;
			bc->cursor0 = cursor_in;
			if (!FLAGS) {
				const unsigned long long admits = p.n - drops;
				bc->drops = drops;
				bc->admits = admits;
				bc->skipped = 0;
				g->tot_drops += drops;
				g->tot_admits += admits;
				bc->gcounts[0] = drops;
				bc->gcounts[1] = admits;
				bc->lcounts[0] = drops;
				bc->lcounts[1] = admits;
				g->cursor = cursor_in + admits; // the emit kernel works from cursor0
				// the next batch's commit may go ahead (a flags batch finishes the cursor in its emit
				// kernel, which then passes the chain on)
				chain_post(p.link, g);
			} else {
				bc->drops = 0; // accumulated by the emit kernel (waiting handles only)
				bc->admits = 0;
				bc->skipped = 0;
			}
		}
	}
}


// Prepare and commit in ONE launch: the state-independent half runs, then the same CTA waits on the
// commit chain and finishes.  With the commits ordered on the device a kernel boundary between the
// two halves buys nothing, and it costs a launch on the host and a launch latency on the device.
template <bool NS, bool FLAGS>
__global__ void __launch_bounds__(RTPB, 1) resolve_kernel(const ResolveParams p) {
	prepare_part<NS, FLAGS>(p);
	__syncthreads(); // prepare's global writes (events, control words) are visible to the whole CTA
	commit_part<NS, FLAGS>(p);
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
// the same with the clocks already in the reference's double milliseconds (hdl.ch_started, lib/connection-fsm.js:523)
__global__ void overloaded_one_ms_kernel(DevState *g, double start_ms, double now_ms, int *out) {
	CodelRegs s;
	s.F = g->F;
	s.D = g->D;
	s.C = g->C;
	s.B = g->B;
	bool drop = codel_step(s, g->target, now_ms, start_ms);
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
