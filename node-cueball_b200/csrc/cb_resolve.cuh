// cb_resolve.cuh -- pass 2: resolve the dropping episodes (one CTA).
//
// Input: per tile, the class-change events the scan kernel found locally, plus the tile's
// record (where the above-target run inherited from earlier tiles closes; what F the tile
// leaves behind; the class of its last claim).
//
// Step A  pairs every closing tile with the tile that opened the run (a block-wide max-scan
//         over "last tile with a CONST claim") and evaluates the run's unclassified stretch:
//         the class there is monotone in the clock, so its (at most two) crossings are
//         lower_bound searches, done by one warp per long run.  This is what replaces a
//         tile-to-tile look-back chain: a run spanning thousands of tiles costs two searches.
// Step B  merges run events and local events into one ordered list (block scan).
// Step C  Between two events every claim has the same canDrop class, so
//         ControlledDelay#overloaded (lib/codel.js:56-86) acts on a whole stretch in closed form:
//           class 0 stretch        : cd_dropping := false, nothing dropped
//           class>0, dropping      : claims with now >= cd_drop_next are dropped (a suffix: the
//                                    clock is monotone) and counted            codel.js:65-68
//           class>0, not dropping  : the state is entered at the stretch's FIRST claim iff
//                                    class==2 or now - drop_next < 100          codel.js:70-83
//         so the sequential recurrence over (drop_next, count, dropping) runs over EVENTS (a few
//         per 100 ms of stream), not over claims.  Its only non-O(1) step, "first claim with
//         now >= drop_next", is a lower_bound on the monotone clock; all of them are speculated
//         in parallel (one warp per event, assuming count==1 on entry, which is what a monotone
//         clock produces except for a rounding coincidence) before the serial chain runs.
//
// If the clock went backwards inside the batch (or a tile overflowed its event slots) the
// decomposition does not apply and warp 0 runs the exact per-claim state machine instead
// (CB_PATH_SEQUENTIAL), writing the outputs itself.
#pragma once
#include "cb_device.cuh"
#include "cb_scan.cuh"

namespace cb {

constexpr int RTPB = 1024;

struct ResolveParams {
	const void *start;
	const void *now;
	const uint8_t *flags;
	unsigned long long n;
	unsigned long long tiles;
	DevState *g;
	const unsigned int *rec_fc;         // per-tile records from the scan kernel
	const unsigned long long *rec_code;
	const unsigned int *rec_cend;
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
};

template <bool NS>
__device__ __forceinline__ double now_at(const void *now, unsigned long long i) {
	using raw_t = typename TimeIn<NS>::raw_t;
	return TimeIn<NS>::to_ms(reinterpret_cast<const raw_t *>(now)[i]);
}

// first index in [lo,hi) with now >= D, else hi.  Warp-cooperative 32-ary search.
template <bool NS>
__device__ unsigned long long warp_lower_bound(const void *now, unsigned long long lo,
    unsigned long long hi, double D, int lane) {
	while (true) {
		unsigned long long s = hi - lo;
		if (s == 0)
			return hi;
		if (s <= 32) {
			bool pr = (lane < s) && (now_at<NS>(now, lo + lane) >= D);
			unsigned int m = __ballot_sync(0xffffffffu, pr);
			return m ? lo + (__ffs(m) - 1) : hi;
		}
		unsigned long long chunk = (s + 31) / 32;
		unsigned long long pidx = lo + (unsigned long long)(lane + 1) * chunk - 1;
		if (pidx > hi - 1)
			pidx = hi - 1;
		bool pr = now_at<NS>(now, pidx) >= D;
		unsigned int m = __ballot_sync(0xffffffffu, pr);
		if (!m)
			return hi;
		int f = __ffs(m) - 1;
		unsigned long long pf = __shfl_sync(0xffffffffu, pidx, f);
		unsigned long long pprev = __shfl_sync(0xffffffffu, pidx, f > 0 ? f - 1 : 0);
		if (f > 0)
			lo = pprev + 1;
		hi = pf; // answer in [lo, pf]; pf itself satisfies the predicate
		if (lo >= hi)
			return hi;
	}
}
// first index in [lo,hi) whose canDrop class w.r.t. F is >= level, else hi (class is monotone in
// the clock inside an above-target run)
template <bool NS>
__device__ unsigned long long warp_first_class(const void *now, unsigned long long lo, unsigned long long hi,
    double F, int level, int lane) {
	while (true) {
		unsigned long long s = hi - lo;
		if (s == 0)
			return hi;
		if (s <= 32) {
			bool pr = (lane < s) && (claim_class(true, now_at<NS>(now, lo + lane), F) >= level);
			unsigned int m = __ballot_sync(0xffffffffu, pr);
			return m ? lo + (__ffs(m) - 1) : hi;
		}
		unsigned long long chunk = (s + 31) / 32;
		unsigned long long pidx = lo + (unsigned long long)(lane + 1) * chunk - 1;
		if (pidx > hi - 1)
			pidx = hi - 1;
		bool pr = claim_class(true, now_at<NS>(now, pidx), F) >= level;
		unsigned int m = __ballot_sync(0xffffffffu, pr);
		if (!m)
			return hi;
		int f = __ffs(m) - 1;
		unsigned long long pf = __shfl_sync(0xffffffffu, pidx, f);
		unsigned long long pprev = __shfl_sync(0xffffffffu, pidx, f > 0 ? f - 1 : 0);
		if (f > 0)
			lo = pprev + 1;
		hi = pf;
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
		g->cursor = cursor;
		g->n = p.n;
		g->drops = drops;
		g->admits = admits;
		g->skipped = skipped;
		g->tot_drops += drops;
		g->tot_admits += admits;
		g->gcounts[0] = drops;
		g->gcounts[1] = admits;
		g->gcounts[2] = p.n;
		g->gcounts[3] = 0;
		g->path = 1;
		g->n_events = 0;
		g->n_intervals = 0;
		g->nonmono = 0;
		g->overflow = 0;
		g->last_marker_p1 = 0;
	}
}

template <bool NS, bool FLAGS>
__global__ void __launch_bounds__(RTPB, 1) resolve_kernel(const ResolveParams p) {
	__shared__ unsigned int s_warp[32];
	__shared__ unsigned int s_base;
	__shared__ unsigned int s_nev;
	__shared__ int s_fallback;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	DevState *g = p.g;

	if (tid == 0) {
		s_fallback = (g->nonmono != 0u) || (g->overflow != 0u);
		s_base = 0;
	}
	__syncthreads();
	if (s_fallback) {
		if (warp == 0)
			sequential_path<NS>(p, lane);
		return;
	}

	// ---- A. pair closing tiles with the run they close; evaluate the inherited stretches ------
	__shared__ int s_wmax[32];
	__shared__ long long s_prevconst; // last tile with a CONST claim seen so far (-1: none yet)
	__shared__ unsigned int s_nheavy;
	if (tid == 0) {
		s_prevconst = -1;
		s_nheavy = 0;
	}
	__syncthreads();
	const unsigned long long carry_code = (g->F != 0.0) ? CODE_CARRY : CODE_ZERO;
	for (unsigned long long tb = 0; tb < p.tiles + 1; tb += RTPB) {
		const unsigned long long t = tb + tid;
		// slot `tiles` is a virtual closing tile at index n (the run still open at batch end)
		const bool is_tail = (t == p.tiles);
		const unsigned int fc = (t < p.tiles) ? p.rec_fc[t] : (is_tail ? 0u : FC_NONE);
		const bool hc = (t <= p.tiles) && fc != FC_NONE;
		// exclusive max-scan of (hc ? t : -1) over the block
		int v = (hc && !is_tail) ? (int)(t - tb) : -1; // block-local index
		int x = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			int y = __shfl_up_sync(0xffffffffu, x, o);
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
		const long long T0 = (ex >= 0) ? (long long)(tb + ex) : s_prevconst;
		if (hc) {
			const unsigned long long code0 = (T0 < 0) ? carry_code : p.rec_code[T0];
			const unsigned int cend0 = (T0 < 0) ? 0u : p.rec_cend[T0];
			const unsigned long long lo = (unsigned long long)(T0 + 1) * TILE;
			const unsigned long long close_idx = is_tail ? p.n : (t * TILE + fc); // first claim after the run
			unsigned int cnt = 0;
			int final_c = (int)cend0;
			if (lo < close_idx) { // unclassified stretch [lo, close_idx)
				double F;
				if (code0 == CODE_ZERO)
					F = __dadd_rn(now_at<NS>(p.now, lo), CODEL_INTERVAL); // the stretch's first claim opens the run
				else if (code0 == CODE_CARRY)
					F = g->F;
				else
					F = __dadd_rn(now_at<NS>(p.now, code0 - 2), CODEL_INTERVAL);
				const int cb = claim_class(true, now_at<NS>(p.now, close_idx - 1), F);
				if (cb > final_c) { // the class rises inside the stretch: find where (one warp, below)
					const unsigned int h = atomicAdd(&s_nheavy, 1u);
					p.hj_lo[h] = lo;
					p.hj_b[h] = close_idx - 1;
					p.hj_F[h] = F;
					p.hj_info[h] = (unsigned int)t | (cend0 << 28);
					final_c = cb;
				}
			}
			if (final_c >= 1 && !is_tail) // the run ends at a CONST claim (class 0)
				p.xev[t * 3 + 2] = (close_idx << 2);
			cnt = (final_c >= 1 && !is_tail) ? 4u : 0u; // bit 2: close event present
			p.xcnt[t] = cnt;
		} else if (t <= p.tiles) {
			p.xcnt[t] = 0;
		}
		__syncthreads();
		if (tid == RTPB - 1) {
			const int blockmax = max(x, wprev);
			if (blockmax >= 0)
				s_prevconst = (long long)(tb + blockmax);
		}
		__syncthreads();
	}
	// long runs: class crossings by binary search, one warp per run
	const unsigned int nheavy = s_nheavy;
	for (unsigned int h = warp; h < nheavy; h += RTPB / 32) {
		const unsigned long long lo = p.hj_lo[h], b = p.hj_b[h];
		const double F = p.hj_F[h];
		const unsigned int info = p.hj_info[h];
		const unsigned int slot = info & 0x0fffffffu;
		const int cend0 = (int)(info >> 28);
		unsigned int bits = p.xcnt[slot];
		unsigned long long e0 = ~0ull, e1 = ~0ull;
		unsigned long long pos = lo;
		int cur = cend0;
		if (cur == 0) {
			pos = warp_first_class<NS>(p.now, lo, b + 1, F, 1, lane);
			cur = claim_class(true, now_at<NS>(p.now, pos), F); // 1 or 2
			e0 = (pos << 2) | (unsigned long long)cur;
		}
		if (cur == 1) {
			const unsigned long long p2 = warp_first_class<NS>(p.now, pos, b + 1, F, 2, lane);
			if (p2 <= b)
				e1 = (p2 << 2) | 2ull;
		}
		if (lane == 0) {
			if (e0 != ~0ull) {
				p.xev[slot * 3 + 0] = e0;
				bits |= 1u;
			}
			if (e1 != ~0ull) {
				p.xev[slot * 3 + 1] = e1;
				bits |= 2u;
			}
			p.xcnt[slot] = bits;
		}
	}
	__syncthreads();

	// ---- B. merge run events and local events into one ordered list -----------------------------
	for (unsigned long long tb = 0; tb < p.tiles + 1; tb += RTPB) {
		unsigned long long t = tb + tid;
		const unsigned int xb = (t <= p.tiles) ? p.xcnt[t] : 0u;
		const unsigned int lc = (t < p.tiles) ? p.ev_count[t] : 0u;
		const unsigned int cnt = __popc(xb) + lc;
		// block exclusive scan of cnt
		unsigned int x = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o)
				x += y;
		}
		if (lane == 31)
			s_warp[warp] = x;
		__syncthreads();
		if (warp == 0) {
			unsigned int w = s_warp[lane];
			unsigned int ws = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				unsigned int y = __shfl_up_sync(0xffffffffu, ws, o);
				if (lane >= o)
					ws += y;
			}
			s_warp[lane] = ws - w; // exclusive
			if (lane == 31)
				s_nev = ws;        // chunk total
		}
		__syncthreads();
		unsigned int off = s_base + s_warp[warp] + (x - cnt);
		if (cnt) {
			// run events of the tile's close come first (crossing(s) <= close index <= local events)
			if ((xb & 1u) && off < p.ev_cap)
				p.evc[off++] = p.xev[t * 3 + 0];
			if ((xb & 2u) && off < p.ev_cap)
				p.evc[off++] = p.xev[t * 3 + 1];
			if ((xb & 4u) && off < p.ev_cap)
				p.evc[off++] = p.xev[t * 3 + 2];
		}
		if (lc) {
			// a tile's events were appended unordered: sort its <= EVK entries by index
			unsigned long long e[EVK];
#pragma unroll
			for (int j = 0; j < EVK; j++)
				e[j] = (j < (int)lc) ? p.ev[t * EVK + j] : ~0ull;
#pragma unroll
			for (int a = 1; a < EVK; a++) {
#pragma unroll
				for (int b = EVK - 1; b >= 1; b--) {
					if (b >= a && e[b] < e[b - 1]) {
						unsigned long long tmp = e[b];
						e[b] = e[b - 1];
						e[b - 1] = tmp;
					}
				}
			}
#pragma unroll
			for (int j = 0; j < EVK; j++)
				if (j < (int)lc && off + j < p.ev_cap)
					p.evc[off + j] = e[j];
		}
		__syncthreads();
		if (tid == 0)
			s_base += s_nev;
		__syncthreads();
	}
	const unsigned int nev = s_base;
	if (nev > p.ev_cap) { // cannot happen with ev_cap = (tiles+1)*(EVK+3); defensive: exact path
		if (warp == 0)
			sequential_path<NS>(p, lane);
		return;
	}

	// ---- 2. per event (one warp each): run end q, now at the event, speculated drop start
	const double D_in = g->D;
	const int B_in = g->B;
	for (unsigned int k = warp; k < nev; k += RTPB / 32) {
		unsigned long long w = p.evc[k];
		int c = (int)(w & 3ull);
		unsigned long long idx = w >> 2;
		if (c == 0)
			continue;
		// the run ends right before the next class-0 event (at most two events ahead)
		unsigned long long q = p.n - 1;
		if (k + 1 < nev) {
			unsigned long long w1 = p.evc[k + 1];
			if ((w1 & 3ull) == 0)
				q = (w1 >> 2) - 1;
			else if (k + 2 < nev) {
				unsigned long long w2 = p.evc[k + 2];
				if ((w2 & 3ull) == 0)
					q = (w2 >> 2) - 1;
			}
		}
		double nowk = now_at<NS>(p.now, idx);
		unsigned long long d;
		if (k == 0 && idx == 0 && B_in) // batch opens inside a dropping episode
			d = warp_lower_bound<NS>(p.now, idx, q + 1, D_in, lane);
		else
			d = warp_lower_bound<NS>(p.now, idx + 1, q + 1, __dadd_rn(nowk, CODEL_INTERVAL), lane);
		if (lane == 0) {
			p.ev_now[k] = nowk;
			p.ev_q[k] = q;
			p.ev_d[k] = d;
		}
	}
	__syncthreads();

	// ---- 3. the serial chain over events (thread 0) -----------------------------------
	if (tid == 0) {
		double D = g->D;
		unsigned long long C = g->C;
		int B = g->B;
		unsigned long long drops = 0;
		unsigned int niv = 0;
		if (nev == 0 || (p.evc[0] >> 2) > 0)
			B = 0; // the batch starts with class-0 claims
		unsigned int k = 0;
		while (k < nev) {
			const unsigned long long w = p.evc[k];
			const int c = (int)(w & 3ull);
			if (c == 0) {
				B = 0;
				k++;
				continue;
			}
			const unsigned long long pidx = w >> 2;
			const bool has2 = (c == 1) && (k + 1 < nev) && ((p.evc[k + 1] & 3ull) == 2);
			const unsigned long long q = p.ev_q[k];
			if (B) { // lib/codel.js:62-68, episode carried in from the previous batch
				const unsigned long long d = p.ev_d[k];
				if (d <= q) {
					p.iv_lo[niv] = d;
					p.iv_hi[niv] = q;
					p.iv_cum[niv] = drops;
					niv++;
					drops += q - d + 1;
					C += q - d + 1;
				}
			} else { // lib/codel.js:70-83
				long long e = -1;
				double now_e = p.ev_now[k];
				unsigned long long dspec = p.ev_d[k];
				bool cond1 = __dsub_rn(now_e, D) < CODEL_INTERVAL;
				if (c == 2 || cond1) {
					e = (long long)pidx;
				} else if (has2) {
					e = (long long)(p.evc[k + 1] >> 2);
					now_e = p.ev_now[k + 1];
					dspec = p.ev_d[k + 1];
					cond1 = __dsub_rn(now_e, D) < CODEL_INTERVAL;
				}
				if (e >= 0) {
					C = cond1 ? (C > 2 ? C - 2 : 1) : 1;
					D = drop_next_after(now_e, C);
					B = 1;
					unsigned long long d = dspec;
					if (C != 1) // speculation assumed count == 1 (drop_next = now + 100)
						d = thread_lower_bound<NS>(p.now, (unsigned long long)e + 1, q + 1, D);
					p.iv_lo[niv] = (unsigned long long)e;
					p.iv_hi[niv] = (unsigned long long)e;
					p.iv_cum[niv] = drops;
					niv++;
					drops += 1;
					if (d <= q) {
						p.iv_lo[niv] = d;
						p.iv_hi[niv] = q;
						p.iv_cum[niv] = drops;
						niv++;
						drops += q - d + 1;
						C += q - d + 1;
					}
				}
			}
			k += has2 ? 2 : 1;
		}
		// ---- final ControlledDelay state -------------------------------------------------
		// F after the batch's last claim: the last tile with a CONST claim fixes it, later (all
		// PASS) tiles inherit it or, if it is 0, the first of them opens the run
		double F;
		{
			const long long Tl = s_prevconst;
			const unsigned long long code0 = (Tl < 0) ? carry_code : p.rec_code[Tl];
			if (code0 == CODE_CARRY)
				F = g->F;
			else if (code0 != CODE_ZERO)
				F = __dadd_rn(now_at<NS>(p.now, code0 - 2), CODEL_INTERVAL);
			else if ((unsigned long long)(Tl + 1) < p.tiles)
				F = __dadd_rn(now_at<NS>(p.now, (unsigned long long)(Tl + 1) * TILE), CODEL_INTERVAL);
			else
				F = 0.0;
		}
		if (FLAGS) {
			const unsigned long long lm = g->last_marker_p1;
			if (lm) {
				g->E = now_at<NS>(p.now, lm - 1);
				g->hasE = 1;
				if (lm == p.n)
					F = 0.0; // trailing empty(): lib/codel.js:93
			}
			g->last_marker_p1 = 0;
		}
		g->F = F;
		g->D = D;
		g->C = C;
		g->B = B;
		g->n = p.n;
		g->path = 0;
		g->n_events = nev;
		g->n_intervals = niv;
		g->gcounts[2] = p.n;
		g->gcounts[3] = drops; // CoDel-level drops (incl. non-waiting handles)
		g->cursor0 = g->cursor;
		if (!FLAGS) {
			const unsigned long long admits = p.n - drops;
			g->drops = drops;
			g->admits = admits;
			g->skipped = 0;
			g->tot_drops += drops;
			g->tot_admits += admits;
			g->gcounts[0] = drops;
			g->gcounts[1] = admits;
			g->cursor += admits; // the emit kernel works from cursor0
		} else {
			g->drops = 0; // accumulated by the emit kernel (waiting handles only)
			g->admits = 0;
			g->skipped = 0;
		}
	}
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
