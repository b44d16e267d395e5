// cb_scan.cuh -- pass 1 of the admission path: the streaming scan kernel.
//
// Reads every claim's (start, now[, flags]) exactly once from HBM (16 B/claim, 17 with
// flags), converts to milliseconds like utils.currentMillis() (lib/utils.js:198-204),
// evaluates ControlledDelay#canDrop (lib/codel.js:34-46) for ALL claims in parallel and
// reduces the stream to the handful of positions where the canDrop class changes.
//
// canDrop's only loop-carried variable is cd_first_above_time (F).  Per claim it is
// rewritten by one of three functions
//      !above                : F' = 0                      CONST(0)
//       above, marker before : F' = now+100                CONST(now+100)
//       above                : F' = F ? F : now+100        PASS(now+100)
// which compose associatively ("last CONST wins, then the first PASS fills a zero"), so
// F before every claim is an exclusive scan: thread-serial over 8 claims, ballot+shuffle
// across the warp, shared memory across the CTA, decoupled look-back across tiles.
// (A CB_F_EMPTY_AFTER marker -- codel.empty(), lib/codel.js:91-94 -- is applied as a reset
// in front of the NEXT claim, which is equivalent and keeps the carried value meaningful.)
#pragma once
#include "cb_device.cuh"

namespace cb {

struct Agg { // a composed F-update: F_out = isconst ? v : (F_in ? F_in : v)
	bool isconst;
	double v;
};
__device__ __forceinline__ Agg compose(const Agg &a, const Agg &b) { // a first, then b
	if (b.isconst)
		return b;
	Agg r;
	r.isconst = a.isconst;
	r.v = (a.v != 0.0) ? a.v : b.v;
	return r;
}
__device__ __forceinline__ double apply_agg(const Agg &a, double Fin) {
	return a.isconst ? a.v : (Fin != 0.0 ? Fin : a.v);
}

__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p));
	return r;
}

// Coalesced load of one warp's 256 consecutive 8-byte elements of two arrays, transposed
// through a conflict-free XOR-swizzled shared-memory stage so that lane L ends up with
// elements [8L, 8L+8).  All 8 global loads are issued before the first shared store.
template <typename T>
__device__ __forceinline__ void load_warp_blocked2(const T *__restrict__ a, const T *__restrict__ b,
    T (&oa)[IPT], T (&ob)[IPT], uint4 *stage, int lane) {
	static_assert(sizeof(T) == 8, "8-byte elements");
	const uint4 *pa = reinterpret_cast<const uint4 *>(a);
	const uint4 *pb = reinterpret_cast<const uint4 *>(b);
	uint4 va[4], vb[4];
#pragma unroll
	for (int r = 0; r < 4; r++)
		va[r] = ld_stream_v4(pa + r * 32 + lane);
#pragma unroll
	for (int r = 0; r < 4; r++)
		vb[r] = ld_stream_v4(pb + r * 32 + lane);
#pragma unroll
	for (int r = 0; r < 4; r++) {
		int g = r * 32 + lane;
		stage[g ^ ((g >> 3) & 7)] = va[r];
	}
	__syncwarp();
#pragma unroll
	for (int k = 0; k < 4; k++) {
		int g = 4 * lane + k;
		uint4 x = stage[g ^ ((g >> 3) & 7)];
		unsigned long long lo = ((unsigned long long)x.y << 32) | x.x;
		unsigned long long hi = ((unsigned long long)x.w << 32) | x.z;
		oa[2 * k] = *reinterpret_cast<T *>(&lo);
		oa[2 * k + 1] = *reinterpret_cast<T *>(&hi);
	}
	__syncwarp();
#pragma unroll
	for (int r = 0; r < 4; r++) {
		int g = r * 32 + lane;
		stage[g ^ ((g >> 3) & 7)] = vb[r];
	}
	__syncwarp();
#pragma unroll
	for (int k = 0; k < 4; k++) {
		int g = 4 * lane + k;
		uint4 x = stage[g ^ ((g >> 3) & 7)];
		unsigned long long lo = ((unsigned long long)x.y << 32) | x.x;
		unsigned long long hi = ((unsigned long long)x.w << 32) | x.z;
		ob[2 * k] = *reinterpret_cast<T *>(&lo);
		ob[2 * k + 1] = *reinterpret_cast<T *>(&hi);
	}
	__syncwarp();
}

struct ScanParams {
	const void *start;
	const void *now;
	const uint8_t *flags; // may be null
	unsigned long long n;
	DevState *g;
	TileDesc *desc;
	unsigned long long epoch;
	unsigned int *ev_count;  // [tiles]
	unsigned long long *ev;  // [tiles][EVK]  (index << 2) | class
	int vec_ok;              // start/now 16-byte aligned
};

template <bool NS, bool FLAGS>
__global__ void __launch_bounds__(TPB, 3) scan_kernel(const ScanParams p) {
	using raw_t = typename TimeIn<NS>::raw_t;
	__shared__ uint4 s_stage[WARPS][128];   // 2 KB per warp
	__shared__ double s_wv[WARPS];
	__shared__ int s_wconst[WARPS];
	__shared__ double s_wlastnow[WARPS];
	__shared__ int s_wlastc[WARPS];
	__shared__ double s_Fc;
	__shared__ int s_cprev;
	__shared__ unsigned int s_evn;
	__shared__ unsigned long long s_ev[EVK];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned long long t = blockIdx.x;
	const unsigned long long ts = t * (unsigned long long)TILE;
	const unsigned long long i0 = ts + (unsigned long long)tid * IPT;
	const raw_t *gstart = reinterpret_cast<const raw_t *>(p.start);
	const raw_t *gnow = reinterpret_cast<const raw_t *>(p.now);
	const double target = p.g->target;
	if (tid == 0)
		s_evn = 0;

	// ---- load this thread's 8 consecutive claims ------------------------------------
	const bool full = (ts + TILE <= p.n);
	int nvalid = IPT;
	raw_t rs[IPT], rn[IPT];
	if (full && p.vec_ok) {
		load_warp_blocked2<raw_t>(gstart + ts + warp * 256, gnow + ts + warp * 256, rs, rn,
		    s_stage[warp], lane);
	} else {
		long long rem = (long long)p.n - (long long)i0;
		nvalid = rem >= IPT ? IPT : (rem > 0 ? (int)rem : 0);
#pragma unroll
		for (int j = 0; j < IPT; j++) {
			if (j < nvalid) {
				rs[j] = gstart[i0 + j];
				rn[j] = gnow[i0 + j];
			} else {
				rs[j] = raw_t(0);
				rn[j] = raw_t(0);
			}
		}
	}
	unsigned long long fl8 = 0x0101010101010101ull; // all waiting, no markers
	unsigned int fprev = 0;                         // flags of claim i0-1
	if (FLAGS) {
		if (nvalid == IPT && ((reinterpret_cast<uintptr_t>(p.flags) & 7) == 0)) {
			fl8 = *reinterpret_cast<const unsigned long long *>(p.flags + i0);
		} else {
			fl8 = 0;
			for (int j = 0; j < nvalid; j++)
				fl8 |= (unsigned long long)p.flags[i0 + j] << (8 * j);
		}
		unsigned int lastb = (unsigned int)(fl8 >> 56);
		fprev = __shfl_up_sync(0xffffffffu, lastb, 1);
		if (lane == 0)
			fprev = (i0 > 0 && nvalid > 0) ? p.flags[i0 - 1] : 0;
	}
	// the claim just before the tile (needed for its class and for the monotone check);
	// issued early so the latency hides behind the conversions below
	raw_t pstart = raw_t(0), pnow = raw_t(0);
	unsigned int pfl = F_WAITING, ppfl = 0;
	if (tid == 0 && t > 0) {
		pstart = gstart[ts - 1];
		pnow = gnow[ts - 1];
		if (FLAGS) {
			pfl = p.flags[ts - 1];
			ppfl = ts >= 2 ? p.flags[ts - 2] : 0;
		}
	}

	// ---- per-claim conversion, above flag, thread aggregate -------------------------
	double nw[IPT];
	unsigned int above = 0, mk = 0; // bit j: claim j above target / carries a marker
	Agg ta;
	ta.isconst = false;
	ta.v = 0.0;
	bool nonmono = false;
#pragma unroll
	for (int j = 0; j < IPT; j++) {
		nw[j] = TimeIn<NS>::to_ms(rn[j]);
		double st = TimeIn<NS>::to_ms(rs[j]);
		if (j < nvalid) {
			bool ab = !(__dsub_rn(nw[j], st) < target);
			bool mprev = FLAGS && (((j == 0 ? fprev : (unsigned int)(fl8 >> (8 * (j - 1)))) & F_EMPTY_AFTER) != 0);
			if (FLAGS && ((fl8 >> (8 * j)) & F_EMPTY_AFTER))
				mk |= 1u << j;
			if (ab)
				above |= 1u << j;
			if (!ab) {
				ta.isconst = true;
				ta.v = 0.0;
			} else if (mprev) {
				ta.isconst = true;
				ta.v = __dadd_rn(nw[j], CODEL_INTERVAL);
			} else if (ta.v == 0.0) {
				ta.v = __dadd_rn(nw[j], CODEL_INTERVAL);
			}
			if (j > 0 && nw[j] < nw[j - 1])
				nonmono = true;
		}
	}

	// ---- warp-level exclusive composition (ballot: last CONST lane before me) --------
	const unsigned int cm = __ballot_sync(0xffffffffu, ta.isconst);
	Agg ex; // composition of lanes < lane
	{
		unsigned int prior = cm & ((1u << lane) - 1u);
		int jl = prior ? (31 - __clz(prior)) : 0;
		double vj = __shfl_sync(0xffffffffu, ta.v, jl);
		int nx = prior ? (jl + 1) : 0; // first PASS lane after the CONST lane (or lane 0)
		double vnx = __shfl_sync(0xffffffffu, ta.v, nx & 31);
		if (prior) {
			ex.isconst = true;
			ex.v = (vj != 0.0) ? vj : ((nx < lane) ? vnx : 0.0);
		} else {
			ex.isconst = false;
			ex.v = (lane > 0) ? vnx : 0.0;
		}
	}
	{
		Agg winc = compose(ex, ta); // inclusive at this lane; lane 31 holds the warp's
		double wv = __shfl_sync(0xffffffffu, winc.v, 31);
		int wc = __shfl_sync(0xffffffffu, (int)winc.isconst, 31);
		double prevnow = __shfl_up_sync(0xffffffffu, nw[IPT - 1], 1);
		if (lane > 0 && nvalid > 0 && nw[0] < prevnow)
			nonmono = true;
		if (lane == 31) {
			s_wv[warp] = wv;
			s_wconst[warp] = wc;
			s_wlastnow[warp] = nw[IPT - 1];
		}
	}
	__syncthreads();

	// ---- CTA-level: exclusive over warps, tile aggregate --------------------------------
	Agg wex;
	wex.isconst = false;
	wex.v = 0.0;
	Agg tile_agg = wex;
#pragma unroll
	for (int w = 0; w < WARPS; w++) {
		Agg a;
		a.isconst = s_wconst[w] != 0;
		a.v = s_wv[w];
		if (w == warp)
			wex = tile_agg;
		tile_agg = compose(tile_agg, a);
	}
	if (lane == 0 && warp > 0 && nvalid > 0 && nw[0] < s_wlastnow[warp - 1])
		nonmono = true;

	// ---- decoupled look-back across tiles (warp 0) ---------------------------------------
	if (warp == 0) {
		double Fc;
		if (t == 0) {
			Fc = p.g->F;
		} else {
			if (lane == 0) {
				p.desc[t].vA = tile_agg.v;
				st_release(&p.desc[t].w, (p.epoch << 3) | ((unsigned long long)tile_agg.isconst << 2) | ST_AGG);
			}
			long long base = (long long)t - 1;
			double next_v0 = 0.0;
			const double Fin = p.g->F;
			while (true) {
				long long k = base - lane;
				double val = 0.0;
				bool term = false;
				if (k >= 0) {
					unsigned long long w;
					do {
						w = ld_acquire(&p.desc[k].w);
					} while ((w >> 3) != p.epoch);
					const unsigned int st = (unsigned int)(w & 3ull);
					const bool isc = ((w >> 2) & 1ull) != 0;
					val = (st == ST_PREFIX) ? ld_relaxed_f64(&p.desc[k].vP) : ld_relaxed_f64(&p.desc[k].vA);
					term = (st == ST_PREFIX) || isc;
				} else if (k == -1) {
					val = Fin;
					term = true;
				}
				const unsigned int tm = __ballot_sync(0xffffffffu, term);
				if (tm) {
					const int f = __ffs(tm) - 1;
					const double vT = __shfl_sync(0xffffffffu, val, f);
					double vn = __shfl_sync(0xffffffffu, val, f > 0 ? f - 1 : 0);
					if (f == 0)
						vn = next_v0;
					Fc = (vT != 0.0) ? vT : vn;
					break;
				}
				next_v0 = __shfl_sync(0xffffffffu, val, 31);
				base -= 32;
			}
		}
		if (lane == 0) {
			p.desc[t].vA = tile_agg.v;
			p.desc[t].vP = apply_agg(tile_agg, Fc);
			st_release(&p.desc[t].w, (p.epoch << 3) | ((unsigned long long)tile_agg.isconst << 2) | ST_PREFIX);
			// class of the claim just before the tile, from F after it (= Fc) -- see DESIGN.md
			int cprev = 0;
			if (t > 0) {
				double pn = TimeIn<NS>::to_ms(pnow), ps = TimeIn<NS>::to_ms(pstart);
				bool pab = !(__dsub_rn(pn, ps) < target);
				bool pm = FLAGS && (ppfl & F_EMPTY_AFTER);
				cprev = (pab && !pm) ? claim_class(true, pn, Fc) : 0;
				if (nvalid > 0 && nw[0] < pn)
					nonmono = true;
			}
			s_Fc = Fc;
			s_cprev = cprev;
		}
	}
	__syncthreads();

	// ---- classify: F before each claim is now known -----------------------------------
	double F = apply_agg(ex, apply_agg(wex, s_Fc));
	unsigned int cls = 0; // 2 bits per claim
	int clast = 0;
#pragma unroll
	for (int j = 0; j < IPT; j++) {
		if (j < nvalid) {
			const bool ab = (above >> j) & 1u;
			const bool mprev = FLAGS && (((j == 0 ? fprev : (unsigned int)(fl8 >> (8 * (j - 1)))) & F_EMPTY_AFTER) != 0);
			const double Fp = mprev ? 0.0 : F;
			const int c = claim_class(ab, nw[j], Fp);
			cls |= (unsigned int)c << (2 * j);
			clast = c;
			F = ab ? ((Fp != 0.0) ? Fp : __dadd_rn(nw[j], CODEL_INTERVAL)) : 0.0;
		}
	}
	// class of the claim before this thread's first claim
	int cbefore = __shfl_up_sync(0xffffffffu, clast, 1);
	if (lane == 31)
		s_wlastc[warp] = clast;
	__syncthreads();
	if (lane == 0)
		cbefore = (warp == 0) ? s_cprev : s_wlastc[warp - 1];

	// ---- events: positions where the class changes (rare) -------------------------------
	if (nvalid > 0) {
		int cp = cbefore;
#pragma unroll
		for (int j = 0; j < IPT; j++) {
			if (j < nvalid) {
				const int c = (cls >> (2 * j)) & 3;
				if (c != cp) {
					unsigned int pos = atomicAdd(&s_evn, 1u);
					if (pos < EVK)
						s_ev[pos] = ((i0 + j) << 2) | (unsigned long long)c;
				}
				cp = c;
			}
		}
		if (FLAGS && mk) {
			int jm = 31 - __clz(mk);
			if (jm < nvalid)
				atomicMax(&p.g->last_marker_p1, i0 + jm + 1);
		}
	}
	if (nonmono)
		p.g->nonmono = 1u;
	__syncthreads();
	if (tid == 0) {
		unsigned int cnt = s_evn;
		if (cnt > EVK) {
			p.g->overflow = 1u;
			cnt = EVK;
		}
		p.ev_count[t] = cnt;
	}
	if (tid < EVK && tid < s_evn)
		p.ev[t * EVK + tid] = s_ev[tid];
}

} // namespace cb
