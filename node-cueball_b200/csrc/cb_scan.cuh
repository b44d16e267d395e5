// cb_scan.cuh -- pass 1 of the admission path: the streaming scan kernel.
//
// Reads every claim's (start, now[, flags]) exactly once from HBM (16 B/claim, 17 with
// flags), evaluates ControlledDelay#canDrop (lib/codel.js:34-46) for ALL claims in parallel
// and reduces the stream to the handful of positions where the canDrop class changes.
//
// canDrop's only loop-carried variable is cd_first_above_time (F).  Per claim it is
// rewritten by one of three functions
//      !above                : F' = 0                      CONST(0)
//       above, marker before : F' = now+100                CONST(now+100)
//       above                : F' = F ? F : now+100        PASS(now+100)
// which compose associatively ("last CONST wins, then the first PASS fills a zero").  Every
// non-zero F is `now[a] + 100` for some claim index a, so compositions are carried as
// INDICES (integer-only, one 64-bit code per warp chunk and per tile) and turned into doubles
// only where a claim's class needs the value.
//
// Layout: a CTA owns a tile of 2048 consecutive claims, a warp a chunk of 256 = 4 rows of 64;
// in a row lane l owns claims 2l and 2l+1, i.e. one fully coalesced 128-bit load per array
// per row, no shared-memory staging.  Phase 1 (integers only) classifies above/below --
// sojourn vs target is decided on the int64 nanoseconds whenever the gap to the target
// exceeds the worst-case rounding of the millisecond doubles, and exactly as
// utils.currentMillis() (lib/utils.js:198-204) otherwise -- and builds the row/warp/tile
// compositions.  Phase 2 touches FP64 only in rows that contain above-target claims.
//
// Tiles are INDEPENDENT: there is no look-back.  A tile classifies only the claims whose F
// is determined inside the tile (everything from its first CONST claim on).  Its leading PASS
// claims belong to an above-target run that was opened in an earlier tile; for those the tile
// just records where the run closes.  The resolve kernel pairs each close with the open it
// belongs to and finds the (at most two) class crossings of such a long run by binary search
// on the monotone clock -- a run spanning thousands of tiles (sustained overload) costs two
// searches instead of a serialised chain of tile-to-tile waits.
// (A CB_F_EMPTY_AFTER marker -- codel.empty(), lib/codel.js:91-94 -- is applied as a reset in
// front of the NEXT claim, which is equivalent and keeps the carried value meaningful.)
#pragma once
#include "cb_device.cuh"
#include <type_traits>

namespace cb {

#ifndef CB_SCAN_MINBLOCKS // resident CTAs per SM the register budget is sized for (tuning knob:
#define CB_SCAN_MINBLOCKS 4 // `make -B EXTRA=-DCB_SCAN_MINBLOCKS=5`; 4 = 64 registers, measured in round 1)
#endif
constexpr int ROWS = 4;            // rows of 64 claims per warp chunk
constexpr int CHUNK = ROWS * 64;   // 256 claims per warp
static_assert(CHUNK * WARPS == TILE, "tile geometry");

// index-coded F: 0 = F is 0, 1 = the batch's carried-in F, a+2 = now[a] + 100
constexpr unsigned long long CODE_ZERO = 0, CODE_CARRY = 1;
constexpr unsigned long long DESC_CODE_MASK = (1ull << 40) - 1;
constexpr unsigned int FC_NONE = 0xffffffffu; // tile record: no CONST claim in the tile
constexpr int EVW = 4;                        // class-change events one warp chunk can hold

// named barriers: arrive without waiting / wait for all TPB threads
#ifndef CB_EMU
__device__ __forceinline__ void bar_arrive(int id) {
	__threadfence_block();
	asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(TPB) : "memory");
}
__device__ __forceinline__ void bar_sync(int id) {
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(TPB) : "memory");
}
#endif

struct IAgg { // composed F-update in index form: F_out = isconst ? code : (F_in ? F_in : code)
	bool isconst;
	unsigned long long code;
};
__device__ __forceinline__ IAgg icompose(const IAgg &a, const IAgg &b) { // a first, then b
	if (b.isconst)
		return b;
	IAgg r;
	r.isconst = a.isconst;
	r.code = (a.code != CODE_ZERO) ? a.code : b.code;
	return r;
}
__device__ __forceinline__ unsigned long long iapply(const IAgg &a, unsigned long long in_code) {
	return a.isconst ? a.code : (in_code != CODE_ZERO ? in_code : a.code);
}

#ifndef CB_EMU
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p));
	return r;
}
#endif

struct ScanParams {
	const void *start;
	const void *now;
	const uint8_t *flags; // may be null
	unsigned long long n;
	BatchCtl *b;   // per-batch control block (nonmono / overflow / last marker)
	double target; // cd_targdelay (ms)
	// per-tile record: first CONST position, F after the tile (index-coded), class of its last claim
	unsigned int *rec_fc;          // [tiles] tile-local position of the first CONST claim, or FC_NONE
	unsigned long long *rec_code;  // [tiles] valid when rec_fc != FC_NONE
	unsigned int *rec_cend;        // [tiles] valid when rec_fc != FC_NONE
	void *tile_now0;               // [tiles] raw clock of the tile's first claim (coarse search index)
	unsigned int *ev_count;   // [tiles]
	unsigned long long *ev;   // [tiles][EVK]  (index << 2) | class
	int vec_ok;               // start/now 16-byte aligned
	long long target_ns;      // target in ns (integer-domain sojourn test)
};

template <bool NS> struct Raw;
template <> struct Raw<true> {
	using t = long long;
	// sojourn vs target in the integer domain; returns 0 below, 1 above, 2 = too close
	static __device__ __forceinline__ int quick(long long st, long long nw, long long tgt) {
		// e = sojourn - target in ns; the double-millisecond comparison of lib/codel.js:35-37 can
		// only disagree with sign(e) when |e| < m, m = 2 + now>>50 ns (> 2.5 ulp(now_ms))
		const unsigned int mh = (unsigned int)(nw >> 50);
		const long long s = nw - st - (tgt - 2) + (long long)mh; // e + m
		if (s < 0)
			return 0; // e < -m: certainly below target
		if ((unsigned long long)s >= 2ull * (mh + 2u))
			return 1; // e >= m: certainly at/above target
		return 2;
	}
	// the same test against two precomputed bounds (lo = tgt - m, hi = tgt + m for a margin m that
	// is at least every claim's own 2 + now>>50): sojourn < lo -> 0, sojourn >= hi -> 1, else 2
	static __device__ __forceinline__ int quick2(long long st, long long nw, long long lo, long long hi) {
		const long long e = nw - st;
		return e < lo ? 0 : (e >= hi ? 1 : 2);
	}
	static __device__ __forceinline__ long long margin_of(long long hi_or) { return 2 + (hi_or >> 18); }
	static __device__ __forceinline__ long long hi_word(long long x) { return x >> 32; }
	static __device__ __forceinline__ bool less(long long a, long long b) { return a < b; }
};
template <> struct Raw<false> {
	using t = double;
	static __device__ __forceinline__ int quick(double, double, long long) { return 2; }
	static __device__ __forceinline__ int quick2(double, double, long long, long long) { return 2; }
	static __device__ __forceinline__ long long margin_of(long long) { return 0; }
	static __device__ __forceinline__ long long hi_word(double) { return 0; }
	static __device__ __forceinline__ bool less(double a, double b) { return a < b; }
};

template <bool NS>
__device__ __forceinline__ bool is_above(typename Raw<NS>::t st, typename Raw<NS>::t nw, long long tgt_ns,
    double target) {
	const int q = Raw<NS>::quick(st, nw, tgt_ns);
	if (q != 2)
		return q == 1;
	return !(__dsub_rn(TimeIn<NS>::to_ms(nw), TimeIn<NS>::to_ms(st)) < target); // lib/codel.js:35-37
}

template <bool NS>
__device__ __forceinline__ bool is_above2(typename Raw<NS>::t st, typename Raw<NS>::t nw, long long lo, long long hi,
    double target) {
	const int q = Raw<NS>::quick2(st, nw, lo, hi);
	if (q != 2)
		return q == 1;
	return !(__dsub_rn(TimeIn<NS>::to_ms(nw), TimeIn<NS>::to_ms(st)) < target); // lib/codel.js:35-37
}

// value of an index-coded F
template <bool NS>
__device__ __forceinline__ double materialise(unsigned long long code, const void *now, const DevState *g) {
	if (code == CODE_ZERO)
		return 0.0;
	if (code == CODE_CARRY)
		return g ? g->F : 0.0;
	using raw_t = typename TimeIn<NS>::raw_t;
	const raw_t v = reinterpret_cast<const raw_t *>(now)[code - 2];
	return __dadd_rn(TimeIn<NS>::to_ms(v), CODEL_INTERVAL);
}

template <bool NS, bool FLAGS>
__global__ void __launch_bounds__(TPB, CB_SCAN_MINBLOCKS) scan_kernel(const ScanParams p) {
	using raw_t = typename Raw<NS>::t;
	__shared__ unsigned long long s_wagg[WARPS]; // isconst << 63 | code
	__shared__ unsigned int s_wfc[WARPS];        // chunk-local position of the warp's first CONST claim
	__shared__ unsigned int s_evn[WARPS];                 // per-warp event counters (no cross-warp init race)
	__shared__ unsigned long long s_ev[WARPS][EVW];
	__shared__ unsigned int s_leftclosed;                 // the claim before the tile is below target

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned long long t = blockIdx.x;
	if (tid == 0 && blockIdx.x == 0)
		p.b->dbg[8] = gtimer(); // timeline stamp (tools/pipeline_timeline.py)
	const unsigned long long ts = t * (unsigned long long)TILE;
	const unsigned long long cs = ts + (unsigned long long)warp * CHUNK; // chunk start
	const raw_t *gstart = reinterpret_cast<const raw_t *>(p.start);
	const raw_t *gnow = reinterpret_cast<const raw_t *>(p.now);
	const double target = p.target;
	const long long tgt_ns = p.target_ns;
	const unsigned int lt_mask = (1u << lane) - 1u;
	if (lane == 0)
		s_evn[warp] = 0;
	__syncwarp();

	// ---- loads: one 128-bit load per array per row (claims 2l, 2l+1) -----------------------
	raw_t rs[ROWS][2], rn[ROWS][2];
	unsigned int fl[ROWS]; // this lane's two flag bytes per row
	const bool full = (cs + CHUNK <= p.n);
	if (full && p.vec_ok) {
		uint4 a[ROWS], b[ROWS];
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			a[r] = ld_stream_v4(gstart + cs + r * 64 + 2 * lane);
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			b[r] = ld_stream_v4(gnow + cs + r * 64 + 2 * lane);
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			unsigned long long x0 = ((unsigned long long)a[r].y << 32) | a[r].x;
			unsigned long long x1 = ((unsigned long long)a[r].w << 32) | a[r].z;
			unsigned long long y0 = ((unsigned long long)b[r].y << 32) | b[r].x;
			unsigned long long y1 = ((unsigned long long)b[r].w << 32) | b[r].z;
			rs[r][0] = *reinterpret_cast<raw_t *>(&x0);
			rs[r][1] = *reinterpret_cast<raw_t *>(&x1);
			rn[r][0] = *reinterpret_cast<raw_t *>(&y0);
			rn[r][1] = *reinterpret_cast<raw_t *>(&y1);
		}
	} else {
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
#pragma unroll
			for (int j = 0; j < 2; j++) {
				const unsigned long long i = cs + r * 64 + 2 * lane + j;
				rs[r][j] = (i < p.n) ? gstart[i] : raw_t(0);
				rn[r][j] = (i < p.n) ? gnow[i] : raw_t(0);
			}
		}
	}
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		fl[r] = 0x0101u; // waiting, no markers
		if (FLAGS) {
			const unsigned long long i = cs + r * 64 + 2 * lane;
			unsigned int f0 = (i < p.n) ? p.flags[i] : 0u;
			unsigned int f1 = (i + 1 < p.n) ? p.flags[i + 1] : 0u;
			fl[r] = f0 | (f1 << 8);
		}
	}
	// The claim just before this warp's chunk (same address in every lane: one broadcast
	// transaction): its class may end an "ok" run exactly at the chunk boundary, its marker
	// resets F in front of our first claim, and the clock must not step back across it.
	const bool have_prev = (cs > 0) && (cs < p.n);
	raw_t prev_now = raw_t(0);
	unsigned int prev_bits = 0; // bit0 above, bit1 marker-before, bit2 own marker
	if (have_prev) {
		const raw_t pstart = gstart[cs - 1];
		prev_now = gnow[cs - 1];
		const unsigned int pf = FLAGS ? p.flags[cs - 1] : 1u, ppf = (FLAGS && cs >= 2) ? p.flags[cs - 2] : 0u;
		prev_bits = (is_above<NS>(pstart, prev_now, tgt_ns, target) ? 1u : 0u) | ((ppf & F_EMPTY_AFTER) ? 2u : 0u) |
		    ((pf & F_EMPTY_AFTER) ? 4u : 0u);
	}

	// ---- phase 1: above bits, CONST/PASS types (integers only) --------------------------------
	// One margin for all of this lane's claims: the OR of their clocks' high words bounds each
	// claim's own now>>50 from above, and a wider margin only sends a claim within a few ns of the
	// target to the exact double comparison.  (A negative clock makes the margin negative: every
	// claim then takes the exact path.)
	long long q_lo, q_hi;
	{
		long long hi_or = 0;
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			hi_or |= Raw<NS>::hi_word(rn[r][0]) | Raw<NS>::hi_word(rn[r][1]);
		const long long m = Raw<NS>::margin_of(hi_or);
		q_lo = tgt_ns - m;
		q_hi = tgt_ns + m;
		if (m < 2) { // negative clocks: no shortcut
			q_lo = (long long)0x8000000000000000ull;
			q_hi = (long long)0x7fffffffffffffffull;
		}
	}
	unsigned int A0[ROWS], A1[ROWS], C0[ROWS], C1[ROWS]; // ballots: above / CONST-type, per claim slot
	int nv[ROWS];                                        // valid claims in the row
	bool nonmono = false;
	unsigned long long last_marker = 0;
	auto phase1 = [&](auto full_tag) {
		constexpr bool FULL = decltype(full_tag)::value; // every claim of the chunk exists
		unsigned int carry_mk = (prev_bits >> 2) & 1u;   // marker of the claim before the row (warp-uniform)
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			const unsigned long long rbase = cs + r * 64;
			if (FULL) {
				nv[r] = 64;
			} else {
				const long long rem = (long long)p.n - (long long)rbase;
				nv[r] = rem >= 64 ? 64 : (rem > 0 ? (int)rem : 0);
			}
			const bool v0 = FULL || (2 * lane < nv[r]), v1 = FULL || (2 * lane + 1 < nv[r]);
			// integer test first; the exact double comparison only if some claim of the row is within
			// the margin of the target (one warp-uniform branch per row instead of one per claim)
			const int q0 = Raw<NS>::quick2(rs[r][0], rn[r][0], q_lo, q_hi);
			const int q1 = Raw<NS>::quick2(rs[r][1], rn[r][1], q_lo, q_hi);
			bool ab0 = v0 && (q0 == 1), ab1 = v1 && (q1 == 1);
			if (__any_sync(0xffffffffu, (v0 && q0 == 2) || (v1 && q1 == 2))) {
				if (v0 && q0 == 2)
					ab0 = !(__dsub_rn(TimeIn<NS>::to_ms(rn[r][0]), TimeIn<NS>::to_ms(rs[r][0])) < target); // lib/codel.js:35-37
				if (v1 && q1 == 2)
					ab1 = !(__dsub_rn(TimeIn<NS>::to_ms(rn[r][1]), TimeIn<NS>::to_ms(rs[r][1])) < target);
			}
			A0[r] = __ballot_sync(0xffffffffu, ab0);
			A1[r] = __ballot_sync(0xffffffffu, ab1);
			if (FLAGS) {
				const unsigned int mk0 = (fl[r] >> 1) & 1u, mk1 = (fl[r] >> 9) & 1u; // own markers
				unsigned int up = __shfl_up_sync(0xffffffffu, mk1, 1);
				if (lane == 0)
					up = carry_mk;
				carry_mk = __shfl_sync(0xffffffffu, mk1, 31);
				if (v1 && mk1)
					last_marker = rbase + 2 * lane + 2;
				else if (v0 && mk0)
					last_marker = rbase + 2 * lane + 1;
				C0[r] = __ballot_sync(0xffffffffu, v0 && (!ab0 || up != 0));
				C1[r] = __ballot_sync(0xffffffffu, v1 && (!ab1 || mk0 != 0));
			} else if (FULL) {
				C0[r] = ~A0[r]; // without markers a claim is CONST-type iff it is below target
				C1[r] = ~A1[r];
			} else {
				C0[r] = __ballot_sync(0xffffffffu, v0) & ~A0[r];
				C1[r] = __ballot_sync(0xffffffffu, v1) & ~A1[r];
			}
			// monotone clock (raw domain: ns-monotone implies ms-monotone)
			// the claim before this lane's pair: lane l-1's second claim; for lane 0 it is lane 31's
			// second claim of the previous row (or the claim before the chunk) -- one rotated shuffle
			// in which lane 31 (read by lane 0 only) offers that value
			const raw_t offer = (lane == 31) ? (r > 0 ? rn[r > 0 ? r - 1 : 0][1] : prev_now) : rn[r][1];
			const raw_t prevn = __shfl_sync(0xffffffffu, offer, (lane + 31) & 31);
			const bool haveprevn = (lane > 0) || (r > 0) || have_prev;
			if ((v1 && Raw<NS>::less(rn[r][1], rn[r][0])) || (v0 && haveprevn && Raw<NS>::less(rn[r][0], prevn)))
				nonmono = true;
		}
	};
	if (full)
		phase1(std::true_type{});
	else
		phase1(std::false_type{});

	// ---- chunk composition in index form, straight from the ballots -------------------------------
	IAgg wagg;
	unsigned int wfc = FC_NONE; // chunk-local position of the first CONST claim
	{
		unsigned int cm_l = 0, c1_l = 0, a0_l = 0, a1_l = 0, cm_f = 0, c0_f = 0;
		int r_l = -1, r_f = -1, nvt = 0;
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			const unsigned int cm = C0[r] | C1[r];
			nvt += nv[r];
			if (cm) {
				if (r_f < 0) {
					r_f = r;
					cm_f = cm;
					c0_f = C0[r];
				}
				r_l = r;
				cm_l = cm;
				c1_l = C1[r];
				a0_l = A0[r];
				a1_l = A1[r];
			}
		}
		if (r_l >= 0) {
			const int L = 31 - __clz(cm_l);
			const int j = (c1_l >> L) & 1u;
			const int pos = r_l * 64 + 2 * L + j;             // last CONST claim of the chunk
			const bool abp = ((j ? a1_l : a0_l) >> L) & 1u;   // CONST and above => a marker precedes it
			wagg.isconst = true;
			wagg.code = abp ? (cs + pos + 2) : CODE_ZERO;
			if (wagg.code == CODE_ZERO && pos + 1 < nvt)
				wagg.code = cs + pos + 1 + 2; // the claim after it is a PASS (above) claim and sets F
			const int Lf = __ffs(cm_f) - 1;
			wfc = r_f * 64 + 2 * Lf + (((c0_f >> Lf) & 1u) ? 0 : 1);
		} else {
			wagg.isconst = false;
			wagg.code = nvt > 0 ? (cs + 2) : CODE_ZERO;
		}
	}
	if (lane == 0) {
		s_wagg[warp] = ((unsigned long long)wagg.isconst << 63) | wagg.code;
		s_wfc[warp] = wfc;
		if (warp == 0)
			s_leftclosed = (t > 0 && !(prev_bits & 1u)) ? 1u : 0u;
	}
	// Only warps whose leading claims inherit F from an earlier warp (and warp 0, which writes the
	// tile's record) wait for the others; everybody else just signals and moves on.
	const bool first_is_pass = nv[0] > 0 && !((C0[0]) & 1u);
	const bool prev_above = have_prev && (prev_bits & 1u) && !(prev_bits & 2u);
	const bool need_wex = (warp == 0) || first_is_pass || prev_above;
	IAgg wex;
	wex.isconst = false;
	wex.code = CODE_ZERO;
	bool left_closed = false;
	if (need_wex) {
		bar_sync(1);
		left_closed = s_leftclosed != 0u;
		// nearest earlier warp with a CONST claim: lane w looks at warp w's aggregate, one ballot
		const unsigned long long xw = (lane < WARPS) ? s_wagg[lane] : 0ull;
		const unsigned int cmask = __ballot_sync(0xffffffffu, (lane < warp) && (xw >> 63));
		if (cmask) {
			const int w = 31 - __clz(cmask);
			wex.isconst = true;
			wex.code = __shfl_sync(0xffffffffu, xw, w) & DESC_CODE_MASK;
			const unsigned long long nxt = __shfl_sync(0xffffffffu, xw, (w + 1) & 31) & DESC_CODE_MASK;
			if (wex.code == CODE_ZERO && w + 1 < warp)
				wex.code = nxt; // the PASS warp after it sets F
		} else if (warp > 0) {
			wex.code = __shfl_sync(0xffffffffu, xw, 0) & DESC_CODE_MASK; // all earlier warps are PASS: the first one's claim
		}
	} else {
		bar_arrive(1);
	}

	// ---- the tile's record (thread 0): where the inherited run closes, what the tile leaves open
	if (tid == 0) {
		IAgg tile_agg;
		tile_agg.isconst = false;
		tile_agg.code = CODE_ZERO;
		unsigned int fc = FC_NONE;
#pragma unroll
		for (int w = 0; w < WARPS; w++) {
			IAgg a;
			const unsigned long long x = s_wagg[w];
			a.isconst = (x >> 63) != 0;
			a.code = x & DESC_CODE_MASK;
			tile_agg = icompose(tile_agg, a);
			if (fc == FC_NONE && s_wfc[w] != FC_NONE)
				fc = w * CHUNK + s_wfc[w];
		}
		if (left_closed)
			fc = 0; // nothing crosses the tile's left edge: it anchors itself (resolve sees an empty stretch)
		p.rec_fc[t] = fc;
		p.rec_code[t] = tile_agg.code;
		reinterpret_cast<raw_t *>(p.tile_now0)[t] = rn[0][0]; // thread 0 owns the tile's first claim
	}

	// ---- phase 2: classes and events, FP64 only in rows with above-target claims ---------------
	// `known`: F before the next claim is determined inside this tile.  Leading PASS claims of
	// the tile (before its first CONST claim) are left to the resolve kernel.
	// `known`: F before the chunk's first claim is determined inside this tile.
	//   - an earlier warp of the tile has a CONST claim (wex.isconst), or
	//   - the tile is left-closed (the claim before it is below target, so F is 0 at its edge and
	//     the tile's first claim opens whatever run there is), or
	//   - this warp needs nothing from before (its first claim is CONST-type, previous claim not ok)
	bool known = wex.isconst || left_closed || !need_wex || (warp == 0 && !first_is_pass && !prev_above);
	const bool prev_ok_possible = known && prev_above;
	double F = 0.0; // F in force before the chunk's first claim (before its marker-before reset)
	if ((wex.isconst || (left_closed && warp > 0)) && (first_is_pass || prev_ok_possible))
		F = materialise<NS>(wex.code, p.now, nullptr); // never CODE_CARRY inside a tile
	int c_last = 0; // class of the claim before the current row
	if (prev_ok_possible && F != 0.0)
		c_last = claim_class(true, TimeIn<NS>::to_ms(prev_now), F);
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		if (nv[r] == 0)
			break;
		const unsigned long long rbase = cs + r * 64;
		const unsigned int cmr = C0[r] | C1[r];
		if (!known && cmr == 0u)
			continue; // still inside the run inherited from an earlier tile: nothing to classify
		// first classifiable position of the row (64 = none masked): the row's first CONST claim
		int fpos = 0;
		if (!known) {
			const int Lf = __ffs(cmr) - 1;
			fpos = 2 * Lf + (((C0[r] >> Lf) & 1u) ? 0 : 1);
		}
		known = true;
		if ((A0[r] | A1[r]) == 0u) { // nobody above target: every class is 0, F ends at 0
			if (c_last != 0 && lane == 0) {
				unsigned int pos = atomicAdd(&s_evn[warp], 1u);
				if (pos < EVW)
					s_ev[warp][pos] = (rbase << 2);
			}
			c_last = 0;
			F = 0.0;
			continue;
		}
		const bool v0 = 2 * lane < nv[r], v1 = 2 * lane + 1 < nv[r];
		const bool ab0 = (A0[r] >> lane) & 1u, ab1 = (A1[r] >> lane) & 1u;
		const bool k0 = (C0[r] >> lane) & 1u, k1 = (C1[r] >> lane) & 1u; // CONST-type
		const double n0 = TimeIn<NS>::to_ms(rn[r][0]), n1 = TimeIn<NS>::to_ms(rn[r][1]);
		// lane summary: F after my two claims if F were 0 before them
		double lv;
		if (k1) {
			lv = ab1 ? __dadd_rn(n1, CODEL_INTERVAL) : 0.0;
		} else if (k0) {
			lv = ab0 ? __dadd_rn(n0, CODEL_INTERVAL) : 0.0;
			if (lv == 0.0 && v1)
				lv = __dadd_rn(n1, CODEL_INTERVAL);
		} else {
			lv = v0 ? __dadd_rn(n0, CODEL_INTERVAL) : 0.0;
		}
		// F before my claim 0: last CONST lane before me, then the PASS lane after it
		const unsigned int cm = C0[r] | C1[r];
		const unsigned int prior = cm & lt_mask;
		const int Lq = prior ? (31 - __clz(prior)) : 0;
		const double vq = __shfl_sync(0xffffffffu, lv, Lq);
		const int nx = prior ? (Lq + 1) : 0;
		const double vnx = __shfl_sync(0xffffffffu, lv, nx & 31);
		double Fp;
		if (prior)
			Fp = (vq != 0.0) ? vq : ((nx < lane) ? vnx : 0.0);
		else
			Fp = (F != 0.0) ? F : ((lane > 0) ? vnx : 0.0);
		// my two claims in order
		int c0 = 0, c1 = 0;
		{
			const double Fuse = (k0 && ab0) ? 0.0 : Fp; // CONST and above = a marker resets F first
			c0 = v0 ? claim_class(ab0, n0, Fuse) : 0;
			Fp = v0 ? (ab0 ? ((Fuse != 0.0) ? Fuse : __dadd_rn(n0, CODEL_INTERVAL)) : 0.0) : Fp;
		}
		{
			const double Fuse = (k1 && ab1) ? 0.0 : Fp;
			c1 = v1 ? claim_class(ab1, n1, Fuse) : c0;
			Fp = v1 ? (ab1 ? ((Fuse != 0.0) ? Fuse : __dadd_rn(n1, CODEL_INTERVAL)) : 0.0) : Fp;
		}
		if (!v0 || 2 * lane < fpos)
			c0 = 0; // unclassified (inherited run): no class, no event
		if (2 * lane + 1 < fpos)
			c1 = 0;
		// events
		int cprev = __shfl_up_sync(0xffffffffu, v1 ? c1 : c0, 1);
		if (lane == 0)
			cprev = c_last;
		if (v0 && c0 != cprev) {
			unsigned int pos = atomicAdd(&s_evn[warp], 1u);
			if (pos < EVW)
				s_ev[warp][pos] = ((rbase + 2 * lane) << 2) | (unsigned long long)c0;
		}
		if (v1 && c1 != c0) {
			unsigned int pos = atomicAdd(&s_evn[warp], 1u);
			if (pos < EVW)
				s_ev[warp][pos] = ((rbase + 2 * lane + 1) << 2) | (unsigned long long)c1;
		}
		// carry to the next row: state after the row's last valid claim
		const int Lend = (nv[r] - 1) >> 1;
		F = __shfl_sync(0xffffffffu, Fp, Lend);
		c_last = __shfl_sync(0xffffffffu, (nv[r] & 1) ? c0 : c1, Lend);
	}

	// class of the tile's last claim (the warp that owns it), for the run the tile leaves open
	{
		unsigned long long te = ts + TILE;
		if (te > p.n)
			te = p.n;
		if (lane == 0 && cs < te && te <= cs + CHUNK)
			p.rec_cend[t] = (unsigned int)c_last;
	}
	if (FLAGS) {
		unsigned long long m = last_marker;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const unsigned long long y = __shfl_xor_sync(0xffffffffu, m, o);
			m = y > m ? y : m;
		}
		if (lane == 0 && m)
			atomicMax(&p.b->last_marker_p1, m);
	}
	if (nonmono)
		p.b->nonmono = 1u;
	// ---- flush: warp 0 gathers every warp's events into the tile's slots ------------------------
	if (warp != 0) {
		bar_arrive(2);
		return;
	}
	bar_sync(2);
	{
		unsigned int cnt = 0;
		bool over = false;
#pragma unroll
		for (int w = 0; w < WARPS; w++) {
			const unsigned int cw = s_evn[w];
			over |= cw > (unsigned int)EVW;
			const unsigned int use = cw > (unsigned int)EVW ? (unsigned int)EVW : cw;
			if ((unsigned int)lane < use) {
				if (cnt + lane < (unsigned int)EVK)
					p.ev[t * EVK + cnt + lane] = s_ev[w][lane];
			}
			cnt += use;
		}
		if (cnt > (unsigned int)EVK) {
			over = true;
			cnt = EVK;
		}
		if (lane == 0) {
			if (over)
				p.b->overflow = 1u;
			p.ev_count[t] = cnt;
			if (t + 1 == gridDim.x)
				p.b->dbg[9] = gtimer(); // the last tile's CTA is among the last to retire
		}
	}
}

} // namespace cb
