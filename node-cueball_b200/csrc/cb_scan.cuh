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
// INDICES (integer-only, one 64-bit code per chunk and per tile) and turned into doubles
// only where a claim's class needs the value.
//
// Execution (round 2): every WARP streams tiles of its own.  A tile is 2048 consecutive claims,
// walked as 8 chunks of 256 = 4 rows of 64; in a row lane l owns claims 2l and 2l+1.  The chunks
// of a warp's tiles flow through a private ring of SC_NST shared-memory stages filled by TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx, issued by the warp's own lane 0 two chunks ahead of
// the one being worked on), so the HBM stream never waits for the arithmetic and no warp ever
// waits for another: there is no CTA-wide barrier, no cross-warp exchange, and what the
// tile-level composition needs from the previous chunk is simply still in registers.  A stage
// also holds the two claims in front of its chunk (16 B): the clock must not step back across a
// chunk boundary and the class of the claim before a tile decides whether the tile anchors itself.
// Phase 1 (integers only) classifies above/below -- sojourn vs target is decided on the int64
// nanoseconds whenever the gap to the target exceeds the worst-case rounding of the millisecond
// doubles, and exactly as utils.currentMillis() (lib/utils.js:198-204) otherwise -- and builds the
// chunk compositions from ballots.  Phase 2 touches FP64 only in rows that contain above-target
// claims whose F is determined inside the tile.
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

constexpr int ROWS = 4;               // rows of 64 claims per chunk
constexpr int CHUNK = ROWS * 64;      // 256 claims: one stage of a warp's ring
constexpr int CHUNKS = TILE / CHUNK;  // 8 chunks per tile
static_assert(CHUNK * CHUNKS == TILE, "tile geometry");

// streaming geometry (tuning knobs: `make -B EXTRA="-DCB_SCAN_WARPS=.. -DCB_SCAN_NST=.. -DCB_SCAN_MINBLOCKS=.."`)
#ifndef CB_SCAN_WARPS
#define CB_SCAN_WARPS 8 // warps per CTA; each streams tiles of its own
#endif
#ifndef CB_SCAN_NST
#define CB_SCAN_NST 3 // chunk stages per warp
#endif
#ifndef CB_SCAN_MINBLOCKS
#define CB_SCAN_MINBLOCKS 2 // resident CTAs per SM the register budget is sized for
#endif
constexpr int SC_WARPS = CB_SCAN_WARPS;
constexpr int SC_TPB = SC_WARPS * 32;
constexpr int SC_NST = CB_SCAN_NST;
constexpr int SC_LEAD = 16;                      // the two claims in front of the chunk
constexpr int SC_ARR = SC_LEAD + CHUNK * 8;      // one array of a stage: lead + 256 x 8 B
constexpr int SC_STAGE = 2 * SC_ARR;             // start array, then now array
constexpr int SC_SMEM = SC_WARPS * SC_NST * SC_STAGE; // dynamic shared memory of a scan CTA

// index-coded F: 0 = F is 0, 1 = the batch's carried-in F, a+2 = now[a] + 100
constexpr unsigned long long CODE_ZERO = 0, CODE_CARRY = 1;
constexpr unsigned long long DESC_CODE_MASK = (1ull << 40) - 1;
constexpr unsigned int FC_NONE = 0xffffffffu; // tile record: no CONST claim in the tile

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

#ifndef CB_EMU // tests/emu/cuda_emu.h supplies host versions
__device__ __forceinline__ unsigned int sc_smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, unsigned int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { // make the initialised barriers visible to the async proxy
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *mbar, unsigned int bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sc_smem_u32(mbar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (16-byte aligned on both sides, size a multiple of 16); streamed data:
// evict-first in L2
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *mbar,
    unsigned long long policy) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
	                 sc_smem_u32(dst)),
	    "l"(src), "r"(bytes), "r"(sc_smem_u32(mbar)), "l"(policy)
	    : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, unsigned int parity) {
	unsigned int done = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred q;\n\t"
		             "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
		             "selp.u32 %0, 1, 0, q;\n\t}"
		             : "=r"(done)
		             : "r"(sc_smem_u32(mbar)), "r"(parity)
		             : "memory");
	}
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
	unsigned long long *ev;   // [tiles][EVK]  (index << 2) | class, in index order
	int vec_ok;               // start/now 16-byte aligned: chunks are staged by TMA bulk copies
	long long target_ns;      // target in ns (integer-domain sojourn test)
};

template <typename T> struct alignas(16) Pair {
	T x, y;
};

template <bool NS> struct Raw;
template <> struct Raw<true> {
	using t = long long;
	static __device__ __forceinline__ long long lowest() { return (long long)0x8000000000000000ull; }
};
template <> struct Raw<false> {
	using t = double;
	static __device__ __forceinline__ double lowest() { return -1.0 / 0.0; }
};

// lib/codel.js:35-37 in the clock's own unit: sojourn >= target, compared as the reference compares it
template <bool NS>
__device__ __forceinline__ bool is_above(typename Raw<NS>::t st, typename Raw<NS>::t nw, double target) {
	return !(__dsub_rn(TimeIn<NS>::to_ms(nw), TimeIn<NS>::to_ms(st)) < target);
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

// ---- one chunk of 256 claims, already in the stage `st` --------------------------------------
// Tile-level state a warp carries from chunk to chunk (all lanes hold the same values).
struct TileState {
	IAgg agg;             // composition of the chunks seen so far
	unsigned int fc;      // tile-local position of the first CONST claim, or FC_NONE
	bool known;           // F in front of the next claim is determined inside this tile
	double F;             // ... its value (meaningful when known)
	int c_last;           // class of the claim before the next row
	unsigned int evn;     // class-change events found so far
	unsigned int carry_mk; // FLAGS: the claim before the next row carries an empty() marker
};

template <bool NS, bool FLAGS, bool FULL>
__device__ __forceinline__ void scan_chunk(const ScanParams &p, const unsigned char *st, unsigned long long t,
    unsigned long long cs, int sub, int lane, unsigned int lt_mask, TileState &T, bool &nonmono,
    unsigned long long &last_marker) {
	using raw_t = typename Raw<NS>::t;
	using pair_t = Pair<raw_t>;
	const raw_t *ss = reinterpret_cast<const raw_t *>(st + SC_LEAD);
	const raw_t *sn = reinterpret_cast<const raw_t *>(st + SC_ARR + SC_LEAD);
	const double target = p.target;
	int nvt = CHUNK; // valid claims in the chunk
	if (!FULL) {
		const unsigned long long rem = p.n - cs;
		nvt = rem >= (unsigned long long)CHUNK ? CHUNK : (int)rem;
	}
	// FLAGS: this lane's two flag bytes per row, straight from global memory (1 B/claim)
	unsigned int fl[ROWS];
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		fl[r] = 0x0101u; // waiting, no markers
		if (FLAGS) {
			const unsigned long long i = cs + r * 64 + 2 * lane;
			const unsigned int f0 = (FULL || i < p.n) ? p.flags[i] : 0u;
			const unsigned int f1 = (FULL || i + 1 < p.n) ? p.flags[i + 1] : 0u;
			fl[r] = f0 | (f1 << 8);
		}
	}

	// ---- phase 1: above bits (integers only on the nanosecond clock) ----------------------------
	// e = sojourn - target in ns.  The double-millisecond comparison of lib/codel.js:35-37 can only
	// disagree with sign(e) when |e| < m_i = 2 + now_i>>50 ns (> 2.5 ulp(now_ms)).  One margin for the
	// chunk: its last clock bounds every now_i from above when the clock is monotone (and when it is
	// not, the batch takes the exact sequential path and nothing computed here is used).  A negative
	// first clock switches the shortcut off.  d = e + m: d < 0 below, d >= 2m above, else exact.
	unsigned int mine = 0; // bit 2r / 2r+1: this lane's claim of row r, slot 0 / 1, is at or above target
	bool near = false;
	long long lo = 0;
	unsigned long long W = 0;
	if (NS) {
		const long long nfirst = *reinterpret_cast<const long long *>(sn);
		const long long nlast = *reinterpret_cast<const long long *>(sn + (nvt - 1));
		const long long m = 2 + (nlast >> 50);
		lo = p.target_ns - m;
		W = 2ull * (unsigned long long)m;
		if (nfirst < 0 || nlast < 0) { // a negative clock: no shortcut, every claim is compared exactly
			W = ~0ull;
			near = true;
		}
	}
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		const pair_t S = *reinterpret_cast<const pair_t *>(ss + r * 64 + 2 * lane);
		const pair_t N = *reinterpret_cast<const pair_t *>(sn + r * 64 + 2 * lane);
		const raw_t P = sn[r * 64 + 2 * lane - 1]; // the claim before this lane's pair (the lead for lane 0 of row 0)
		const bool v0 = FULL || (r * 64 + 2 * lane < nvt), v1 = FULL || (r * 64 + 2 * lane + 1 < nvt);
		bool ab0, ab1;
		if (NS) {
			const long long d0 = (long long)N.x - (long long)S.x - lo, d1 = (long long)N.y - (long long)S.y - lo;
			ab0 = d0 >= 0; // provisional for a near claim: redone below
			ab1 = d1 >= 0;
			near |= (v0 && (unsigned long long)d0 < W) || (v1 && (unsigned long long)d1 < W);
		} else {
			ab0 = is_above<NS>(S.x, N.x, target);
			ab1 = is_above<NS>(S.y, N.y, target);
		}
		mine |= ((v0 && ab0) ? 1u : 0u) << (2 * r);
		mine |= ((v1 && ab1) ? 1u : 0u) << (2 * r + 1);
		// monotone clock (raw domain: ns-monotone implies ms-monotone)
		if ((v0 && N.x < P) || (v1 && N.y < N.x))
			nonmono = true;
	}
	if (NS && __any_sync(0xffffffffu, near)) { // some claim within the margin of the target: exact comparison
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			const pair_t S = *reinterpret_cast<const pair_t *>(ss + r * 64 + 2 * lane);
			const pair_t N = *reinterpret_cast<const pair_t *>(sn + r * 64 + 2 * lane);
			const bool v0 = FULL || (r * 64 + 2 * lane < nvt), v1 = FULL || (r * 64 + 2 * lane + 1 < nvt);
			const long long d0 = (long long)N.x - (long long)S.x - lo, d1 = (long long)N.y - (long long)S.y - lo;
			if (v0 && ((unsigned long long)d0 < W || W == ~0ull)) {
				const bool ab = is_above<NS>(S.x, N.x, target); // lib/codel.js:35-37
				mine = (mine & ~(1u << (2 * r))) | ((ab ? 1u : 0u) << (2 * r));
			}
			if (v1 && ((unsigned long long)d1 < W || W == ~0ull)) {
				const bool ab = is_above<NS>(S.y, N.y, target);
				mine = (mine & ~(2u << (2 * r))) | ((ab ? 2u : 0u) << (2 * r));
			}
		}
	}
	// ---- uniform chunks (nearly all of them): no ballots, no per-row work ----------------------------
	if (FULL && !FLAGS) {
		if (__all_sync(0xffffffffu, mine == 0u)) {
			// every claim is below target: CONST(0) throughout, class 0, F ends at 0
			T.agg.isconst = true;
			T.agg.code = CODE_ZERO;
			if (T.fc == FC_NONE)
				T.fc = (unsigned int)(sub * CHUNK);
			if (T.c_last != 0) { // the run that was open closes at the chunk's first claim
				if (lane == 0 && T.evn < (unsigned int)EVK)
					p.ev[t * EVK + T.evn] = (cs << 2);
				T.evn++;
			}
			T.known = true;
			T.c_last = 0;
			T.F = 0.0;
			return;
		}
		if (!T.known && __all_sync(0xffffffffu, mine == 0xffu)) {
			// every claim is above target and the run was opened in an earlier tile: PASS throughout
			if (T.agg.code == CODE_ZERO)
				T.agg.code = cs + 2; // the chunk's first claim fills a zero F
			return;
		}
	}
	unsigned int A0[ROWS], A1[ROWS]; // ballots: claim slot 0 / 1 is at or above target
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		A0[r] = __ballot_sync(0xffffffffu, (mine >> (2 * r)) & 1u);
		A1[r] = __ballot_sync(0xffffffffu, (mine >> (2 * r + 1)) & 1u);
	}
	// ---- CONST / PASS types ------------------------------------------------------------------------
	unsigned int C0[ROWS], C1[ROWS]; // ballots: the claim is CONST-type (below target, or a marker precedes it)
	int nv[ROWS];                    // valid claims in the row
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		const unsigned long long rbase = cs + r * 64;
		if (FULL) {
			nv[r] = 64;
		} else {
			const int rem = nvt - r * 64;
			nv[r] = rem >= 64 ? 64 : (rem > 0 ? rem : 0);
		}
		const bool v0 = FULL || (2 * lane < nv[r]), v1 = FULL || (2 * lane + 1 < nv[r]);
		if (FLAGS) {
			const unsigned int mk0 = (fl[r] >> 1) & 1u, mk1 = (fl[r] >> 9) & 1u; // own markers
			unsigned int up = __shfl_up_sync(0xffffffffu, mk1, 1);
			if (lane == 0)
				up = T.carry_mk;
			T.carry_mk = __shfl_sync(0xffffffffu, mk1, 31);
			if (v1 && mk1)
				last_marker = rbase + 2 * lane + 2;
			else if (v0 && mk0)
				last_marker = rbase + 2 * lane + 1;
			const bool ab0 = (A0[r] >> lane) & 1u, ab1 = (A1[r] >> lane) & 1u;
			C0[r] = __ballot_sync(0xffffffffu, v0 && (!ab0 || up != 0));
			C1[r] = __ballot_sync(0xffffffffu, v1 && (!ab1 || mk0 != 0));
		} else if (FULL) {
			C0[r] = ~A0[r]; // without markers a claim is CONST-type iff it is below target
			C1[r] = ~A1[r];
		} else {
			C0[r] = __ballot_sync(0xffffffffu, v0) & ~A0[r];
			C1[r] = __ballot_sync(0xffffffffu, v1) & ~A1[r];
		}
	}

	// ---- chunk composition in index form, straight from the ballots -------------------------------
	{
		IAgg wagg;
		unsigned int wfc = FC_NONE; // chunk-local position of the first CONST claim
		unsigned int cm_l = 0, c1_l = 0, a0_l = 0, a1_l = 0, cm_f = 0, c0_f = 0;
		int r_l = -1, r_f = -1;
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			const unsigned int cm = C0[r] | C1[r];
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
			wagg.code = cs + 2; // nvt > 0: the chunk's first claim fills a zero F
		}
		T.agg = icompose(T.agg, wagg);
		if (T.fc == FC_NONE && wfc != FC_NONE)
			T.fc = (unsigned int)(sub * CHUNK) + wfc;
	}

	// ---- phase 2: classes and events, FP64 only in rows with above-target claims ---------------
	// Leading PASS claims of the tile (before its first CONST claim) are left to the resolve kernel.
#pragma unroll
	for (int r = 0; r < ROWS; r++) {
		if (!FULL && nv[r] == 0)
			break;
		const unsigned long long rbase = cs + r * 64;
		const unsigned int cmr = C0[r] | C1[r];
		if (!T.known && cmr == 0u)
			continue; // still inside the run inherited from an earlier tile: nothing to classify
		// first classifiable position of the row: the row's first CONST claim
		int fpos = 0;
		if (!T.known) {
			const int Lf = __ffs(cmr) - 1;
			fpos = 2 * Lf + (((C0[r] >> Lf) & 1u) ? 0 : 1);
		}
		T.known = true;
		if ((A0[r] | A1[r]) == 0u) { // nobody above target: every class is 0, F ends at 0
			if (T.c_last != 0) {
				if (lane == 0 && T.evn < (unsigned int)EVK)
					p.ev[t * EVK + T.evn] = (rbase << 2);
				T.evn++;
			}
			T.c_last = 0;
			T.F = 0.0;
			continue;
		}
		const pair_t N = *reinterpret_cast<const pair_t *>(sn + r * 64 + 2 * lane);
		const bool v0 = FULL || (2 * lane < nv[r]), v1 = FULL || (2 * lane + 1 < nv[r]);
		const bool ab0 = (A0[r] >> lane) & 1u, ab1 = (A1[r] >> lane) & 1u;
		const bool k0 = (C0[r] >> lane) & 1u, k1 = (C1[r] >> lane) & 1u; // CONST-type
		const double n0 = TimeIn<NS>::to_ms(N.x), n1 = TimeIn<NS>::to_ms(N.y);
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
		const unsigned int prior = cmr & lt_mask;
		const int Lq = prior ? (31 - __clz(prior)) : 0;
		const double vq = __shfl_sync(0xffffffffu, lv, Lq);
		const int nx = prior ? (Lq + 1) : 0;
		const double vnx = __shfl_sync(0xffffffffu, lv, nx & 31);
		double Fp;
		if (prior)
			Fp = (vq != 0.0) ? vq : ((nx < lane) ? vnx : 0.0);
		else
			Fp = (T.F != 0.0) ? T.F : ((lane > 0) ? vnx : 0.0);
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
		// events, in index order
		int cprev = __shfl_up_sync(0xffffffffu, v1 ? c1 : c0, 1);
		if (lane == 0)
			cprev = T.c_last;
		const bool e0 = v0 && c0 != cprev, e1 = v1 && c1 != c0;
		const unsigned int E0 = __ballot_sync(0xffffffffu, e0), E1 = __ballot_sync(0xffffffffu, e1);
		if (E0 | E1) {
			unsigned int pos = T.evn + __popc(E0 & lt_mask) + __popc(E1 & lt_mask);
			if (e0) {
				if (pos < (unsigned int)EVK)
					p.ev[t * EVK + pos] = ((rbase + 2 * lane) << 2) | (unsigned long long)c0;
				pos++;
			}
			if (e1 && pos < (unsigned int)EVK)
				p.ev[t * EVK + pos] = ((rbase + 2 * lane + 1) << 2) | (unsigned long long)c1;
			T.evn += __popc(E0) + __popc(E1);
		}
		// carry to the next row: state after the row's last valid claim
		const int Lend = (nv[r] - 1) >> 1;
		T.F = __shfl_sync(0xffffffffu, Fp, Lend);
		T.c_last = __shfl_sync(0xffffffffu, (nv[r] & 1) ? c0 : c1, Lend);
	}
}

template <bool NS, bool FLAGS>
__global__ void __launch_bounds__(SC_TPB, CB_SCAN_MINBLOCKS) scan_kernel(const ScanParams p) {
	using raw_t = typename Raw<NS>::t;
	CB_DYN_SHARED(unsigned char, s_dyn); // SC_WARPS x SC_NST stages
	__shared__ __align__(8) unsigned long long s_mbar[SC_WARPS * SC_NST];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0 && blockIdx.x == 0)
		p.b->dbg[8] = gtimer(); // timeline stamp (tools/pipeline_timeline.py)
	// chunk and tile numbers fit 32 bits (n < 2^39); only addresses are 64-bit
	const unsigned int tiles = (unsigned int)((p.n + TILE - 1) / TILE);
	const unsigned int nfull = (unsigned int)(p.n / CHUNK);                 // chunks with all 256 claims
	const unsigned int nchunks = (unsigned int)((p.n + CHUNK - 1) / CHUNK); // chunks with any claim
	// warp gw walks tiles gw, gw + GW, gw + 2 GW, ...; consecutive tiles go to consecutive CTAs, so a
	// batch with fewer tiles than warps still spreads over every SM
	const unsigned int gw = (unsigned int)warp * gridDim.x + blockIdx.x;
	const unsigned int GW = gridDim.x * SC_WARPS;
	if (gw >= tiles)
		return;
	unsigned char *wbase = s_dyn + (size_t)warp * SC_NST * SC_STAGE;
	unsigned long long *mbar = s_mbar + warp * SC_NST;
	const unsigned int lt_mask = (1u << lane) - 1u;
	const bool vec = p.vec_ok != 0;
	unsigned long long policy = 0;
	if (lane == 0) {
#pragma unroll
		for (int s = 0; s < SC_NST; s++)
			mbar_init(&mbar[s], 1);
		mbar_init_fence();
		policy = l2_evict_first_policy();
	}
	__syncwarp();

	// lane 0: start the bulk copies of chunk `c` (a batch-wide chunk number) into stage sg; a partial
	// chunk or unaligned arrays are filled by hand when their turn comes
	auto issue = [&](unsigned int c, int sg) {
		if (!vec || c >= nfull)
			return;
		unsigned char *st = wbase + sg * SC_STAGE;
		const unsigned int lead = c ? (unsigned int)SC_LEAD : 0u; // the two claims in front (none before claim 0)
		const unsigned int bytes = (unsigned int)(CHUNK * 8) + lead;
		const size_t off = (size_t)c * (CHUNK * 8) - lead;
		mbar_expect_tx(&mbar[sg], 2u * bytes);
		bulk_g2s(st + SC_LEAD - lead, reinterpret_cast<const unsigned char *>(p.start) + off, bytes, &mbar[sg], policy);
		bulk_g2s(st + SC_ARR + SC_LEAD - lead, reinterpret_cast<const unsigned char *>(p.now) + off, bytes, &mbar[sg], policy);
	};
	// the chunk SC_NST ahead of (tile t, chunk sub) in this warp's walk
	auto ahead = [&](unsigned int t, int sub) {
		const int s2 = sub + SC_NST;
		return s2 < CHUNKS ? t * CHUNKS + (unsigned int)s2
		                   : (t + GW < tiles ? (t + GW) * CHUNKS + (unsigned int)(s2 - CHUNKS) : 0xffffffffu);
	};
	static_assert(SC_NST <= CHUNKS, "the prologue stays inside the first tile");
	if (lane == 0) {
#pragma unroll
		for (int s = 0; s < SC_NST; s++)
			issue(gw * CHUNKS + (unsigned int)s, s);
	}

	unsigned int phases = 0; // parity each stage's barrier completes next
	int sg = 0;
	bool nonmono = false;
	unsigned long long last_marker = 0;
	const raw_t *gstart = reinterpret_cast<const raw_t *>(p.start);
	const raw_t *gnow = reinterpret_cast<const raw_t *>(p.now);
	for (unsigned int t = gw; t < tiles; t += GW) {
		const unsigned long long ts = (unsigned long long)t * TILE;
		TileState T;
		T.agg.isconst = false;
		T.agg.code = CODE_ZERO;
		T.fc = FC_NONE;
		T.known = false;
		T.F = 0.0;
		T.c_last = 0;
		T.evn = 0;
		T.carry_mk = 0;
		bool left_closed = false;
		raw_t first_now = raw_t(0);
#pragma unroll 1
		for (int sub = 0; sub < CHUNKS; sub++) {
			const unsigned int c = t * CHUNKS + (unsigned int)sub;
			if (c >= nchunks)
				break; // only in the batch's last tile: nothing follows
			const unsigned long long cs = ts + (unsigned long long)sub * CHUNK;
			unsigned char *st = wbase + sg * SC_STAGE;
			const bool full = c < nfull;
			if (vec && full) {
				mbar_wait(&mbar[sg], (phases >> sg) & 1u);
				phases ^= 1u << sg;
				if (c == 0 && lane == 0) { // nothing precedes claim 0: a clock no claim can step back from
					*reinterpret_cast<raw_t *>(st + SC_ARR + SC_LEAD - 8) = Raw<NS>::lowest();
					mbar_init_fence(); // ordered before the bulk copy that later refills this stage (async proxy)
				}
			} else {
				// by hand: the two claims in front, the valid claims, zeros behind them
				raw_t *ds = reinterpret_cast<raw_t *>(st), *dn = reinterpret_cast<raw_t *>(st + SC_ARR);
				for (int j = lane; j < CHUNK + 2; j += 32) {
					const long long i = (long long)cs - 2 + j;
					const bool ok = i >= 0 && (unsigned long long)i < p.n;
					ds[j] = ok ? gstart[i] : raw_t(0);
					dn[j] = ok ? gnow[i] : (i < 0 ? Raw<NS>::lowest() : raw_t(0));
				}
			}
			__syncwarp();
			if (sub == 0) {
				// The claim just before the tile: if it is below target F is 0 at the tile's edge and the
				// tile's first claim opens whatever run there is; its marker resets F in front of our
				// first claim.
				const raw_t *ss = reinterpret_cast<const raw_t *>(st + SC_LEAD);
				const raw_t *sn = reinterpret_cast<const raw_t *>(st + SC_ARR + SC_LEAD);
				first_now = sn[0];
				if (t > 0) {
					left_closed = !is_above<NS>(ss[-1], sn[-1], p.target);
					if (FLAGS)
						T.carry_mk = (p.flags[ts - 1] & F_EMPTY_AFTER) ? 1u : 0u;
				}
				T.known = left_closed;
			}
			if (full)
				scan_chunk<NS, FLAGS, true>(p, st, t, cs, sub, lane, lt_mask, T, nonmono, last_marker);
			else
				scan_chunk<NS, FLAGS, false>(p, st, t, cs, sub, lane, lt_mask, T, nonmono, last_marker);
			__syncwarp(); // every lane is done with the stage: refill it with the chunk SC_NST ahead
			if (lane == 0)
				issue(ahead(t, sub), sg);
			sg = (sg + 1 == SC_NST) ? 0 : sg + 1;
		}
		// ---- the tile's record: where the inherited run closes, what the tile leaves open ----------
		if (lane == 0) {
			p.rec_fc[t] = left_closed ? 0u : T.fc; // left-closed: nothing crosses the tile's left edge, it anchors itself
			p.rec_code[t] = T.agg.code;
			p.rec_cend[t] = (unsigned int)T.c_last;
			reinterpret_cast<raw_t *>(p.tile_now0)[t] = first_now;
			p.ev_count[t] = T.evn > (unsigned int)EVK ? (unsigned int)EVK : T.evn;
			if (T.evn > (unsigned int)EVK)
				p.b->overflow = 1u;
			if (t + 1 == tiles)
				p.b->dbg[9] = gtimer(); // the last tile's warp is among the last to retire
		}
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
}

} // namespace cb
