// cb_emit.cuh -- pass 3: write the per-claim outputs (decision 1 B + slot 4 B [+ backend 4 B]).
//
// The drop set of a batch is a short sorted list of index intervals (pass 2), so this pass
// never re-reads the timestamps: without flags a claim's admit rank is closed-form
//      admits_before(i) = i - drops_before(i)
// and its slot is (cursor0 + admits_before(i)) mod nslots -- the j-th admitted claim takes
// the j-th slot of the cyclic idle queue (lib/pool.js:747,756,934-935).
// With a flags array (cancelled handles, lib/pool.js:738-740) the admit rank is a prefix sum
// of (waiting && !drop): block scan + decoupled look-back over per-tile admit counts.
//
// Persistent CTAs stride over the tiles (4096 claims), so the per-CTA set-up -- the interval
// list staged in shared memory, the TMA bulk copy of the slot->backend table, the control
// block reads -- is paid once per CTA, not once per tile.  A tile is classified against the
// interval list with two ballots: no drop inside (the common case), entirely dropped (inside
// one episode), or mixed (an episode boundary crosses it; only these tiles walk the list per
// claim).  Every thread writes four groups of 4 consecutive claims, so each store instruction
// of a warp covers one contiguous 128 B (decisions) or 512 B (slots / backends) span.
#pragma once
#include "cb_device.cuh"
#include "cb_peers.cuh"

namespace cb {

constexpr int ETPB = 256;               // threads per emit CTA
constexpr int EGW = 4;                  // consecutive claims per group
constexpr int EG = ETILE / (ETPB * EGW); // groups per thread per tile (4)
constexpr int ESTR = ETPB * EGW;        // 1024: group g of thread t starts at ts + g*ESTR + t*EGW
constexpr int EWARPS = ETPB / 32;
constexpr int IVS_MAX = 64;             // interval lists up to this length are staged in shared memory
static_assert(EG * ESTR == ETILE, "emit geometry");

struct EmitParams {
	const uint8_t *flags;
	unsigned long long n;
	unsigned long long tiles;
	DevState *g;
	BatchCtl *bc;
	const unsigned long long *iv_lo, *iv_hi, *iv_cum;
	uint8_t *decision;
	int32_t *slot;
	int32_t *backend;
	const int32_t *slot_backend; // device table, nslots entries (padded to 16 B)
	int nslots;
	unsigned long long slot_magic; // floor((2^64-1) / nslots): division-free modulo
	int table_smem_bytes;          // 0 = read the table through L1/L2 instead
	SumDesc *sdesc;
	unsigned long long epoch;
	ChainLink link; // a flags batch passes the commit chain on from here (its cursor is final then)
	unsigned int *ticket;          // CTA completion counter (the last CTA finalises a flags batch and posts the counts)
	PeerLink peer;                 // count boards (peer.mine == nullptr: none): the grid then has ONE extra CTA, the last,
	                               // which handles no tile: it posts the batch's counts and reads the other ranks' rows
	int vec_ok;                    // outputs suitably aligned for vector stores
};

#ifndef CB_EMU
__device__ __forceinline__ unsigned int smem_u32(const void *p) {
	return (unsigned int)__cvta_generic_to_shared(p);
}
#endif
// x mod S without a 64-bit division: q = mulhi(x, floor((2^64-1)/S)) is floor(x/S) or one less
__device__ __forceinline__ unsigned int fastmod(unsigned long long x, unsigned int S, unsigned long long magic) {
	const unsigned long long q = __umul64hi(x, magic);
	unsigned long long r = x - q * (unsigned long long)S;
	while (r >= S)
		r -= S;
	return (unsigned int)r;
}
#ifndef CB_EMU
__device__ __forceinline__ void wait_table(unsigned long long *mbar) {
	unsigned int done = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred q;\n\t"
		             "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\t"
		             "selp.u32 %0, 1, 0, q;\n\t}"
		             : "=r"(done)
		             : "r"(smem_u32(mbar))
		             : "memory");
	}
}
// one thread: start the TMA bulk copy of the slot->backend table into shared memory
__device__ __forceinline__ void stage_table_async(unsigned char *dst, const void *src, unsigned int bytes,
    unsigned long long *mbar) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(dst)),
	    "l"(src), "r"(bytes), "r"(smem_u32(mbar))
	    : "memory");
}
#endif

// drop bits of the 4 claims [i0, i0+4) and the number of (CoDel-level) drops before i0
__device__ __forceinline__ unsigned int drops_general(const unsigned long long *iv_lo, const unsigned long long *iv_hi,
    const unsigned long long *iv_cum, unsigned int m, unsigned int jstart, unsigned long long i0, int nvalid,
    unsigned long long total_drops, unsigned long long &drops_before) {
	unsigned int j = jstart;
	while (j < m && iv_hi[j] < i0) // first interval with hi >= i0 (few steps from the tile's first)
		j++;
	if (j >= m) {
		drops_before = total_drops;
		return 0u;
	}
	unsigned long long ilo = iv_lo[j], ihi = iv_hi[j];
	drops_before = iv_cum[j] + (i0 > ilo ? i0 - ilo : 0ull);
	unsigned int bits = 0;
#pragma unroll
	for (int k = 0; k < EGW; k++) {
		const unsigned long long i = i0 + k;
		while (j < m && i > ihi) {
			j++;
			if (j < m) {
				ilo = iv_lo[j];
				ihi = iv_hi[j];
			}
		}
		if (k < nvalid && j < m && i >= ilo)
			bits |= 1u << k;
	}
	return bits;
}

template <bool FLAGS>
__global__ void __launch_bounds__(ETPB, 4) emit_kernel(const EmitParams p) {
	CB_DYN_SHARED(unsigned char, s_dyn); // backend table stage
	__shared__ __align__(8) unsigned long long s_mbar;
	__shared__ unsigned int s_wsum[EG][EWARPS];
	__shared__ unsigned long long s_tile_excl;
	__shared__ unsigned long long s_ivlo[IVS_MAX], s_ivhi[IVS_MAX], s_ivcum[IVS_MAX];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	DevState *g = p.g;
	BatchCtl *bc = p.bc;
	if (tid == 0 && blockIdx.x == 0)
		bc->dbg[10] = gtimer(); // timeline stamp (tools/pipeline_timeline.py)
	// Count boards: the remote reads of the other ranks' rows take NVLink round trips.  They run in a CTA of
	// their own beside the tile CTAs, so nothing waits for them (in the resolve kernel they delayed this kernel's
	// start, and the pipelined step with it: 58 -> 62-66 us on two GPUs).  A batch with waiting flags knows its
	// counts only when its last tile CTA is done and posts from there.
	const unsigned int tile_ctas = gridDim.x - (p.peer.mine ? 1u : 0u);
	if (p.peer.mine && blockIdx.x == tile_ctas) {
		if (tid == 0 && (!FLAGS || bc->path == 1))
			peer_post(p.peer, bc);
		return;
	}
	if (bc->path == 1) // the sequential path already wrote every output
		return;
	const bool want_backend = (p.backend != nullptr) && (p.nslots > 0);
	const bool stage_table = want_backend && (p.table_smem_bytes > 0);
	const unsigned int m = bc->n_intervals;
	const unsigned long long total_drops = bc->gcounts[3];
	const unsigned long long cursor0 = bc->cursor0;
	const unsigned int S = (unsigned int)p.nslots;

	// ---- per-CTA set-up --------------------------------------------------------------------
	if (stage_table && tid == 0) // TMA bulk copy of the slot->backend table
		stage_table_async(s_dyn, p.slot_backend, (unsigned int)p.table_smem_bytes, &s_mbar);
	const bool iv_smem = m <= IVS_MAX; // short interval lists live in shared memory
	if (iv_smem) {
		for (unsigned int k = tid; k < m; k += ETPB) {
			s_ivlo[k] = p.iv_lo[k];
			s_ivhi[k] = p.iv_hi[k];
			s_ivcum[k] = p.iv_cum[k];
		}
	}
	__syncthreads(); // also: the mbarrier is initialised before anyone polls it
	const unsigned long long *ivlo = iv_smem ? s_ivlo : p.iv_lo;
	const unsigned long long *ivhi = iv_smem ? s_ivhi : p.iv_hi;
	const unsigned long long *ivcum = iv_smem ? s_ivcum : p.iv_cum;
	const int32_t *tb = reinterpret_cast<const int32_t *>(s_dyn);
	bool table_ready = false;

	for (unsigned long long t = blockIdx.x; t < p.tiles; t += tile_ctas) {
		const unsigned long long ts = t * (unsigned long long)ETILE;
		unsigned long long te = ts + ETILE;
		if (te > p.n)
			te = p.n;
		const bool full_tile = (te - ts) == (unsigned long long)ETILE;

		// ---- classify the tile against the interval list (warp-uniform) -----------------------
		unsigned int kase = 0, jtile = 0;
		unsigned long long dtile = total_drops;
		{
			unsigned int lo = 0, hi = m;
			if (iv_smem) { // sorted: count the intervals that end before the tile (two per lane)
				const unsigned int b0 = __ballot_sync(0xffffffffu, (unsigned int)lane < m && s_ivhi[lane] < ts);
				const unsigned int b1 = __ballot_sync(0xffffffffu,
				    (unsigned int)lane + 32u < m && s_ivhi[(lane + 32) & (IVS_MAX - 1)] < ts);
				lo = __popc(b0) + __popc(b1);
			} else {
				while (lo < hi) { // first interval with hi >= ts
					const unsigned int mid = (lo + hi) >> 1;
					if (ivhi[mid] >= ts)
						hi = mid;
					else
						lo = mid + 1;
				}
			}
			jtile = lo;
			if (lo < m) {
				const unsigned long long ilo = ivlo[lo], ihi = ivhi[lo], cum = ivcum[lo];
				if (ilo >= te) {
					kase = 0;
					dtile = cum;
				} else if (ilo <= ts && ihi >= te - 1) {
					kase = 1;
					dtile = cum + (ts - ilo);
				} else {
					kase = 2;
					dtile = 0;
				}
			}
		}

		if (!FLAGS && kase != 2 && full_tile && p.vec_ok) {
			// ---- fast path: the whole tile is admitted, or the whole tile timed out -------------
			unsigned int base = 0;
			if (kase == 0 && S > 0)
				base = fastmod(cursor0 + (ts - dtile), S, p.slot_magic);
			if (want_backend && stage_table && !table_ready && kase == 0) {
				wait_table(&s_mbar);
				table_ready = true;
			}
			unsigned int *dec4 = p.decision ? reinterpret_cast<unsigned int *>(p.decision + ts) : nullptr;
			int4 *sl4 = p.slot ? reinterpret_cast<int4 *>(p.slot + ts) : nullptr;
			int4 *be4 = p.backend ? reinterpret_cast<int4 *>(p.backend + ts) : nullptr;
#pragma unroll
			for (int gI = 0; gI < EG; gI++) {
				const unsigned int q4 = gI * ETPB + tid; // index of the group of 4 inside the tile
				int4 sv = make_int4(-1, -1, -1, -1), bv = make_int4(-1, -1, -1, -1);
				if (kase == 0 && S > 0) {
					unsigned int s0 = base + q4 * EGW;
					if (S >= (unsigned int)ETILE + EGW) { // at most one wrap inside the tile
						if (s0 >= S)
							s0 -= S;
					} else {
						s0 %= S;
					}
					unsigned int s1 = s0 + 1, s2 = s0 + 2, s3 = s0 + 3;
					if (S >= EGW) {
						s1 = s1 >= S ? s1 - S : s1;
						s2 = s2 >= S ? s2 - S : s2;
						s3 = s3 >= S ? s3 - S : s3;
					} else {
						s1 %= S;
						s2 %= S;
						s3 %= S;
					}
					sv = make_int4((int)s0, (int)s1, (int)s2, (int)s3);
					if (be4) {
						if (stage_table)
							bv = make_int4(tb[s0], tb[s1], tb[s2], tb[s3]);
						else
							bv = make_int4(__ldg(p.slot_backend + s0), __ldg(p.slot_backend + s1),
							    __ldg(p.slot_backend + s2), __ldg(p.slot_backend + s3));
					}
				}
				if (dec4)
					dec4[q4] = kase == 1 ? 0x01010101u : 0u;
				if (sl4)
					sl4[q4] = sv;
				if (be4)
					be4[q4] = bv;
			}
			continue;
		}

		// ---- general path: partial tiles, episode boundaries, flags -----------------------------
		unsigned int dropbits[EG], waitbits[EG], admitbits[EG], validbits[EG];
		unsigned long long rank0[EG]; // admits before the group's first claim (whole batch)
		unsigned long long gi0[EG];
		int nval[EG];
#pragma unroll
		for (int gI = 0; gI < EG; gI++) {
			gi0[gI] = ts + (unsigned long long)gI * ESTR + (unsigned long long)tid * EGW;
			long long rem = (long long)p.n - (long long)gi0[gI];
			nval[gI] = rem >= EGW ? EGW : (rem > 0 ? (int)rem : 0);
			validbits[gI] = (1u << nval[gI]) - 1u;
			unsigned long long db = 0;
			if (kase == 0) {
				dropbits[gI] = 0;
				db = dtile;
			} else if (kase == 1) {
				dropbits[gI] = validbits[gI];
				db = dtile + (gi0[gI] - ts);
			} else {
				dropbits[gI] = drops_general(ivlo, ivhi, ivcum, m, jtile, gi0[gI], nval[gI], total_drops, db);
			}
			waitbits[gI] = 0xfu;
			if (FLAGS) {
				waitbits[gI] = 0;
				if (nval[gI] == EGW && ((reinterpret_cast<uintptr_t>(p.flags) & 3) == 0)) {
					const unsigned int f4 = *reinterpret_cast<const unsigned int *>(p.flags + gi0[gI]);
#pragma unroll
					for (int k = 0; k < EGW; k++)
						waitbits[gI] |= ((f4 >> (8 * k)) & 1u) << k;
				} else {
					for (int k = 0; k < nval[gI]; k++)
						waitbits[gI] |= (unsigned int)(p.flags[gi0[gI] + k] & 1u) << k;
				}
			}
			admitbits[gI] = waitbits[gI] & ~dropbits[gI] & validbits[gI];
			rank0[gI] = gi0[gI] - db; // closed form (no flags)
		}

		if (FLAGS) {
			// admit rank = prefix sum of (waiting && !drop): group 0 of every thread, then group 1, ...
			unsigned int a[EG], x[EG];
#pragma unroll
			for (int gI = 0; gI < EG; gI++) {
				a[gI] = __popc(admitbits[gI]);
				x[gI] = a[gI];
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					unsigned int y = __shfl_up_sync(0xffffffffu, x[gI], o);
					if (lane >= o)
						x[gI] += y;
				}
				if (lane == 31)
					s_wsum[gI][warp] = x[gI];
			}
			__syncthreads();
			unsigned int wex[EG], gbase[EG], total = 0;
#pragma unroll
			for (int gI = 0; gI < EG; gI++) {
				unsigned int tot = 0;
				wex[gI] = 0;
#pragma unroll
				for (int w = 0; w < EWARPS; w++) {
					if (w == warp)
						wex[gI] = tot;
					tot += s_wsum[gI][w];
				}
				gbase[gI] = total;
				total += tot;
			}
			// per-tile counters of real timeouts / skipped handles
			unsigned int dwait = 0, skip = 0;
#pragma unroll
			for (int gI = 0; gI < EG; gI++) {
				dwait += __popc(waitbits[gI] & dropbits[gI] & validbits[gI]);
				skip += __popc(~waitbits[gI] & validbits[gI]);
			}
			dwait = __reduce_add_sync(0xffffffffu, dwait);
			skip = __reduce_add_sync(0xffffffffu, skip);
			if (lane == 0) {
				if (dwait)
					atomicAdd(&bc->drops, (unsigned long long)dwait);
				if (skip)
					atomicAdd(&bc->skipped, (unsigned long long)skip);
			}
			// decoupled look-back over tile admit totals (warp 0)
			if (warp == 0) {
				unsigned long long excl = 0;
				if (t > 0) {
					if (lane == 0) {
						p.sdesc[t].agg = total;
						st_release(&p.sdesc[t].w, (p.epoch << 2) | ST_AGG);
					}
					long long base = (long long)t - 1;
					while (true) {
						long long k = base - lane;
						unsigned long long val = 0;
						bool term = false;
						if (k >= 0) {
							unsigned long long w;
							do {
								w = ld_acquire(&p.sdesc[k].w);
							} while ((w >> 2) != p.epoch);
							term = (w & 3ull) == ST_PREFIX;
							val = term ? ld_relaxed_u64(&p.sdesc[k].incl) : ld_relaxed_u64(&p.sdesc[k].agg);
						} else if (k == -1) {
							term = true;
						}
						const unsigned int tm = __ballot_sync(0xffffffffu, term);
						unsigned long long v = val;
						if (tm) {
							const int f = __ffs(tm) - 1; // nearest terminator; sum lanes 0..f
							v = (lane <= f) ? val : 0ull;
						}
#pragma unroll
						for (int o = 16; o > 0; o >>= 1)
							v += __shfl_xor_sync(0xffffffffu, v, o);
						excl += v;
						if (tm)
							break;
						base -= 32;
					}
				}
				if (lane == 0) {
					p.sdesc[t].agg = total;
					p.sdesc[t].incl = excl + total;
					st_release(&p.sdesc[t].w, (p.epoch << 2) | ST_PREFIX);
					s_tile_excl = excl;
				}
			}
			__syncthreads();
#pragma unroll
			for (int gI = 0; gI < EG; gI++)
				rank0[gI] = s_tile_excl + gbase[gI] + wex[gI] + (x[gI] - a[gI]);
		}

		// ---- slots / backends, vector stores ------------------------------------------------------
#pragma unroll
		for (int gI = 0; gI < EG; gI++) {
			if (nval[gI] == 0)
				continue;
			unsigned int scur = 0;
			if (S > 0)
				scur = fastmod(cursor0 + rank0[gI], S, p.slot_magic);
			int32_t sl[EGW];
			unsigned int d4 = 0;
#pragma unroll
			for (int k = 0; k < EGW; k++) {
				const bool adm = (admitbits[gI] >> k) & 1u;
				const unsigned int dk = ((waitbits[gI] >> k) & 1u) ? ((dropbits[gI] >> k) & 1u) : 2u;
				d4 |= dk << (8 * k);
				sl[k] = (adm && S > 0) ? (int32_t)scur : -1;
				if (adm) {
					scur++;
					if (scur == S)
						scur = 0;
				}
			}
			const bool vec = (nval[gI] == EGW) && p.vec_ok;
			if (p.decision) {
				if (vec)
					*reinterpret_cast<unsigned int *>(p.decision + gi0[gI]) = d4;
				else
					for (int k = 0; k < nval[gI]; k++)
						p.decision[gi0[gI] + k] = (uint8_t)(d4 >> (8 * k));
			}
			if (p.slot) {
				if (vec)
					*reinterpret_cast<int4 *>(p.slot + gi0[gI]) = make_int4(sl[0], sl[1], sl[2], sl[3]);
				else
					for (int k = 0; k < nval[gI]; k++)
						p.slot[gi0[gI] + k] = sl[k];
			}
			if (p.backend) {
				int32_t be[EGW];
				if (stage_table) {
					if (!table_ready) { // wait for the bulk copy (phase 0)
						wait_table(&s_mbar);
						table_ready = true;
					}
#pragma unroll
					for (int k = 0; k < EGW; k++)
						be[k] = (sl[k] >= 0) ? tb[sl[k]] : -1;
				} else {
#pragma unroll
					for (int k = 0; k < EGW; k++)
						be[k] = (sl[k] >= 0) ? __ldg(p.slot_backend + sl[k]) : -1;
				}
				if (vec)
					*reinterpret_cast<int4 *>(p.backend + gi0[gI]) = make_int4(be[0], be[1], be[2], be[3]);
				else
					for (int k = 0; k < nval[gI]; k++)
						p.backend[gi0[gI] + k] = be[k];
			}
		}
		if (FLAGS)
			__syncthreads(); // s_wsum / s_tile_excl are reused by the next tile
	}

	if (tid == 0 && blockIdx.x == tile_ctas - 1)
		bc->dbg[11] = gtimer();
	// ---- the last CTA to finish: a flags batch's totals (its admit cursor is final only now) and the
	// counts for the other ranks' boards ---------------------------------------------------------------
	if (FLAGS) {
		__threadfence();
		__syncthreads();
		if (tid == 0) {
			unsigned int done = atomicAdd(p.ticket, 1u);
			if (done == tile_ctas - 1) {
				*p.ticket = 0;
				__threadfence();
				{
					const unsigned long long admits = ld_relaxed_u64(&p.sdesc[p.tiles - 1].incl);
					const unsigned long long drops = atomicAdd(&bc->drops, 0ull);
					bc->admits = admits;
					g->cursor = bc->cursor0 + admits;
					g->tot_drops += drops;
					g->tot_admits += admits;
					bc->gcounts[0] = drops;
					bc->gcounts[1] = admits;
					bc->lcounts[0] = drops;
					bc->lcounts[1] = admits;
					chain_post(p.link, g); // the next batch's commit may read the cursor now
				}
				if (p.peer.mine)
					peer_post(p.peer, bc);
			}
		}
	}
}

} // namespace cb
