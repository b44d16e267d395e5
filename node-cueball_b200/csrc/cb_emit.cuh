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
// The slot -> backend table (built from the preference list and planRebalance) is staged in
// shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) per CTA when `backend` is
// requested; it is at most nslots*4 bytes.
#pragma once
#include "cb_device.cuh"

namespace cb {

struct EmitParams {
	const uint8_t *flags;
	unsigned long long n;
	unsigned long long tiles;
	DevState *g;
	const unsigned long long *iv_lo, *iv_hi, *iv_cum;
	uint8_t *decision;
	int32_t *slot;
	int32_t *backend;
	const int32_t *slot_backend; // device table, nslots entries (padded to 16 B)
	int nslots;
	int table_smem_bytes;        // 0 = read the table through L1/L2 instead
	SumDesc *sdesc;
	unsigned long long epoch;
	unsigned int *ticket;        // CTA completion counter (flags variant finalisation)
	int vec_ok;                  // outputs suitably aligned for vector stores
};

__device__ __forceinline__ unsigned int smem_u32(const void *p) {
	return (unsigned int)__cvta_generic_to_shared(p);
}

template <bool FLAGS>
__global__ void __launch_bounds__(TPB, 4) emit_kernel(const EmitParams p) {
	extern __shared__ __align__(16) unsigned char s_dyn[]; // backend table stage
	__shared__ __align__(8) unsigned long long s_mbar;
	__shared__ unsigned int s_wsum[WARPS];
	__shared__ unsigned long long s_tile_excl;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	DevState *g = p.g;
	if (g->path == 1) // the sequential path already wrote every output
		return;
	const unsigned long long t = blockIdx.x;
	const unsigned long long i0 = t * (unsigned long long)TILE + (unsigned long long)tid * IPT;
	const bool want_backend = (p.backend != nullptr) && (p.nslots > 0);
	const bool stage_table = want_backend && (p.table_smem_bytes > 0);

	// ---- kick off the TMA bulk copy of the slot->backend table ------------------------
	if (stage_table) {
		if (tid == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar)));
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_mbar)),
			    "r"((unsigned int)p.table_smem_bytes)
			    : "memory");
			asm volatile(
			    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
			        smem_u32(s_dyn)),
			    "l"(p.slot_backend), "r"((unsigned int)p.table_smem_bytes), "r"(smem_u32(&s_mbar))
			    : "memory");
		}
		__syncthreads(); // the barrier must be initialised before anyone polls it
	}

	long long rem = (long long)p.n - (long long)i0;
	const int nvalid = rem >= IPT ? IPT : (rem > 0 ? (int)rem : 0);

	// ---- drop flags from the interval list ----------------------------------------------
	const unsigned int m = g->n_intervals;
	unsigned int dropbits = 0;
	unsigned long long drops_before = 0; // drops (CoDel level) among claims < i0
	if (m > 0 && nvalid > 0) {
		// first interval with hi >= i0
		unsigned int lo = 0, hi = m;
		while (lo < hi) {
			unsigned int mid = (lo + hi) >> 1;
			if (p.iv_hi[mid] >= i0)
				hi = mid;
			else
				lo = mid + 1;
		}
		unsigned int j = lo;
		if (j < m) {
			unsigned long long ilo = p.iv_lo[j], ihi = p.iv_hi[j];
			drops_before = p.iv_cum[j] + (i0 > ilo ? i0 - ilo : 0ull);
#pragma unroll
			for (int k = 0; k < IPT; k++) {
				const unsigned long long i = i0 + k;
				while (j < m && i > ihi) {
					j++;
					if (j < m) {
						ilo = p.iv_lo[j];
						ihi = p.iv_hi[j];
					}
				}
				if (j < m && i >= ilo)
					dropbits |= 1u << k;
			}
		} else {
			drops_before = g->gcounts[3];
		}
	}

	// ---- waiting flags, admit rank ----------------------------------------------------------
	unsigned int waitbits = 0xffu;
	if (FLAGS) {
		waitbits = 0;
		if (nvalid == IPT && ((reinterpret_cast<uintptr_t>(p.flags) & 7) == 0)) {
			unsigned long long f8 = *reinterpret_cast<const unsigned long long *>(p.flags + i0);
#pragma unroll
			for (int k = 0; k < IPT; k++)
				waitbits |= (unsigned int)((f8 >> (8 * k)) & 1ull) << k;
		} else {
			for (int k = 0; k < nvalid; k++)
				waitbits |= (unsigned int)(p.flags[i0 + k] & 1u) << k;
		}
	}
	const unsigned int validbits = (nvalid == IPT) ? 0xffu : ((1u << nvalid) - 1u);
	const unsigned int admitbits = waitbits & ~dropbits & validbits;
	unsigned long long rank0; // admits before claim i0 (whole batch)
	if (!FLAGS) {
		rank0 = i0 - drops_before;
	} else {
		// block exclusive scan of per-thread admit counts
		unsigned int a = __popc(admitbits), x = a;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o)
				x += y;
		}
		if (lane == 31)
			s_wsum[warp] = x;
		__syncthreads();
		unsigned int wex = 0, total = 0;
#pragma unroll
		for (int w = 0; w < WARPS; w++) {
			if (w == warp)
				wex = total;
			total += s_wsum[w];
		}
		// per-tile counters of real timeouts / skipped handles
		unsigned int dwait = __popc(waitbits & dropbits & validbits);
		unsigned int skip = __popc(~waitbits & validbits);
		dwait = __reduce_add_sync(0xffffffffu, dwait);
		skip = __reduce_add_sync(0xffffffffu, skip);
		if (lane == 0) {
			if (dwait)
				atomicAdd(&g->drops, (unsigned long long)dwait);
			if (skip)
				atomicAdd(&g->skipped, (unsigned long long)skip);
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
					if (tm) {
						const int f = __ffs(tm) - 1; // nearest terminator; sum lanes 0..f
						unsigned long long v = (lane <= f) ? val : 0ull;
#pragma unroll
						for (int o = 16; o > 0; o >>= 1)
							v += __shfl_xor_sync(0xffffffffu, v, o);
						excl += v;
						break;
					}
					unsigned long long v = val;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1)
						v += __shfl_xor_sync(0xffffffffu, v, o);
					excl += v;
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
		rank0 = s_tile_excl + wex + (x - a);
	}

	// ---- slots / backends -----------------------------------------------------------------
	int32_t sl[IPT];
	uint8_t dec[IPT];
	unsigned int scur = 0;
	if (p.nslots > 0)
		scur = (unsigned int)((g->cursor0 + rank0) % (unsigned long long)p.nslots);
#pragma unroll
	for (int k = 0; k < IPT; k++) {
		const bool adm = (admitbits >> k) & 1u;
		dec[k] = ((waitbits >> k) & 1u) ? (uint8_t)((dropbits >> k) & 1u) : (uint8_t)2;
		sl[k] = (adm && p.nslots > 0) ? (int32_t)scur : -1;
		if (adm) {
			scur++;
			if (scur == (unsigned int)p.nslots)
				scur = 0;
		}
	}
	if (nvalid == IPT && p.vec_ok) {
		if (p.decision) {
			unsigned long long d8 = 0;
#pragma unroll
			for (int k = 0; k < IPT; k++)
				d8 |= (unsigned long long)dec[k] << (8 * k);
			*reinterpret_cast<unsigned long long *>(p.decision + i0) = d8;
		}
		if (p.slot) {
			int4 *o = reinterpret_cast<int4 *>(p.slot + i0);
			o[0] = make_int4(sl[0], sl[1], sl[2], sl[3]);
			o[1] = make_int4(sl[4], sl[5], sl[6], sl[7]);
		}
	} else {
		for (int k = 0; k < nvalid; k++) {
			if (p.decision)
				p.decision[i0 + k] = dec[k];
			if (p.slot)
				p.slot[i0 + k] = sl[k];
		}
	}
	if (p.backend) {
		int32_t be[IPT];
		if (stage_table) {
			// wait for the bulk copy (phase 0)
			unsigned int done = 0;
			while (!done) {
				asm volatile("{\n\t.reg .pred q;\n\t"
				             "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\t"
				             "selp.u32 %0, 1, 0, q;\n\t}"
				             : "=r"(done)
				             : "r"(smem_u32(&s_mbar))
				             : "memory");
			}
			const int32_t *tb = reinterpret_cast<const int32_t *>(s_dyn);
#pragma unroll
			for (int k = 0; k < IPT; k++)
				be[k] = (sl[k] >= 0) ? tb[sl[k]] : -1;
		} else {
#pragma unroll
			for (int k = 0; k < IPT; k++)
				be[k] = (sl[k] >= 0) ? __ldg(p.slot_backend + sl[k]) : -1;
		}
		if (nvalid == IPT && p.vec_ok) {
			int4 *o = reinterpret_cast<int4 *>(p.backend + i0);
			o[0] = make_int4(be[0], be[1], be[2], be[3]);
			o[1] = make_int4(be[4], be[5], be[6], be[7]);
		} else {
			for (int k = 0; k < nvalid; k++)
				p.backend[i0 + k] = be[k];
		}
	}

	// ---- flags variant: the last CTA to finish publishes the batch totals -----------------
	if (FLAGS) {
		__threadfence();
		__syncthreads();
		if (tid == 0) {
			unsigned int done = atomicAdd(p.ticket, 1u);
			if (done == (unsigned int)(p.tiles - 1)) {
				*p.ticket = 0;
				__threadfence();
				const unsigned long long admits = ld_relaxed_u64(&p.sdesc[p.tiles - 1].incl);
				const unsigned long long drops = atomicAdd(&g->drops, 0ull);
				g->admits = admits;
				g->cursor = g->cursor0 + admits;
				g->tot_drops += drops;
				g->tot_admits += admits;
				g->gcounts[0] = drops;
				g->gcounts[1] = admits;
			}
		}
	}
}

} // namespace cb
