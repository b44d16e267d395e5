// cb_topology.cuh -- backend preference list, planRebalance and the slot table, on device.
//
// These run once per topology change (resolver 'added'/'removed', decoherence reshuffle,
// rebalance), not per claim: single-CTA kernels whose job is to keep the whole claim path -- including the table
// the emit kernel stages through TMA -- resident on the GPU and bit-identical to
//   lib/pool.js:285-313,501-519   (preference list)
//   lib/utils.js:239-393          (planRebalance)
#pragma once
#include "cb_device.cuh"

namespace cb {

#ifdef CB_TOPO_TPB // the host emulation of tests/emu runs the same code with a narrower CTA
constexpr int TOPO_TPB = CB_TOPO_TPB;
#else
constexpr int TOPO_TPB = 1024;
#endif

// insert `key` at index idx of keys[0..len): parallel shift-right by one, whole CTA
__device__ void cta_insert(int32_t *keys, int len, int idx, int32_t key) {
	// chunks are walked back to front so a chunk's destination never clobbers unread data
	int nmove = len - idx;
	int nchunks = (nmove + TOPO_TPB - 1) / TOPO_TPB;
	for (int c = nchunks - 1; c >= 0; c--) {
		int e = idx + c * TOPO_TPB + (int)threadIdx.x;
		int32_t v = 0;
		bool act = e < len;
		if (act)
			v = keys[e];
		__syncthreads();
		if (act)
			keys[e + 1] = v;
		__syncthreads();
	}
	if (threadIdx.x == 0)
		keys[idx] = key;
	__syncthreads();
}

// lib/pool.js:287-288 for keys 0..B-1 in resolver order; us[k] replaces Math.random()
__global__ void __launch_bounds__(TOPO_TPB, 1) pref_build_kernel(int32_t *keys, int B, const double *us) {
	for (int k = 0; k < B; k++) {
		int idx = (int)floor(__dmul_rn(us[k], (double)(k + 1)));
		cta_insert(keys, k, idx, k);
	}
}
// one 'added' event: lib/pool.js:285-291
__global__ void __launch_bounds__(TOPO_TPB, 1) pref_insert_kernel(int32_t *keys, int *len, int32_t key, double u) {
	int n = *len;
	__syncthreads();
	int idx = (int)floor(__dmul_rn(u, (double)(n + 1)));
	cta_insert(keys, n, idx, key);
	if (threadIdx.x == 0)
		*len = n + 1;
}
// one 'removed' event: lib/pool.js:293-296 (indexOf + splice); *status = -1 if absent
__global__ void __launch_bounds__(TOPO_TPB, 1) pref_remove_kernel(int32_t *keys, int *len, int32_t key, int *status) {
	__shared__ int s_idx;
	int n = *len;
	if (threadIdx.x == 0)
		s_idx = 0x7fffffff;
	__syncthreads();
	for (int e = threadIdx.x; e < n; e += TOPO_TPB)
		if (keys[e] == key)
			atomicMin(&s_idx, e);
	__syncthreads();
	int idx = s_idx;
	if (idx == 0x7fffffff) {
		if (threadIdx.x == 0)
			*status = -1;
		return;
	}
	// shift left, front to back
	int nmove = n - idx - 1;
	int nchunks = (nmove + TOPO_TPB - 1) / TOPO_TPB;
	for (int c = 0; c < nchunks; c++) {
		int e = idx + 1 + c * TOPO_TPB + (int)threadIdx.x;
		int32_t v = 0;
		bool act = e < n;
		if (act)
			v = keys[e];
		__syncthreads();
		if (act)
			keys[e - 1] = v;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		*len = n - 1;
		*status = 0;
	}
}
// reshuffle: lib/pool.js:501-519 (pop the last key, random re-insert)
__global__ void __launch_bounds__(TOPO_TPB, 1) pref_reshuffle_kernel(int32_t *keys, const int *len, double u) {
	int n = *len;
	if (n <= 1)
		return;
	int32_t taken = keys[n - 1];
	__syncthreads();
	int idx = (int)floor(__dmul_rn(u, (double)(n - 1 + 1)));
	cta_insert(keys, n - 1, idx, taken);
}

// ---- block-wide helpers (TOPO_TPB threads) ----------------------------------------------------
// sum of one long long per thread; every thread gets the total.  s_w: 32 long longs of shared memory.
__device__ __forceinline__ long long topo_block_sum(long long v, long long *s_w) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads(); // s_w may still be read from the previous call
	if (lane == 0)
		s_w[warp] = v;
	__syncthreads();
	long long t = 0;
	for (int w = 0; w < TOPO_TPB / 32; w++)
		t += s_w[w];
	return t;
}
// exclusive prefix sum of one int per thread; `total` = the block's sum (all threads).
__device__ __forceinline__ int topo_block_excl_scan(int v, long long *s_w, int &total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= o)
			x += y;
	}
	__syncthreads();
	if (lane == 31)
		s_w[warp] = x;
	__syncthreads();
	int before = 0, t = 0;
	for (int w = 0; w < TOPO_TPB / 32; w++) {
		if (w < warp)
			before += (int)s_w[w];
		t += (int)s_w[w];
	}
	total = t;
	return before + x - v;
}

// utils.planRebalance, lib/utils.js:239-393, one CTA.
//   1. The round-robin over `target` slots (:266-290) in closed form: position k of the preference list is
//      visited v_k = floor(target / B) + [k < target mod B] times; an alive backend takes v_k slots (one in
//      singleton mode), a dead one exactly one (its monitor) and asks for v_k replacements.  Per key, in
//      parallel; `done` and `replacements` are block sums.  Without dead backends -- every slot-table build of
//      the claim path -- this is the whole computation.
//   2. Only when replacements are wanted (dead backends): the replacement walk (:293-366), whose early exits
//      depend on the running totals, by one thread exactly as the reference writes it; `empties.length > 0`
//      (:331-339) is two running counters instead of the reference's O(B) filter per step.
//   3. plan.remove (reverse key order, oldest first, :372-382) and plan.add (forward key order, :383-390) from
//      exclusive scans of the per-key differences; every output element is placed by its own thread (binary
//      search over the offsets), so a table of 16 384 slots over 4 096 backends is filled in a few passes.
// Keys are positions 0..B-1; see include/cueball_b200.h for the output format.  scratch: 2 (B + 1) ints.
__global__ void __launch_bounds__(TOPO_TPB, 1) plan_rebalance_kernel(const int32_t *have, const uint8_t *dead, int B,
    int target, int max, int singleton, int32_t *want, int32_t *add, int32_t *n_add, int32_t *rm_key,
    int32_t *rm_ord, int32_t *n_rm, int32_t *scratch) {
	__shared__ long long s_w[32];
	const int tid = threadIdx.x;
	// ---- 1. the first loop in closed form -----------------------------------------------------------
	const int q = B > 0 ? target / B : 0, r = B > 0 ? target % B : 0;
	long long p_done = 0, p_repl = 0, p_alive = 0, p_aundef = 0, p_dundef = 0;
	for (int k = tid; k < B; k += TOPO_TPB) {
		const int v = q + (k < r ? 1 : 0);
		int w;
		if (!dead[k]) {
			w = singleton ? (v > 0 ? 1 : 0) : v;
			p_alive++;
			if (v == 0)
				p_aundef++;
		} else {
			w = v > 0 ? 1 : 0;
			p_repl += v;
			if (v == 0)
				p_dundef++;
		}
		p_done += w;
		want[k] = v > 0 ? w : -1; // -1: wantedSpares[k] === undefined
	}
	long long done = topo_block_sum(p_done, s_w);
	long long replacements = topo_block_sum(p_repl, s_w);
	const int alive = (int)topo_block_sum(p_alive, s_w);
	int alive_undef = (int)topo_block_sum(p_aundef, s_w);
	int dead_undef = (int)topo_block_sum(p_dundef, s_w);
	__syncthreads(); // want[] is in place for the thread that walks the replacements
	// ---- 2. replacements for dead backends (:293-366) ---------------------------------------------------
	if (tid == 0 && B > 0 && replacements > 0) {
		int ring = r;
#define CB_DEFINE(k)                       \
	if (want[k] < 0) {                     \
		want[k] = 0;                       \
		if (dead[k])                       \
			dead_undef--;                  \
		else                               \
			alive_undef--;                 \
	}
		if (done + replacements > max) // :293-294
			replacements = max - done;
		for (long long i = 0; i < replacements; ++i) { // :300-366
			int k = ring;
			ring = (ring + 1 == B) ? 0 : ring + 1;
			CB_DEFINE(k);
			if (!dead[k]) {
				if (singleton) {
					if (want[k] == 0) {
						want[k] = 1;
						++done;
						continue;
					}
				} else {
					++want[k];
					++done;
					continue;
				}
			}
			const long long count = done + replacements - i;                          // :330
			const int empties = singleton ? alive_undef : (alive + dead_undef);      // :331-339
			if (count + 1 <= max) {                                                   // :342-348
				if (want[k] == 0) {
					want[k] = 1;
					++done;
				}
				if (empties > 0)
					++replacements;
			} else if (count <= max && empties > 0) {                                 // :354-355
				++replacements;
			} else if (count <= max) {                                                // :358-362
				if (want[k] == 0) {
					want[k] = 1;
					++done;
				}
			} else {                                                                  // :365-367
				break;
			}
		}
#undef CB_DEFINE
	}
	__syncthreads();
	// ---- 3. the plan's lists ------------------------------------------------------------------------------
	int32_t *off_add = scratch, *off_rm = scratch + (B + 1);
	int carry_a = 0, carry_r = 0;
	for (int base = 0; base < B; base += TOPO_TPB) {
		const int k = base + tid;   // forward key order
		const int kr = B - 1 - k;   // reverse key order
		int a = 0, d = 0;
		if (k < B) {
			int w = want[k];
			if (w < 0) {
				w = 0;
				want[k] = 0;
			}
			a = w > have[k] ? w - have[k] : 0;
			int wr = want[kr];
			if (wr < 0)
				wr = 0;
			d = have[kr] > wr ? have[kr] - wr : 0;
		}
		int ta, tr;
		const int ea = topo_block_excl_scan(a, s_w, ta);
		const int er = topo_block_excl_scan(d, s_w, tr);
		if (k < B) {
			off_add[k] = carry_a + ea;
			off_rm[k] = carry_r + er; // indexed by reverse position j = B-1-key
		}
		carry_a += ta;
		carry_r += tr;
	}
	if (tid == 0) {
		off_add[B] = carry_a;
		off_rm[B] = carry_r;
		if (n_add)
			*n_add = carry_a;
		if (n_rm)
			*n_rm = carry_r;
	}
	__syncthreads();
	const int na = carry_a, nr = carry_r;
	if (add) {
		for (int s = tid; s < na; s += TOPO_TPB) { // the key whose span of plan.add holds element s
			int lo = 0, hi = B; // last k with off_add[k] <= s
			while (hi - lo > 1) {
				const int mid = (lo + hi) >> 1;
				if (off_add[mid] <= s)
					lo = mid;
				else
					hi = mid;
			}
			add[s] = lo;
		}
	}
	if (rm_key) {
		for (int s = tid; s < nr; s += TOPO_TPB) {
			int lo = 0, hi = B;
			while (hi - lo > 1) {
				const int mid = (lo + hi) >> 1;
				if (off_rm[mid] <= s)
					lo = mid;
				else
					hi = mid;
			}
			rm_key[s] = B - 1 - lo;
			rm_ord[s] = s - off_rm[lo]; // oldest connection of the key first
		}
	}
}

// slot table: slot s belongs to backend keys[add[s]] (plan.add order = addConnection
// order, lib/pool.js:632-654); pad the tail to a 16-byte multiple for the TMA bulk copy
__global__ void slot_table_kernel(const int32_t *keys, const int32_t *add, const int32_t *n_add,
    int32_t *slot_backend, int cap, int *nslots_out) {
	int na = *n_add;
	for (int s = threadIdx.x + blockIdx.x * blockDim.x; s < cap; s += blockDim.x * gridDim.x)
		slot_backend[s] = (s < na) ? keys[add[s]] : -1;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		*nslots_out = na;
}

// CB_BALANCE_SRV_WEIGHTED (extension, no reference oracle -- see DESIGN.md): only the
// backends of the lowest priority value take slots; the S slots are apportioned by weight
// with the largest-remainder method, ties to the earlier position in the preference list;
// a tier whose weights are all zero is treated as uniform.  want[] is by preference position.
__global__ void weighted_want_kernel(const int32_t *keys, int B, const int32_t *prio,
    const int32_t *weight, int S, int32_t *want, int32_t *add, int32_t *n_add) {
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	int best = 0x7fffffff;
	for (int k = 0; k < B; k++)
		best = min(best, prio[keys[k]]);
	long long wsum = 0;
	int members = 0;
	for (int k = 0; k < B; k++)
		if (prio[keys[k]] == best) {
			wsum += weight[keys[k]];
			members++;
		}
	long long given = 0;
	for (int k = 0; k < B; k++) {
		want[k] = 0;
		if (prio[keys[k]] != best)
			continue;
		long long w = wsum > 0 ? weight[keys[k]] : 1;
		long long tot = wsum > 0 ? wsum : members;
		want[k] = (int32_t)(((long long)S * w) / tot);
		given += want[k];
	}
	// largest remainders, stable in preference order
	for (long long left = S - given; left > 0; left--) {
		long long bestrem = -1;
		int bk = -1;
		for (int k = 0; k < B; k++) {
			if (prio[keys[k]] != best)
				continue;
			long long w = wsum > 0 ? weight[keys[k]] : 1;
			long long tot = wsum > 0 ? wsum : members;
			long long rem = ((long long)S * w) % tot;
			// a position that already received its bonus slot is skipped
			long long base = ((long long)S * w) / tot;
			if (want[k] > base)
				continue;
			if (rem > bestrem) {
				bestrem = rem;
				bk = k;
			}
		}
		if (bk < 0)
			break;
		want[bk]++;
	}
	int na = 0;
	for (int k = 0; k < B; k++)
		for (int c = 0; c < want[k]; c++)
			add[na++] = k;
	*n_add = na;
}

// ----------------------------------------------------------------------------- utils.shuffle
// lib/utils.js:207-217: i = n; while (i > 0) { j = floor(random() * i); --i; swap(a[i], a[j]); } with the
// draws injected (us[t] is the t-th Math.random()).  The loop is a chain of n transpositions, but its result
// has a per-position form: step i (the one that swaps a[i] with a[j_i], j_i <= i) is the LAST one that touches
// position i, so out[i] is what sat at j_i just before it.  Walking the earlier steps (i+1, i+2, ... n-1)
// backwards in time, the element under the finger moves only when a step's j equals the finger's position
// (then it came from that step's i).  Every output position follows its own finger: n independent walks over
// the shared list of j's, which the CTA stages through shared memory tile by tile -- O(n^2 / threads), no
// serial chain.  j_i = floor(us[n-1-i] * (i+1)).
constexpr int SHUF_TILE = 2048;
__device__ __forceinline__ int shuffle_j(const double *us, int n, int i) {
	return (int)floor(__dmul_rn(us[n - 1 - i], (double)(i + 1)));
}
__global__ void __launch_bounds__(TOPO_TPB) shuffle_kernel(const int32_t *in, int32_t *out, const double *us, int n) {
	__shared__ int32_t sj[SHUF_TILE];
	for (int p0 = (int)blockIdx.x * TOPO_TPB; p0 < n; p0 += (int)gridDim.x * TOPO_TPB) { // uniform in the CTA
		const int p = p0 + (int)threadIdx.x;
		int pos = p < n ? shuffle_j(us, n, p) : -1;
		for (int base = p0 + 1; base < n; base += SHUF_TILE) {
			const int lim = n - base < SHUF_TILE ? n - base : SHUF_TILE;
			__syncthreads();
			for (int k = (int)threadIdx.x; k < lim; k += TOPO_TPB)
				sj[k] = shuffle_j(us, n, base + k);
			__syncthreads();
			if (p < n) {
				int k = p + 1 - base;
				for (k = k < 0 ? 0 : k; k < lim; k++)
					if (sj[k] == pos)
						pos = base + k;
			}
		}
		if (p < n)
			out[p] = in[pos];
	}
}

} // namespace cb
