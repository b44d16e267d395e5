// cb_topology.cuh -- backend preference list, planRebalance and the slot table, on device.
//
// These run once per topology change (resolver 'added'/'removed', decoherence reshuffle,
// rebalance), not per claim, and are inherently sequential in the reference, so they are
// small single-CTA kernels whose job is to keep the whole claim path -- including the table
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

// utils.planRebalance, lib/utils.js:239-393.  One thread: the algorithm is a sequential
// round-robin with data-dependent early exits; `empties.length > 0` (:331-339) is kept as
// two running counters instead of the reference's O(B) filter per replacement.
// Keys are positions 0..B-1; see include/cueball_b200.h for the output format.
__global__ void plan_rebalance_kernel(const int32_t *have, const uint8_t *dead, int B, int target,
    int max, int singleton, int32_t *want, int32_t *add, int32_t *n_add, int32_t *rm_key,
    int32_t *rm_ord, int32_t *n_rm) {
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	long long done = 0, replacements = 0;
	int ring = 0;
	int alive = 0, alive_undef = 0, dead_undef = 0;
	for (int k = 0; k < B; k++) {
		want[k] = -1; // wantedSpares[k] === undefined
		if (dead[k])
			dead_undef++;
		else {
			alive++;
			alive_undef++;
		}
	}
#define CB_DEFINE(k)                       \
	if (want[k] < 0) {                     \
		want[k] = 0;                       \
		if (dead[k])                       \
			dead_undef--;                  \
		else                               \
			alive_undef--;                 \
	}
	if (B > 0) {
		for (int i = 0; i < target; ++i) { // :266-290
			int k = ring;
			ring = (ring + 1 == B) ? 0 : ring + 1;
			CB_DEFINE(k);
			if (!dead[k]) {
				if (singleton) {
					if (want[k] == 0) {
						want[k] = 1;
						++done;
					}
				} else {
					++want[k];
					++done;
				}
				continue;
			}
			if (want[k] == 0) {
				want[k] = 1;
				++done;
			}
			++replacements;
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
	}
#undef CB_DEFINE
	int na = 0, nr = 0;
	for (int k = 0; k < B; k++)
		if (want[k] < 0)
			want[k] = 0;
	for (int k = B - 1; k >= 0; k--) { // :372-382 remove: reverse key order, oldest first
		int h = have[k], ord = 0;
		while (h > want[k]) {
			if (rm_key) {
				rm_key[nr] = k;
				rm_ord[nr] = ord;
			}
			nr++;
			ord++;
			--h;
		}
	}
	for (int k = 0; k < B; k++) { // :383-390 add: forward key order
		int h = have[k];
		while (h < want[k]) {
			if (add)
				add[na] = k;
			na++;
			++h;
		}
	}
	if (n_add)
		*n_add = na;
	if (n_rm)
		*n_rm = nr;
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

} // namespace cb
