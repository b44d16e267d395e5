/* tests/emu/topology_emu.cpp -- the topology kernels of csrc/cb_topology.cuh (preference list insert /
 * remove / reshuffle, planRebalance, slot table) compiled for the host under tests/emu/cuda_emu.h
 * (64-thread CTA via CB_TOPO_TPB) behind plain C functions.  TEST INFRASTRUCTURE. */
#define CB_TOPO_TPB 64
#include "../../node-cueball_b200/csrc/cb_topology.cuh"
#include "../../node-cueball_b200/csrc/cb_resolver.cuh"

#include <vector>

using namespace cb;

namespace {
template <typename F> void one_cta(int threads, F call) {
	gridDim.x = 1;
	blockIdx.x = 0;
	emu::launch(threads, call);
}
} // namespace

extern "C" {
void emu_pref_build(int32_t *keys, int B, const double *us) {
	one_cta(TOPO_TPB, [&] { pref_build_kernel(keys, B, us); });
}
void emu_pref_insert(int32_t *keys, int *len, int32_t key, double u) {
	one_cta(TOPO_TPB, [&] { pref_insert_kernel(keys, len, key, u); });
}
int emu_pref_remove(int32_t *keys, int *len, int32_t key) {
	int status = 0;
	one_cta(TOPO_TPB, [&] { pref_remove_kernel(keys, len, key, &status); });
	return status;
}
void emu_pref_reshuffle(int32_t *keys, const int *len, double u) {
	one_cta(TOPO_TPB, [&] { pref_reshuffle_kernel(keys, len, u); });
}
void emu_plan_rebalance(const int32_t *have, const uint8_t *dead, int B, int target, int max, int singleton,
    int32_t *want, int32_t *add, int32_t *n_add, int32_t *rm_key, int32_t *rm_ord, int32_t *n_rm) {
	std::vector<int32_t> scratch(2 * (size_t)(B + 1) + 2);
	int32_t *scr = scratch.data();
	one_cta(TOPO_TPB, [&] { plan_rebalance_kernel(have, dead, B, target, max, singleton, want, add, n_add, rm_key, rm_ord, n_rm, scr); });
}
void emu_shuffle(const int32_t *in, int32_t *out, const double *us, int n, int ctas) {
	gridDim.x = (unsigned)ctas;
	for (unsigned b = 0; b < (unsigned)ctas; b++) {
		blockIdx.x = b;
		emu::launch(TOPO_TPB, [&] { shuffle_kernel(in, out, us, n); });
	}
}
/* CB_BALANCE_SRV_WEIGHTED (extension): want[] by preference position and the add list, as rebuild_slots() launches it */
void emu_weighted_want(const int32_t *keys, int B, const int32_t *prio, const int32_t *weight, int S, int32_t *want,
    int32_t *add, int32_t *n_add) {
	one_cta(1, [&] { weighted_want_kernel(keys, B, prio, weight, S, want, add, n_add); });
}
/* srvKey of n messages (csrc/cb_resolver.cuh: SHA-1 + base64, one thread per message), 128-thread CTAs as
 * cb_srv_keys launches them */
void emu_srv_keys(const unsigned char *msgs, const unsigned int *offs, int n, char *keys) {
	const unsigned ctas = (unsigned)((n + 127) / 128);
	gridDim.x = ctas;
	for (unsigned b = 0; b < ctas; b++) {
		blockIdx.x = b;
		emu::launch(128, [&] { srv_key_kernel(msgs, offs, n, keys); });
	}
}
void emu_slot_table(const int32_t *keys, const int32_t *add, const int32_t *n_add, int32_t *slot_backend, int cap,
    int *nslots_out) {
	gridDim.x = 4;
	for (unsigned b = 0; b < 4; b++) {
		blockIdx.x = b;
		emu::launch(64, [&] { slot_table_kernel(keys, add, n_add, slot_backend, cap, nslots_out); });
	}
}
}
