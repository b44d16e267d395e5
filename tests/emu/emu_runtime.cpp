/* tests/emu/emu_runtime.cpp -- runtime of the host emulation (see cuda_emu.h).  TEST INFRASTRUCTURE. */
#include "cuda_emu.h"

thread_local emu_dim3 threadIdx, blockIdx;
emu_dim3 blockDim, gridDim;
namespace emu {
Barrier block_barrier;
Barrier warp_barrier[64];
unsigned long long warp_slot[64][32];
unsigned char dyn_smem[96 * 1024];

namespace {
std::mutex nb_m;
std::condition_variable nb_cv;
int nb_count[16];
unsigned long long nb_gen[16];
int nb_threads = 0;
} // namespace

/* bar.arrive id, n / bar.sync id, n: the barrier completes when n threads have arrived or synced */
void named_arrive(int id) {
	std::unique_lock<std::mutex> lk(nb_m);
	if (++nb_count[id] == nb_threads) {
		nb_count[id] = 0;
		nb_gen[id]++;
		nb_cv.notify_all();
	}
}
void named_sync(int id) {
	std::unique_lock<std::mutex> lk(nb_m);
	const unsigned long long g = nb_gen[id];
	if (++nb_count[id] == nb_threads) {
		nb_count[id] = 0;
		nb_gen[id]++;
		nb_cv.notify_all();
	} else {
		nb_cv.wait(lk, [&] { return nb_gen[id] != g; });
	}
}

void launch(int threads, const std::function<void()> &body) {
	block_barrier.reset(threads);
	for (int w = 0; w < threads / 32; w++)
		warp_barrier[w].reset(32);
	for (int i = 0; i < 16; i++)
		nb_count[i] = 0;
	nb_threads = threads;
	blockDim.x = threads;
	const emu_dim3 bi = blockIdx; /* the launching thread sets blockIdx before the call */
	std::vector<std::thread> ts;
	for (int t = 0; t < threads; t++)
		ts.emplace_back([t, bi, &body] {
			threadIdx.x = (unsigned)t;
			blockIdx = bi;
			body();
		});
	for (auto &t : ts)
		t.join();
}
} // namespace emu
