/* tests/emu/emu_runtime.cpp -- runtime of the host emulation (see cuda_emu.h).  TEST INFRASTRUCTURE. */
#include "cuda_emu.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>

thread_local emu_dim3 threadIdx, blockIdx;
emu_dim3 blockDim, gridDim;
namespace emu {
/* a barrier that does not complete is a bug in the kernel under test (threads disagree about reaching
 * it) or in the emulation: say where, instead of hanging the test run */
int Barrier::stall_seconds() {
	const char *e = std::getenv("CB_EMU_STALL_SECONDS");
	return e ? std::atoi(e) : 120;
}
static std::atomic<int> g_stalled{0};
bool Barrier::stall_flag() { return g_stalled.load() != 0; }
void Barrier::report_stall(const char *what) {
	static std::mutex m;
	{
		std::lock_guard<std::mutex> lk(m);
		const int first = g_stalled.fetch_add(1) == 0;
		std::fprintf(stderr, "[emu] %s stalled: block %u thread %u (warp %u lane %u)%s\n", what, blockIdx.x, threadIdx.x,
		    threadIdx.x >> 5, threadIdx.x & 31, first ? "  <- first to give up" : "");
		void *bt[24];
		const int n = backtrace(bt, 24);
		backtrace_symbols_fd(bt + 2, n > 8 ? 6 : n - 2, 2);
		if (first) {
			const unsigned w0 = (threadIdx.x >> 5) * 32;
			std::fprintf(stderr, "[emu] warp collectives done / exited, lanes 0..31 of this warp:");
			for (unsigned l = 0; l < 32; l++)
				std::fprintf(stderr, " %u%s", dbg_collectives[w0 + l], dbg_exited[w0 + l] ? "x" : "");
			std::fprintf(stderr, "\n");
		}
		std::fflush(stderr);
	}
	std::this_thread::sleep_for(std::chrono::seconds(4)); /* let the other stuck threads report too */
	std::_Exit(86);
}
Barrier block_barrier;
Barrier warp_barrier[64];
unsigned long long warp_slot[64][32];
alignas(128) unsigned char dyn_smem[232 * 1024];
unsigned int dbg_collectives[2048];
int dbg_exited[2048];

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
		int waited = 0;
		while (!nb_cv.wait_for(lk, std::chrono::seconds(1), [&] { return nb_gen[id] != g; })) {
			if (++waited >= Barrier::stall_seconds() || Barrier::stall_flag()) {
				lk.unlock();
				Barrier::report_stall("named barrier");
			}
		}
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
			dbg_collectives[t] = 0;
			dbg_exited[t] = 0;
			body();
			dbg_exited[t] = 1;
		});
	for (auto &t : ts)
		t.join();
}
} // namespace emu
