/* tests/emu/cuda_emu.h -- a minimal host emulation of the CUDA execution model for SINGLE-CTA
 * kernels: every CUDA thread is an OS thread, __syncthreads() is a barrier over the block,
 * warp shuffles exchange values through per-warp slots between warp barriers, __shared__ variables
 * are function-local statics (one block runs at a time).  It exists so that the control flow and
 * arithmetic of small bookkeeping kernels (csrc/cb_queue.cuh) can be checked against the oracle
 * on a machine without a GPU.  It says nothing about memory-model or performance behaviour on the
 * device.  TEST INFRASTRUCTURE: never part of the product library. */
#pragma once
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__

struct emu_dim3 {
	unsigned int x = 1, y = 1, z = 1;
};
extern thread_local emu_dim3 threadIdx, blockIdx;
extern emu_dim3 blockDim, gridDim;

namespace emu {
class Barrier {
public:
	void reset(int n) {
		n_ = n;
		count_ = 0;
		gen_ = 0;
	}
	void wait() {
		std::unique_lock<std::mutex> lk(m_);
		const unsigned long long g = gen_;
		if (++count_ == n_) {
			count_ = 0;
			gen_++;
			cv_.notify_all();
		} else {
			cv_.wait(lk, [&] { return gen_ != g; });
		}
	}

private:
	std::mutex m_;
	std::condition_variable cv_;
	int n_ = 1, count_ = 0;
	unsigned long long gen_ = 0;
};
extern Barrier block_barrier;
extern Barrier warp_barrier[64];
extern unsigned long long warp_slot[64][32];

template <typename T> inline T shfl_from(T v, int src_lane, bool use_own) {
	static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long raw = 0;
	std::memcpy(&raw, &v, sizeof(T));
	warp_slot[warp][lane] = raw;
	warp_barrier[warp].wait();
	unsigned long long got = use_own ? raw : warp_slot[warp][src_lane & 31];
	warp_barrier[warp].wait();
	T out;
	std::memcpy(&out, &got, sizeof(T));
	return out;
}
/* run `body` once per thread of a `threads`-wide block (threads % 32 == 0) */
void launch(int threads, const std::function<void()> &body);
} // namespace emu

inline void __syncthreads() { emu::block_barrier.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier[threadIdx.x >> 5].wait(); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int delta) {
	const int lane = threadIdx.x & 31;
	return emu::shfl_from(v, lane - delta, lane < delta);
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_from(v, src, false); }
inline unsigned int __ballot_sync(unsigned, bool pred) {
	unsigned int m = 0;
	for (int l = 0; l < 32; l++)
		m |= (emu::shfl_from<unsigned int>(pred ? 1u : 0u, l, false) & 1u) << l;
	return m;
}
inline void __threadfence() {}
inline void __threadfence_system() {}
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline unsigned int atomicOr(unsigned int *p, unsigned int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
	return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
