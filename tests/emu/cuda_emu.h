/* tests/emu/cuda_emu.h -- a minimal host emulation of the CUDA execution model for SINGLE-CTA
 * kernels: every CUDA thread is an OS thread, __syncthreads() is a barrier over the block,
 * warp shuffles exchange values through per-warp slots between warp barriers, __shared__ variables
 * are function-local statics (one block runs at a time).  It exists so that the control flow and
 * arithmetic of small bookkeeping kernels (csrc/cb_queue.cuh) can be checked against the oracle
 * on a machine without a GPU.  It says nothing about memory-model or performance behaviour on the
 * device.  TEST INFRASTRUCTURE: never part of the product library. */
#pragma once
#include <chrono>
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
			int waited = 0;
			while (!cv_.wait_for(lk, std::chrono::seconds(1), [&] { return gen_ != g; })) {
				if (++waited >= stall_seconds() || stall_flag())
					report_stall("barrier");
			}
		}
	}
	static int stall_seconds();
	static bool stall_flag();
	static void report_stall(const char *what); /* prints the stuck thread's stack and aborts */

private:
	std::mutex m_;
	std::condition_variable cv_;
	int n_ = 1, count_ = 0;
	unsigned long long gen_ = 0;
};
extern Barrier block_barrier;
extern Barrier warp_barrier[64];
extern unsigned long long warp_slot[64][32];
extern unsigned int dbg_collectives[2048]; /* warp collectives executed per thread of the running block */
extern int dbg_exited[2048];

/* every lane publishes one 64-bit value; returns a pointer to the warp's 32 slots, valid until
 * the matching warp_release() */
inline const unsigned long long *warp_publish(unsigned long long raw) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	dbg_collectives[threadIdx.x]++;
	warp_slot[warp][lane] = raw;
	warp_barrier[warp].wait();
	return warp_slot[warp];
}
inline void warp_release() { warp_barrier[threadIdx.x >> 5].wait(); }

template <typename T> inline T shfl_from(T v, int src_lane, bool use_own) {
	static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long raw = 0;
	std::memcpy(&raw, &v, sizeof(T));
	dbg_collectives[threadIdx.x]++;
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
	const unsigned long long *sl = emu::warp_publish(pred ? 1ull : 0ull);
	unsigned int m = 0;
	for (int l = 0; l < 32; l++)
		m |= (unsigned int)(sl[l] & 1ull) << l;
	emu::warp_release();
	return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
	return emu::shfl_from(v, (int)((threadIdx.x & 31) ^ (unsigned)lanemask), false);
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, int delta) {
	const int lane = threadIdx.x & 31;
	return emu::shfl_from(v, lane + delta, lane + delta > 31);
}
inline unsigned int __reduce_add_sync(unsigned, unsigned int v) {
	const unsigned long long *sl = emu::warp_publish(v);
	unsigned int s = 0;
	for (int l = 0; l < 32; l++)
		s += (unsigned int)sl[l];
	emu::warp_release();
	return s;
}
inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned int)v); }
inline int __ffs(unsigned int x) { return __builtin_ffs((int)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(unsigned int x) { return __builtin_popcount(x); }
inline int __clz(unsigned int x) { return x ? __builtin_clz(x) : 32; }
inline int __clz(int x) { return __clz((unsigned int)x); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
	return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T> inline T __ldg(const T *p) { return *p; }
inline void __threadfence_block() {}
#define __align__(n) __attribute__((aligned(n)))
struct alignas(16) uint4 {
	unsigned int x, y, z, w;
};
struct alignas(16) int4 {
	int x, y, z, w;
};
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int atomicMax(int *p, int v) {
	int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
	}
	return old;
}
inline int atomicMin(int *p, int v) {
	int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
	}
	return old;
}
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
	unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
	}
	return old;
}
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) {
	return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
}

/* dynamic shared memory: one buffer per (sequentially emulated) block */
namespace emu {
extern unsigned char dyn_smem[232 * 1024];
/* named barriers (bar.arrive / bar.sync id, nthreads): arrivals are counted, a sync also waits */
void named_arrive(int id);
void named_sync(int id);
} // namespace emu
#define CB_EMU 1
#define CB_EMU_LOCKSTEP() __syncthreads()
#define CB_DYN_SHARED(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem)

/* host versions of the PTX helpers of csrc/cb_device.cuh, cb_scan.cuh, cb_emit.cuh */
namespace cb {
inline unsigned long long gtimer() { return 0; }
inline unsigned long long ld_acquire(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return ld_acquire(p); }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { st_release(p, v); }
inline void st_relaxed_sys(unsigned long long *p, unsigned long long v) { st_release(p, v); }
struct PeerWord;
template <typename W> inline void st_word(W *p, unsigned long long tag, unsigned long long v) {
	p->v = v;
	__atomic_store_n(&p->tag, tag, __ATOMIC_RELEASE);
}
template <typename W> inline W ld_word(const W *p) {
	W w;
	w.tag = __atomic_load_n(&p->tag, __ATOMIC_ACQUIRE);
	w.v = p->v;
	return w;
}
inline double ld_relaxed_f64(const double *p) { return *(const volatile double *)p; }
inline unsigned long long ld_relaxed_u64(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
/* mbarrier + TMA bulk copy (csrc/cb_scan.cuh): the barrier word holds the pending transaction bytes in its low half
 * and the number of completed phases in its high half; a bulk copy is a memcpy that completes its bytes at once */
inline void mbar_init(unsigned long long *mbar, unsigned int) { __atomic_store_n(mbar, 0ull, __ATOMIC_RELEASE); }
inline void mbar_init_fence() {}
inline void mbar_expect_tx(unsigned long long *mbar, unsigned int bytes) { __atomic_fetch_add(mbar, (unsigned long long)bytes, __ATOMIC_ACQ_REL); }
inline void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *mbar, unsigned long long) {
	std::memcpy(dst, src, bytes);
	const unsigned long long left = __atomic_sub_fetch(mbar, (unsigned long long)bytes, __ATOMIC_ACQ_REL);
	if ((left & 0xffffffffull) == 0)
		__atomic_fetch_add(mbar, 1ull << 32, __ATOMIC_ACQ_REL); /* the phase completes */
}
inline unsigned long long l2_evict_first_policy() { return 0; }
inline void mbar_wait(unsigned long long *mbar, unsigned int parity) { /* phase `parity` has completed */
	int spins = 0;
	while ((((__atomic_load_n(mbar, __ATOMIC_ACQUIRE) >> 32) & 1ull) == (unsigned long long)parity)) {
		std::this_thread::yield();
		if (++spins > 200000000)
			emu::Barrier::report_stall("mbarrier");
	}
}
inline void wait_table(unsigned long long *) {}
inline void stage_table_async(unsigned char *dst, const void *src, unsigned int bytes, unsigned long long *) {
	std::memcpy(dst, src, bytes);
}
} // namespace cb
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
inline unsigned int atomicCAS(unsigned int *p, unsigned int cmp, unsigned int v) {
	__atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
	return cmp; /* the value found */
}
inline unsigned int atomicExch(unsigned int *p, unsigned int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
