// cb_device.cuh -- shared device-side definitions for the claim-queue admission kernels
// (sm_100a).  See DESIGN.md for the decomposition; reference citations are file:line in
// TritonDataCenter/node-cueball v2.10.3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr int TILE = 2048;        // claims per scan tile (scan and prepare agree on it)
constexpr int ETILE = 4096;       // claims per emit tile
constexpr int TPB = 256;         // threads of a scan CTA: 8 warps x 256 claims
constexpr int WARPS = TPB / 32;
constexpr int EVK = 8;          // class-change events a tile can hold before "overflow"
constexpr double CODEL_INTERVAL = 100.0; // lib/codel.js:16

// per-claim flag bits (include/cueball_b200.h)
constexpr unsigned F_WAITING = 1u, F_EMPTY_AFTER = 2u;

// ---------------------------------------------------------------------------------------
// Device-resident state.  DevState is the pool's persistent state (one per context): the
// ControlledDelay fields and the admit cursor; consecutive batches chain through it without a
// host round trip.  BatchCtl is the per-batch control block: up to CB_LANES batches are in
// flight at once (scans, prepares and emits of different batches overlap; only the commits are
// ordered), each on its own lane with its own BatchCtl and scratch.
struct DevState {
	// ControlledDelay, lib/codel.js:24-32 (+ cd_last_empty :92)
	double target;            // cd_targdelay
	double F;                 // cd_first_above_time (0 = unset)
	double D;                 // cd_drop_next
	unsigned long long C;     // cd_count
	int B;                    // cd_dropping
	int hasE;                 // cd_last_empty defined?
	double E;                 // cd_last_empty
	unsigned long long cursor; // admit cursor (slot = cursor mod nslots)
	unsigned long long tot_drops, tot_admits;
};
struct BatchCtl {
	unsigned long long n, drops, admits, skipped;
	unsigned long long cursor0;    // admit cursor before this batch (read by the emit kernel)
	unsigned long long gcounts[4]; // {drops, admits, n, CoDel-level drops}: [0..2] = allreduce buffer
	unsigned long long lcounts[4]; // this rank's {drops, admits, n} of the batch (posted to the peer boards)
	// set by the scan kernel, consumed + reset by the resolve kernel
	unsigned int nonmono;          // clock went backwards inside the batch
	unsigned int overflow;         // a tile found more than EVK events
	unsigned long long last_marker_p1; // 1 + index of the last CB_F_EMPTY_AFTER claim
	int path;                      // CB_PATH_*
	unsigned int n_events;
	unsigned int n_intervals;
	unsigned int pad;
	unsigned long long dbg[16];    // resolve-kernel phase timestamps (globaltimer ns), measurement only
	// prepare kernel -> commit kernel
	unsigned int nev_prep;         // events prepared (without the carried run's)
	unsigned int cj_exists;        // the batch begins inside a stretch inherited from the previous batch
	unsigned int cj_closes;        // ... and a CONST claim closes it inside this batch
	unsigned int anchored;         // some tile of the batch fixes first_above_time
	unsigned long long cj_close;   // index of the first claim after that stretch (n if it never closes)
	double F_anchor;               // first_above_time after the batch when anchored
	double now0p100;               // now[0] + 100
	double E_val;                  // cd_last_empty if the batch carries an empty() marker
	unsigned int hasE_new, trailing_marker;
};

constexpr unsigned ST_AGG = 1u, ST_PREFIX = 2u;

// Tile descriptor of the admit-count look-back (emit kernel, flags variant).
struct alignas(32) SumDesc {
	unsigned long long w; // (epoch << 2) | status
	unsigned long long agg;
	unsigned long long incl;
	unsigned long long pad;
};

// The commit chain.  Batches commit their ControlledDelay state strictly in queue order; the
// order is kept on the device: a commit kernel spins until `*wait_word == wait_val`, and the batch
// publishes `*post_word = post_val` once its state (and admit cursor) is final.
//   one GPU:        wait_word == post_word == the context's chain word; the state stays in DevState.
//   one queue over R GPUs (ring): wait_word is this rank's mailbox, written by the previous rank over
//   NVLink peer memory; the state arrives in the mailbox (in_state) and is pushed on to the next
//   rank's mailbox (post_state) before its word is released at system scope.
// F-chain (commit_fchain_kernel): first_above_time evolves independently of (drop_next, count,
// dropping), and a batch that contains any below-target claim fixes its outgoing value without
// knowing the incoming one.  The commit kernel of a batch therefore publishes its outgoing value in
// its prologue, into the slot its successor reads, and the successor resolves its inherited stretch
// before the ControlledDelay state itself arrives.  A mailbox receives one value per step, so the
// slot of queue position s in a ring of R is (s / R) % FSLOTS.
constexpr int FSLOTS = 16; // more than the steps that can be in flight
struct FSlot {
	unsigned long long tag; // queue position of the batch this value is the INPUT of
	double F;
};
struct ChainMem {
	unsigned long long seq; // the chain word / mailbox sequence number
	unsigned long long err; // hand-over timeouts (a broken chain must not hang the device)
	DevState st;            // ring mode: the state handed to this rank
	FSlot fs[FSLOTS];       // F-chain slots written by the predecessor (this context itself without a ring)
};
struct ChainLink {
	unsigned long long *wait_word, *post_word, *err;
	unsigned long long wait_val, post_val, timeout_ns;
	const DevState *in_state;
	DevState *post_state;
};

// dynamic shared memory of a kernel (the host emulation of tests/emu substitutes a plain buffer)
#ifndef CB_DYN_SHARED
#define CB_DYN_SHARED(T, name) extern __shared__ __align__(16) T name[]
#endif

// Points where the code relies on the lanes of a warp (or the warps of a CTA) having executed the
// preceding loads together before one of them goes on to overwrite what was read: true of converged
// SIMT execution, not of the host emulation, whose threads are free-running.  Expands to nothing in
// the CUDA build; the emulation makes it a block barrier.  (Round-2 item: pass such values through
// shared memory instead, which removes the reliance altogether.)
#ifndef CB_EMU_LOCKSTEP
#define CB_EMU_LOCKSTEP()
#endif

#ifndef CB_EMU // tests/emu/cuda_emu.h supplies host versions of the PTX helpers below
__device__ __forceinline__ unsigned long long gtimer() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
	asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double *p) {
	double v;
	asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
#endif // CB_EMU

__device__ __forceinline__ void chain_copy_state(DevState *dst, const DevState *src) {
	dst->F = src->F;
	dst->D = src->D;
	dst->C = src->C;
	dst->B = src->B;
	dst->hasE = src->hasE;
	dst->E = src->E;
	dst->cursor = src->cursor;
	dst->tot_drops = src->tot_drops;
	dst->tot_admits = src->tot_admits;
}
// one thread: wait for the previous batch of the queue; in ring mode take over its state
__device__ __forceinline__ void chain_wait(const ChainLink &ln, DevState *g) {
	const unsigned long long t0 = gtimer();
	if (ln.in_state) {
		while (ld_acquire_sys(ln.wait_word) != ln.wait_val) {
			if (gtimer() - t0 > ln.timeout_ns) {
				atomicAdd(ln.err, 1ull);
				break;
			}
		}
		chain_copy_state(g, ln.in_state);
	} else {
		while (ld_acquire(ln.wait_word) != ln.wait_val) {
			if (gtimer() - t0 > ln.timeout_ns) {
				atomicAdd(ln.err, 1ull);
				break;
			}
		}
	}
}
// one thread, after the batch's last write to DevState
__device__ __forceinline__ void chain_post(const ChainLink &ln, const DevState *g) {
	if (ln.post_state) {
		chain_copy_state(ln.post_state, g); // peer memory (NVLink) unless the ring has one member
		__threadfence_system();
		st_release_sys(ln.post_word, ln.post_val);
	} else {
		__threadfence();
		st_release(ln.post_word, ln.post_val);
	}
}

// ---------------------------------------------------------------------------------------
// utils.currentMillis(), lib/utils.js:198-204:  secs*1000 + ns/1000000 with each operation
// rounded once (no FMA contraction).  `t` is process.hrtime() as total nanoseconds >= 0.
//
// The quotient ns/1e6 (0 <= ns < 1e9, an integer) must be the correctly rounded IEEE
// quotient.  It is computed with a reciprocal multiply plus one residual correction
// (Markstein); tools/verify_div1e6.c proves exhaustively, for all 10^9 dividends, that the
// result is bit-identical to the IEEE division.  This keeps the FP64 pipe (the scarce one
// for this kernel) at 3 ops instead of the ~25 of a generic division.
__device__ __forceinline__ double div_1e6(double x) {
	const double R = 1e-6; // RN(1/1e6)
	double q = __dmul_rn(x, R);
	double r = __fma_rn(-q, 1000000.0, x);
	return __fma_rn(r, R, q);
}
__device__ __forceinline__ double ns_to_ms(long long t) {
	unsigned long long ut = (unsigned long long)t;
	unsigned long long secs = ut / 1000000000ull;
	unsigned int ns = (unsigned int)(ut - secs * 1000000000ull);
	double secs2ms = __dmul_rn((double)secs, 1000.0);
	double ns2ms = div_1e6((double)ns);
	return __dadd_rn(secs2ms, ns2ms);
}

template <bool NS> struct TimeIn;
template <> struct TimeIn<true> {
	using raw_t = long long;
	static __device__ __forceinline__ double to_ms(long long v) { return ns_to_ms(v); }
};
template <> struct TimeIn<false> {
	using raw_t = double;
	static __device__ __forceinline__ double to_ms(double v) { return v; }
};

// ControlledDelay#canDrop's three-way outcome for a claim given the first_above_time in
// force (lib/codel.js:34-46), refined by the re-entry test of :72:
//   0 = not ok to drop, 1 = ok, 2 = ok and (now - first_above_time >= INTERVAL)
__device__ __forceinline__ int claim_class(bool above, double now, double F) {
	if (!above || F == 0.0 || !(now >= F))
		return 0;
	return (__dsub_rn(now, F) >= CODEL_INTERVAL) ? 2 : 1;
}

// getDropNext(), lib/codel.js:48-50
__device__ __forceinline__ double drop_next_after(double now, unsigned long long count) {
	double step = __ddiv_rn(CODEL_INTERVAL, __dsqrt_rn((double)count));
	return __dadd_rn(now, step);
}

// One step of ControlledDelay#overloaded (lib/codel.js:56-86) on explicit state.
struct CodelRegs {
	double F, D;
	unsigned long long C;
	int B;
};
__device__ __forceinline__ bool codel_step(CodelRegs &s, double target, double now, double start) {
	bool ok = false;
	double soj = __dsub_rn(now, start);
	if (soj < target) {
		s.F = 0.0;
	} else if (s.F == 0.0) {
		s.F = __dadd_rn(now, CODEL_INTERVAL);
	} else if (now >= s.F) {
		ok = true;
	}
	bool drop = false;
	if (s.B) {
		if (!ok) {
			s.B = 0;
		} else if (now >= s.D) {
			drop = true;
			s.C += 1;
		}
	} else if (ok && ((__dsub_rn(now, s.D) < CODEL_INTERVAL) ||
	    (__dsub_rn(now, s.F) >= CODEL_INTERVAL))) {
		drop = true;
		s.B = 1;
		if (__dsub_rn(now, s.D) < CODEL_INTERVAL)
			s.C = s.C > 2 ? s.C - 2 : 1;
		else
			s.C = 1;
		s.D = drop_next_after(now, s.C);
	}
	return drop;
}

} // namespace cb
