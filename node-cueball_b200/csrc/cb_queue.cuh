// cb_queue.cuh -- the pending-claim queue (`p_waiters`) resident in HBM: enqueue side of
// ConnectionPool#claim, hard timeouts, and the slot-idle drain as ONE closed-form event
// (SURVEY.md 8(f) row 1).  Executable specification, checked on CPU against the pinned model of the reference (tests/):
// tools/proto_queue.py; these kernels follow it step by step.
//
// STATUS: compiled for sm_100a, NOT yet validated on hardware (the round's GPU budget was spent
// before this row was reached); its GPU tests are skipped unless CB_TEST_UNVALIDATED=1.
//
// Reference (file:line in node-cueball):
//   claim(): claimTimeout = getMaxIdle()      lib/pool.js:874-885, lib/codel.js:109-118
//            idle slots first, else enqueue   lib/pool.js:929-966  (+ counters :887,962-963)
//   hard timeout of a waiting handle          lib/connection-fsm.js:626-642
//   slot turns idle: the drain loop           lib/pool.js:733-757, lib/codel.js:34-94
//
// Layout: a power-of-two ring of structure-of-arrays entries, FIFO order = position order:
// start_ns (int64), deadline_ns (int64), ticket (int64), waiting (uint8).  Positions are absolute
// (monotone); entry of position q lives at q & mask.  One CTA per operation: the operations are
// latency-bound bookkeeping (a bisection, a closed form, a prefix sum over the handles one event
// consumes), not bandwidth work.
#pragma once
#include "cb_device.cuh"

namespace cb {

#ifdef CB_QTPB // the host emulation of tests/emu runs the same code with a narrower CTA
constexpr int QTPB = CB_QTPB;
#else
constexpr int QTPB = 1024;
#endif
constexpr long long Q_ADMIT = 0, Q_DROP = 1, Q_SKIPPED = 2, Q_EXPIRED = 3;
constexpr unsigned Q_ERR_NONMONO = 1u, Q_ERR_FULL = 2u, Q_ERR_RESULTS = 4u;

struct QueueDev { // mirrors cb_queue_info in include/cueball_b200.h
	unsigned long long head, tail, next_ticket, idle;
	unsigned long long c_claim, c_queued, c_maxq, c_timeout; // lib/pool.js:269-283
	unsigned long long n_res;
	unsigned int err, pad;
};
struct QueueParams {
	QueueDev *q;
	DevState *g;
	long long *start_ns, *deadline_ns, *ticket;
	unsigned char *waiting;
	unsigned long long mask;    // capacity - 1
	long long *res;             // [res_cap][3] (ticket, code, slot)
	unsigned long long res_cap;
	long long nslots;
};

// block-wide exclusive prefix sum of one unsigned per thread; returns the exclusive value and the
// block total through `total` (all threads).  s_w: 32 unsigneds of shared memory.
__device__ __forceinline__ unsigned int block_excl_scan(unsigned int v, unsigned int *s_w, unsigned int &total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned int x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= o)
			x += y;
	}
	__syncthreads(); // s_w may still be read from the previous call
	if (lane == 31)
		s_w[warp] = x;
	__syncthreads();
	unsigned int w = (threadIdx.x < QTPB / 32) ? s_w[threadIdx.x] : 0u;
	if (warp == 0) {
		unsigned int ws = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned int y = __shfl_up_sync(0xffffffffu, ws, o);
			if (lane >= o)
				ws += y;
		}
		s_w[lane] = ws; // inclusive over warps
	}
	__syncthreads();
	total = s_w[QTPB / 32 - 1];
	const unsigned int before = warp > 0 ? s_w[warp - 1] : 0u;
	return before + x - v;
}

__device__ __forceinline__ void q_record(const QueueParams &p, unsigned long long idx, long long ticket, long long code,
    long long slot) {
	if (idx < p.res_cap) {
		p.res[3 * idx] = ticket;
		p.res[3 * idx + 1] = code;
		p.res[3 * idx + 2] = slot;
	}
}

// ---- hard timeouts due at clock T (lib/connection-fsm.js:626-642), in queue (= ticket) order ------
__device__ void q_expire_block(const QueueParams &p, long long T_ns, double target, unsigned int *s_w,
    unsigned long long *s_u64) {
	QueueDev *q = p.q;
	const unsigned long long h = q->head, t = q->tail;
	if (threadIdx.x == 0) {
		// only handles with start <= T - 3*target can be due (getMaxIdle never returns less): they are a
		// prefix of the queue because enqueue clocks are non-decreasing
		const long long horizon = T_ns - (long long)(target * 3.0 * 1000000.0);
		unsigned long long lo = h, hi = t;
		while (lo < hi) {
			const unsigned long long mid = lo + (hi - lo) / 2;
			if (p.start_ns[mid & p.mask] <= horizon)
				lo = mid + 1;
			else
				hi = mid;
		}
		s_u64[0] = lo;
	}
	__syncthreads();
	const unsigned long long hi = s_u64[0];
	unsigned long long base = q->n_res, expired = 0;
	for (unsigned long long c = h; c < hi; c += QTPB) {
		const unsigned long long pos = c + threadIdx.x;
		bool due = false;
		if (pos < hi) {
			const unsigned long long k = pos & p.mask;
			due = p.waiting[k] != 0 && p.deadline_ns[k] <= T_ns;
		}
		unsigned int total;
		const unsigned int ex = block_excl_scan(due ? 1u : 0u, s_w, total);
		if (due) {
			const unsigned long long k = pos & p.mask;
			p.waiting[k] = 0;
			q_record(p, base + expired + ex, p.ticket[k], Q_EXPIRED, -1);
		}
		expired += total;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (base + expired > p.res_cap)
			atomicOr(&q->err, Q_ERR_RESULTS);
		q->n_res = base + expired;
		q->c_timeout += expired; // lib/connection-fsm.js:629-630
	}
	__syncthreads();
}

__global__ void __launch_bounds__(QTPB, 1) queue_expire_kernel(const QueueParams p, long long T_ns) {
	__shared__ unsigned int s_w[32];
	__shared__ unsigned long long s_u64[4];
	q_expire_block(p, T_ns, p.g->target, s_w, s_u64);
}

// ---- claim(): n arrivals, clock = each claim's own start (lib/pool.js:874-966) ---------------------
__global__ void __launch_bounds__(QTPB, 1) queue_push_kernel(const QueueParams p, const long long *starts,
    unsigned long long n, double *timeouts_out) {
	__shared__ unsigned int s_w[32];
	__shared__ unsigned long long s_u64[4];
	QueueDev *q = p.q;
	DevState *g = p.g;
	const double target = g->target;
	if (n == 0)
		return;
	q_expire_block(p, starts[0], target, s_w, s_u64); // timers that came due before the first arrival
	const unsigned long long tail = q->tail, head = q->head, t0 = q->next_ticket, base = q->n_res;
	const unsigned long long idle = q->idle, cursor = g->cursor;
	const unsigned long long take = idle < n ? idle : n; // lib/pool.js:934-951: idle slots first
	const unsigned long long m = n - take;
	const int hasE = g->hasE;
	const double E = g->E;
	const bool full = (tail - head) + m > p.mask + 1;
	bool nonmono = false;
	for (unsigned long long i = threadIdx.x; i < n; i += QTPB) {
		const long long st = starts[i];
		const double now = ns_to_ms(st);
		// getMaxIdle, lib/codel.js:109-118 (`undefined < x` is false)
		const double bound = target * 10.0;
		const double timeout = (hasE && E < __dsub_rn(now, bound)) ? target * 3.0 : bound;
		if (timeouts_out)
			timeouts_out[i] = timeout;
		const long long prev = i > 0 ? starts[i - 1] : ((tail > head) ? p.start_ns[(tail - 1) & p.mask] : st);
		if (st < prev)
			nonmono = true;
		if (i < take) {
			q_record(p, base + i, (long long)(t0 + i), Q_ADMIT,
			    p.nslots > 0 ? (long long)((cursor + i) % (unsigned long long)p.nslots) : -1);
		} else if (!full) {
			const unsigned long long k = (tail + (i - take)) & p.mask;
			p.start_ns[k] = st;
			p.deadline_ns[k] = st + (long long)(timeout * 1000000.0);
			p.ticket[k] = (long long)(t0 + i);
			p.waiting[k] = 1;
		}
	}
	if (nonmono)
		atomicOr(&q->err, Q_ERR_NONMONO);
	__syncthreads();
	if (threadIdx.x == 0) {
		if (full) {
			atomicOr(&q->err, Q_ERR_FULL);
			return; // nothing was enqueued; the host reports the error
		}
		if (base + take > p.res_cap)
			atomicOr(&q->err, Q_ERR_RESULTS);
		q->n_res = base + take;
		q->idle = idle - take;
		g->cursor = cursor + take;
		g->tot_admits += take;
		q->tail = tail + m;
		q->next_ticket = t0 + n;
		q->c_claim += n;  // lib/pool.js:887
		q->c_queued += m; // lib/pool.js:963
		if (m && tail + m - head > q->c_maxq) // lib/pool.js:962 (the queue only grows inside a push)
			q->c_maxq = tail + m - head;
	}
}

// ---- handle.cancel() while waiting (lib/connection-fsm.js:580-587) ---------------------------------
__global__ void __launch_bounds__(QTPB, 1) queue_cancel_kernel(const QueueParams p, const long long *tickets,
    unsigned long long n) {
	const QueueDev *q = p.q;
	const unsigned long long h = q->head, t = q->tail;
	for (unsigned long long i = threadIdx.x; i < n; i += QTPB) {
		const long long tk = tickets[i];
		unsigned long long lo = h, hi = t; // tickets increase along the queue
		while (lo < hi) {
			const unsigned long long mid = lo + (hi - lo) / 2;
			if (p.ticket[mid & p.mask] < tk)
				lo = mid + 1;
			else
				hi = mid;
		}
		if (lo < t && p.ticket[lo & p.mask] == tk)
			p.waiting[lo & p.mask] = 0;
	}
}

// ---- k slots turn idle at clock T: lib/pool.js:733-757 for every slot, every clock read = T ---------
__global__ void __launch_bounds__(QTPB, 1) queue_idle_kernel(const QueueParams p, long long T_ns, unsigned long long k) {
	__shared__ unsigned int s_w[32];
	__shared__ unsigned long long s_u64[4];
	__shared__ unsigned long long s_P, s_drop_from, s_L;
	__shared__ int s_drop_first;
	QueueDev *q = p.q;
	DevState *g = p.g;
	const double target = g->target;
	q_expire_block(p, T_ns, target, s_w, s_u64);
	const double T = ns_to_ms(T_ns);
	const unsigned long long h = q->head, t = q->tail, Q = t - h;
	const double F0 = g->F, D0 = g->D;
	const unsigned long long C0 = g->C;
	const int B0 = g->B;
	unsigned long long left = k;
	if (Q > 0 && k > 0) {
		// -- P: the claims at or above target are a prefix (sojourn = T - start is non-increasing);
		//    bisection with the exact comparison of lib/codel.js:35-37
		if (threadIdx.x == 0) {
			unsigned long long lo = 0, hi = Q;
			while (lo < hi) {
				const unsigned long long mid = lo + (hi - lo) / 2;
				const double st = ns_to_ms(p.start_ns[(h + mid) & p.mask]);
				if (!(__dsub_rn(T, st) < target))
					lo = mid + 1; // still above target
				else
					hi = mid;
			}
			s_P = lo;
		}
		__syncthreads();
		const unsigned long long P = s_P;
		// -- the drop pattern along the prefix (lib/codel.js:34-86 at constant now); every thread
		//    evaluates the same scalars
		const bool ok = (P > 0) && (F0 != 0.0) && (T >= F0);          // lib/codel.js:39-42
		const double F1 = (P > 0 && F0 == 0.0) ? __dadd_rn(T, CODEL_INTERVAL) : F0;
		const bool near_prev = __dsub_rn(T, D0) < CODEL_INTERVAL;
		const bool enter = !B0 && ok && (near_prev || (__dsub_rn(T, F0) >= CODEL_INTERVAL)); // lib/codel.js:70-72
		const bool all_drop = (P > 0) && B0 && ok && (T >= D0);      // lib/codel.js:65-68, drop_next stays put
		unsigned long long C1 = C0;
		double D1 = D0;
		bool drop_first = false;
		unsigned long long drop_from = Q; // drops = {0} if drop_first, plus [drop_from, P)
		if (all_drop) {
			drop_from = 0;
		} else if (P > 0 && enter) {
			drop_first = true;
			C1 = near_prev ? (C0 > 2 ? C0 - 2 : 1) : 1;             // lib/codel.js:76-81
			D1 = drop_next_after(T, C1);                            // lib/codel.js:83
			if (T >= D1)
				drop_from = 1; // only if 100/sqrt(count) vanishes against T
		}
		// -- dequeue until k handles have been admitted: handle i is dequeued iff fewer than k of
		//    the handles before it were admitted; one record per dequeued handle, in order
		const unsigned long long base = q->n_res, cursor = g->cursor;
		unsigned long long cum = 0, L = Q, ndrop_w = 0;
		if (threadIdx.x == 0)
			s_L = Q;
		for (unsigned long long c = 0; c < Q && cum < k; c += QTPB) {
			const unsigned long long i = c + threadIdx.x;
			bool in = i < Q, w = false, dropped = false;
			long long tk = 0;
			if (in) {
				const unsigned long long kk = (h + i) & p.mask;
				w = p.waiting[kk] != 0;
				tk = p.ticket[kk];
				dropped = (drop_first && i == 0) || (i >= drop_from && i < P);
			}
			const bool adm = in && w && !dropped;
			unsigned int total;
			const unsigned int ex = block_excl_scan(adm ? 1u : 0u, s_w, total);
			const bool popped = in && (cum + ex < k);
			if (popped) {
				if (!w)
					q_record(p, base + i, tk, Q_SKIPPED, -1);      // lib/pool.js:738-740
				else if (dropped)
					q_record(p, base + i, tk, Q_DROP, -1);         // lib/pool.js:742-745
				else
					q_record(p, base + i, tk, Q_ADMIT,             // lib/pool.js:747
					    p.nslots > 0 ? (long long)((cursor + cum + ex) % (unsigned long long)p.nslots) : -1);
				if (adm && cum + ex + 1 == k)
					s_L = i + 1; // the handle that takes the last slot
			}
			unsigned int dtotal;
			(void)block_excl_scan((popped && w && dropped) ? 1u : 0u, s_w, dtotal);
			ndrop_w += dtotal;
			cum += total;
		}
		__syncthreads();
		L = s_L;
		const unsigned long long admitted = cum < k ? cum : k;
		left = k - admitted;
		const unsigned long long pa = L < P ? L : P; // above-target claims actually dequeued
		if (threadIdx.x == 0) {
			double F = F0, D = D0;
			unsigned long long C = C0;
			int B = B0;
			if (pa > 0) {
				F = F1;
				if (B0 && !ok) {
					B = 0;                                      // lib/codel.js:63-64
				} else if (all_drop) {
					C = C0 + pa;                                // every dequeued claim of the prefix counted
				} else if (enter) {
					B = 1;
					C = C1 + ((drop_from == 1 && pa > 1) ? pa - 1 : 0);
					D = D1;
				}
			}
			if (L > pa) { // a below-target claim was dequeued: lib/codel.js:37-38, 63-64
				F = 0.0;
				B = 0;
			}
			g->F = F;
			g->D = D;
			g->C = C;
			g->B = B;
			g->cursor = cursor + admitted;
			g->tot_admits += admitted;
			g->tot_drops += ndrop_w;
			q->head = h + L;
			if (base + L > p.res_cap)
				atomicOr(&q->err, Q_ERR_RESULTS);
			q->n_res = base + L;
			q->c_timeout += ndrop_w; // lib/connection-fsm.js:633-638
		}
	}
	__syncthreads();
	if (threadIdx.x == 0 && left > 0) { // the queue ran empty with slots still idle: lib/pool.js:751-757
		g->E = T;
		g->hasE = 1;
		g->F = 0.0;
		q->idle += left;
	}
}

} // namespace cb
