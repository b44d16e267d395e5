/* tests/emu/queue_emu.cpp -- the kernels of node-cueball_b200/csrc/cb_queue.cuh compiled for the
 * host under tests/emu/cuda_emu.h (64 OS threads standing for the CTA), behind a small C interface
 * shaped like the cb_queue_* entry points, so that tests/test_queue_kernels_emulated.py can run
 * them against the oracle without a GPU.  TEST INFRASTRUCTURE. */
#define CB_QTPB 64
#include "../../node-cueball_b200/csrc/cb_queue.cuh"

#include <cstdlib>

using namespace cb;

struct EmuQueue {
	QueueDev q;
	DevState g;
	std::vector<long long> start, deadline, ticket, res;
	std::vector<unsigned char> waiting;
	long long nslots;
	QueueParams params() {
		QueueParams p;
		p.q = &q;
		p.g = &g;
		p.start_ns = start.data();
		p.deadline_ns = deadline.data();
		p.ticket = ticket.data();
		p.waiting = waiting.data();
		p.mask = start.size() - 1;
		p.res = res.data();
		p.res_cap = res.size() / 3;
		p.nslots = nslots;
		return p;
	}
};

extern "C" {
EmuQueue *emu_queue_create(double target, long long nslots, unsigned long long cap_pow2, unsigned long long res_cap,
    unsigned long long idle0) {
	EmuQueue *e = new EmuQueue();
	std::memset(&e->q, 0, sizeof(e->q));
	std::memset(&e->g, 0, sizeof(e->g));
	e->g.target = target;
	e->q.idle = idle0;
	e->start.assign(cap_pow2, 0);
	e->deadline.assign(cap_pow2, 0);
	e->ticket.assign(cap_pow2, 0);
	e->waiting.assign(cap_pow2, 0);
	e->res.assign(res_cap * 3, 0);
	e->nslots = nslots;
	return e;
}
void emu_queue_destroy(EmuQueue *e) { delete e; }
void emu_queue_set_state(EmuQueue *e, double F, double D, unsigned long long C, int B, int hasE, double E) {
	e->g.F = F;
	e->g.D = D;
	e->g.C = C;
	e->g.B = B;
	e->g.hasE = hasE;
	e->g.E = E;
}
void emu_queue_get_state(EmuQueue *e, double *F, double *D, unsigned long long *C, int *B, int *hasE, double *E,
    unsigned long long *cursor) {
	*F = e->g.F;
	*D = e->g.D;
	*C = e->g.C;
	*B = e->g.B;
	*hasE = e->g.hasE;
	*E = e->g.E;
	*cursor = e->g.cursor;
}
void emu_queue_push(EmuQueue *e, const long long *starts, unsigned long long n, double *timeouts) {
	const QueueParams p = e->params();
	emu::launch(QTPB, [&] { queue_push_kernel(p, starts, n, timeouts); });
}
void emu_queue_cancel(EmuQueue *e, const long long *tickets, unsigned long long n) {
	const QueueParams p = e->params();
	emu::launch(QTPB, [&] { queue_cancel_kernel(p, tickets, n); });
}
void emu_queue_expire(EmuQueue *e, long long T_ns) {
	const QueueParams p = e->params();
	emu::launch(QTPB, [&] { queue_expire_kernel(p, T_ns); });
}
void emu_queue_idle(EmuQueue *e, long long T_ns, unsigned long long k) {
	const QueueParams p = e->params();
	emu::launch(QTPB, [&] { queue_idle_kernel(p, T_ns, k); });
}
unsigned long long emu_queue_results(EmuQueue *e, long long *out, unsigned long long cap) {
	unsigned long long n = e->q.n_res;
	if (n > cap)
		n = cap;
	std::memcpy(out, e->res.data(), n * 3 * sizeof(long long));
	e->q.n_res = 0;
	return n;
}
void emu_queue_info(EmuQueue *e, QueueDev *out) { *out = e->q; }
/* what the host side of the library does without a kernel: the slot table's size after a topology
 * change (queue_params() reads the context's nslots) and cb_queue_trim_idle's bookkeeping word */
void emu_queue_set_nslots(EmuQueue *e, long long nslots) { e->nslots = nslots; }
unsigned long long emu_queue_trim_idle(EmuQueue *e, unsigned long long k) {
	const unsigned long long take = k < e->q.idle ? k : e->q.idle;
	e->q.idle -= take;
	return take;
}
}
