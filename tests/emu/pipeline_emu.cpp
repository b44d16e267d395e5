/* tests/emu/pipeline_emu.cpp -- the batch pipeline's kernels (csrc/cb_scan.cuh, cb_resolve.cuh,
 * cb_emit.cuh: scan -> prepare -> commit -> emit) compiled for the host under tests/emu/cuda_emu.h
 * and run one CTA at a time, behind one C function shaped like cb_admit_batch_dev, so that
 * tests/test_pipeline_kernels_emulated.py can compare them with the oracle without a GPU.
 * The flags variant of the emit kernel is run as ONE persistent CTA (its CTAs otherwise wait for one
 * another, which would need them to run concurrently).  TEST INFRASTRUCTURE. */
#include "../../node-cueball_b200/csrc/cb_scan.cuh"
#include "../../node-cueball_b200/csrc/cb_resolve.cuh"
#include "../../node-cueball_b200/csrc/cb_emit.cuh"

#include <cstdio>
#include <cstdlib>

using namespace cb;

namespace {
template <typename T> T *zalloc(size_t n) { return static_cast<T *>(std::calloc(n + 8, sizeof(T))); }

/* the scan's persistent grid: a few CTAs, so that warps walk several tiles each (the library uses the SM count) */
unsigned scan_grid(size_t tiles) {
	const char *e = std::getenv("CB_EMU_SCAN_GRID");
	const size_t g = e ? (size_t)std::atoi(e) : 3;
	return (unsigned)(tiles < g ? tiles : g);
}
template <typename F> void run_grid(unsigned grid, int threads, F kernel_call) {
	gridDim.x = grid;
	for (unsigned b = 0; b < grid; b++) {
		if (std::getenv("CB_EMU_TRACE"))
			std::fprintf(stderr, "[emu] block %u/%u x %d threads\n", b, grid, threads);
		blockIdx.x = b;
		emu::launch(threads, kernel_call);
	}
}
} // namespace

struct EmuPipe {
	DevState g;
	ChainMem chain;
	unsigned long long seq;
	int fseed;   /* the state was set from the host since the last batch */
	EmuPipe *next; /* ring: the context whose mailbox this one writes (null: no ring) */
	int ring_n, ring_rank;
};

template <bool NS>
static int pipe_admit(EmuPipe *e, const void *start, const void *now, const unsigned char *flags,
    unsigned long long n, int nslots, unsigned char *decision, int *slot, unsigned long long *info) {
	if (n == 0)
		return 0;
	const size_t tiles = (n + TILE - 1) / TILE, etiles = (n + ETILE - 1) / ETILE;
	const size_t cap = tiles + 2, evcap = (cap + 1) * (EVK + 3) + EV_FRONT;
	BatchCtl *ctl = zalloc<BatchCtl>(1);
	unsigned int *ev_count = zalloc<unsigned int>(cap), *rec_fc = zalloc<unsigned int>(cap),
	             *rec_cend = zalloc<unsigned int>(cap), *xcnt = zalloc<unsigned int>(cap + 1),
	             *hj_info = zalloc<unsigned int>(cap + 1), *ticket = zalloc<unsigned int>(16);
	unsigned long long *ev = zalloc<unsigned long long>(cap * EVK), *evc = zalloc<unsigned long long>(evcap),
	                   *ev_q = zalloc<unsigned long long>(evcap), *ev_d = zalloc<unsigned long long>(evcap),
	                   *rec_code = zalloc<unsigned long long>(cap), *xev = zalloc<unsigned long long>((cap + 1) * 3),
	                   *hj_lo = zalloc<unsigned long long>(cap + 1), *hj_b = zalloc<unsigned long long>(cap + 1),
	                   *tile_now0 = zalloc<unsigned long long>(cap), *iv_lo = zalloc<unsigned long long>(evcap + 8),
	                   *iv_hi = zalloc<unsigned long long>(evcap + 8), *iv_cum = zalloc<unsigned long long>(evcap + 8);
	double *ev_now = zalloc<double>(evcap), *hj_F = zalloc<double>(cap + 1);
	SumDesc *sdesc = zalloc<SumDesc>(cap);

	ScanParams sp;
	sp.start = start;
	sp.now = now;
	sp.flags = flags;
	sp.n = n;
	sp.b = ctl;
	sp.target = e->g.target;
	sp.rec_fc = rec_fc;
	sp.rec_code = rec_code;
	sp.rec_cend = rec_cend;
	sp.tile_now0 = tile_now0;
	sp.ev_count = ev_count;
	sp.ev = ev;
	sp.vec_ok = ((((uintptr_t)start | (uintptr_t)now) & 15) == 0);
	sp.target_ns = (long long)(e->g.target * 1000000.0);

	ResolveParams rp;
	rp.start = start;
	rp.now = now;
	rp.flags = flags;
	rp.n = n;
	rp.tiles = tiles;
	rp.g = &e->g;
	rp.bc = ctl;
	rp.rec_fc = rec_fc;
	rp.rec_code = rec_code;
	rp.rec_cend = rec_cend;
	rp.tile_now0 = tile_now0;
	rp.link.err = &e->chain.err;
	rp.link.wait_word = &e->chain.seq;
	rp.link.timeout_ns = ~0ull;
	if (e->ring_n > 0) { /* as enqueue_batch() does in ring mode */
		rp.link.wait_val = e->seq * (unsigned long long)e->ring_n + (unsigned long long)e->ring_rank;
		rp.link.in_state = &e->chain.st;
		rp.link.post_word = &e->next->chain.seq;
		rp.link.post_state = &e->next->chain.st;
	} else {
		rp.link.wait_val = e->seq;
		rp.link.in_state = nullptr;
		rp.link.post_word = &e->chain.seq;
		rp.link.post_state = nullptr;
	}
	rp.link.post_val = rp.link.wait_val + 1;
	rp.xev = xev;
	rp.xcnt = xcnt;
	rp.hj_lo = hj_lo;
	rp.hj_b = hj_b;
	rp.hj_F = hj_F;
	rp.hj_info = hj_info;
	rp.ev_count = ev_count;
	rp.ev = ev;
	rp.evc = evc;
	rp.ev_now = ev_now;
	rp.ev_q = ev_q;
	rp.ev_d = ev_d;
	rp.ev_cap = evcap;
	rp.iv_lo = iv_lo;
	rp.iv_hi = iv_hi;
	rp.iv_cum = iv_cum;
	rp.decision = decision;
	rp.slot = slot;
	rp.backend = nullptr;
	rp.slot_backend = nullptr;
	rp.nslots = nslots;
	rp.f_in = e->chain.fs;
	rp.f_out = e->ring_n > 0 ? e->next->chain.fs : e->chain.fs;
	rp.f_ring = e->ring_n > 0 ? (unsigned long long)e->ring_n : 1ull;

	EmitParams ep;
	ep.flags = flags;
	ep.n = n;
	ep.tiles = etiles;
	ep.g = &e->g;
	ep.bc = ctl;
	ep.iv_lo = iv_lo;
	ep.iv_hi = iv_hi;
	ep.iv_cum = iv_cum;
	ep.decision = decision;
	ep.slot = slot;
	ep.backend = nullptr;
	ep.slot_backend = nullptr;
	ep.nslots = nslots;
	ep.slot_magic = nslots > 0 ? (~0ull) / (unsigned long long)nslots : 0ull;
	ep.table_smem_bytes = 0;
	ep.sdesc = sdesc;
	ep.epoch = 1;
	ep.link = rp.link;
	ep.peer.mine = nullptr; /* no peer boards in the emulation */
	ep.peer.peers = nullptr;
	ep.peer.R = 0;
	ep.peer.me = 0;
	ep.peer.step = 0;
	ep.peer.timeout_ns = 0;
	ep.peer.dbg = 0;
	ep.ticket = ticket;
	ep.vec_ok = (((uintptr_t)decision & 7) == 0) && (((uintptr_t)slot & 15) == 0);

	if (e->fseed) {
		if (e->ring_n == 0 || (e->ring_rank == 0 && e->seq == 0)) {
			DevState *gp = &e->g;
			FSlot *fs = e->chain.fs;
			const unsigned long long sq = rp.link.wait_val, rn = rp.f_ring;
			run_grid(1, 32, [=] {
				if (threadIdx.x == 0)
					fchain_seed_kernel(gp, fs, sq, rn);
			});
		}
		e->fseed = 0;
	}
	if (flags) {
		/* the flags variant of the emit kernel chains tile admit counts through a look-back; with ONE
		 * persistent CTA walking the tiles in order every look-back finds its predecessors done */
		run_grid(scan_grid(tiles), SC_TPB, [&] { scan_kernel<NS, true>(sp); });
		run_grid(1, RTPB, [&] { resolve_kernel<NS, true>(rp); });
		run_grid(1, ETPB, [&] { emit_kernel<true>(ep); });
	} else {
		run_grid(scan_grid(tiles), SC_TPB, [&] { scan_kernel<NS, false>(sp); });
		run_grid(1, RTPB, [&] { resolve_kernel<NS, false>(rp); });
		const unsigned egrid = (unsigned)(etiles < 8 ? etiles : 8); /* persistent CTAs striding over the emit tiles */
		run_grid(egrid, ETPB, [&] { emit_kernel<false>(ep); });
	}
	e->seq++;
	if (info) {
		info[0] = (unsigned long long)ctl->path;
		info[1] = ctl->n_events;
		info[2] = ctl->n_intervals;
		info[3] = ctl->drops;
	}
	void *fr[] = {ctl, ev_count, rec_fc, rec_cend, xcnt, hj_info, ticket, ev, evc, ev_q, ev_d, rec_code, xev, hj_lo, hj_b,
	    tile_now0, iv_lo, iv_hi, iv_cum, ev_now, hj_F, sdesc};
	for (void *p : fr)
		std::free(p);
	return 0;
}


extern "C" {
EmuPipe *emu_pipe_create(double target, int) { /* (second argument: unused, kept for the callers' signature) */
	EmuPipe *e = static_cast<EmuPipe *>(std::calloc(1, sizeof(EmuPipe)));
	e->g.target = target;
	e->fseed = 1;
	return e;
}
void emu_pipe_destroy(EmuPipe *e) { std::free(e); }
void emu_pipe_set_state(EmuPipe *e, double F, double D, unsigned long long C, int B, int hasE, double E,
    unsigned long long cursor) {
	e->g.F = F;
	e->g.D = D;
	e->g.C = C;
	e->g.B = B;
	e->g.hasE = hasE;
	e->g.E = E;
	e->g.cursor = cursor;
	e->fseed = 1;
}
/* cb_ring_connect: this context drains segment (k, rank) of every step k and hands over to `next` */
void emu_pipe_ring(EmuPipe *e, EmuPipe *next, int nranks, int rank) {
	e->next = next;
	e->ring_n = nranks;
	e->ring_rank = rank;
	e->seq = 0;
	std::memset(&e->chain, 0, sizeof(e->chain));
	e->chain.st = e->g;
	e->fseed = 1;
}
void emu_pipe_get_state(EmuPipe *e, double *F, double *D, unsigned long long *C, int *B, int *hasE, double *E,
    unsigned long long *cursor, unsigned long long *tot_drops, unsigned long long *tot_admits) {
	*F = e->g.F;
	*D = e->g.D;
	*C = e->g.C;
	*B = e->g.B;
	*hasE = e->g.hasE;
	*E = e->g.E;
	*cursor = e->g.cursor;
	*tot_drops = e->g.tot_drops;
	*tot_admits = e->g.tot_admits;
}

/* one batch, nanosecond clocks, no flags; mirrors enqueue_batch() of csrc/cueball_b200.cu.
 * info[0] = path, [1] = events, [2] = intervals, [3] = drops */
int emu_pipe_admit(EmuPipe *e, const long long *start, const long long *now, const unsigned char *flags,
    unsigned long long n, int nslots, unsigned char *decision, int *slot, unsigned long long *info) {
	return pipe_admit<true>(e, start, now, flags, n, nslots, decision, slot, info);
}
/* the scalar entry points (cb_codel_overloaded / _ms / _empty / _max_idle): one-thread kernels on the context's state */
int emu_pipe_overloaded(EmuPipe *e, long long start_ns, long long now_ns) {
	int out = 0;
	DevState *g = &e->g;
	run_grid(1, 1, [&] { overloaded_one_kernel(g, start_ns, now_ns, &out); });
	e->fseed = 1; /* the state changed outside the batch chain */
	return out;
}
int emu_pipe_overloaded_ms(EmuPipe *e, double start_ms, double now_ms) {
	int out = 0;
	DevState *g = &e->g;
	run_grid(1, 1, [&] { overloaded_one_ms_kernel(g, start_ms, now_ms, &out); });
	e->fseed = 1;
	return out;
}
void emu_pipe_empty(EmuPipe *e, long long now_ns) {
	DevState *g = &e->g;
	run_grid(1, 1, [&] { empty_kernel(g, now_ns); });
	e->fseed = 1;
}
double emu_pipe_max_idle(EmuPipe *e, long long now_ns) {
	double out = 0.0;
	const DevState *g = &e->g;
	run_grid(1, 1, [&] { max_idle_kernel(g, now_ns, &out); });
	return out;
}
/* the _ms entry points: clocks already in double milliseconds */
int emu_pipe_admit_ms(EmuPipe *e, const double *start, const double *now, const unsigned char *flags,
    unsigned long long n, int nslots, unsigned char *decision, int *slot, unsigned long long *info) {
	return pipe_admit<false>(e, start, now, flags, n, nslots, decision, slot, info);
}
}
