// cb_peers.cuh -- per-batch GLOBAL drop/admit counts over the GPUs of one box without a collective.
//
// Every rank owns a board in its HBM.  After a batch's counts are final the rank stores them, tagged
// with the batch number, into its row of EVERY rank's board (NVLink peer stores; the payload first,
// then the tag at system scope).  A rank then finds the global counts of step s by reading its own
// board: once all R rows carry tag s+1 the sums are the counts of step s over the whole sharded
// queue, and the running totals are the pool's global 'claim-timeout' counter (lib/pool.js:269-283)
// as of that step.  Nothing on the data path waits for this: the post is done by one thread of the
// last CTA of the batch's emit kernel (no extra launch), and ranks never block one another (a rank that runs
// ahead is only held once it is PB_SLOTS batches in front of the slowest reader: flow control through
// the ack words, which each rank stores into its peers' boards the same way).
// The literal "single NCCL allreduce" of the north star stays available (cb_allreduce_counts).
#pragma once
#include "cb_device.cuh"

namespace cb {

constexpr int PB_MAXR = 16;  // ranks (GPUs of one NVSwitch domain used by one sharded pool)
constexpr int PB_SLOTS = 64; // batches a rank may run ahead of the slowest reader

struct PeerRec {
	unsigned long long tag; // step + 1 once the payload is in place
	unsigned long long drops, admits, n;
};
struct PeerBoard {
	PeerRec rec[PB_MAXR][PB_SLOTS]; // rec[r][s % PB_SLOTS]: rank r's counts of its batch s (stored by rank r)
	PeerRec glob[PB_SLOTS];         // glob[s % PB_SLOTS]: running global totals after step s (this rank's finaliser)
	unsigned long long ack[PB_MAXR]; // ack[r]: steps rank r has finalised (stored by rank r)
	unsigned long long gdone;        // steps this rank has finalised
	unsigned long long tot[3];       // running global totals {drops, admits, n} after gdone steps
	unsigned int lock;               // finaliser lock (post kernels of different lanes may overlap)
	unsigned int pad;
	unsigned long long err;          // flow-control / wait timeouts
};
struct PeerLink {
	PeerBoard *mine;
	PeerBoard *const *peers; // [R] device pointers (peer memory; peers[me] == mine)
	int R, me;
	unsigned long long step; // this rank's batch number
	unsigned long long timeout_ns;
};

// finalise every step whose R rows are in place (one thread at a time per board)
__device__ __forceinline__ void peer_finalise(const PeerLink &pl) {
	PeerBoard *b = pl.mine;
	if (atomicCAS(&b->lock, 0u, 1u) != 0u)
		return; // another post kernel of this rank is at it
	__threadfence();
	unsigned long long g = b->gdone;
	const unsigned long long g0 = g;
	while (true) {
		unsigned long long d = 0, a = 0, n = 0;
		bool all = true;
		for (int r = 0; r < pl.R; r++) {
			const PeerRec *q = &b->rec[r][g % PB_SLOTS];
			if (ld_acquire_sys(&q->tag) != g + 1) {
				all = false;
				break;
			}
			d += ld_relaxed_u64(&q->drops);
			a += ld_relaxed_u64(&q->admits);
			n += ld_relaxed_u64(&q->n);
		}
		if (!all)
			break;
		b->tot[0] += d;
		b->tot[1] += a;
		b->tot[2] += n;
		PeerRec *o = &b->glob[g % PB_SLOTS];
		o->drops = b->tot[0];
		o->admits = b->tot[1];
		o->n = b->tot[2];
		__threadfence();
		st_release(&o->tag, g + 1);
		g++;
	}
	if (g != g0) {
		b->gdone = g;
		__threadfence();
		for (int r = 0; r < pl.R; r++) // flow control: this rank has consumed steps < g
			st_release_sys(&pl.peers[r]->ack[pl.me], g);
	}
	__threadfence();
	atomicExch(&b->lock, 0u);
}

// one thread, once the batch's counts are final (the last CTA of its emit kernel): publish them to every rank,
// then finalise whatever is complete on this rank's board
__device__ __forceinline__ void peer_post(const PeerLink &pl, const BatchCtl *bc) {
	const unsigned long long s = pl.step;
	const unsigned long long drops = bc->lcounts[0], admits = bc->lcounts[1], n = bc->lcounts[2];
	const unsigned long long t0 = gtimer();
	for (int r = 0; r < pl.R; r++) {
		if (s >= (unsigned long long)PB_SLOTS) { // the slot's previous use (step s - PB_SLOTS) must have been consumed by rank r
			while (ld_acquire_sys(&pl.mine->ack[r]) + PB_SLOTS <= s) {
				if (gtimer() - t0 > pl.timeout_ns) {
					atomicAdd(&pl.mine->err, 1ull);
					break;
				}
			}
		}
		PeerRec *d = &pl.peers[r]->rec[pl.me][s % PB_SLOTS];
		d->drops = drops;
		d->admits = admits;
		d->n = n;
	}
	__threadfence_system(); // every payload before any tag
	for (int r = 0; r < pl.R; r++)
		st_release_sys(&pl.peers[r]->rec[pl.me][s % PB_SLOTS].tag, s + 1);
	peer_finalise(pl);
}

// one thread: wait (bounded) until this rank has finalised `steps` steps, then copy the totals out
__global__ void peer_counts_kernel(const PeerLink pl, unsigned long long steps, unsigned long long *out4) {
	if (threadIdx.x != 0)
		return;
	const unsigned long long t0 = gtimer();
	PeerBoard *b = pl.mine;
	while (true) {
		peer_finalise(pl);
		if (ld_acquire(&b->gdone) >= steps)
			break;
		if (gtimer() - t0 > pl.timeout_ns) {
			atomicAdd(&b->err, 1ull);
			break;
		}
	}
	const unsigned long long g = ld_acquire(&b->gdone);
	out4[0] = g;
	if (steps > 0 && g >= steps && steps + PB_SLOTS > g) { // the requested step's record is still on the board
		const PeerRec *o = &b->glob[(steps - 1) % PB_SLOTS];
		out4[1] = o->drops;
		out4[2] = o->admits;
		out4[3] = o->n;
	} else { // the latest totals
		out4[1] = b->tot[0];
		out4[2] = b->tot[1];
		out4[3] = b->tot[2];
	}
}

} // namespace cb
