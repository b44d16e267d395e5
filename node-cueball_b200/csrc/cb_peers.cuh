// cb_peers.cuh -- per-batch GLOBAL drop/admit counts over the GPUs of one box without a collective.
//
// Every rank owns a board in its HBM and stores each batch's counts, tagged with the batch number, into
// it (local stores).  A rank finds the global counts of step s by reading row s of EVERY rank's board over
// NVLink peer memory (remote loads): once all R rows carry tag s+1 the sums are the counts of step s over
// the whole sharded queue, and the running totals are the pool's global 'claim-timeout' counter
// (lib/pool.js:269-283) as of that step.  Nothing on the data path waits for this: post and read are done
// by one thread of the batch's resolve kernel after the commit chain has moved on (no extra launch, no
// fence), and ranks never block one another (a rank that runs ahead is only held once it is PB_SLOTS
// batches in front of the slowest reader: flow control through ack words).
// The literal "single NCCL allreduce" of the north star stays available (cb_allreduce_counts).
#pragma once
#include "cb_device.cuh"

namespace cb {

constexpr int PB_MAXR = 16;  // ranks (GPUs of one NVSwitch domain used by one sharded pool)
constexpr int PB_SLOTS = 64; // batches a rank may run ahead of the slowest reader

// PULL, not push.  A rank writes its records into its OWN board (local stores); whoever wants the sums
// reads the other ranks' rows over NVLink (remote LOADS).  Round 2 measured the alternative on 2 GPUs at
// C3: any remote STORE issued from a kernel of the batch pipeline (from the resolve kernel, from the emit
// kernel's last CTA or from a kernel of its own; with system-scope fences or without any) grew the 58-us
// step to 72-82 us, while the peer mapping alone cost nothing.  The only remote stores left are the
// flow-control acks, one per rank every PB_ACK_EVERY steps.
// A rank's record of a batch is ONE self-tagged 16-byte word {step+1 : 32 | drops : 32, admits : 32 | n : 32}
// (a batch has fewer than 2^32 claims), written and read by ONE 128-bit access, so a word whose tag reads
// step+1 carries that step's counts: no payload-then-flag ordering, no fences, one remote load per rank per step.
struct alignas(16) PeerWord {
	unsigned long long tag; // step + 1
	unsigned long long v;
};
struct alignas(64) PeerRec {
	PeerWord drops, admits, n;
	PeerWord pad;
};
constexpr int PB_ACK_EVERY = 16; // a reader tells the writers how far it has read once per this many steps
struct PeerBoard {
	PeerWord own[PB_SLOTS];          // own[s % PB_SLOTS]: THIS rank's counts of its batch s (stored locally, read by everyone)
	PeerRec glob[PB_SLOTS];          // glob[s % PB_SLOTS]: running global totals after step s (this rank's finaliser)
	unsigned long long ack[PB_MAXR]; // ack[r]: steps rank r has read from this board (stored by rank r, rarely)
	unsigned long long gdone;        // steps this rank has finalised
	unsigned long long acked;        // value of gdone last sent to the other ranks
	unsigned long long tot[3];       // running global totals {drops, admits, n} after gdone steps
	unsigned int lock;               // finaliser lock (resolve kernels of different lanes may overlap)
	unsigned int pad;
	unsigned long long err;          // flow-control / wait timeouts
};
struct PeerLink {
	PeerBoard *mine;
	PeerBoard *const *peers; // [R] device pointers (peer memory; peers[me] == mine)
	int R, me;
	unsigned long long step; // this rank's batch number
	unsigned long long timeout_ns;
	unsigned int dbg;        // measurement only (CB_PEERS_DEBUG): 1 = nothing posted, 4 = no finalise
};

#ifndef CB_EMU
__device__ __forceinline__ void st_word(PeerWord *p, unsigned long long tag, unsigned long long v) {
	asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(tag), "l"(v) : "memory");
}
__device__ __forceinline__ PeerWord ld_word(const PeerWord *p) {
	PeerWord w;
	asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.tag), "=l"(w.v) : "l"(p) : "memory");
	return w;
}
#endif

// finalise every step that every rank has posted (one thread at a time per board): remote loads of the other
// ranks' rows, all issued before the first is looked at
__device__ __forceinline__ void peer_finalise(const PeerLink &pl, bool flush_acks) {
	PeerBoard *b = pl.mine;
	if (atomicCAS(&b->lock, 0u, 1u) != 0u)
		return; // another kernel of this rank is at it
	__threadfence();
	unsigned long long g = b->gdone;
	while (true) {
		PeerWord w[PB_MAXR];
		for (int r = 0; r < pl.R; r++)
			w[r] = ld_word(&pl.peers[r]->own[g % PB_SLOTS]);
		unsigned long long d = 0, a = 0, n = 0;
		bool all = true;
		for (int r = 0; r < pl.R; r++) {
			all = all && (unsigned int)(w[r].tag >> 32) == (unsigned int)(g + 1);
			d += w[r].tag & 0xffffffffull;
			a += w[r].v >> 32;
			n += w[r].v & 0xffffffffull;
		}
		if (!all)
			break;
		b->tot[0] += d;
		b->tot[1] += a;
		b->tot[2] += n;
		PeerRec *o = &b->glob[g % PB_SLOTS];
		st_word(&o->drops, g + 1, b->tot[0]);
		st_word(&o->admits, g + 1, b->tot[1]);
		st_word(&o->n, g + 1, b->tot[2]);
		g++;
	}
	b->gdone = g;
	if (g >= b->acked + PB_ACK_EVERY || (flush_acks && g != b->acked)) {
		// flow control: this rank has read steps < g of every board.  The value stored depends on the loads
		// above having returned, and a slot is reused only PB_SLOTS steps later: a plain store is enough
		for (int r = 0; r < pl.R; r++)
			st_relaxed_sys(&pl.peers[r]->ack[pl.me], g);
		b->acked = g;
	}
	__threadfence();
	atomicExch(&b->lock, 0u);
}

// one thread, once the batch's counts are final -- the resolve kernel after it has passed the commit chain on,
// or the last CTA of the emit kernel for a batch with waiting flags: store them into this rank's own board,
// then finalise whatever every rank has posted
__device__ __forceinline__ void peer_post(const PeerLink &pl, const BatchCtl *bc) {
	const unsigned long long s = pl.step;
	const unsigned long long t0 = gtimer();
	if (s >= (unsigned long long)PB_SLOTS) { // the slot's previous use (step s - PB_SLOTS) must have been read by every rank
		for (int r = 0; r < pl.R; r++) {
			while (ld_relaxed_u64(&pl.mine->ack[r]) + PB_SLOTS <= s) {
				if (gtimer() - t0 > pl.timeout_ns) {
					atomicAdd(&pl.mine->err, 1ull);
					break;
				}
			}
		}
	}
	st_word(&pl.mine->own[s % PB_SLOTS], ((s + 1) << 32) | (bc->lcounts[0] & 0xffffffffull),
	    (bc->lcounts[1] << 32) | (bc->lcounts[2] & 0xffffffffull));
	if (!(pl.dbg & 4u))
		peer_finalise(pl, false);
}

// one thread: wait (bounded) until this rank has finalised `steps` steps, then copy the totals out
__global__ void peer_counts_kernel(const PeerLink pl, unsigned long long steps, unsigned long long *out4) {
	if (threadIdx.x != 0)
		return;
	const unsigned long long t0 = gtimer();
	PeerBoard *b = pl.mine;
	while (true) {
		peer_finalise(pl, true);
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
		out4[1] = ld_word(&o->drops).v;
		out4[2] = ld_word(&o->admits).v;
		out4[3] = ld_word(&o->n).v;
	} else { // the latest totals
		out4[1] = b->tot[0];
		out4[2] = b->tot[1];
		out4[3] = b->tot[2];
	}
}

} // namespace cb
