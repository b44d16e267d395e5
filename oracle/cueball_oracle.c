/*
 * oracle/cueball_oracle.c -- plain-C restatement of node-cueball's claim-queue
 * hot path.  TEST INFRASTRUCTURE (see cueball_oracle.h for the rules and the
 * parity status).  Every function cites the reference lines it follows; the
 * code is written from the semantics, not transcribed.
 */
#include "cueball_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ clock */

/* lib/utils.js:198-204: secs*1000 + ns/1e6, three separately rounded ops */
double
cbo_current_millis(int64_t hr_secs, int64_t hr_ns)
{
	volatile double secs2ms = (double)hr_secs * 1000.0;
	volatile double ns2ms = (double)hr_ns / 1000000.0;
	return (secs2ms + ns2ms);
}

/* process.hrtime() returns [whole seconds, remaining nanoseconds] */
double
cbo_ns_to_ms(int64_t t)
{
	int64_t secs = t / 1000000000LL;
	int64_t ns = t - secs * 1000000000LL;
	return (cbo_current_millis(secs, ns));
}

/* ------------------------------------------------------------------ CoDel */

void
cbo_codel_init(cbo_codel *cd, double target) /* lib/codel.js:24-32 */
{
	cd->cd_targdelay = target;
	cd->cd_first_above_time = 0;
	cd->cd_drop_next = 0;
	cd->cd_count = 0;
	cd->cd_dropping = 0;
	cd->cd_has_last_empty = 0; /* never initialised by the ctor */
	cd->cd_last_empty = 0;
}

int
cbo_codel_can_drop(cbo_codel *cd, double now, double start) /* lib/codel.js:34-46 */
{
	double sojourn = now - start;

	if (sojourn < cd->cd_targdelay) {
		cd->cd_first_above_time = 0;
	} else if (cd->cd_first_above_time == 0) {
		cd->cd_first_above_time = now + CBO_CODEL_INTERVAL;
	} else if (now >= cd->cd_first_above_time) {
		return (1);
	}
	return (0);
}

double
cbo_codel_get_drop_next(const cbo_codel *cd, double now) /* lib/codel.js:48-50 */
{
	double step = CBO_CODEL_INTERVAL / sqrt(cd->cd_count);
	return (now + step);
}

/* lib/codel.js:56-86; the clock read (:58) is the `now` argument */
int
cbo_codel_overloaded(cbo_codel *cd, double now, double start)
{
	int ok = cbo_codel_can_drop(cd, now, start);
	int drop = 0;

	if (cd->cd_dropping) {
		if (!ok) {
			cd->cd_dropping = 0;
		} else if (now >= cd->cd_drop_next) {
			/* drop_next is NOT advanced here */
			drop = 1;
			cd->cd_count += 1;
		}
	} else if (ok && ((now - cd->cd_drop_next < CBO_CODEL_INTERVAL) ||
	    (now - cd->cd_first_above_time >= CBO_CODEL_INTERVAL))) {
		drop = 1;
		cd->cd_dropping = 1;
		if (now - cd->cd_drop_next < CBO_CODEL_INTERVAL)
			cd->cd_count = cd->cd_count > 2 ? cd->cd_count - 2 : 1;
		else
			cd->cd_count = 1;
		cd->cd_drop_next = cbo_codel_get_drop_next(cd, now);
	}
	return (drop);
}

void
cbo_codel_empty(cbo_codel *cd, double now) /* lib/codel.js:91-94 */
{
	cd->cd_has_last_empty = 1;
	cd->cd_last_empty = now;
	cd->cd_first_above_time = 0;
}

double
cbo_codel_get_max_idle(const cbo_codel *cd, double now) /* lib/codel.js:109-118 */
{
	double bound = cd->cd_targdelay * 10;

	/* `undefined < x` is false in JS */
	if (cd->cd_has_last_empty && cd->cd_last_empty < (now - bound))
		return (cd->cd_targdelay * 3);
	return (bound);
}

/* ------------------------------------------------------------------ Queue */

struct cbo_qnode { /* lib/queue.js:13-18 */
	void *n_value;
	struct cbo_qnode *n_next, *n_prev;
	struct cbo_queue *n_queue;
};
struct cbo_queue { /* lib/queue.js:39-45 */
	struct cbo_qnode q_head, q_tail;
	size_t q_len;
};

cbo_queue *
cbo_queue_new(void)
{
	cbo_queue *q = calloc(1, sizeof (*q));
	if (q == NULL)
		return (NULL);
	q->q_head.n_next = &q->q_tail;
	q->q_tail.n_prev = &q->q_head;
	q->q_head.n_queue = q->q_tail.n_queue = q;
	return (q);
}

void
cbo_queue_free(cbo_queue *q)
{
	if (q == NULL)
		return;
	while (!cbo_queue_is_empty(q)) {
		cbo_qnode *n = q->q_head.n_next;
		cbo_qnode_remove(n);
		free(n);
	}
	free(q);
}

int
cbo_queue_is_empty(const cbo_queue *q) /* lib/queue.js:46-48 */
{
	return (q->q_head.n_next == &q->q_tail);
}

size_t
cbo_queue_length(const cbo_queue *q) /* lib/queue.js:71-75 */
{
	return (q->q_len);
}

cbo_qnode *
cbo_queue_push(cbo_queue *q, void *v) /* lib/queue.js:52-56 + :19-27 */
{
	cbo_qnode *n = malloc(sizeof (*n));
	cbo_qnode *prev = q->q_tail.n_prev;
	if (n == NULL)
		return (NULL);
	n->n_value = v;
	n->n_queue = q;
	n->n_next = &q->q_tail;
	n->n_prev = prev;
	prev->n_next = n;
	q->q_tail.n_prev = n;
	q->q_len++;
	return (n);
}

void
cbo_qnode_remove(cbo_qnode *n) /* lib/queue.js:28-37 */
{
	cbo_qnode *next = n->n_next, *prev = n->n_prev;
	n->n_next = n->n_prev = NULL;
	prev->n_next = next;
	next->n_prev = prev;
	n->n_queue->q_len--;
}

void *
cbo_queue_shift(cbo_queue *q) /* lib/queue.js:57-62; callers check length first */
{
	cbo_qnode *n = q->q_head.n_next;
	void *v = n->n_value;
	cbo_qnode_remove(n);
	free(n);
	return (v);
}

/* ------------------------------------------------------------- drain loop */

typedef struct cbo_handle { /* the CueBallClaimHandle fields the loop touches */
	double ch_started; /* lib/connection-fsm.js:523 */
	int waiting;       /* isInState('waiting') */
	size_t idx;
} cbo_handle;

static inline void
cbo_emit(int waiting, int drop, int64_t nslots, uint64_t *cursor, uint8_t *decision,
    int32_t *slot, size_t i, cbo_batch_stats *st)
{
	uint8_t d;
	int32_t s = -1;

	if (!waiting) { /* lib/pool.js:738-740 (after overloaded() has run) */
		d = CBO_SKIPPED;
		st->skipped++;
	} else if (drop) { /* lib/pool.js:742-745 */
		d = CBO_DROP;
		st->drops++;
	} else { /* lib/pool.js:747; slot model SURVEY.md A.4 */
		d = CBO_ADMIT;
		if (nslots > 0)
			s = (int32_t)(*cursor % (uint64_t)nslots);
		(*cursor)++;
		st->admits++;
	}
	if (decision != NULL)
		decision[i] = d;
	if (slot != NULL)
		slot[i] = s;
}

/* faithful variant: handles on a lib/queue.js-style linked list */
int
cbo_drain_batch(cbo_codel *cd, const int64_t *start_ns, const int64_t *now_ns,
    const uint8_t *flags, size_t n, int64_t nslots, uint64_t cursor,
    uint8_t *decision, int32_t *slot, cbo_batch_stats *st)
{
	cbo_queue *waiters = cbo_queue_new();
	cbo_handle *hs = malloc((n ? n : 1) * sizeof (*hs));
	cbo_batch_stats local = { 0, 0, 0, 0 };
	size_t i;

	if (waiters == NULL || hs == NULL) {
		free(hs);
		cbo_queue_free(waiters);
		return (-1);
	}
	for (i = 0; i < n; i++) { /* claim(): lib/pool.js:914 + :960 */
		hs[i].ch_started = cbo_ns_to_ms(start_ns[i]);
		hs[i].waiting = flags ? (flags[i] & CBO_F_WAITING) != 0 : 1;
		hs[i].idx = i;
		cbo_queue_push(waiters, &hs[i]);
	}
	while (cbo_queue_length(waiters) > 0) { /* lib/pool.js:733 */
		cbo_handle *h = cbo_queue_shift(waiters); /* :734 */
		double now = cbo_ns_to_ms(now_ns[h->idx]); /* codel.js:58 */
		int drop = cbo_codel_overloaded(cd, now, h->ch_started); /* :735-736 */
		cbo_emit(h->waiting, drop, nslots, &cursor, decision, slot, h->idx, &local);
		if (flags && (flags[h->idx] & CBO_F_EMPTY_AFTER))
			cbo_codel_empty(cd, now); /* lib/pool.js:751-753 */
	}
	local.cursor = cursor;
	if (st != NULL)
		*st = local;
	free(hs);
	cbo_queue_free(waiters);
	return (0);
}

int
cbo_drain_batch_flat(cbo_codel *cd, const int64_t *start_ns, const int64_t *now_ns,
    const uint8_t *flags, size_t n, int64_t nslots, uint64_t cursor,
    uint8_t *decision, int32_t *slot, cbo_batch_stats *st)
{
	cbo_batch_stats local = { 0, 0, 0, 0 };
	size_t i;

	for (i = 0; i < n; i++) {
		double start = cbo_ns_to_ms(start_ns[i]);
		double now = cbo_ns_to_ms(now_ns[i]);
		int drop = cbo_codel_overloaded(cd, now, start);
		int waiting = flags ? (flags[i] & CBO_F_WAITING) != 0 : 1;
		cbo_emit(waiting, drop, nslots, &cursor, decision, slot, i, &local);
		if (flags && (flags[i] & CBO_F_EMPTY_AFTER))
			cbo_codel_empty(cd, now);
	}
	local.cursor = cursor;
	if (st != NULL)
		*st = local;
	return (0);
}

int
cbo_drain_batch_ms(cbo_codel *cd, const double *start_ms, const double *now_ms,
    const uint8_t *flags, size_t n, int64_t nslots, uint64_t cursor,
    uint8_t *decision, int32_t *slot, cbo_batch_stats *st)
{
	cbo_batch_stats local = { 0, 0, 0, 0 };
	size_t i;

	for (i = 0; i < n; i++) {
		int drop = cbo_codel_overloaded(cd, now_ms[i], start_ms[i]);
		int waiting = flags ? (flags[i] & CBO_F_WAITING) != 0 : 1;
		cbo_emit(waiting, drop, nslots, &cursor, decision, slot, i, &local);
		if (flags && (flags[i] & CBO_F_EMPTY_AFTER))
			cbo_codel_empty(cd, now_ms[i]);
	}
	local.cursor = cursor;
	if (st != NULL)
		*st = local;
	return (0);
}

/* ---------------------------------------------------------- planRebalance */

/*
 * lib/utils.js:239-393.  `keys.shift(); keys.push(k)` is a rotation, kept as a
 * ring cursor; wantedSpares[k] === undefined is want[k] == -1 while planning.
 */
int
cbo_plan_rebalance(const int32_t *have, const uint8_t *dead, int32_t B,
    int32_t target, int32_t max, int singleton, int32_t *want,
    int32_t *add, int32_t *n_add, int32_t *rm_key, int32_t *rm_ord, int32_t *n_rm)
{
	int32_t ring = 0, i, k, kk;
	int64_t done = 0, replacements = 0;
	int32_t na = 0, nr = 0;

	if (target < 0 || max < target) /* :248-249 */
		return (-1);
	for (k = 0; k < B; k++)
		want[k] = -1;

	if (B > 0) {
		for (i = 0; i < target; ++i) { /* :266-290 */
			k = ring;
			ring = (ring + 1) % B;
			if (want[k] < 0)
				want[k] = 0;
			if (!dead[k]) {
				if (singleton) {
					if (want[k] == 0) {
						want[k] = 1;
						++done;
					}
				} else {
					++want[k];
					++done;
				}
				continue;
			}
			if (want[k] == 0) {
				want[k] = 1;
				++done;
			}
			++replacements;
		}

		if (done + replacements > max) /* :293-294 */
			replacements = max - done;

		for (i = 0; i < replacements; ++i) { /* :300-366 */
			int64_t count;
			int32_t empties = 0;

			k = ring;
			ring = (ring + 1) % B;
			if (want[k] < 0)
				want[k] = 0;
			if (!dead[k]) {
				if (singleton) {
					if (want[k] == 0) {
						want[k] = 1;
						++done;
						continue;
					}
				} else {
					++want[k];
					++done;
					continue;
				}
			}
			count = done + replacements - i; /* :330 */
			for (kk = 0; kk < B; kk++) { /* :331-339 */
				if (singleton) {
					if (!dead[kk] && want[kk] < 0)
						empties++;
				} else {
					if (!dead[kk] || want[kk] < 0)
						empties++;
				}
			}
			if (count + 1 <= max) { /* :342-348 */
				if (want[k] == 0) {
					want[k] = 1;
					++done;
				}
				if (empties > 0)
					++replacements;
			} else if (count <= max && empties > 0) { /* :354-355 */
				++replacements;
			} else if (count <= max) { /* :358-362 */
				if (want[k] == 0) {
					want[k] = 1;
					++done;
				}
			} else { /* :365-367 */
				break;
			}
		}
	}

	for (k = 0; k < B; k++)
		if (want[k] < 0)
			want[k] = 0;

	/* :372-382 remove: reverse key order, oldest connections first */
	for (k = B - 1; k >= 0; k--) {
		int32_t h = have[k], ord = 0;
		while (h > want[k]) {
			if (rm_key != NULL) {
				rm_key[nr] = k;
				rm_ord[nr] = ord;
			}
			nr++;
			ord++;
			--h;
		}
	}
	/* :383-390 add: forward key order */
	for (k = 0; k < B; k++) {
		int32_t h = have[k];
		while (h < want[k]) {
			if (add != NULL)
				add[na] = k;
			na++;
			++h;
		}
	}
	if (n_add != NULL)
		*n_add = na;
	if (n_rm != NULL)
		*n_rm = nr;
	return (0);
}

/* ------------------------------------------------------- preference list */

void
cbo_pref_insert(int32_t *keys, int32_t *len, int32_t key, double u) /* lib/pool.js:287-288 */
{
	int32_t idx = (int32_t)floor(u * (double)(*len + 1));
	memmove(&keys[idx + 1], &keys[idx], (size_t)(*len - idx) * sizeof (int32_t));
	keys[idx] = key;
	(*len)++;
}

int
cbo_pref_remove(int32_t *keys, int32_t *len, int32_t key) /* lib/pool.js:294-296 */
{
	int32_t i;
	for (i = 0; i < *len; i++) {
		if (keys[i] == key) {
			memmove(&keys[i], &keys[i + 1], (size_t)(*len - i - 1) * sizeof (int32_t));
			(*len)--;
			return (0);
		}
	}
	return (-1);
}

void
cbo_pref_reshuffle(int32_t *keys, int32_t len, double u) /* lib/pool.js:501-519 */
{
	int32_t taken, n;
	if (len <= 1)
		return;
	taken = keys[len - 1];
	n = len - 1;
	cbo_pref_insert(keys, &n, taken, u);
}

int
cbo_build_slots(int32_t B, int32_t S, const double *us, int32_t *keys_out,
    int32_t *slot_backend_out)
{
	int32_t len = 0, k, na = 0, nr = 0, rc;
	int32_t *have = calloc((size_t)(B > 0 ? B : 1), sizeof (int32_t));
	uint8_t *dead = calloc((size_t)(B > 0 ? B : 1), 1);
	int32_t *want = calloc((size_t)(B > 0 ? B : 1), sizeof (int32_t));
	int32_t *add = calloc((size_t)(S > 0 ? S : 1), sizeof (int32_t));

	if (!have || !dead || !want || !add) {
		free(have); free(dead); free(want); free(add);
		return (-1);
	}
	for (k = 0; k < B; k++) /* resolver 'added' in key order */
		cbo_pref_insert(keys_out, &len, k, us[k]);
	/* conns built in p_keys order (lib/pool.js:562-565): positions, not ids */
	rc = cbo_plan_rebalance(have, dead, B, S, S, 0, want, add, &na, NULL, NULL, &nr);
	for (k = 0; k < na; k++)
		slot_backend_out[k] = keys_out[add[k]];
	free(have); free(dead); free(want); free(add);
	return (rc == 0 ? na : -1);
}

void
cbo_shuffle(int32_t *arr, int32_t n, const double *us) /* lib/utils.js:207-217 */
{
	int32_t i = n, d = 0;
	while (i > 0) {
		int32_t j = (int32_t)floor(us[d++] * (double)i);
		int32_t tmp;
		--i;
		tmp = arr[i];
		arr[i] = arr[j];
		arr[j] = tmp;
	}
}

/* ------------------------------------------------------------------------------------------
 * Event-driven pool model: ConnectionPool#claim + the slot-idle handler + the handle's hard
 * timeout, over the linked-list Queue above (lib/queue.js-faithful).
 */
typedef struct pool_handle {
	double ch_started;   /* lib/connection-fsm.js:523 */
	int64_t deadline_ns; /* started + claimTimeout */
	int64_t ticket;
	int state; /* 0 waiting, 1 claimed, 2 failed, 3 cancelled */
} pool_handle;

int cbo_pool_sim(cbo_codel *cd, int64_t nslots, uint64_t idle0, uint64_t cursor0, const uint8_t *ev_type,
    const int64_t *ev_a, const int64_t *ev_b, size_t n_events, int64_t *res, double *timeouts_ms,
    cbo_pool_result *out) {
	size_t nclaims = 0, e;
	for (e = 0; e < n_events; e++)
		if (ev_type[e] == 0)
			nclaims++;
	pool_handle *handles = (pool_handle *)calloc(nclaims ? nclaims : 1, sizeof(pool_handle));
	cbo_queue *waiters = cbo_queue_new();
	if (!handles || !waiters)
		return -1;
	uint64_t idle = idle0, cursor = cursor0, nres = 0;
	size_t nh = 0, first_live = 0;
	cbo_pool_counters c = {0, 0, 0, 0};
#define ADMIT(h)                                                                     \
	do {                                                                         \
		(h)->state = 1;                                                      \
		res[3 * nres] = (h)->ticket;                                         \
		res[3 * nres + 1] = 0;                                               \
		res[3 * nres + 2] = nslots > 0 ? (int64_t)(cursor % (uint64_t)nslots) : -1; \
		nres++;                                                              \
		cursor++;                                                            \
	} while (0)
	for (e = 0; e < n_events; e++) {
		const int ty = ev_type[e];
		const int64_t a = ev_a[e], b = ev_b[e];
		if (ty == 0 || ty == 1) {
			/* hard timeouts that have come due, lib/connection-fsm.js:626-642 */
			size_t t;
			for (t = first_live; t < nh; t++) {
				pool_handle *h = &handles[t];
				if (h->state != 0) {
					if (t == first_live)
						first_live++;
					continue;
				}
				if (a >= h->deadline_ns) {
					h->state = 2;
					c.claim_timeout++; /* connection-fsm.js:629-630 */
					res[3 * nres] = h->ticket;
					res[3 * nres + 1] = 3;
					res[3 * nres + 2] = -1;
					nres++;
				}
			}
		}
		if (ty == 0) {
			const double now = cbo_ns_to_ms(a);
			const double timeout = cbo_codel_get_max_idle(cd, now); /* pool.js:880 */
			c.claim++;                                              /* pool.js:887 */
			pool_handle *h = &handles[nh];
			h->ticket = (int64_t)nh;
			h->ch_started = now; /* connection-fsm.js:523 */
			h->state = 0;
			h->deadline_ns = a + (int64_t)(timeout * 1000000.0);
			timeouts_ms[nh] = timeout;
			nh++;
			if (idle > 0) { /* pool.js:934-951 */
				idle--;
				ADMIT(h);
			} else {
				cbo_queue_push(waiters, h); /* pool.js:960 */
				if (cbo_queue_length(waiters) > c.max_claim_queue) /* pool.js:962 */
					c.max_claim_queue = cbo_queue_length(waiters);
				c.queued_claim++; /* pool.js:963 */
			}
		} else if (ty == 1) {
			const double now = cbo_ns_to_ms(a);
			int64_t s;
			for (s = 0; s < b; s++) { /* one handler run per slot */
				int handed = 0;
				while (cbo_queue_length(waiters) > 0) { /* pool.js:733 */
					pool_handle *hdl = (pool_handle *)cbo_queue_shift(waiters);
					const int drop = cbo_codel_overloaded(cd, now, hdl->ch_started); /* pool.js:735-736 */
					if (hdl->state != 0) { /* pool.js:738-740 */
						res[3 * nres] = hdl->ticket;
						res[3 * nres + 1] = 2;
						res[3 * nres + 2] = -1;
						nres++;
						continue;
					}
					if (drop) { /* pool.js:742-745 */
						hdl->state = 2;
						c.claim_timeout++; /* connection-fsm.js:633-638 */
						res[3 * nres] = hdl->ticket;
						res[3 * nres + 1] = 1;
						res[3 * nres + 2] = -1;
						nres++;
						continue;
					}
					ADMIT(hdl); /* pool.js:747-748 */
					handed = 1;
					break;
				}
				if (!handed) {
					cbo_codel_empty(cd, now); /* pool.js:751-753 */
					idle++;                   /* pool.js:756 */
				}
			}
		} else {
			if (a >= 0 && (size_t)a < nh && handles[a].state == 0)
				handles[a].state = 3; /* connection-fsm.js:580-587 */
		}
	}
#undef ADMIT
	out->n_res = nres;
	out->idle = idle;
	out->cursor = cursor;
	out->queued = cbo_queue_length(waiters);
	out->counters = c;
	while (!cbo_queue_is_empty(waiters))
		cbo_queue_shift(waiters);
	cbo_queue_free(waiters);
	free(handles);
	return 0;
}

/* ------------------------------------------------------------------ low-pass filter, lib/pool.js:44-100 */
void
cbo_lpf_taps(double *taps)
{
	/* genTaps(128, -0.2), lib/pool.js:64-76: taps[i] = exp(tc * i), then each divided by their sum */
	const double tc = -0.2;
	double sum = 0.0;
	int i;
	for (i = 0; i < CBO_LPF_TAPS; ++i) {
		taps[i] = exp(tc * i);
		sum += taps[i];
	}
	for (i = 0; i < CBO_LPF_TAPS; ++i)
		taps[i] /= sum;
}

size_t
cbo_lpf_run(int64_t p_spares, int64_t max, const uint8_t *ty, const int64_t *a, const int64_t *b,
    const int64_t *c, const int64_t *d, size_t n_events, double *out)
{
	double taps[CBO_LPF_TAPS], buf[CBO_LPF_TAPS];
	int ptr = 0;		/* f_ptr, lib/pool.js:80 */
	size_t e, k = 0;
	int i, j;
	cbo_lpf_taps(taps);
	memset(buf, 0, sizeof (buf));
	for (e = 0; e < n_events; e++) {
		if (ty[e] == 0) {
			/* the sampler, lib/pool.js:251-258, into FIRFilter.put, :82-87 */
			const int64_t busy0 = a[e] - b[e];
			buf[ptr++] = (double)(busy0 + p_spares);
			if (ptr == CBO_LPF_TAPS)
				ptr = 0;
			continue;
		}
		/* FIRFilter.get, lib/pool.js:88-100: newest sample times taps[0], walking back round the ring */
		double acc = 0.0;
		i = ptr - 1;
		if (i < 0)
			i += CBO_LPF_TAPS;
		for (j = 0; j < CBO_LPF_TAPS; ++j) {
			acc += buf[i] * taps[j];
			if (--i < 0)
				i += CBO_LPF_TAPS;
		}
		/* _rebalance, lib/pool.js:563-586 */
		int64_t spares = b[e] + c[e] - d[e];
		if (spares < 0)
			spares = 0;
		int64_t busy = a[e] - spares;
		if (busy < 0)
			busy = 0;
		int64_t extras = d[e] - c[e];
		if (extras < 0)
			extras = 0;
		double target = (double)(busy + extras + p_spares);
		const double min = ceil(acc);
		double clamped = 0.0;
		if (target < min * 1.05) {
			target = min;
			clamped = 1.0;
		}
		if (target > (double)max)
			target = (double)max;
		out[3 * k] = target;
		out[3 * k + 1] = clamped;
		out[3 * k + 2] = acc;
		k++;
	}
	return (k);
}
