/*
 * vb_broker.c -- see vb_broker.h.  One process-shared mutex, three conditions:
 *   work  : "a request is queued"        (requesters -> broker)
 *   space : "a slot was released"        (requesters -> requesters waiting for a slot)
 *   done  : "a batch has been answered"  (broker -> the requesters of that batch; each checks its own slot)
 * A slot goes FREE -> QUEUED (requester: payload copied in, slot number appended to the queue) -> RUNNING (broker: taken
 * into the batch) -> DONE (broker: results and status in the slot) -> FREE (requester: results copied out).  A requester
 * with a timeout withdraws a request that is still QUEUED, and ABANDONs one that is RUNNING (the broker frees it).
 * The broker never holds the mutex across the library call, so requests keep queueing while a batch runs -- they are
 * the next batch, which is what makes the batches grow with the load (the window only matters at low load, where it
 * trades latency for sharing).
 */
#include <errno.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "vb_broker.h"

enum { SLOT_FREE = 0, SLOT_QUEUED, SLOT_RUNNING, SLOT_DONE, SLOT_ABANDONED };

struct VbBrokerShared
{
	pthread_mutex_t mu;
	pthread_cond_t work, space, done;
	VbBrokerConfig cfg;
	VbBrokerStats stats;
	int			queued;			/* entries of queue[] */
	int			stopping;
	/* offsets from the start of the block (the block may sit at different addresses in different processes) */
	size_t		off_state, off_rc, off_queue, off_query, off_ids, off_dist;
};

struct VbBroker
{
	VbBrokerShared *sh;
	vb_ivf	   *ix;
	pthread_t	thread;
	int			own_block;
	/* the serving thread's staging (a server keeps these page-locked: the batch is DMA'd from here) */
	int		   *batch;
	char	   *q_stage;
	int64_t    *id_stage;
	double	   *d_stage;
};

static size_t
align64(size_t x)
{
	return (x + 63) & ~(size_t) 63;
}

#define SH_PTR(sh, type, off) ((type *) ((char *) (sh) + (sh)->off))

size_t
VbBrokerSharedSize(const VbBrokerConfig *cfg)
{
	size_t		n = (size_t) cfg->max_batch;

	return align64(sizeof(VbBrokerShared)) + align64(sizeof(int) * n) * 3 + align64(cfg->query_bytes * n) +
		align64(sizeof(int64_t) * n * (size_t) cfg->k) + align64(sizeof(double) * n * (size_t) cfg->k);
}

VbBrokerShared *
VbBrokerSharedInit(void *block, const VbBrokerConfig *cfg)
{
	if (!block || !cfg || cfg->max_batch < 1 || cfg->k < 1 || cfg->probes < 1 || cfg->query_bytes == 0)
		return NULL;
	VbBrokerShared *sh = block;
	size_t		n = (size_t) cfg->max_batch, off = align64(sizeof(VbBrokerShared));
	pthread_mutexattr_t ma;
	pthread_condattr_t ca;

	memset(sh, 0, sizeof(*sh));
	sh->cfg = *cfg;
	sh->off_state = off;
	off += align64(sizeof(int) * n);
	sh->off_rc = off;
	off += align64(sizeof(int) * n);
	sh->off_queue = off;
	off += align64(sizeof(int) * n);
	sh->off_query = off;
	off += align64(cfg->query_bytes * n);
	sh->off_ids = off;
	off += align64(sizeof(int64_t) * n * (size_t) cfg->k);
	sh->off_dist = off;
	memset(SH_PTR(sh, int, off_state), 0, sizeof(int) * n);
	pthread_mutexattr_init(&ma);
	pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
	pthread_mutex_init(&sh->mu, &ma);
	pthread_mutexattr_destroy(&ma);
	pthread_condattr_init(&ca);
	pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
	pthread_cond_init(&sh->work, &ca);
	pthread_cond_init(&sh->space, &ca);
	pthread_cond_init(&sh->done, &ca);
	pthread_condattr_destroy(&ca);
	return sh;
}

static void
deadline_after(struct timespec *ts, long usec)
{
	clock_gettime(CLOCK_REALTIME, ts);
	ts->tv_nsec += (usec % 1000000L) * 1000L;
	ts->tv_sec += usec / 1000000L + ts->tv_nsec / 1000000000L;
	ts->tv_nsec %= 1000000000L;
}

static void *
broker_main(void *arg)
{
	VbBroker   *b = arg;
	VbBrokerShared *sh = b->sh;
	const int	k = sh->cfg.k;
	const size_t qb = sh->cfg.query_bytes;
	int		   *state = SH_PTR(sh, int, off_state), *rcs = SH_PTR(sh, int, off_rc), *queue = SH_PTR(sh, int, off_queue);
	char	   *queries = SH_PTR(sh, char, off_query);
	int64_t    *ids = SH_PTR(sh, int64_t, off_ids);
	double	   *dist = SH_PTR(sh, double, off_dist);

	pthread_mutex_lock(&sh->mu);
	for (;;)
	{
		while (sh->queued == 0 && !sh->stopping)
			pthread_cond_wait(&sh->work, &sh->mu);
		if (sh->queued == 0 && sh->stopping)
			break;
		/* the first request of a batch waits for company, at most window_us, unless every slot is queued already */
		if (sh->queued < sh->cfg.max_batch && sh->cfg.window_us > 0 && !sh->stopping)
		{
			struct timespec until;

			deadline_after(&until, sh->cfg.window_us);
			while (sh->queued < sh->cfg.max_batch && !sh->stopping)
				if (pthread_cond_timedwait(&sh->work, &sh->mu, &until) == ETIMEDOUT)
					break;
		}
		const int	n = sh->queued;

		for (int i = 0; i < n; i++)
		{
			b->batch[i] = queue[i];
			state[queue[i]] = SLOT_RUNNING;
		}
		sh->queued = 0;
		pthread_mutex_unlock(&sh->mu);

		/* one call for the whole batch: every probed list is read once for all the scans that probe it */
		for (int i = 0; i < n; i++)
			memcpy(b->q_stage + qb * (size_t) i, queries + qb * (size_t) b->batch[i], qb);
		int			rc = vb_ivf_search(b->ix, b->q_stage, n, sh->cfg.probes, k, b->id_stage, b->d_stage);

		if (rc == VB_OK)
			for (int i = 0; i < n; i++)
			{
				memcpy(ids + (size_t) b->batch[i] * k, b->id_stage + (size_t) i * k, sizeof(int64_t) * (size_t) k);
				memcpy(dist + (size_t) b->batch[i] * k, b->d_stage + (size_t) i * k, sizeof(double) * (size_t) k);
			}
		pthread_mutex_lock(&sh->mu);
		for (int i = 0; i < n; i++)
		{
			rcs[b->batch[i]] = rc;
			if (state[b->batch[i]] == SLOT_ABANDONED)
			{
				/* its requester gave up waiting (timeout): nobody will collect the answer, the slot is free again */
				state[b->batch[i]] = SLOT_FREE;
				pthread_cond_signal(&sh->space);
			}
			else
				state[b->batch[i]] = SLOT_DONE;
		}
		sh->stats.requests += n;
		sh->stats.batches += 1;
		if (n > sh->stats.largest)
			sh->stats.largest = n;
		if (rc != VB_OK)
			sh->stats.failed += n;
		pthread_cond_broadcast(&sh->done);
	}
	pthread_mutex_unlock(&sh->mu);
	return NULL;
}

VbBroker *
VbBrokerServe(VbBrokerShared *sh, vb_ivf *ix)
{
	if (!sh || !ix)
		return NULL;
	VbBroker   *b = calloc(1, sizeof(VbBroker));

	if (!b)
		return NULL;
	const size_t n = (size_t) sh->cfg.max_batch;

	b->sh = sh;
	b->ix = ix;
	b->batch = calloc(n, sizeof(int));
	b->q_stage = malloc(sh->cfg.query_bytes * n);
	b->id_stage = malloc(sizeof(int64_t) * n * (size_t) sh->cfg.k);
	b->d_stage = malloc(sizeof(double) * n * (size_t) sh->cfg.k);
	if (!b->batch || !b->q_stage || !b->id_stage || !b->d_stage || pthread_create(&b->thread, NULL, broker_main, b) != 0)
	{
		free(b->batch);
		free(b->q_stage);
		free(b->id_stage);
		free(b->d_stage);
		free(b);
		return NULL;
	}
	return b;
}

VbBroker *
VbBrokerStart(vb_ivf *ix, const VbBrokerConfig *cfg)
{
	if (!ix || !cfg || cfg->max_batch < 1 || cfg->k < 1 || cfg->probes < 1 || cfg->query_bytes == 0)
		return NULL;
	void	   *block = malloc(VbBrokerSharedSize(cfg));
	VbBrokerShared *sh = VbBrokerSharedInit(block, cfg);
	VbBroker   *b = sh ? VbBrokerServe(sh, ix) : NULL;

	if (!b)
	{
		free(block);
		return NULL;
	}
	b->own_block = 1;
	return b;
}

VbBrokerShared *
VbBrokerBlock(VbBroker *b)
{
	return b->sh;
}

int
VbBrokerRequest(VbBrokerShared *sh, const void *query, int64_t *out_ids, double *out_dist, int timeout_ms)
{
	int		   *state = SH_PTR(sh, int, off_state), *rcs = SH_PTR(sh, int, off_rc), *queue = SH_PTR(sh, int, off_queue);
	const int	k = sh->cfg.k, n = sh->cfg.max_batch;
	struct timespec until;
	int			slot = -1, rc;

	if (timeout_ms > 0)
		deadline_after(&until, (long) timeout_ms * 1000L);
	pthread_mutex_lock(&sh->mu);
	for (;;)
	{
		if (sh->stopping)
		{
			pthread_mutex_unlock(&sh->mu);
			return VB_ESTATE;
		}
		for (int i = 0; i < n; i++)
			if (state[i] == SLOT_FREE)
			{
				slot = i;
				break;
			}
		if (slot >= 0)
			break;
		if (timeout_ms > 0)
		{
			if (pthread_cond_timedwait(&sh->space, &sh->mu, &until) == ETIMEDOUT)
			{
				pthread_mutex_unlock(&sh->mu);
				return VB_ECUDA;
			}
		}
		else
			pthread_cond_wait(&sh->space, &sh->mu);
	}
	memcpy(SH_PTR(sh, char, off_query) + sh->cfg.query_bytes * (size_t) slot, query, sh->cfg.query_bytes);
	state[slot] = SLOT_QUEUED;
	queue[sh->queued++] = slot;
	pthread_cond_signal(&sh->work);
	while (state[slot] != SLOT_DONE)
	{
		if (timeout_ms > 0)
		{
			if (pthread_cond_timedwait(&sh->done, &sh->mu, &until) == ETIMEDOUT && state[slot] != SLOT_DONE)
			{
				/* no answer: a request still queued is withdrawn; one already taken is abandoned (the broker frees the slot) */
				if (state[slot] == SLOT_QUEUED)
				{
					for (int i = 0; i < sh->queued; i++)
						if (queue[i] == slot)
						{
							memmove(queue + i, queue + i + 1, sizeof(int) * (size_t) (sh->queued - i - 1));
							sh->queued--;
							break;
						}
					state[slot] = SLOT_FREE;
					pthread_cond_signal(&sh->space);
				}
				else
					state[slot] = SLOT_ABANDONED;
				pthread_mutex_unlock(&sh->mu);
				return VB_ECUDA;
			}
		}
		else
			pthread_cond_wait(&sh->done, &sh->mu);
	}
	rc = rcs[slot];
	if (rc == VB_OK)
	{
		memcpy(out_ids, SH_PTR(sh, int64_t, off_ids) + (size_t) slot * k, sizeof(int64_t) * (size_t) k);
		memcpy(out_dist, SH_PTR(sh, double, off_dist) + (size_t) slot * k, sizeof(double) * (size_t) k);
	}
	state[slot] = SLOT_FREE;
	pthread_cond_signal(&sh->space);
	pthread_mutex_unlock(&sh->mu);
	return rc;
}

int
VbBrokerSearch(VbBroker *b, const void *query, int64_t *ids, double *dist)
{
	return VbBrokerRequest(b->sh, query, ids, dist, 0);
}

void
VbBrokerGetStats(VbBrokerShared *sh, VbBrokerStats *out)
{
	pthread_mutex_lock(&sh->mu);
	*out = sh->stats;
	pthread_mutex_unlock(&sh->mu);
}

void
VbBrokerStop(VbBroker *b)
{
	VbBrokerShared *sh = b->sh;

	pthread_mutex_lock(&sh->mu);
	sh->stopping = 1;
	pthread_cond_broadcast(&sh->work);
	pthread_cond_broadcast(&sh->space);
	pthread_mutex_unlock(&sh->mu);
	pthread_join(b->thread, NULL);
	free(b->batch);
	free(b->q_stage);
	free(b->id_stage);
	free(b->d_stage);
	if (b->own_block)
	{
		pthread_mutex_destroy(&sh->mu);
		pthread_cond_destroy(&sh->work);
		pthread_cond_destroy(&sh->space);
		pthread_cond_destroy(&sh->done);
		free(sh);
	}
	free(b);
}
