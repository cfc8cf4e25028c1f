/*
 * vb_broker.c -- see vb_broker.h.  One mutex, three conditions:
 *   work  : "a request is queued"        (requesters -> broker)
 *   space : "the queue was drained"      (broker -> requesters waiting for a slot)
 *   done  : "a batch has been answered"  (broker -> the requesters of that batch; each checks its own flag)
 * The broker never holds the mutex across the library call, so requests keep queueing while a batch runs -- they are
 * the next batch, which is what makes the batches grow with the load (no window is needed once the GPU is the
 * bottleneck; the window only matters at low load, where it trades latency for sharing).
 */
#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "vb_broker.h"

typedef struct VbBrokerRequest
{
	const void *query;
	int64_t    *ids;
	double	   *dist;
	int			rc;
	int			done;
} VbBrokerRequest;

struct VbBroker
{
	vb_ivf	   *ix;
	VbBrokerConfig cfg;
	pthread_mutex_t mu;
	pthread_cond_t work, space, done;
	pthread_t	thread;
	VbBrokerRequest **queue;	/* [max_batch] pointers to the requesters' own records */
	int			queued;
	int			stopping;
	VbBrokerStats stats;
	/* the broker thread's staging (a server keeps these page-locked: the batch is DMA'd from here) */
	VbBrokerRequest **batch;
	char	   *q_stage;
	int64_t    *id_stage;
	double	   *d_stage;
};

static void
deadline_after(struct timespec *ts, int usec)
{
	clock_gettime(CLOCK_REALTIME, ts);
	ts->tv_nsec += (long) (usec % 1000000) * 1000L;
	ts->tv_sec += usec / 1000000 + ts->tv_nsec / 1000000000L;
	ts->tv_nsec %= 1000000000L;
}

static void *
broker_main(void *arg)
{
	VbBroker   *b = arg;
	const int	k = b->cfg.k;
	const size_t qb = b->cfg.query_bytes;

	pthread_mutex_lock(&b->mu);
	for (;;)
	{
		while (b->queued == 0 && !b->stopping)
			pthread_cond_wait(&b->work, &b->mu);
		if (b->queued == 0 && b->stopping)
			break;
		/* the first request of a batch waits for company, at most window_us, unless the batch is full already */
		if (b->queued < b->cfg.max_batch && b->cfg.window_us > 0 && !b->stopping)
		{
			struct timespec until;

			deadline_after(&until, b->cfg.window_us);
			while (b->queued < b->cfg.max_batch && !b->stopping)
				if (pthread_cond_timedwait(&b->work, &b->mu, &until) == ETIMEDOUT)
					break;
		}
		const int	n = b->queued;

		memcpy(b->batch, b->queue, sizeof(VbBrokerRequest *) * (size_t) n);
		b->queued = 0;
		pthread_cond_broadcast(&b->space);
		pthread_mutex_unlock(&b->mu);

		/* one call for the whole batch: every probed list is read once for all the scans that probe it */
		for (int i = 0; i < n; i++)
			memcpy(b->q_stage + qb * (size_t) i, b->batch[i]->query, qb);
		int			rc = vb_ivf_search(b->ix, b->q_stage, n, b->cfg.probes, k, b->id_stage, b->d_stage);

		for (int i = 0; i < n; i++)
		{
			if (rc == VB_OK)
			{
				memcpy(b->batch[i]->ids, b->id_stage + (size_t) i * k, sizeof(int64_t) * (size_t) k);
				memcpy(b->batch[i]->dist, b->d_stage + (size_t) i * k, sizeof(double) * (size_t) k);
			}
		}
		pthread_mutex_lock(&b->mu);
		for (int i = 0; i < n; i++)
		{
			b->batch[i]->rc = rc;
			b->batch[i]->done = 1;
		}
		b->stats.requests += n;
		b->stats.batches += 1;
		if (n > b->stats.largest)
			b->stats.largest = n;
		if (rc != VB_OK)
			b->stats.failed += n;
		pthread_cond_broadcast(&b->done);
	}
	pthread_mutex_unlock(&b->mu);
	return NULL;
}

VbBroker *
VbBrokerStart(vb_ivf *ix, const VbBrokerConfig *cfg)
{
	if (!ix || !cfg || cfg->max_batch < 1 || cfg->k < 1 || cfg->probes < 1 || cfg->query_bytes == 0)
		return NULL;
	VbBroker   *b = calloc(1, sizeof(VbBroker));

	if (!b)
		return NULL;
	b->ix = ix;
	b->cfg = *cfg;
	b->queue = calloc((size_t) cfg->max_batch, sizeof(VbBrokerRequest *));
	b->batch = calloc((size_t) cfg->max_batch, sizeof(VbBrokerRequest *));
	b->q_stage = malloc(cfg->query_bytes * (size_t) cfg->max_batch);
	b->id_stage = malloc(sizeof(int64_t) * (size_t) cfg->max_batch * (size_t) cfg->k);
	b->d_stage = malloc(sizeof(double) * (size_t) cfg->max_batch * (size_t) cfg->k);
	pthread_mutex_init(&b->mu, NULL);
	pthread_cond_init(&b->work, NULL);
	pthread_cond_init(&b->space, NULL);
	pthread_cond_init(&b->done, NULL);
	if (!b->queue || !b->batch || !b->q_stage || !b->id_stage || !b->d_stage ||
		pthread_create(&b->thread, NULL, broker_main, b) != 0)
	{
		free(b->queue);
		free(b->batch);
		free(b->q_stage);
		free(b->id_stage);
		free(b->d_stage);
		free(b);
		return NULL;
	}
	return b;
}

int
VbBrokerSearch(VbBroker *b, const void *query, int64_t *ids, double *dist)
{
	VbBrokerRequest r = {query, ids, dist, VB_OK, 0};

	pthread_mutex_lock(&b->mu);
	while (b->queued == b->cfg.max_batch && !b->stopping)
		pthread_cond_wait(&b->space, &b->mu);
	if (b->stopping)
	{
		pthread_mutex_unlock(&b->mu);
		return VB_ESTATE;
	}
	b->queue[b->queued++] = &r;
	pthread_cond_signal(&b->work);
	while (!r.done)
		pthread_cond_wait(&b->done, &b->mu);
	pthread_mutex_unlock(&b->mu);
	return r.rc;
}

void
VbBrokerGetStats(VbBroker *b, VbBrokerStats *out)
{
	pthread_mutex_lock(&b->mu);
	*out = b->stats;
	pthread_mutex_unlock(&b->mu);
}

void
VbBrokerStop(VbBroker *b)
{
	pthread_mutex_lock(&b->mu);
	b->stopping = 1;
	pthread_cond_broadcast(&b->work);
	pthread_cond_broadcast(&b->space);
	pthread_mutex_unlock(&b->mu);
	pthread_join(b->thread, NULL);
	pthread_mutex_destroy(&b->mu);
	pthread_cond_destroy(&b->work);
	pthread_cond_destroy(&b->space);
	pthread_cond_destroy(&b->done);
	free(b->queue);
	free(b->batch);
	free(b->q_stage);
	free(b->id_stage);
	free(b->d_stage);
	free(b);
}
