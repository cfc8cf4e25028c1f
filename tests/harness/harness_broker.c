/*
 * harness_broker.c -- drives pgvector_b200/ext/vb_broker.c the way a set of backends would: `threads` requesters, each
 * issuing its share of the queries one scan at a time through VbBrokerSearch.  TEST INFRASTRUCTURE.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#include "vb_broker.h"

typedef struct
{
	VbBroker   *b;
	const char *queries;
	size_t		qb;
	int			nq, first, stride, k;
	int64_t    *ids;
	double	   *dist;
	int			rc;
} Requester;

static void *
requester_main(void *arg)
{
	Requester  *r = arg;

	for (int q = r->first; q < r->nq; q += r->stride)
	{
		int			rc = VbBrokerSearch(r->b, r->queries + r->qb * (size_t) q, r->ids + (size_t) q * r->k, r->dist + (size_t) q * r->k);

		if (rc != 0)
			r->rc = rc;
	}
	return NULL;
}

/* returns 0, or the first error a requester saw; stats4 = {requests, batches, largest, failed} */
int
hb_broker_run(vb_ivf *ix, const void *queries, int nq, size_t query_bytes, int threads, int probes, int k, int max_batch,
			  int window_us, int64_t *out_ids, double *out_dist, int64_t *stats4)
{
	VbBrokerConfig cfg = {max_batch, window_us, probes, k, query_bytes};
	VbBroker   *b = VbBrokerStart(ix, &cfg);

	if (!b)
		return -100;
	pthread_t  *th = calloc((size_t) threads, sizeof(pthread_t));
	Requester  *rq = calloc((size_t) threads, sizeof(Requester));
	int			rc = 0;

	for (int t = 0; t < threads; t++)
	{
		rq[t] = (Requester) {b, queries, query_bytes, nq, t, threads, k, out_ids, out_dist, 0};
		pthread_create(&th[t], NULL, requester_main, &rq[t]);
	}
	for (int t = 0; t < threads; t++)
	{
		pthread_join(th[t], NULL);
		if (rq[t].rc != 0 && rc == 0)
			rc = rq[t].rc;
	}
	VbBrokerStats st;

	VbBrokerGetStats(b, &st);
	stats4[0] = st.requests;
	stats4[1] = st.batches;
	stats4[2] = st.largest;
	stats4[3] = st.failed;
	VbBrokerStop(b);
	free(th);
	free(rq);
	return rc;
}
