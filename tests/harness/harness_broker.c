/*
 * harness_broker.c -- drives pgvector_b200/ext/vb_broker.c the way a set of backends would: `threads` requesters, each
 * issuing its share of the queries one scan at a time through VbBrokerSearch.  TEST INFRASTRUCTURE.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

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

	VbBrokerGetStats(VbBrokerBlock(b), &st);
	stats4[0] = st.requests;
	stats4[1] = st.batches;
	stats4[2] = st.largest;
	stats4[3] = st.failed;
	VbBrokerStop(b);
	free(th);
	free(rq);
	return rc;
}

/*
 * The same with PROCESSES as requesters (what backends are): the request block and the result area live in anonymous
 * shared memory, `procs` children are forked BEFORE the serving thread starts (a child is single-threaded and never
 * touches the library or the heap: VbBrokerRequest works on the block and its own stack), each issues its share of the
 * scans with a 30 s timeout and exits at the first error.  Returns 0, -100 / -101 for setup failures, or the number of
 * children that failed.
 */
int
hb_broker_run_fork(vb_ivf *ix, const void *queries, int nq, size_t query_bytes, int procs, int probes, int k, int max_batch,
				   int window_us, int64_t *out_ids, double *out_dist, int64_t *stats4)
{
	VbBrokerConfig cfg = {max_batch, window_us, probes, k, query_bytes};
	const size_t blk = VbBrokerSharedSize(&cfg);
	const size_t res = (sizeof(int64_t) + sizeof(double)) * (size_t) nq * (size_t) k;
	char	   *map = mmap(NULL, blk + res + 64, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);

	if (map == MAP_FAILED)
		return -100;
	VbBrokerShared *sh = VbBrokerSharedInit(map, &cfg);
	int64_t    *r_ids = (int64_t *) (map + ((blk + 63) & ~(size_t) 63));
	double	   *r_dist = (double *) (r_ids + (size_t) nq * k);

	if (!sh)
	{
		munmap(map, blk + res + 64);
		return -101;
	}
	pid_t	   *pids = calloc((size_t) procs, sizeof(pid_t));
	int			failed = 0;

	for (int p = 0; p < procs; p++)
	{
		pids[p] = fork();
		if (pids[p] == 0)
		{
			for (int q = p; q < nq; q += procs)
				if (VbBrokerRequest(sh, (const char *) queries + query_bytes * (size_t) q, r_ids + (size_t) q * k,
									r_dist + (size_t) q * k, 30000) != 0)
					_exit(1);
			_exit(0);
		}
		if (pids[p] < 0)
			failed++;
	}
	VbBroker   *b = VbBrokerServe(sh, ix);

	for (int p = 0; p < procs; p++)
	{
		int			status = 0;

		if (pids[p] > 0 && (waitpid(pids[p], &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0))
			failed++;
	}
	VbBrokerStats st;

	VbBrokerGetStats(sh, &st);
	stats4[0] = st.requests;
	stats4[1] = st.batches;
	stats4[2] = st.largest;
	stats4[3] = st.failed;
	if (b)
		VbBrokerStop(b);
	else
		failed++;
	memcpy(out_ids, r_ids, sizeof(int64_t) * (size_t) nq * (size_t) k);
	memcpy(out_dist, r_dist, sizeof(double) * (size_t) nq * (size_t) k);
	free(pids);
	munmap(map, blk + res + 64);
	return failed;
}
