/*
 * vb_broker.h -- batches out of one-scan-per-backend.
 *
 * PostgreSQL runs one index scan per backend (amcanparallel = false, src/ivfflat.c:266); the library is 20x faster per
 * query when the scans of many backends share one vb_ivf_search call (each probed list is then read once per batch,
 * DESIGN.md section 5).  The broker is the piece between the two: requesters hand it ONE query each and block; it
 * collects what arrives within a short window (or until a batch is full), issues one batched call, and hands every
 * requester its own k results.
 *
 * This file is the batching logic with POSIX threads as requesters (what the harness runs); in a server the requesters
 * are backends and the three primitives map one to one: the mutex-protected slot array -> a ring in shared memory, the
 * condition variables -> the broker's and the backends' latches (SetLatch / WaitLatch, so query cancel works), the
 * broker thread -> a background worker that owns the CUDA context and the index images.
 */
#ifndef VB_BROKER_H
#define VB_BROKER_H

#include <stddef.h>
#include <stdint.h>

#include "vecb200.h"

typedef struct VbBroker VbBroker;

typedef struct VbBrokerConfig
{
	int			max_batch;		/* queries per call at most (the headline uses 2048) */
	int			window_us;		/* how long the first request of a batch waits for company */
	int			probes;			/* ivfflat.probes of the scans this broker serves */
	int			k;				/* results per scan (LIMIT) */
	size_t		query_bytes;	/* payload bytes of one query (dim * 4 for vector) */
} VbBrokerConfig;

typedef struct VbBrokerStats
{
	int64_t		requests;		/* scans served */
	int64_t		batches;		/* vb_ivf_search calls issued */
	int64_t		largest;		/* largest batch */
	int64_t		failed;			/* scans that got an error back */
} VbBrokerStats;

/* starts the broker thread for one loaded index; NULL when out of memory / the thread cannot start */
extern VbBroker *VbBrokerStart(vb_ivf *ix, const VbBrokerConfig *cfg);

/*
 * one scan: blocks until the batch this query joined has run.  ids / dist: k entries (id -1 / +inf past the candidates).
 * Returns the status of the batched call (VB_OK, or the library's error: every scan of a failed batch gets it), or
 * VB_ESTATE when the broker is stopping.  Thread-safe.
 */
extern int	VbBrokerSearch(VbBroker *b, const void *query, int64_t *ids, double *dist);

extern void VbBrokerGetStats(VbBroker *b, VbBrokerStats *out);

/* serves what is queued, then stops the thread and frees the broker; later VbBrokerSearch calls are invalid */
extern void VbBrokerStop(VbBroker *b);

#endif
