/*
 * vb_broker.h -- batches out of one-scan-per-backend.
 *
 * PostgreSQL runs one index scan per backend (amcanparallel = false, src/ivfflat.c:266) and backends are PROCESSES; the
 * library is 20x faster per query when the scans of many backends share one vb_ivf_search call (each probed list is then
 * read once per batch).  The broker is the piece between the two: requesters hand it ONE query each
 * and block; it collects what arrives within a short window (or until a batch is full), issues one batched call, and
 * hands every requester its own k results.  Only the broker touches the library (one CUDA context, one index image).
 *
 * Everything the two sides share lives in ONE block of shared memory (VbBrokerShared: a ring of request slots, each
 * with the query payload and the result area, as INTEGRATION.md describes; process-shared mutex and conditions), so a
 * requester can be a thread of the broker's process or a process that inherited / attached the mapping.  In a server
 * the block is a shmem request of the extension, the conditions are latches (SetLatch / WaitLatch, so query cancel
 * works) and the broker is a background worker; the protocol is this file's.
 */
#ifndef VB_BROKER_H
#define VB_BROKER_H

#include <stddef.h>
#include <stdint.h>

#include "vecb200.h"

typedef struct VbBrokerShared VbBrokerShared;	/* the shared block */
typedef struct VbBroker VbBroker;				/* the serving side: the block + the index + the serving thread */

typedef struct VbBrokerConfig
{
	int			max_batch;		/* queries per call at most = number of request slots (the headline uses 2048) */
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

/* bytes of the shared block for a configuration */
extern size_t VbBrokerSharedSize(const VbBrokerConfig *cfg);

/* initialises a block of VbBrokerSharedSize() bytes (zeroed or not) that every requester can address: call it ONCE,
 * before the requesters exist (postmaster-time shmem init / before fork); returns NULL on a bad configuration */
extern VbBrokerShared *VbBrokerSharedInit(void *block, const VbBrokerConfig *cfg);

/* serving side: starts the thread that answers the block's requests from `ix` (the caller's process owns the library) */
extern VbBroker *VbBrokerServe(VbBrokerShared *sh, vb_ivf *ix);

/* convenience for one process: allocates the block on the heap and serves it */
extern VbBroker *VbBrokerStart(vb_ivf *ix, const VbBrokerConfig *cfg);
extern VbBrokerShared *VbBrokerBlock(VbBroker *b);

/*
 * one scan, from any thread or process that can address the block: blocks until the batch this query joined has run.
 * ids / dist: k entries (id -1 / +inf past the candidates).  Returns the status of the batched call (VB_OK, or the
 * library's error: every scan of a failed batch gets it), VB_ESTATE when the broker is stopping, VB_ECUDA when no
 * answer came within timeout_ms (0 = wait for ever; a dead broker must not hang its backends).
 */
extern int	VbBrokerRequest(VbBrokerShared *sh, const void *query, int64_t *ids, double *dist, int timeout_ms);
extern int	VbBrokerSearch(VbBroker *b, const void *query, int64_t *ids, double *dist);

extern void VbBrokerGetStats(VbBrokerShared *sh, VbBrokerStats *out);

/* serves what is queued, then stops the thread; frees the block only if VbBrokerStart allocated it */
extern void VbBrokerStop(VbBroker *b);

#endif
