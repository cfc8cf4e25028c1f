/*
 * pgv_pairingheap.h -- restatement of PostgreSQL's lib/pairingheap.c semantics.
 * TEST INFRASTRUCTURE ONLY (see pgv_oracle.h).
 *
 * PostgreSQL core is NOT under /root/reference (third-party dependency of the
 * reference, any of PG 13-20; Docker default 17).  The published algorithm is
 * restated here because it decides tie order at the reference's call sites:
 * src/ivfscan.c:86-115 and src/hnswutils.c:828-829,876-877,890-891,942,954-955,969-981.
 * Parity on tie order is UNPINNED (no reference test fixes it).
 *
 * Rules restated:
 *   - the root is the node that compares GREATEST under the comparator;
 *   - merge(a,b): if cmp(a,b) < 0 swap; b becomes the FIRST child of a
 *     (on a tie the first argument stays on top);
 *   - add(h,n): root = merge(root, n);
 *   - remove_first: two-pass merge of the root's children: pass 1 walks
 *     siblings left to right merging adjacent pairs and PREPENDING each result
 *     to a list (an odd last child is prepended unmerged); pass 2 folds that
 *     list from its head.
 */
#ifndef PGV_PAIRINGHEAP_H
#define PGV_PAIRINGHEAP_H

#include <stddef.h>

typedef struct ph_node
{
	struct ph_node *first_child;
	struct ph_node *next_sibling;
	struct ph_node *prev_or_parent;
}			ph_node;

typedef int (*ph_comparator) (const ph_node *a, const ph_node *b, void *arg);

typedef struct
{
	ph_comparator cmp;
	void	   *arg;
	ph_node    *root;
}			ph_heap;

static inline void
ph_init(ph_heap *h, ph_comparator cmp, void *arg)
{
	h->cmp = cmp;
	h->arg = arg;
	h->root = NULL;
}

static inline int
ph_is_empty(const ph_heap *h)
{
	return h->root == NULL;
}

static inline ph_node *
ph_merge(ph_heap *h, ph_node *a, ph_node *b)
{
	if (a == NULL)
		return b;
	if (b == NULL)
		return a;
	if (h->cmp(a, b, h->arg) < 0)
	{
		ph_node    *t = a;

		a = b;
		b = t;
	}
	/* b becomes the first child of a */
	if (a->first_child)
		a->first_child->prev_or_parent = b;
	b->prev_or_parent = a;
	b->next_sibling = a->first_child;
	a->first_child = b;
	return a;
}

static inline void
ph_add(ph_heap *h, ph_node *n)
{
	n->first_child = NULL;
	h->root = ph_merge(h, h->root, n);
	h->root->prev_or_parent = NULL;
	h->root->next_sibling = NULL;
}

static inline ph_node *
ph_first(ph_heap *h)
{
	return h->root;
}

static inline ph_node *
ph_merge_children(ph_heap *h, ph_node *children)
{
	ph_node    *curr,
			   *next,
			   *pairs,
			   *newroot;

	if (children == NULL || children->next_sibling == NULL)
		return children;

	next = children;
	pairs = NULL;
	for (;;)
	{
		curr = next;
		if (curr == NULL)
			break;
		if (curr->next_sibling == NULL)
		{
			/* odd last child goes on the list unmerged */
			curr->next_sibling = pairs;
			pairs = curr;
			break;
		}
		next = curr->next_sibling->next_sibling;
		curr = ph_merge(h, curr, curr->next_sibling);
		curr->next_sibling = pairs;
		pairs = curr;
	}

	newroot = pairs;
	next = pairs->next_sibling;
	while (next)
	{
		curr = next;
		next = curr->next_sibling;
		newroot = ph_merge(h, newroot, curr);
	}
	return newroot;
}

static inline ph_node *
ph_remove_first(ph_heap *h)
{
	ph_node    *result = h->root;
	ph_node    *children = result->first_child;

	h->root = ph_merge_children(h, children);
	if (h->root)
	{
		h->root->prev_or_parent = NULL;
		h->root->next_sibling = NULL;
	}
	return result;
}

#define ph_container(type, member, ptr) ((type *) ((char *) (ptr) - offsetof(type, member)))

#endif
