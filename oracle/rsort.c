/*
 * oracle/rsort.c -- TEST INFRASTRUCTURE ONLY (see ora.h).
 *
 * The reference sorts with an in-place MSD byte radix sort ("American flag" cycle-leader
 * permutation) that falls back to insertion sort for ranges of <= 64 items
 * (ksort.h:109-162, instantiated in misc.c:4-8).  For 64-bit keys the result is the unique
 * sorted order, so ora_sort64 may be any correct sort.  radix_sort_mp128x sorts 16-byte items
 * by .x only and is NOT stable: the order of equal keys is whatever the cycle-leader
 * permutation leaves behind, and chain backtracking (chain.c:40), chain ordering
 * (chain.c:98) and region ordering (hit.c:119,264) observe it.  ora_sort128x therefore
 * restates the permutation exactly:
 *   pass(range, shift): histogram of digit (key>>shift)&255; bucket k owns [b_k, e_k);
 *     for k = 0..255: while bucket k has an unplaced slot at b_k:
 *        if the item there belongs to k: advance b_k;
 *        else lift it, and repeatedly drop the lifted item at the head of ITS bucket, lifting
 *             what was there, until an item for bucket k is in hand; put it at b_k, advance.
 *   after the pass each bucket is recursed into with shift-8 (if shift > 0): more than 64
 *   items -> pass(); 2..64 items -> insertion sort; top level: <= 64 items -> insertion sort.
 */
#include <stdlib.h>
#include <string.h>
#include "ora.h"

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
	return x < y ? -1 : x > y;
}

void ora_sort64(uint64_t *beg, uint64_t *end)
{
	if (end - beg > 1) qsort(beg, (size_t)(end - beg), sizeof(uint64_t), cmp_u64);
}

static void ins128(ora128_t *beg, ora128_t *end)
{
	ora128_t *i, *j;
	for (i = beg + 1; i < end; ++i) {
		if (i->x < (i-1)->x) {
			ora128_t t = *i;
			for (j = i; j > beg && t.x < (j-1)->x; --j) *j = *(j-1);
			*j = t;
		}
	}
}

static void flag128(ora128_t *beg, ora128_t *end, int shift)
{
	ora128_t *head[256], *tail[256], *p;
	size_t cnt[256];
	int k;
	memset(cnt, 0, sizeof(cnt));
	for (p = beg; p != end; ++p) ++cnt[p->x >> shift & 255];
	for (k = 0, p = beg; k < 256; ++k) head[k] = p, p += cnt[k], tail[k] = p;
	for (k = 0; k < 256;) {
		if (head[k] == tail[k]) { ++k; continue; }
		int d = (int)(head[k]->x >> shift & 255);
		if (d == k) { ++head[k]; continue; }
		ora128_t hand = *head[k];
		do {
			ora128_t nxt = *head[d];
			*head[d]++ = hand;
			hand = nxt;
			d = (int)(hand.x >> shift & 255);
		} while (d != k);
		*head[k]++ = hand;
	}
	if (shift > 0) {
		int s = shift > 8 ? shift - 8 : 0;
		for (k = 0, p = beg; k < 256; ++k) {
			ora128_t *q = tail[k];
			if (q - p > 64) flag128(p, q, s);
			else if (q - p > 1) ins128(p, q);
			p = q;
		}
	}
}

void ora_sort128x(ora128_t *beg, ora128_t *end)
{
	if (end - beg <= 64) ins128(beg, end);
	else flag128(beg, end, 56);
}
