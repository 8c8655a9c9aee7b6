// flagsort.hpp -- in-place MSD byte radix sort with the reference's tie order.
//
// Why this exists: the reference sorts 16-byte records by their first 8 bytes with an UNSTABLE
// "American flag" sort (ksort.h:109-162, instantiated as radix_sort_mp128x in misc.c:8).  The
// order in which equal keys come out is observable downstream (chain backtracking chain.c:40,
// chain order chain.c:98, region order hit.c:119/264), so a bit-exact hot path has to leave
// ties in exactly the same places.  The routine below produces that permutation.  It is
// written once as host+device code: the host uses it for region bookkeeping, the chaining
// kernel uses it (one thread per problem) for the backtrack order.
//
//   pass(range, shift): histogram of digit (key>>shift)&255 -> bucket k owns [head_k, tail_k)
//     for k = 0..255: while head_k != tail_k:
//        item at head_k already has digit k -> ++head_k
//        otherwise lift it and chase: drop the item in hand at the head of ITS bucket, pick up
//        what was there, until the item in hand has digit k; store it at head_k, ++head_k
//   then every bucket is refined on the next lower byte: > 64 items -> pass(), 2..64 items ->
//   insertion sort; ranges of <= 64 items at top level are insertion-sorted directly.
//
// A pass whose items all share one digit leaves the range untouched, so such passes are
// skipped (the key bytes above the highest differing byte never move anything): identical
// output, and the common case "scores < 65536" costs two passes instead of eight.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MPB_HD __host__ __device__
#else
#define MPB_HD
#endif

namespace mpb {

template <class T, class KeyFn>
MPB_HD inline void insertion_sort_by(T *beg, T *end, KeyFn key)
{
	for (T *i = beg + 1; i < end; ++i) {
		if (key(*i) < key(*(i - 1))) {
			T hold = *i, *j = i;
			for (; j > beg && key(hold) < key(*(j - 1)); --j) *j = *(j - 1);
			*j = hold;
		}
	}
}

// One explicit stack frame per pending range: recursion depth is at most 8 but the fan-out is
// 256, so an explicit work list with bounded size is used instead of recursion (device friendly).
template <class T>
struct FlagRange { T *beg, *end; int shift; };

template <class T, class KeyFn>
MPB_HD inline void flag_pass(T *beg, T *end, int shift, KeyFn key, T **tail_out /*[256]*/)
{
	T *head[256];
	uint32_t cnt[256];
	for (int k = 0; k < 256; ++k) cnt[k] = 0;
	for (T *p = beg; p != end; ++p) ++cnt[(key(*p) >> shift) & 255];
	{
		T *p = beg;
		for (int k = 0; k < 256; ++k) { head[k] = p; p += cnt[k]; tail_out[k] = p; }
	}
	for (int k = 0; k < 256;) {
		if (head[k] == tail_out[k]) { ++k; continue; }
		int d = (int)((key(*head[k]) >> shift) & 255);
		if (d == k) { ++head[k]; continue; }
		T hand = *head[k];
		do {
			T next = *head[d];
			*head[d]++ = hand;
			hand = next;
			d = (int)((key(hand) >> shift) & 255);
		} while (d != k);
		*head[k]++ = hand;
	}
}

// Sort [beg,end) by key(.) (a 64-bit unsigned key) with the reference's tie order.
// `stack` must provide room for 8*256 pending ranges in the worst case; callers on the host pass
// a std::vector-backed buffer, the device passes a per-thread global scratch slab.
template <class T, class KeyFn>
MPB_HD inline void flag_sort_by(T *beg, T *end, KeyFn key, FlagRange<T> *stack)
{
	if (end - beg <= 64) { insertion_sort_by(beg, end, key); return; }
	int top = 0;
	stack[top++] = FlagRange<T>{beg, end, 56};
	T *tails[256];
	while (top > 0) {
		FlagRange<T> r = stack[--top];
		// skip passes that cannot move anything: all items share the digit at r.shift
		uint64_t lo = ~0ULL, hi = 0;
		for (T *p = r.beg; p != r.end; ++p) { uint64_t k = key(*p); lo = k < lo ? k : lo; hi = k > hi ? k : hi; }
		int shift = r.shift;
		while (shift > 0 && ((lo >> shift) & 255) == ((hi >> shift) & 255) && (lo >> shift >> 8) == (hi >> shift >> 8)) shift -= 8;
		// NB: the skip is only valid while ALL higher bytes agree, which holds here because every
		// range on the stack came out of one bucket of the byte above it.
		if (shift == 0 && ((lo & 255) == (hi & 255)) && (lo >> 8) == (hi >> 8)) continue; // all keys equal: untouched
		flag_pass(r.beg, r.end, shift, key, tails);
		if (shift > 0) {
			// push in reverse so that buckets are refined in ascending order (order does not matter
			// for the result -- buckets are disjoint -- but keeps the stack small)
			T *p = r.end;
			for (int k = 255; k >= 0; --k) {
				T *q = tails[k], *b = k ? tails[k - 1] : r.beg;
				(void)p;
				if (q - b > 64) stack[top++] = FlagRange<T>{b, q, shift - 8};
				else if (q - b > 1) insertion_sort_by(b, q, key);
			}
		}
	}
}

} // namespace mpb
