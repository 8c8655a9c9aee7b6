/*
 * oracle/sketch.c -- TEST INFRASTRUCTURE ONLY (see ora.h).
 *
 * Restatement of the seed extraction of the reference (sketch.c) and of the per-query seed
 * lookup / anchor expansion (map.c:126-177).
 *   - k-mers are over a 4-bit reduced amino-acid alphabet (13 classes, codes >= 14 are
 *     stop/unknown and break the window), hashed with an invertible 4k-bit mixer and kept iff
 *     the low mod_bit bits of the hash are zero ("mod sampling", sketch.c:32-34).
 *   - genome side: stop-to-stop ORFs of >= min_aa_len codons in each of the three frames of
 *     the given strand; every k-mer inside an ORF is emitted with the block (pos >> bbit) of
 *     its last base; the list is sorted and exact duplicates removed (sketch.c:62-100).
 */
#include <stdlib.h>
#include <string.h>
#include "ora.h"

uint32_t ora_hash32_mask(uint32_t x, uint32_t mask) /* sketch.c:7-16 */
{
	x = (x + ~(x << 15)) & mask;
	x ^= x >> 10;
	x = (x + (x << 3)) & mask;
	x ^= x >> 6;
	x = (x + ~(x << 11)) & mask;
	x ^= x >> 16;
	return x;
}

int32_t ora_sketch_prot(const ora_tab_t *tab, const char *seq, int32_t len, int32_t kmer, int32_t mod_bit, uint64_t *out)
{
	const uint32_t mask = (1U << kmer * 4) - 1, mod = (1U << mod_bit) - 1;
	uint32_t w = 0;
	int32_t i, run = 0, n = 0;
	for (i = 0; i < len; ++i) {
		uint32_t c = tab->aa13[(uint8_t)seq[i]];
		if (c >= 14) { w = 0, run = 0; continue; }
		w = (w << 4 | c) & mask;
		if (++run >= kmer) {
			uint32_t h = ora_hash32_mask(w, mask);
			if ((h & mod) == 0) out[n++] = (uint64_t)(h >> mod_bit) << 32 | (uint32_t)i;
		}
	}
	return n;
}

/* all k-mers of one ORF [st,en) (en-st a multiple of 3), sketch.c:40-60 */
static int64_t orf_kmers(const ora_tab_t *tab, const uint8_t *seq, int64_t st, int64_t en, int32_t kmer, int32_t mod_bit,
                         int32_t bbit, int64_t boff, uint64_t *out, int64_t n)
{
	const uint32_t mask = (1U << kmer * 4) - 1, mod = (1U << mod_bit) - 1;
	uint32_t w = 0;
	int32_t run = 0;
	int64_t i;
	for (i = st; i < en; i += 3) {
		uint32_t cod = (uint32_t)seq[i] << 4 | (uint32_t)seq[i+1] << 2 | seq[i+2];
		w = (w << 4 | tab->codon13[cod]) & mask;
		if (++run >= kmer) {
			uint32_t h = ora_hash32_mask(w, mask);
			if ((h & mod) == 0) out[n++] = (uint64_t)(h >> mod_bit) << 32 | (uint64_t)(((i + 2) >> bbit) + boff);
		}
	}
	return n;
}

int64_t ora_sketch_nt4(const ora_tab_t *tab, const uint8_t *seq, int64_t len, int32_t min_aa_len, int32_t kmer, int32_t mod_bit,
                       int32_t bbit, int64_t boff, uint64_t *out)
{
	int64_t end[3] = { -1, -1, -1 }, cnt[3] = { 0, 0, 0 }, i, n = 0, m;
	int32_t f, g, run = 0;
	uint32_t cod = 0;
	for (i = 0; i < len; ++i) {
		f = (int32_t)((i + 1) % 3); /* frame of the codon ending at i */
		if (seq[i] < 4) {
			cod = (cod << 2 | seq[i]) & 0x3f;
			if (++run < 3) continue;
			if (tab->codon[cod] >= 20) { /* stop closes this frame's run */
				if (cnt[f] >= min_aa_len) n = orf_kmers(tab, seq, end[f] + 1 - cnt[f] * 3, end[f] + 1, kmer, mod_bit, bbit, boff, out, n);
				cnt[f] = 0, end[f] = -1;
			} else end[f] = i, ++cnt[f];
		} else { /* ambiguous base closes all three */
			for (g = 0; g < 3; ++g) {
				if (cnt[g] >= min_aa_len) n = orf_kmers(tab, seq, end[g] + 1 - cnt[g] * 3, end[g] + 1, kmer, mod_bit, bbit, boff, out, n);
				cnt[g] = 0, end[g] = -1;
			}
			run = 0, cod = 0;
		}
	}
	for (g = 0; g < 3; ++g)
		if (cnt[g] >= min_aa_len) n = orf_kmers(tab, seq, end[g] + 1 - cnt[g] * 3, end[g] + 1, kmer, mod_bit, bbit, boff, out, n);
	if (n <= 1) return n;
	ora_sort64(out, out + n);
	for (i = 1, m = 0; i < n; ++i)
		if (out[m] != out[i]) out[++m] = out[i];
	return m + 1;
}

/* map.c:126-141: adaptive occurrence cap, box-plot rule on the bucket sizes of this query's seeds */
static int32_t occ_cap(const int64_t *ki, int64_t n_kb, int64_t n_bucket, int32_t n, const uint64_t *sd)
{
	uint64_t *c = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n), q25, q75;
	int32_t i, r;
	for (i = 0; i < n; ++i) {
		int64_t b = (int64_t)(sd[i] >> 32), e = b + 1 < n_bucket ? ki[b+1] : n_kb;
		c[i] = (uint64_t)(e - ki[b]);
	}
	ora_sort64(c, c + n);
	q25 = c[(int64_t)(n * .25 + .499)];
	q75 = c[(int64_t)(n * .75 + .499)];
	free(c);
	r = (int32_t)(q75 + (q75 - q25) * 1.5 + 10.);
	return r;
}

uint64_t *ora_seed_anchors(const ora_tab_t *tab, const int64_t *ki, int64_t n_kb, const uint32_t *kb, int32_t kmer, int32_t mod_bit,
                           int32_t max_occ_cap, const char *seq, int32_t len, int64_t *n_a_)
{
	const int64_t n_bucket = (int64_t)1 << (kmer * 4 - mod_bit);
	uint64_t *sd = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(len + 1)), *a;
	int32_t n_sd, i, max_occ = max_occ_cap;
	int64_t n_a = 0, k = 0, j;
	n_sd = ora_sketch_prot(tab, seq, len, kmer, mod_bit, sd);
	ora_sort64(sd, sd + n_sd);
	if (n_sd >= 8) { /* map.c:158-161 */
		max_occ = occ_cap(ki, n_kb, n_bucket, n_sd, sd);
		if (max_occ > max_occ_cap) max_occ = max_occ_cap;
	}
	for (i = 0; i < n_sd; ++i) {
		int64_t b = (int64_t)(sd[i] >> 32), e = b + 1 < n_bucket ? ki[b+1] : n_kb;
		if (e - ki[b] <= max_occ) n_a += e - ki[b];
	}
	a = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n_a + 1));
	for (i = 0; i < n_sd; ++i) {
		int64_t b = (int64_t)(sd[i] >> 32), e = b + 1 < n_bucket ? ki[b+1] : n_kb;
		if (e - ki[b] <= max_occ)
			for (j = ki[b]; j < e; ++j) a[k++] = (uint64_t)kb[j] << 32 | (uint32_t)sd[i];
	}
	free(sd);
	ora_sort64(a, a + n_a);
	*n_a_ = n_a;
	return a;
}
