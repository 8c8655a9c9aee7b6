/*
 * oracle/chain.c -- TEST INFRASTRUCTURE ONLY (see ora.h).
 *
 * Restatement of the reference's anchor chaining (chain.c) and of the second-round refinement
 * core that is built on it (map.c:41-97).
 *
 * Anchor = x<<32 | y, sorted ascending; x = genome block id (first round, bbit=8) or
 * nucleotide end position (refinement, bbit=0); y = residue end position (bit 31 clear).
 *   fill      : f[i] = max(kmer, max_j f[j] + sc(i,j)) scanning j = i-1 .. st with the
 *               minimap2-style max_skip heuristic driven by t[] marks, the "hi/hf" rescue
 *               of the best-so-far anchor and the max_iter cap (chain.c:181-209).
 *   backtrack : ends with f >= min_sc sorted by f with the reference's UNSTABLE radix sort,
 *               peeled best-first; anchors of rejected chains stay marked (chain.c:26-75).
 *   compact   : chains reversed to ascending order, then ordered by first target coordinate
 *               with the same unstable sort (chain.c:77-110).
 */
#include <stdlib.h>
#include <string.h>
#include "ora.h"

static inline float log2_approx(float x) /* mppriv.h:91-99; valid for x >= 2 */
{
	union { float f; uint32_t i; } z = { x };
	float r = (float)((int32_t)((z.i >> 23) & 255) - 128);
	z.i &= ~(255U << 23);
	z.i += 127U << 23;
	r += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return r;
}

int32_t ora_comput_sc(const ora_chain_par_t *p, uint64_t ai, uint64_t aj) /* chain.c:112-151 */
{
	int32_t dq = (int32_t)ai - (int32_t)aj, dq3 = dq * 3, dr3, dd, dds = 0, sc;
	if (dq <= 0 || dq3 > p->max_dist_x || dq > p->max_dist_y) return INT32_MIN;
	if (p->bbit > 0) { /* block resolution: the gap is known only up to +-one block */
		int32_t bs = 1 << p->bbit;
		dr3 = (int32_t)(((ai >> 32) - (aj >> 32)) << p->bbit);
		if (dq3 < dr3 - bs) dd = dr3 - bs - dq3, dds = -dd;
		else if (dq3 > dr3 + bs) dd = dq3 - (dr3 + bs), dds = dd;
		else dd = 0;
	} else {
		dr3 = (int32_t)((ai >> 32) - (aj >> 32));
		if (dr3 == 0) return INT32_MIN;
		dd = dr3 > dq3 ? dr3 - dq3 : dq3 - dr3;
		dds = dq3 - dr3;
	}
	if (dd > p->bw) return INT32_MIN;
	if (p->bbit > 0) sc = p->kmer < dq ? p->kmer : dq;
	else if (p->kmer <= dq && p->kmer * 3 <= dr3) sc = p->kmer;
	else {
		int32_t dr = dr3 / 3, g = dr < dq ? dr : dq;
		sc = g < p->kmer ? g : p->kmer;
		if (dr3 - dr * 3 != 0) --sc;
	}
	if (dd > 0) {
		float lin = (float)dd * .33334f;
		float lg = dd >= 2 ? p->chn_coef_log * (log2_approx((float)(dd + 1)) - 1.0f) + 1.0f : (float)dd;
		if (p->is_spliced && dds < 0) sc -= (int32_t)(lin < lg ? lin : lg);
		else sc -= (int32_t)(lin + lg);
	}
	if (p->bbit > 0 && ai >> 32 == aj >> 32 && dd == 0) sc += 2; /* MP_BLOCK_BONUS, miniprot.h:23 */
	return sc;
}

/* chain.c:8-24 */
static int64_t bk_end(int32_t max_drop, const ora128_t *z, const int32_t *f, const int64_t *p, int32_t *t, int64_t k)
{
	int64_t i = (int64_t)z[k].y, end_i = -1, max_i = i;
	int32_t max_s = 0;
	if (i < 0 || t[i] != 0) return i;
	do {
		int32_t s;
		t[i] = 2;
		end_i = i = p[i];
		s = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
		if (s > max_s) max_s = s, max_i = i;
		else if (max_s - s > max_drop) break;
	} while (i >= 0 && t[i] == 0);
	for (i = (int64_t)z[k].y; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
	return max_i;
}

uint64_t *ora_chain(const ora_chain_par_t *par, int64_t n, const uint64_t *a, int32_t *n_u_, uint64_t **u_)
{
	ora_chain_par_t P = *par;
	int32_t *f, *t, *v, n_u = 0, hf = 0, max_drop = P.bw, pass;
	int64_t *p, i, j, k, st = 0, hi = -1, n_z = 0, n_v = 0;
	uint64_t *u = 0, *b, *u2;
	ora128_t *z, *w;

	*n_u_ = 0, *u_ = 0;
	if (n == 0 || a == 0) return 0;
	if (P.max_dist_x < P.bw) P.max_dist_x = P.bw;
	if (P.max_dist_y < P.bw && !P.is_spliced) P.max_dist_y = P.bw;
	if (P.is_spliced) max_drop = INT32_MAX;
	p = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
	f = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	v = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	t = (int32_t*)calloc((size_t)n, sizeof(int32_t));

	for (i = 0; i < n; ++i) { /* fill */
		int64_t max_j = -1;
		int32_t max_f = P.kmer, n_skip = 0;
		while (st < i && (int64_t)(((a[i] >> 32) - (a[st] >> 32)) << P.bbit) > P.max_dist_x) ++st;
		if (hi >= 0 && hi >= st) {
			int32_t sc = hf + ora_comput_sc(&P, a[i], a[hi]); /* hf >= 0, so INT32_MIN + hf cannot overflow */
			if (sc > max_f) max_f = sc, max_j = hi;
		} else hf = 0, hi = -1;
		if (i - st > P.max_iter) st = i - P.max_iter;
		for (j = i - 1; j >= st; --j) {
			int32_t sc = ora_comput_sc(&P, a[i], a[j]);
			if (sc == INT32_MIN) continue;
			sc += f[j];
			if (sc > max_f) {
				max_f = sc, max_j = j;
				if (n_skip > 0) --n_skip;
			} else if (t[j] == (int32_t)i) {
				if (++n_skip > P.max_skip) break;
			}
			if (p[j] >= 0) t[p[j]] = (int32_t)i;
		}
		f[i] = max_f, p[i] = max_j;
		if (hf < max_f) hf = max_f, hi = i;
	}

	/* backtrack (chain.c:26-75) */
	for (i = 0; i < n; ++i) if (f[i] >= P.min_sc) ++n_z;
	if (n_z == 0) { free(p); free(f); free(v); free(t); return 0; }
	z = (ora128_t*)malloc(sizeof(ora128_t) * (size_t)n_z);
	for (i = 0, k = 0; i < n; ++i) if (f[i] >= P.min_sc) z[k].x = (uint64_t)(int64_t)f[i], z[k++].y = (uint64_t)i;
	ora_sort128x(z, z + n_z);
	for (pass = 0; pass < 2; ++pass) { /* the reference runs the identical loop twice: count, then fill */
		memset(t, 0, sizeof(int32_t) * (size_t)n);
		n_v = 0, n_u = 0;
		for (k = n_z - 1; k >= 0; --k) {
			int64_t n_v0 = n_v, end_i;
			int32_t sc;
			if (t[z[k].y] != 0) continue;
			end_i = bk_end(max_drop, z, f, p, t, k);
			for (i = (int64_t)z[k].y; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
			sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
			if (sc >= P.min_sc && n_v > n_v0 && n_v - n_v0 >= P.min_cnt) {
				if (pass) u[n_u] = (uint64_t)sc << 32 | (uint64_t)(n_v - n_v0);
				++n_u;
			} else n_v = n_v0;
		}
		if (pass == 0) u = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n_u + 1));
	}
	free(z); free(p); free(f); free(t);
	*n_u_ = n_u, *u_ = u;
	if (n_u == 0) { free(v); return 0; }

	/* compact (chain.c:77-110) */
	b = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n_v);
	for (i = 0, k = 0; i < n_u; ++i) {
		int64_t k0 = k, ni = (int32_t)u[i];
		for (j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
	}
	free(v);
	w = (ora128_t*)malloc(sizeof(ora128_t) * (size_t)n_u);
	for (i = k = 0; i < n_u; ++i) {
		w[i].x = b[k] >> 32, w[i].y = (uint64_t)k << 32 | (uint64_t)i;
		k += (int32_t)u[i];
	}
	ora_sort128x(w, w + n_u);
	u2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n_u);
	{
		uint64_t *c = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n_v);
		for (i = k = 0; i < n_u; ++i) {
			int32_t src = (int32_t)w[i].y, cnt = (int32_t)u[src];
			u2[i] = u[src];
			memcpy(&c[k], &b[w[i].y >> 32], sizeof(uint64_t) * (size_t)cnt);
			k += cnt;
		}
		free(b);
		b = c;
	}
	memcpy(u, u2, sizeof(uint64_t) * (size_t)n_u);
	free(u2); free(w);
	return b;
}

/* map.c:41-97: sketch the window and the protein with all k-mers (mod_bit 0), join equal hashes
 * (groups with n1*n2 <= max_ava), chain at base resolution, keep the best-scoring chain */
uint64_t *ora_refine(const ora_tab_t *tab, const ora_chain_par_t *par, int32_t min_aa_len, int32_t max_ava, const uint8_t *nt,
                     int64_t l_nt, const char *aa, int32_t l_aa, int32_t *n_best, int32_t *sc_best)
{
	uint64_t *sd = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(l_nt + l_aa + 2)), *a, *b, *u = 0, *out;
	int64_t n_sd, i, k, j, n_a = 0;
	int32_t n_q, n_u = 0, best = 0, mx;
	*n_best = 0, *sc_best = 0;
	n_sd = ora_sketch_nt4(tab, nt, l_nt, min_aa_len, par->kmer, 0, 0, 0, sd);
	n_q = ora_sketch_prot(tab, aa, l_aa, par->kmer, 0, sd + n_sd);
	for (i = 0; i < n_q; ++i) sd[n_sd + i] |= 1ULL << 31;
	n_sd += n_q;
	ora_sort64(sd, sd + n_sd);
	a = 0;
	for (int pass = 0; pass < 2; ++pass) {
		n_a = 0;
		for (k = 0, i = 1; i <= n_sd; ++i) {
			if (i == n_sd || sd[k] >> 32 != sd[i] >> 32) {
				int64_t n1, n2, i1, i2;
				for (j = k; j < i; ++j) if (sd[j] >> 31 & 1) break;
				n1 = j - k, n2 = i - k - n1;
				if (n1 > 0 && n2 > 0 && n1 * n2 <= max_ava) {
					if (pass)
						for (i1 = k; i1 < k + n1; ++i1)
							for (i2 = k + n1; i2 < i; ++i2)
								a[n_a++] = (uint64_t)((uint32_t)sd[i1]) << 32 | ((uint32_t)sd[i2] << 1 >> 1);
					else n_a += n1 * n2;
				}
				k = i;
			}
		}
		if (!pass) a = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n_a + 1));
	}
	free(sd);
	ora_sort64(a, a + n_a);
	b = ora_chain(par, n_a, a, &n_u, &u);
	free(a);
	if (n_u == 0) { free(u); free(b); return 0; }
	mx = (int32_t)(u[0] >> 32);
	for (i = 1; i < n_u; ++i) if (mx < (int32_t)(u[i] >> 32)) mx = (int32_t)(u[i] >> 32), best = (int32_t)i;
	for (i = k = 0; i < best; ++i) k += (uint32_t)u[i];
	*n_best = (int32_t)(uint32_t)u[best], *sc_best = mx;
	out = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)*n_best);
	memcpy(out, b + k, sizeof(uint64_t) * (size_t)*n_best);
	free(u); free(b);
	return out;
}
