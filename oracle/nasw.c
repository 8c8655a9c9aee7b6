/*
 * oracle/nasw.c -- TEST INFRASTRUCTURE ONLY (see ora.h).
 *
 * Scalar restatement of ns_global_gs16b (reference nasw-sse.c:340-551) in LOGICAL column
 * order.  The reference runs a Farrar-striped 8 x int16 SSE kernel; what is restated here is
 * its observable semantics, including the parts that depend on the striping:
 *   - all arithmetic saturates to int16 (nasw-sse.c:360-402, _mm_adds/_mm_subs_epi16);
 *   - the row is padded to 8*slen columns (slen = ceil(al/8)) with profile -32768
 *     (nasw-sse.c:212-224); padding columns take part in the extension row maximum;
 *   - the first pass restarts the insertion chain at each of the 8 segment starts
 *     (column % slen == 0), which decides the state nibble and bit 4 of the traceback word;
 *   - the lazy-F loop (nasw-sse.c:408-422 / 521-537) is equivalent to one left-to-right pass
 *     over logical columns that raises H and sets bit 9.  PRECONDITION: gap
 *     open go >= 1.  The loop stops when "I - ge <= max(H, I) - go - ge" holds in all lanes; with
 *     go > 0 that is only true where I did not raise H, so the stop loses nothing.  With go == 0
 *     it is true at once and the reference's scores depend on the stripe layout (only the first
 *     column of each segment sees the carried-over insertion): not restated here, refused by the
 *     product (backend.cu bad_scoring); tests/test_oracle_pin.py pins both facts.
 * Cell recurrences: nasw-sse.c:15-22.  Backtrack: nasw-sse.c:40-89.  Sequence preparation:
 * nasw-sse.c:91-210.  Extension bookkeeping: nasw-sse.c:423-443.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "ora.h"

#define NEG (-32768)

int16_t *ora_nasw_dbg_H = 0;
uint16_t *ora_nasw_dbg_tb = 0;

static inline int32_t sat16(int32_t x) { return x < -32768 ? -32768 : (x > 32767 ? 32767 : x); }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

/* nasw-sse.c:330-338: bit-trick log2, FP32, each operation rounded separately */
static float log2_approx(float x)
{
	union { float f; uint32_t i; } z = { x };
	float r = (float)((int32_t)((z.i >> 23) & 255) - 128);
	z.i &= ~(255U << 23);
	z.i += 127U << 23;
	r += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return r;
}

/* nasw-sse.c:91-104: amino acid of the codon ENDING at i; X until three clean bases are seen */
static void fill_nas(const ora_tab_t *tab, const uint8_t *ns, int32_t nl, uint8_t *nas)
{
	int32_t i, run = 0;
	uint8_t cod = 0, x = tab->aa20['X'];
	for (i = 0; i < nl; ++i) {
		uint8_t c = tab->nt4[ns[i]];
		nas[i] = x;
		if (c < 4) {
			cod = (uint8_t)((cod << 2 | c) & 0x3f);
			if (++run >= 3) nas[i] = tab->codon[cod];
		} else cod = 0, run = 0;
	}
}

/* nasw-sse.c:106-155: forward orientation (global alignment and right extension) */
static void prep_forward(const ora_tab_t *tab, const ora_nasw_par_t *p, const uint8_t *ns, int32_t nl, const char *as, int32_t al,
                         const uint8_t *ss, uint8_t *nas, uint8_t *aas, int8_t *don, int8_t *acc)
{
	int32_t i, j;
	uint8_t *c = nas; /* nt4 codes first, overwritten by fill_nas at the end */
	for (j = 0; j < al; ++j) aas[j] = tab->aa20[(uint8_t)as[j]];
	for (i = 0; i < nl; ++i) c[i] = tab->nt4[ns[i]];
	for (i = 0; i <= nl; ++i) don[i] = acc[i] = (int8_t)p->sp[3];
	for (i = 0; i < nl - 3; ++i) {
		int32_t t = 3;
		if (c[i+1] == 2 && c[i+2] == 3) /* GT */
			t = (i + 3 < nl && (c[i+3] == 0 || c[i+3] == 2)) ? (c[i] == 2 ? -1 : 4) : 0;
		else if (c[i+1] == 2 && c[i+2] == 1 && c[i] == 2) t = 1; /* GGC */
		else if (c[i+1] == 0 && c[i+2] == 3) t = 2; /* AT */
		don[i] = (int8_t)(t < 0 ? 0 : p->sp[t]);
	}
	for (i = 1; i < nl; ++i) {
		int32_t t = 3, pen = 0;
		if (c[i-1] == 0 && c[i] == 2) { /* AG */
			t = (i >= 2 && (c[i-2] == 1 || c[i-2] == 3)) ? -1 : 0;
			for (j = i - 4; j >= 0 && j > i - 7; --j)
				if (c[j] != 1 && c[j] != 3) pen += p->sp[5];
		} else if (c[i-1] == 0 && c[i] == 1) t = 2; /* AC */
		acc[i] = (int8_t)(t < 0 ? 0 : p->sp[t]);
		if (t == -1 || t == 0) acc[i] = (int8_t)(acc[i] + pen);
	}
	if (ss) { /* nasw-sse.c:138-152 */
		int32_t cap = (p->io + 1) / 2 - 1;
		for (i = 1; i < nl; ++i) {
			int32_t s = (int8_t)(ss[i] >> 1) - 64;
			if (s > cap) s = cap;
			if (ss[i] == 0xff) don[i-1] = (int8_t)(don[i-1] - p->sp_null_bonus), acc[i-1] = (int8_t)(acc[i-1] - p->sp_null_bonus);
			else if (ss[i] & 1) acc[i-1] = (int8_t)(acc[i-1] - s);
			else don[i-1] = (int8_t)(don[i-1] - s);
		}
	}
	fill_nas(tab, ns, nl, nas);
}

/* nasw-sse.c:157-210: left extension; both sequences reversed, splice rules mirrored */
static void prep_left(const ora_tab_t *tab, const ora_nasw_par_t *p, const uint8_t *ns, int32_t nl, const char *as, int32_t al,
                      const uint8_t *ss, uint8_t *nas, uint8_t *aas, int8_t *don, int8_t *acc)
{
	int32_t i, j;
	uint8_t *c = nas, x = tab->aa20['X'];
	for (j = 0; j < al; ++j) aas[al - 1 - j] = tab->aa20[(uint8_t)as[j]];
	for (i = 0; i < nl; ++i) c[nl - 1 - i] = tab->nt4[ns[i]];
	for (i = 0; i <= nl; ++i) don[i] = acc[i] = (int8_t)p->sp[3];
	for (i = 0; i < nl - 3; ++i) { /* mirrored acceptor */
		int32_t t = 3, pen = 0;
		if (c[i+1] == 2 && c[i+2] == 0) { /* GA = reversed AG */
			t = (i + 3 < nl && (c[i+3] == 1 || c[i+3] == 3)) ? -1 : 0;
			for (j = i + 5; j < nl && j < i + 8; ++j)
				if (c[j] != 1 && c[j] != 3) pen += p->sp[5];
		} else if (c[i+1] == 1 && c[i+2] == 0) t = 2; /* CA */
		don[i] = (int8_t)(t < 0 ? 0 : p->sp[t]);
		if (t == -1 || t == 0) don[i] = (int8_t)(don[i] + pen);
	}
	for (i = 1; i < nl; ++i) { /* mirrored donor */
		int32_t t = 3;
		if (c[i-1] == 3 && c[i] == 2) /* TG = reversed GT */
			t = (i >= 2 && (c[i-2] == 0 || c[i-2] == 2)) ? ((i + 1 < nl && c[i+1] == 2) ? -1 : 4) : 0;
		else if (c[i-1] == 1 && c[i] == 2 && i + 1 < nl && c[i+1] == 1) t = 1; /* CGG */
		else if (c[i-1] == 3 && c[i] == 0) t = 2; /* TA */
		acc[i] = (int8_t)(t < 0 ? 0 : p->sp[t]);
	}
	if (ss) { /* nasw-sse.c:189-203 */
		int32_t cap = (p->io + 1) / 2 - 1;
		for (i = 0; i < nl; ++i) {
			int32_t s = (int8_t)(ss[i] >> 1) - 64;
			if (s > cap) s = cap;
			if (ss[i] == 0xff) don[nl-i-1] = (int8_t)(don[nl-i-1] - p->sp_null_bonus), acc[nl-i-1] = (int8_t)(acc[nl-i-1] - p->sp_null_bonus);
			else if (ss[i] & 1) don[nl-i-1] = (int8_t)(don[nl-i-1] - s);
			else acc[nl-i-1] = (int8_t)(acc[nl-i-1] - s);
		}
	}
	fill_nas(tab, ns, nl, nas);
	for (i = 0; i < nl >> 1; ++i) { uint8_t t = nas[i]; nas[i] = nas[nl-1-i]; nas[nl-1-i] = t; }
	if (nl >= 2) {
		memmove(nas + 2, nas, (size_t)(nl - 2));
		nas[0] = nas[1] = x;
	}
}

/* nasw.h:141-151: append one operation, merging runs except for F and G */
static void cigar_push(ora_nasw_rst_t *r, uint32_t op, int32_t len)
{
	if (r->n_cigar == 0 || op != (r->cigar[r->n_cigar-1] & 0xf) || op == 10 || op == 11) {
		if (r->n_cigar == r->m_cigar) {
			r->m_cigar += (r->m_cigar >> 1) + 8;
			r->cigar = (uint32_t*)realloc(r->cigar, sizeof(uint32_t) * (size_t)r->m_cigar);
		}
		r->cigar[r->n_cigar++] = (uint32_t)len << 4 | op;
	} else r->cigar[r->n_cigar-1] += (uint32_t)len << 4;
}

/* nasw-sse.c:40-89 with tb[] in row-major logical order */
static void backtrack(const uint16_t *tb, int32_t W, int32_t nl, int32_t al, ora_nasw_rst_t *r)
{
	static const uint8_t st2op[10] = { 0, 1, 2, 3, 12, 13, 10, 10, 11, 11 };
	int32_t i = nl - 1, j = al - 1, last = 0, k;
	while (i >= 2 && j >= 0) {
		int32_t x = tb[(size_t)i * W + j], state, ext;
		if (x >> 9 & 1) x = 1 | (x >> 4 << 4);
		state = last == 0 ? (x & 0xf) : last;
		ext = (state >= 1 && state <= 5) ? (x >> (state + 3) & 1) : 0;
		cigar_push(r, st2op[state], (state == 7 || state == 9) ? 2 : 1);
		switch (state) {
		case 0: i -= 3, --j; break;
		case 1: --j; break;
		case 2: i -= 3; break;
		case 3: --i; break;
		case 4: case 5: --i; if (!ext) --j; break;
		case 6: --i; break;
		case 7: i -= 2; break;
		case 8: --i, --j; break;
		case 9: i -= 2, --j; break;
		}
		last = (state >= 1 && state <= 5 && ext) ? state : 0;
	}
	if (j > 0) cigar_push(r, 1, j);
	if (i >= 0) {
		int32_t l = (i + 1) / 3 * 3, t = (i + 1) % 3;
		if (l > 0) cigar_push(r, 2, l);
		if (t != 0) cigar_push(r, 10, t);
	}
	for (k = 0; k < r->n_cigar >> 1; ++k) {
		uint32_t t = r->cigar[k]; r->cigar[k] = r->cigar[r->n_cigar-1-k]; r->cigar[r->n_cigar-1-k] = t;
	}
	for (k = 0; k < r->n_cigar; ++k) { /* nasw-sse.c:30-38 */
		uint32_t op = r->cigar[k] & 0xf;
		if ((op == 12 || op == 13) && r->cigar[k] >> 4 < 3) r->cigar[k] = r->cigar[k] >> 4 << 4 | 11;
	}
}

void ora_nasw(const ora_tab_t *tab, const ora_nasw_par_t *p, const uint8_t *ns, int32_t nl, const char *as, int32_t al,
              const uint8_t *ss, ora_nasw_rst_t *r)
{
	const int32_t is_ext = !!(p->flag & 6), want_tb = (p->flag & 1) && !is_ext;
	const int32_t slen = (al + 7) / 8, W = slen * 8;
	const int32_t go = p->go, ge = p->ge, io = p->io, fs = p->fs;
	int32_t i, j;
	uint8_t *nas = (uint8_t*)malloc((size_t)nl + 1), *aas = (uint8_t*)malloc((size_t)al + 1);
	int8_t *don = (int8_t*)malloc((size_t)nl + 1), *acc = (int8_t*)malloc((size_t)nl + 1);
	int16_t *buf = (int16_t*)malloc(sizeof(int16_t) * (size_t)(W + 1) * 12), *H, *H1, *H2, *H3, *Hm, *D, *D1, *D2, *D3, *A, *B, *C;
	uint16_t *tb = 0;
	int32_t max_sc = INT32_MIN, max_log = INT32_MIN, max_i = -1;

	r->n_cigar = 0, r->nt_len = nl, r->aa_len = al, r->score = INT32_MIN;
	if (p->flag & 2) prep_left(tab, p, ns, nl, as, al, ss, nas, aas, don, acc);
	else prep_forward(tab, p, ns, nl, as, al, ss, nas, aas, don, acc);
	H = buf, H1 = H + W + 1, H2 = H1 + W + 1, H3 = H2 + W + 1, Hm = H3 + W + 1;
	D = Hm + W + 1, D1 = D + W + 1, D2 = D1 + W + 1, D3 = D2 + W + 1, A = D3 + W + 1, B = A + W + 1, C = B + W + 1;
	for (j = 0; j < (W + 1) * 12; ++j) buf[j] = NEG;
	if (want_tb) tb = (uint16_t*)calloc((size_t)(nl > 0 ? nl : 1) * (size_t)(W > 0 ? W : 1), sizeof(uint16_t));

	for (i = 2; i < nl; ++i) {
		const int8_t *srow = p->mat + nas[i] * 22;
		const int32_t gei = nas[i] == 20 ? fs : ge;
		const int32_t dim1 = don[i-1], di = don[i], dip1 = don[i+1], ai = acc[i], aim1 = acc[i-1], aim2 = acc[i-2];
		/* boundary column -1 (nasw-sse.c:253-271): proper values only while i == 2 */
		const int32_t b3 = i == 2 ? 0 : NEG, b2 = i == 2 ? -fs : NEG, b1 = i == 2 ? -fs : NEG;
		int32_t I = NEG, last_h = NEG, It;
		int16_t *tmp;
		for (j = 0; j < W; ++j) { /* first pass */
			int32_t h3l = j ? H3[j-1] : b3, h2l = j ? H2[j-1] : b2, h1l = j ? H1[j-1] : b1;
			int32_t s = j < al ? srow[aas[j]] : NEG;
			int32_t h, t, u, v, y = 0, z = 0;
			if (j % slen == 0) I = NEG, last_h = NEG; /* striping: each segment restarts its insertion chain */
			h = sat16(h3l + s);
			t = sat16(last_h - go);
			if (I > t) z |= 1 << 4;
			I = sat16(imax(t, I) - ge);
			if (I > h) y = 1;
			h = imax(h, I);
			u = sat16(H3[j] - go), v = D3[j];
			if (v > u) z |= 1 << 5;
			t = sat16(imax(u, v) - gei);
			D[j] = (int16_t)t;
			if (t > h) y = 2;
			h = imax(h, t);
			u = sat16(H1[j] - io), v = A[j];
			t = sat16(u - dim1);
			if (v > t) z |= 1 << 6;
			t = imax(t, v), A[j] = (int16_t)t;
			t = sat16(t - ai);
			if (t > h) y = 3;
			h = imax(h, t);
			u = sat16(h1l - io), v = B[j];
			t = sat16(u - di);
			if (v > t) z |= 1 << 7;
			t = imax(t, v), B[j] = (int16_t)t;
			t = sat16(t - aim2);
			if (t > h) y = 4;
			h = imax(h, t);
			v = C[j];
			t = sat16(u - dip1);
			if (v > t) z |= 1 << 8;
			t = imax(t, v), C[j] = (int16_t)t;
			t = sat16(t - aim1);
			if (t > h) y = 5;
			h = imax(h, t);
			t = sat16(H1[j] - fs); if (t > h) y = 6; h = imax(h, t);
			t = sat16(H2[j] - fs); if (t > h) y = 7; h = imax(h, t);
			t = sat16(h1l - fs);   if (t > h) y = 8; h = imax(h, t);
			t = sat16(h2l - fs);   if (t > h) y = 9; h = imax(h, t);
			H[j] = (int16_t)h;
			if (tb) tb[(size_t)i * W + j] = (uint16_t)(z | y);
			last_h = h;
		}
		for (j = 1, It = NEG; j < W; ++j) { /* lazy-F, closed form */
			It = sat16(imax(sat16(H[j-1] - go), It) - ge);
			if (It > H[j]) {
				H[j] = (int16_t)It;
				if (tb) tb[(size_t)i * W + j] |= 1 << 9;
			}
		}
		if (ora_nasw_dbg_H) memcpy(ora_nasw_dbg_H + (size_t)i * W, H, sizeof(int16_t) * (size_t)W);
		if (!tb) { /* nasw-sse.c:423-433 (executed for every non-traceback call) */
			int32_t mx = NEG, end_sc, tsc, tlog, x = i - al * 3;
			for (j = 0; j < W; ++j) mx = imax(mx, H[j]);
			end_sc = H[al-1] + p->end_bonus;
			tsc = imax(mx, end_sc);
			tlog = tsc - (x < 2 ? 0 : (int32_t)(p->ie_coef * log2_approx((float)x) + .5f));
			if (tlog > max_log) {
				max_sc = tsc, max_log = tlog, max_i = i;
				memcpy(Hm, H, sizeof(int16_t) * (size_t)W);
			}
			tmp = H3, H3 = H2, H2 = H1, H1 = H, H = tmp;
			tmp = D3, D3 = D2, D2 = D1, D1 = D, D = tmp;
			if (max_log - tlog > p->xdrop) break;
		} else {
			tmp = H3, H3 = H2, H2 = H1, H1 = H, H = tmp;
			tmp = D3, D3 = D2, D2 = D1, D1 = D, D = tmp;
		}
	}
	if (is_ext) { /* nasw-sse.c:435-443 */
		for (j = 0; j < al; ++j) {
			int32_t sc = Hm[j];
			if (j == al - 1) sc += p->end_bonus;
			if (sc == max_sc) break;
		}
		/* the reference asserts j < al here; report aa_len = al+1 instead of aborting */
		r->nt_len = max_i + 1, r->aa_len = j + 1, r->score = max_sc;
	} else r->score = al > 0 ? H1[al-1] : NEG;
	if (tb) {
		if (ora_nasw_dbg_tb) memcpy(ora_nasw_dbg_tb, tb, sizeof(uint16_t) * (size_t)nl * (size_t)W);
		r->cigar = 0, r->m_cigar = 0;
		backtrack(tb, W, nl, al, r);
		free(tb);
	}
	free(buf); free(nas); free(aas); free(don); free(acc);
}
