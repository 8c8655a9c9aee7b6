// paf.cpp -- PAF line writer with the cs tag, byte-compatible with the reference (format.c:102-187,333-358).
// Host code (output formatting is outside the accelerated path), but it must emit identical bytes
// because PAF equality is how the whole path is judged.
#include <ctype.h>
#include <stdio.h>
#include "internal.hpp"

namespace mpb {

void Str::reserve(int64_t extra)
{
	if (l + extra + 1 > m) {
		m = (l + extra + 1) * 3 / 2 + 64;
		s = (char*)realloc(s, (size_t)m);
	}
}
void Str::put(const char *p, int64_t n) { reserve(n); memcpy(s + l, p, (size_t)n); l += n; s[l] = 0; }
void Str::puti(int64_t v)
{
	char tmp[24];
	int n = 0;
	uint64_t x = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
	do { tmp[n++] = (char)('0' + x % 10); x /= 10; } while (x);
	if (v < 0) tmp[n++] = '-';
	reserve(n);
	while (n) s[l++] = tmp[--n];
	s[l] = 0;
}

static inline uint8_t codon_aa(uint8_t n1, uint8_t n2, uint8_t n3)
{
	return (n1 > 3 || n2 > 3 || n3 > 3) ? ns_tab_aa20[(uint8_t)'X'] : ns_tab_codon[n1 << 4 | n2 << 2 | n3];
}

// cs:Z: difference string (format.c:102-187): ":n" identical codons, "*acgX" substitution (codon + residue),
// "+XYZ" residues without codon, "-acg" bases without residue, "~gt123ag" intron with its 2+2 boundary bases
static void write_cs(Str &o, const mp_idx_t *mi, const char *aa /* from r->qs */, const mp_reg1_t *r)
{
	static const char lc[] = "acgtn";
	const mp_extra_t *e = r->p;
	if (!e) return;
	auto nt = [&](int64_t i) -> uint8_t { return nt_at_v(mi->nt, r->vid, r->vs + i); };
	int32_t nl = 0, al = 0;
	o.puts("cs:Z:");
	for (int32_t k = 0; k < e->n_cigar; ++k) {
		const int32_t op = (int32_t)(e->cigar[k] & 0xf), len = (int32_t)(e->cigar[k] >> 4);
		if (op == NS_CIGAR_M) {
			int32_t same = 0;
			for (int32_t l = 0; l < len; ++l) {
				const uint8_t b0 = nt(nl + l * 3), b1 = nt(nl + l * 3 + 1), b2 = nt(nl + l * 3 + 2);
				if (codon_aa(b0, b1, b2) != ns_tab_aa20[(uint8_t)aa[al + l]]) {
					if (same > 0) o.putc(':'), o.puti(same);
					o.putc('*'), o.putc(lc[b0]), o.putc(lc[b1]), o.putc(lc[b2]), o.putc((char)toupper(aa[al + l]));
					same = 0;
				} else ++same;
			}
			if (same > 0) o.putc(':'), o.puti(same);
			nl += len * 3, al += len;
		} else if (op == NS_CIGAR_I) {
			o.putc('+');
			for (int32_t j = 0; j < len; ++j) o.putc((char)toupper(aa[al + j]));
			al += len;
		} else if (op == NS_CIGAR_D || op == NS_CIGAR_F) {
			const int32_t n = op == NS_CIGAR_D ? len * 3 : len;
			o.putc('-');
			for (int32_t i = 0; i < n; ++i) o.putc(lc[nt(nl + i)]);
			nl += n;
		} else if (op == NS_CIGAR_G) {
			o.putc('*');
			for (int32_t i = 0; i < len; ++i) o.putc(lc[nt(nl + i)]);
			o.putc((char)toupper(aa[al]));
			nl += len, ++al;
		} else if (op == NS_CIGAR_N || op == NS_CIGAR_U || op == NS_CIGAR_V) {
			const int32_t head = op == NS_CIGAR_N ? 0 : op == NS_CIGAR_U ? 1 : 2, tail = head ? 3 - head : 0;
			if (head) {
				o.putc('*');
				for (int32_t i = 0; i < head; ++i) o.putc(lc[nt(nl + i)]);
				o.putc((char)toupper(aa[al]));
			}
			o.putc('~'), o.putc(lc[nt(nl + head)]), o.putc(lc[nt(nl + head + 1)]);
			o.puti(len - (head + tail));
			o.putc(lc[nt(nl + len - tail - 2)]), o.putc(lc[nt(nl + len - tail - 1)]);
			if (tail) {
				o.putc('-');
				for (int32_t i = 0; i < tail; ++i) o.putc(lc[nt(nl + len - tail + i)]);
			}
			if (head) ++al;
			nl += len;
		}
	}
}

void format_hit(Str &o, const mp_idx_t *mi, const mp_mapopt_t *opt, const char *qname, int32_t qlen, const char *qseq, const mp_reg1_t *r)
{
	if (opt->flag & (MP_F_GFF | MP_F_GTF)) o.puts("##PAF\t");
	o.puts(qname), o.putc('\t'), o.puti(qlen);
	if (!r) { o.puts("\t0\t0\t*\t*\t0\t0\t0\t0\t0\t0\n"); return; }
	const mp_ctg_t *c = &mi->nt->ctg[r->vid >> 1];
	o.putc('\t'), o.puti(r->qs), o.putc('\t'), o.puti(r->qe), o.putc('\t'), o.putc("+-"[r->vid & 1]), o.putc('\t');
	o.puts(c->name), o.putc('\t'), o.puti(c->len), o.putc('\t');
	if (r->vid & 1) o.puti(c->len - r->ve), o.putc('\t'), o.puti(c->len - r->vs);
	else o.puti(r->vs), o.putc('\t'), o.puti(r->ve);
	o.putc('\t');
	if (r->p) {
		const mp_extra_t *e = r->p;
		o.puti(e->n_iden * 3), o.putc('\t'), o.puti(e->blen), o.puts("\t0\tAS:i:"), o.puti(e->dp_score);
		o.puts("\tms:i:"), o.puti(e->dp_max), o.puts("\tnp:i:"), o.puti(e->n_plus), o.puts("\tfs:i:"), o.puti(e->n_fs);
		o.puts("\tst:i:"), o.puti(e->n_stop), o.puts("\tda:i:"), o.puti(e->dist_start), o.puts("\tdo:i:"), o.puti(e->dist_stop);
		o.puts("\tcg:Z:");
		for (int32_t k = 0; k < e->n_cigar; ++k) o.puti(e->cigar[k] >> 4), o.putc(NS_CIGAR_STR[e->cigar[k] & 0xf]);
	} else o.puti(r->chn_sc), o.putc('\t'), o.puti(r->chn_sc_ungap), o.putc('\t'), o.puti(r->cnt);
	if (!(opt->flag & MP_F_NO_CS)) {
		o.putc('\t');
		write_cs(o, mi, qseq + r->qs, r);
	}
	o.putc('\n');
}

} // namespace mpb
