// nasw_pair.cuh -- per-thread logic of the PAIR-LANE nasw kernels ("v4"), host + device.
//
// Two protein columns per thread, packed as int16x2 in one 32-bit register, on Blackwell's packed DPX instructions
// (VIADDMNMX.S16x2 = max(a + b, c) per half, VIMNMX3.S16x2, VIMNMX.S16x2 with one predicate per half): a cell pair costs
// what one 32-bit cell costs in nasw_core.cuh, a warp covers 64 columns instead of 32, and the score that crosses from row
// to row goes through ONE instruction (see "critical path").  Same wavefront as the block-wide kernels: three nucleotide
// rows per macro-step, every column one macro-step (three rows) behind the column to its left.  Thread x owns columns 2x
// (low half) and 2x+1 (high half); at macro-step T the low half works on triple T - 2x (rows 3m+2 .. 3m+4), the high half on
// triple T - 2x - 1, i.e. on the rows its own low half finished one macro-step earlier.  What a half needs from the column to
// its left therefore is: low half <- the left thread's high half (warp shuffle), high half <- the thread's own low half, both
// from the previous macro-step; one byte permute per exchanged register builds the packed input.
//
// Value domain.  The reference computes in saturating int16 (floor -32768, nasw-sse.c:360-402).  Here every score is
// stored BIASED: v' = v + 16384, the floor of the packed "relu" forms (0) stands for -16384, and only H is clamped (once per
// cell, free in the last instruction).  Exactness argument: call a value LEGIT if it derives from the
// origin H(-1,-1) = 0 through the recurrences, JUNK if it derives from the "minus infinity" initial state only.  Every legit
// value is at least  L = Smin - go - ge*al - (io + dmax + amax) - 2*fs  (reach any cell by insertions along row 2, then
// one intron or two frameshifts down), every junk value at most  floor + Smax*al  (junk only grows through match scores along
// a diagonal), both in the reference (floor -32768) and here (floor -16384).  As long as  -16384 + Smax*al < L  and the
// largest legit score stays below 16383 - ge*al, every maximum and every comparison that involves a legit value has the same
// outcome in both, legit values are equal, and nothing that only junk decides is ever read (row maxima of an extension
// always contain a legit cell; the backtrack only visits cells whose winning candidate is legit).  The host checks the two
// inequalities per problem (pair_eligible) and sends the rare problem that violates them (al > ~1000) to the 32-bit kernels
// of nasw_core.cuh, which reproduce the floor itself.
//
// Critical path.  H(i,j) = max(T, H(i-1,j) - K_i) with K_i = min(io + donor[i-1] + acceptor[i], fs): the intron-open-and-close
// term and the frameshift term are the only ones that need the row above, both have the form "H(i-1,j) minus a row constant",
// so the prep kernel combines the constants and everything else (T) is computed off the chain.
#pragma once
#include <stdint.h>
#include "nasw_core.cuh"

namespace nsw {

constexpr int PAIR_BIAS = 16384;     // stored value = score + PAIR_BIAS
constexpr int PAIR_DEAD = -30000;    // "profile" of padding / dead columns: keeps the match term below the junk floor (no int16 wrap: stored values are >= 0)
constexpr int PAIR_CB = 10;          // column-code bits of the travelling row maximum (columns <= 1023)
constexpr int PAIR_MAX_AL = 1000;

// ---- packed int16x2 arithmetic (device: one DPX instruction each; host: the same semantics for the CPU emulation)
NSW_HD int imin(int a, int b) { return a < b ? a : b; }
NSW_HD uint32_t pk(int lo, int hi) { return (uint32_t)(uint16_t)(int16_t)lo | (uint32_t)(uint16_t)(int16_t)hi << 16; }
NSW_HD uint32_t pk2(int v) { return pk(v, v); }
NSW_HD int lo16(uint32_t x) { return (int)(int16_t)(x & 0xffff); }
NSW_HD int hi16(uint32_t x) { return (int)(int16_t)(x >> 16); }
#ifdef __CUDA_ARCH__
NSW_HD uint32_t vam(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }            // max(a + b, c)
NSW_HD uint32_t vam_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }  // max(a + b, c, 0)
NSW_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
NSW_HD uint32_t vmax(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
NSW_HD uint32_t vadd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
// prmt.b32 in its default mode: selector nibble = source byte 0..7 (a, then b); bit 3 of a nibble replicates the SIGN of that byte
// instead of copying it (PTX ISA).  __byte_perm() only promises the three low bits, hence the instruction itself.
NSW_HD uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}
// max per half; *ge_lo / *ge_hi = (a >= b) of that half (VIMNMX.S16x2 with predicate outputs)
NSW_HD uint32_t vbmax(uint32_t a, uint32_t b, bool *ge_hi, bool *ge_lo) { return __vibmax_s16x2(a, b, ge_hi, ge_lo); }
NSW_HD uint32_t vmaxu(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
#else
NSW_HD int w16(int v) { return (int)(int16_t)(uint16_t)(v & 0xffff); } // wrap to int16 like the hardware does
NSW_HD uint32_t vam(uint32_t a, uint32_t b, uint32_t c) { return pk(imax(w16(lo16(a) + lo16(b)), lo16(c)), imax(w16(hi16(a) + hi16(b)), hi16(c))); }
NSW_HD uint32_t vam_relu(uint32_t a, uint32_t b, uint32_t c) { return pk(imax(imax(w16(lo16(a) + lo16(b)), lo16(c)), 0), imax(imax(w16(hi16(a) + hi16(b)), hi16(c)), 0)); }
NSW_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return pk(imax(imax(lo16(a), lo16(b)), lo16(c)), imax(imax(hi16(a), hi16(b)), hi16(c))); }
NSW_HD uint32_t vmax(uint32_t a, uint32_t b) { return pk(imax(lo16(a), lo16(b)), imax(hi16(a), hi16(b))); }
NSW_HD uint32_t vadd(uint32_t a, uint32_t b) { return pk(w16(lo16(a) + lo16(b)), w16(hi16(a) + hi16(b))); }
NSW_HD uint32_t vbmax(uint32_t a, uint32_t b, bool *ge_hi, bool *ge_lo)
{
	*ge_lo = lo16(a) >= lo16(b), *ge_hi = hi16(a) >= hi16(b);
	return vmax(a, b);
}
NSW_HD uint32_t vmaxu(uint32_t a, uint32_t b)
{
	const uint32_t al = a & 0xffff, bl = b & 0xffff, ah = a >> 16, bh = b >> 16;
	return (al > bl ? al : bl) | (ah > bh ? ah : bh) << 16;
}
NSW_HD uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel)
{
	const uint64_t v = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int k = 0; k < 4; ++k) {
		const uint32_t s = sel >> (4 * k) & 0xf;
		uint32_t byte = (uint32_t)(v >> (8 * (s & 7))) & 0xff;
		if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00; // replicate the sign of the selected byte
		r |= byte << (8 * k);
	}
	return r;
}
#endif

// Constants of one problem in packed form.
struct PairPar {
	uint32_t ngo, nfs; // -go, -fs in both halves
	int go, ge, fs, end_bonus;
};

// May this problem run on the pair-lane kernels?  (value-domain argument in the header; smin / smax = extreme entries of the
// substitution matrix, dmax / amax = largest donor / acceptor penalty, all as the kernels will see them)
inline bool pair_eligible(int al, int go, int ge, int io, int fs, int end_bonus, int smin, int smax, int dmax, int amax)
{
	if (al > PAIR_MAX_AL || al < 1) return false;
	if (go < 0 || ge < 0 || io < 0 || fs < 0 || go > 2000 || ge > 200 || io > 2000 || fs > 2000 || dmax > 120 || amax > 120 || dmax < 0 || amax < 0) return false;
	if (smax < 0) smax = 0;
	if (smin > 0) smin = 0;
	const long long legit_lo = (long long)smin - go - (long long)ge * al - (io + dmax + amax) - 2LL * fs - 64;      // lowest legit score
	const long long junk_hi = -(long long)PAIR_BIAS + (long long)smax * al + (end_bonus > 0 ? end_bonus : 0) + 64;  // highest junk score
	const long long legit_hi = (long long)smax * al + (end_bonus > 0 ? end_bonus : 0) + (long long)ge * al + 64;    // highest legit score (+ shifted insertion chain)
	return junk_hi < legit_lo && legit_hi < 32767 - PAIR_BIAS;
}

// The extremes pair_eligible() wants, from the scoring matrix (22 x 22) and the six splice-model penalties as the kernels see them.
struct PairLimits { int smin, smax, dmax, amax; };
inline PairLimits pair_limits(const int8_t *mat, const int32_t *sp)
{
	PairLimits l;
	l.smin = 127, l.smax = -128;
	for (int k = 0; k < 484; ++k) l.smin = mat[k] < l.smin ? mat[k] : l.smin, l.smax = mat[k] > l.smax ? mat[k] : l.smax;
	auto mx = [](int a, int b) { return a > b ? a : b; };
	l.dmax = mx(mx(mx(sp[0], sp[1]), mx(sp[2], sp[3])), mx(sp[4], 0));          // nasw-sse.c:120-127
	l.amax = mx(mx(sp[0] + 3 * mx(sp[5], 0), sp[2]), mx(sp[3], 0));              // nasw-sse.c:128-137
	int dmin = sp[0];
	for (int k = 1; k < 6; ++k) dmin = sp[k] < dmin ? sp[k] : dmin;
	if (dmin < 0) l.dmax = 1 << 20; // negative splice penalties: not a case the value-domain argument covers
	return l;
}

// ---- row records -------------------------------------------------------------------------------------------------------------
// One record per TRIPLE of rows m (rows i0 = 3m+2 .. i0+2), holding what a thread needs at the macro-step in which its low half
// is on triple m and its high half on triple m-1: every entry is a pair (low half: row of triple m, high half: the row three
// above).  24 words = six 16-byte fields:
//   w[ 0.. 4]  PD(k) = -(io + donor[k]),  k = i0-1 .. i0+3     (row i uses PD(i-1), PD(i), PD(i+1): nasw-sse.c:373-392)
//   w[ 5.. 9]  PA(k) = -acceptor[k],      k = i0-2 .. i0+2     (row i uses PA(i), PA(i-2), PA(i-1))
//   w[10..12]  NK(i) = -min(io + donor[i-1] + acceptor[i], fs)   i = i0 .. i0+2
//   w[13..15]  NG(i) = -(gap extension of row i: fs on stop-codon rows, nasw-sse.c:263)
//   w[16..21]  profile byte offsets (NOT packed): low half row i0+r at w[16+2r], high half row i0+r-3 at w[17+2r]
//   w[22..23]  unused
constexpr int PAIR_REC_WORDS = 24;
struct PairRec { uint32_t w[PAIR_REC_WORDS]; };

// rw(k) = row word (row_pack) of row k, clamped by the caller; prof_stride = bytes between profile rows of consecutive amino acids
template <class RowWord>
NSW_HD PairRec make_pair_rec(const RowWord &rw, int m, int io, int ge, int fs, int prof_stride, int prof_hi_base)
{
	PairRec r;
	const int i0 = 3 * m + 2;
	auto don = [&](int k) { return row_don(rw(k)); };
	auto acc = [&](int k) { return row_acc(rw(k)); };
	auto nas = [&](int k) { return row_nas(rw(k)); };
	for (int k = 0; k < 5; ++k) {
		r.w[k] = pk(-(io + don(i0 - 1 + k)), -(io + don(i0 - 4 + k)));
		r.w[5 + k] = pk(-acc(i0 - 2 + k), -acc(i0 - 5 + k));
	}
	for (int q = 0; q < 3; ++q) {
		const int i = i0 + q, j = i - 3;
		r.w[10 + q] = pk(-imin(io + don(i - 1) + acc(i), fs), -imin(io + don(j - 1) + acc(j), fs));
		r.w[13 + q] = pk(-(nas(i) == 20 ? fs : ge), -(nas(j) == 20 ? fs : ge));
		r.w[16 + 2 * q] = (uint32_t)(nas(i) * prof_stride);
		r.w[17 + 2 * q] = (uint32_t)(nas(j) * prof_stride + prof_hi_base);
	}
	r.w[22] = r.w[23] = 0;
	return r;
}

// ---- geometry of a problem on the kernels: one warp = 32 column pairs = up to 64 padded columns
constexpr int PAIR_MAX_W8 = 64;
NSW_HD int pair_triples(int nl) { return nl > 2 ? (nl - 2 + 2) / 3 : 0; }
NSW_HD int pair_rec_slots(int nl) { return (pair_triples(nl) + 2) / 2 + 1; }                  // records 0 .. triples, per parity of the triple index
// Device layout of the records of one problem, in 16-byte fields: [parity of the triple index][block of 32 records][field 0..5][32 records].
// The lanes of a warp are on triples two apart, i.e. on CONSECUTIVE records of one parity: every load of a field reads (at most two
// runs of) consecutive 16-byte words, and the six fields of a record sit at constant 512-byte distances.
NSW_HD int pair_rec_blocks(int nl) { return (pair_rec_slots(nl) + 31) / 32; }
NSW_HD int64_t pair_rec_index(int nb, int m) { const int k = m >> 1; return (int64_t)(m & 1) * nb * 192 + (int64_t)(k >> 5) * 192 + (k & 31); } // field f: + 32 f
NSW_HD int pair_n_macro(int nl, int W8) { const int M = pair_triples(nl); return M > 0 ? (M + W8 + 2) & ~1 : 0; } // even: the loops are unrolled by two

// ---- geometry of one thread -------------------------------------------------------------------------------------------------
struct PairGeo {
	int x;          // thread index within the problem's first..last thread (column pair)
	int col;        // first column of the pair (2x + column base of the warp)
	int nl, al, W8; // rows, residues, padded columns (8 * ceil(al / 8), even)
	bool first;     // col == 0: the constant boundary is to its left
};

// One thread of the pair-lane wavefront, score-only (extension) form.  Env must provide
//   uint32_t prof(uint32_t byte_off)   one word of the packed profile in shared memory (the thread's own base already added)
// Everything lives in registers under compile-time slot rotation.
struct PairLane {
	// packed state of the two columns
	uint32_t H[3], D[3], A, B, C;   // H / D of the row three above (before a row) resp. of the row itself (after it), intron states
	// per-thread constants
	uint32_t kq, ngej;              // shifted insertion chain: q(j) = it(j) + ge*j = max(H(i,j-1) + kq_j, q(j-1)), kq_j = ge*j - go - ge; ngej = -ge*j
	int code_lo, code_hi;           // row-maximum candidate = H' * 2^PAIR_CB + code (column code + end bonus), 0 code for padding columns
	uint32_t xmask;                 // 0 for the first thread of the problem (nothing to its left), ~0 otherwise
	uint32_t sel;                   // byte-permute selector of the exchange (the first thread takes zeros for its low half)
	// outputs of the last macro-step: what the column pair to the right (and the own high half) need next
	uint32_t oH[3], oQ[3];
	int oXlo[3], oXhi[3];

	NSW_HD void init(const PairGeo &g, const PairPar &p)
	{
		for (int k = 0; k < 3; ++k) H[k] = D[k] = 0, oH[k] = oQ[k] = 0, oXlo[k] = oXhi[k] = 0;
		A = B = C = 0;
		const int j0 = g.col, j1 = g.col + 1;
		kq = pk(p.ge * j0 - p.go - p.ge, p.ge * j1 - p.go - p.ge);
		ngej = pk(-p.ge * j0, -p.ge * j1);
		const int cm = (1 << PAIR_CB) - 1;
		code_lo = (j0 < g.al ? cm - j0 : 0) + (j0 == g.al - 1 ? p.end_bonus : 0) * (1 << PAIR_CB);
		code_hi = (j1 < g.al ? cm - j1 : 0) + (j1 == g.al - 1 ? p.end_bonus : 0) * (1 << PAIR_CB);
		xmask = g.first ? 0u : 0xffffffffu;
		// low half <- high half of the left thread's register (bytes 2,3 of operand a), high half <- low half of the own register
		// (bytes 0,1 of operand b = selector values 4,5).  First thread: low half <- sign replication of a non-negative byte = 0.
		sel = g.first ? 0x54bbu : 0x5432u;
	}

	// the packed left-neighbour register of one exchanged value: `shuffled` = the left thread's register, `own` = this thread's
	NSW_HD uint32_t left_of(uint32_t shuffled, uint32_t own) const { return bperm(shuffled, own, sel); }

	// One row of the pair, straight-line (both halves real rows).  R = position of the row in the macro-step (slot rotation).
	//   l0      H of the left columns for this row          l1..l3  for the rows 1..3 above
	//   lq      shifted insertion chain of the left columns  lx_lo   travelling row maximum left of the low column (already masked)
	//   rec     the pair record of this macro-step            xlo_prev  the own low column's row maximum for the row the HIGH half is on
	template <int R, class Env>
	NSW_HD void row(const PairPar &p, const PairRec &rec, const Env &env, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t lq, int lx_lo, int xlo_prev)
	{
		constexpr int h1 = (R + 2) % 3, h2 = (R + 1) % 3;
		const uint32_t ncA = rec.w[R], ncB = rec.w[R + 1], ncC = rec.w[R + 2];      // PD(i-1), PD(i), PD(i+1)
		const uint32_t naA = rec.w[5 + R + 2], naB = rec.w[5 + R], naC = rec.w[5 + R + 1]; // PA(i), PA(i-2), PA(i-1)
		const uint32_t s_lo = env.prof(rec.w[16 + 2 * R]), s_hi = env.prof(rec.w[17 + 2 * R]);
		uint32_t m = vadd(vadd(l3, s_lo), s_hi);                    // match: H(i-3, j-1) + s
		const uint32_t q = vam(l0, kq, lq);                         // insertion chain (shifted by ge*j)
		const uint32_t dn = vadd(vam(H[R], p.ngo, D[R]), rec.w[13 + R]); // deletion: max(H(i-3) - go, D(i-3)) - gei
		const uint32_t an = vam(H[h1], ncA, A);                     // intron states (phase 0 from the row above, 1 and 2 from the left column)
		const uint32_t bn = vam(l1, ncB, B);
		const uint32_t cn = vam(l1, ncC, C);
		const uint32_t f = vmax3(H[h2], l1, l2);                    // frameshifts from (i-2,j), (i-1,j-1), (i-2,j-1); (i-1,j) is inside NK
		uint32_t c1 = vam(q, ngej, m);
		uint32_t c2 = vam(A, naA, dn);                              // the OLD phase-0 state closes here; the new one is inside NK
		c1 = vam(bn, naB, c1);
		c2 = vam(cn, naC, c2);
		c1 = vam(f, p.nfs, c1);
		const uint32_t h = vam_relu(H[h1], rec.w[10 + R], vmax(c1, c2)); // the one instruction on the row-to-row chain
		H[R] = h, D[R] = dn, A = an, B = bn, C = cn;
		oH[R] = h, oQ[R] = q;
		// travelling row maximum with the column of its first occurrence: 32-bit, one column after the other
		const int vlo = (int)(h & 0xffffu), vhi = (int)(h >> 16);
		oXlo[R] = imax(lx_lo, vlo * (1 << PAIR_CB) + code_lo);
		oXhi[R] = imax(xlo_prev, vhi * (1 << PAIR_CB) + code_hi);
	}

	// A row in which the halves are not both real (ramp-up, ramp-down, tiny problems): same arithmetic, results committed per
	// half.  keep = 0xffff per half whose row is real.  bnd: the thread is the first of the problem and its low half is on row 2,
	// the one row that sees the boundary values H(-1,-1) = 0, H(0,-1) = H(1,-1) = -fs (nasw-sse.c:253-258).
	template <int R, class Env>
	NSW_HD void row_masked(const PairPar &p, const PairRec &rec, const Env &env, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t lq, int lx_lo, int xlo_prev,
	                       uint32_t keep, bool bnd)
	{
		if (bnd) {
			l3 = (l3 & 0xffff0000u) | (uint32_t)(uint16_t)(PAIR_BIAS);
			l2 = (l2 & 0xffff0000u) | (uint32_t)(uint16_t)(PAIR_BIAS - p.fs);
			l1 = (l1 & 0xffff0000u) | (uint32_t)(uint16_t)(PAIR_BIAS - p.fs);
		}
		const uint32_t sH = H[R], sD = D[R], sA = A, sB = B, sC = C, sOH = oH[R], sOQ = oQ[R];
		const int sXlo = oXlo[R], sXhi = oXhi[R];
		row<R>(p, rec, env, l0, l1, l2, l3, lq, lx_lo, xlo_prev);
		H[R] = (H[R] & keep) | (sH & ~keep), D[R] = (D[R] & keep) | (sD & ~keep);
		A = (A & keep) | (sA & ~keep), B = (B & keep) | (sB & ~keep), C = (C & keep) | (sC & ~keep);
		oH[R] = (oH[R] & keep) | (sOH & ~keep), oQ[R] = (oQ[R] & keep) | (sOQ & ~keep);
		if (!(keep & 0xffffu)) oXlo[R] = sXlo;
		if (!(keep >> 16)) oXhi[R] = sXhi;
	}
};

// The same thread in traceback (global alignment) form: every comparison of the reference's first pass (nasw-sse.c:448-520)
// is one packed max with a predicate per half; a false predicate ("the new candidate is strictly larger") raises the state
// nibble to the candidate's code resp. leaves an "extension" bit clear.  The state codes grow in evaluation order, so "the
// last candidate that beat the running maximum" is a maximum over codes.  Extension bits are collected INVERTED (a set bit =
// "the freshly opened gap / intron won or tied") and flipped once when the word is formed.  The two 16-bit words of a pair sit
// in one register and leave with one 32-bit store (wavefront-major layout of nasw_kernels.cu).
//   exchanged values: H, Q = true insertion chain (unshifted here), F = first-pass H, S = first-pass (segment-restarted) insertion chain
struct PairLaneTb {
	uint32_t H[3], D[3], A, B, C;
	uint32_t nge;
	uint32_t segmask;      // 0 in a half whose column starts a stripe segment of the reference layout (column % slen == 0): its
	                       // first-pass insertion chain starts from "minus infinity" (nasw-sse.c:455-457)
	uint32_t sel;
	uint32_t oH[3], oQ[3], oF[3], oS[3];

	NSW_HD void init(const PairGeo &g, const PairPar &p)
	{
		for (int k = 0; k < 3; ++k) H[k] = D[k] = 0, oH[k] = oQ[k] = oF[k] = oS[k] = 0;
		A = B = C = 0;
		const int j0 = g.col, j1 = g.col + 1, slen = g.W8 / 8;
		nge = pk2(-p.ge);
		segmask = (slen > 0 && j0 % slen == 0 ? 0u : 0xffffu) | (slen > 0 && j1 % slen == 0 ? 0u : 0xffff0000u);
		sel = g.first ? 0x54bbu : 0x5432u;
	}
	NSW_HD uint32_t left_of(uint32_t shuffled, uint32_t own) const { return bperm(shuffled, own, sel); }

	// returns the two traceback words of the pair (low half = low column)
	template <int R, class Env>
	NSW_HD uint32_t row(const PairPar &p, const PairRec &rec, const Env &env, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t lq, uint32_t lf, uint32_t ls)
	{
		constexpr int h1 = (R + 2) % 3, h2 = (R + 1) % 3;
		const uint32_t ncA = rec.w[R], ncB = rec.w[R + 1], ncC = rec.w[R + 2];
		const uint32_t naA = rec.w[5 + R + 2], naB = rec.w[5 + R], naC = rec.w[5 + R + 1];
		const uint32_t s_lo = env.prof(rec.w[16 + 2 * R]), s_hi = env.prof(rec.w[17 + 2 * R]);
		bool ph, pl;
		uint32_t Y = 0, Z = 0; // state nibbles (bits 0..3 of each half), inverted extension bits (bits 4..9)
#define NSW_PAIR_STATE(k) { if (!pl) Y = vmaxu(Y, (uint32_t)(k)); if (!ph) Y = vmaxu(Y, (uint32_t)(k) << 16); }
#define NSW_PAIR_EXT(b) { if (pl) Z |= 1u << (b); if (ph) Z |= 1u << ((b) + 16); }
		// every add below is the reference's saturating add/sub with the floor at the stored 0: junk values then sit at the same
		// offsets from the floor as in the reference, so even comparisons between two junk values come out the same
		uint32_t h = vam_relu(vadd(l3, s_lo), s_hi, 0u), t, u;                   // match (state 0)
		t = vam_relu(lf & segmask, p.ngo, 0u);                                   // insertion, first pass (state 1, bit 4)
		u = vbmax(t, ls & segmask, &ph, &pl); NSW_PAIR_EXT(4)
		const uint32_t sn = vam_relu(u, nge, 0u);
		h = vbmax(h, sn, &ph, &pl); NSW_PAIR_STATE(1)
		t = vam_relu(H[R], p.ngo, 0u);                                           // deletion (state 2, bit 5)
		u = vbmax(t, D[R], &ph, &pl); NSW_PAIR_EXT(5)
		const uint32_t dn = vam_relu(u, rec.w[13 + R], 0u);
		h = vbmax(h, dn, &ph, &pl); NSW_PAIR_STATE(2)
		t = vam_relu(H[h1], ncA, 0u);                                            // phase-0 intron (state 3, bit 6)
		const uint32_t an = vbmax(t, A, &ph, &pl); NSW_PAIR_EXT(6)
		h = vbmax(h, vam_relu(an, naA, 0u), &ph, &pl); NSW_PAIR_STATE(3)
		t = vam_relu(l1, ncB, 0u);                                               // phase-1 intron (state 4, bit 7)
		const uint32_t bn = vbmax(t, B, &ph, &pl); NSW_PAIR_EXT(7)
		h = vbmax(h, vam_relu(bn, naB, 0u), &ph, &pl); NSW_PAIR_STATE(4)
		t = vam_relu(l1, ncC, 0u);                                               // phase-2 intron (state 5, bit 8)
		const uint32_t cn = vbmax(t, C, &ph, &pl); NSW_PAIR_EXT(8)
		h = vbmax(h, vam_relu(cn, naC, 0u), &ph, &pl); NSW_PAIR_STATE(5)
		h = vbmax(h, vam_relu(H[h1], p.nfs, 0u), &ph, &pl); NSW_PAIR_STATE(6)    // frameshifts (states 6..9)
		h = vbmax(h, vam_relu(H[h2], p.nfs, 0u), &ph, &pl); NSW_PAIR_STATE(7)
		h = vbmax(h, vam_relu(l1, p.nfs, 0u), &ph, &pl); NSW_PAIR_STATE(8)
		h = vbmax(h, vam_relu(l2, p.nfs, 0u), &ph, &pl); NSW_PAIR_STATE(9)
		const uint32_t q = vam_relu(vam_relu(l0, p.ngo, lq), nge, 0u);           // true insertion chain (lazy-F in closed form, bit 9)
		const uint32_t hf = vbmax(h, q, &ph, &pl); NSW_PAIR_EXT(9)
#undef NSW_PAIR_STATE
#undef NSW_PAIR_EXT
		H[R] = hf, D[R] = dn, A = an, B = bn, C = cn;
		oH[R] = hf, oQ[R] = q, oF[R] = h, oS[R] = sn;
		return (Z ^ 0x03f003f0u) | Y;
	}

	template <int R, class Env>
	NSW_HD uint32_t row_masked(const PairPar &p, const PairRec &rec, const Env &env, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t lq, uint32_t lf, uint32_t ls,
	                           uint32_t keep, bool bnd)
	{
		if (bnd) {
			l3 = (l3 & 0xffff0000u) | (uint32_t)(uint16_t)(PAIR_BIAS);
			l2 = (l2 & 0xffff0000u) | (uint32_t)(uint16_t)(PAIR_BIAS - p.fs);
			l1 = (l1 & 0xffff0000u) | (uint32_t)(uint16_t)(PAIR_BIAS - p.fs);
		}
		const uint32_t sH = H[R], sD = D[R], sA = A, sB = B, sC = C, sOH = oH[R], sOQ = oQ[R], sOF = oF[R], sOS = oS[R];
		const uint32_t w = row<R>(p, rec, env, l0, l1, l2, l3, lq, lf, ls);
		H[R] = (H[R] & keep) | (sH & ~keep), D[R] = (D[R] & keep) | (sD & ~keep);
		A = (A & keep) | (sA & ~keep), B = (B & keep) | (sB & ~keep), C = (C & keep) | (sC & ~keep);
		oH[R] = (oH[R] & keep) | (sOH & ~keep), oQ[R] = (oQ[R] & keep) | (sOQ & ~keep);
		oF[R] = (oF[R] & keep) | (sOF & ~keep), oS[R] = (oS[R] & keep) | (sOS & ~keep);
		return w;
	}
};

} // namespace nsw
