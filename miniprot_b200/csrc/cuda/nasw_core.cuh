// nasw_core.cuh -- per-lane logic of the nasw DP kernels, written once as host+device code.
//
// The kernels (nasw_kernels.cu) run an anti-diagonal wavefront: lane l of a warp owns C consecutive
// protein columns [l*C, l*C+C) and at step t works on nucleotide row i = t - l + 2, so every value it needs
// from the column to its left was produced by lane l-1 one step earlier and arrives through one warp
// shuffle.  Everything a lane does inside one step lives in the functions below; the kernel adds only the
// shuffles and the loop.  tests/hostcheck/emu_nasw.cpp compiles this same header for the CPU and steps
// 32 lanes in lockstep, which lets the CPU test-suite check the exact device arithmetic against the
// oracle without a GPU (test infrastructure only; the product never runs it).
//
// Semantics restated from the reference (nasw-sse.c:340-551):
//  * int16 saturating arithmetic: all values live in [-32768, 32767]; x - y is max(x - y, -32768).  The upper
//    bound is never reached (documented precondition nasw.h:111), the floor is reproduced exactly.
//  * the row is padded to W8 = 8*ceil(al/8) columns whose profile is -32768; they never feed real columns
//    but do take part in the extension row maximum.
//  * score-only mode needs only the true insertion chain It(j) = sat(max(sat(H(j-1)-go), It(j-1)) - ge).
//  * traceback mode additionally tracks the reference's FIRST-PASS values: the striped SSE kernel restarts the
//    insertion chain at every segment start (column % slen == 0) and its lazy-F loop then raises H and sets
//    bit 9; the state nibble and bit 4 of the traceback word come from the first pass.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NSW_HD __host__ __device__ __forceinline__
#else
#define NSW_HD inline
#endif

namespace nsw {

constexpr int NEG = -32768;

NSW_HD int imax(int a, int b) { return a > b ? a : b; }
NSW_HD int subs(int x, int y) { return imax(x - y, NEG); }           // _mm_subs_epi16 (floor only)
NSW_HD int adds(int x, int y) { return imax(x + y, NEG); }           // _mm_adds_epi16 (floor only)

// row word: everything row i contributes, produced by the prep kernel (nasw-sse.c:91-210)
//   bits 0..7  nas[i]   amino acid (0..21) of the codon ending at i
//   bits 8..15 donor[i] (int8)     bits 16..23 acceptor[i] (int8)
NSW_HD uint32_t row_pack(int nas, int don, int acc) { return (uint32_t)(nas & 0xff) | (uint32_t)(don & 0xff) << 8 | (uint32_t)(acc & 0xff) << 16; }
NSW_HD int row_nas(uint32_t w) { return (int)(w & 0xff); }
NSW_HD int row_don(uint32_t w) { return (int)(int8_t)(w >> 8 & 0xff); }
NSW_HD int row_acc(uint32_t w) { return (int)(int8_t)(w >> 16 & 0xff); }

struct Par {          // scalar parameters of one problem
	int go, ge, io, fs;
	int gei_stop;     // ge used on rows whose codon is a stop: fs (nasw-sse.c:263)
};

// Row record: everything row i contributes to the recurrences, with the constants already combined (written by the prep
// kernel, 32 B per row, read by the DP kernels with two 128-bit loads one step ahead):
//   cA = io + donor[i-1], cB = io + donor[i], cC = io + donor[i+1]   (A/B/C intron-open terms, nasw-sse.c:373-392)
//   aA = acceptor[i], aB = acceptor[i-2], aC = acceptor[i-1]         gei = fs on stop-codon rows else ge (nasw-sse.c:263)
// sat(sat(x - io) - d) == sat(x - (io + d)) because both penalties are >= 0 (floor clamping composes); the device path
// has no --spsc input, which is the only source of negative donor/acceptor values.
struct RowRec { int cA, cB, cC, gei, aA, aB, aC, nas; };

NSW_HD RowRec make_row_rec(const Par &p, uint32_t w_m2, uint32_t w_m1, uint32_t w_0, uint32_t w_p1)
{
	RowRec r;
	r.nas = row_nas(w_0);
	r.gei = r.nas == 20 ? p.fs : p.ge;
	r.cA = p.io + row_don(w_m1), r.cB = p.io + row_don(w_0), r.cC = p.io + row_don(w_p1);
	r.aA = row_acc(w_0), r.aC = row_acc(w_m1), r.aB = row_acc(w_m2);
	return r;
}

NSW_HD int imax3(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
	return __vimax3_s32(a, b, c);
#else
	return imax(imax(a, b), c);
#endif
}

// ------------------------------------------------------------------------------------------------
// score-only cell (extension mode), nasw-sse.c:355-404 + closed-form lazy-F.
//   in : h1,h2,h3 = H(i-1..i-3, j); d3 = D(i-3,j); a,b,c = running intron states of column j;
//        l0 = H(i,j-1) (final), l1,l2,l3 = H(i-1..i-3, j-1); it = It(i,j-1); s = profile(nas_i, j)
//   out: returns H(i,j); updates d_new, a, b, c, it (-> It(i,j))
// Every stored value stays >= -32768, so the per-operation floor of the SSE kernel is needed only where a value is
// carried (it, d_new) or first formed (the match term); max(h, x) absorbs it everywhere else.  17 integer ops.
// ------------------------------------------------------------------------------------------------
NSW_HD int cell_score(const Par &p, const RowRec &r, int s, int h1, int h2, int h3, int d3, int &d_new, int &a, int &b, int &c,
                      int l0, int l1, int l2, int l3, int &it)
{
	int h = adds(l3, s);
	it = subs(imax(l0 - p.go, it), p.ge);
	d_new = subs(imax(h3 - p.go, d3), r.gei);
	a = imax(h1 - r.cA, a);
	b = imax(l1 - r.cB, b);
	c = imax(l1 - r.cC, c);
	const int fg = imax(imax3(h1, h2, l1), l2) - p.fs;
	h = imax3(h, it, d_new);
	h = imax(h, a - r.aA);
	h = imax(h, b - r.aB);
	h = imax(h, c - r.aC);
	h = imax(h, fg);
	return h;
}

// ------------------------------------------------------------------------------------------------
// traceback cell, nasw-sse.c:448-520 + lazy-F (:521-537) in closed form.
//   extra in : f0 = first-pass H(i,j-1) and iseg = first-pass insertion chain at j-1 (both -32768 when
//              column j starts a stripe segment), it = true insertion chain
//   out      : hfirst (first-pass H(i,j)), returns final H(i,j), tb word (10 bits)
// Comparisons see exactly the saturated values the SSE kernel compares.
// ------------------------------------------------------------------------------------------------
NSW_HD int cell_trace(const Par &p, const RowRec &r, int s, int h1, int h2, int h3, int d3, int &d_new, int &a, int &b, int &c,
                      int l0, int f0, int l1, int l2, int l3, int &iseg, int &it, int &hfirst, uint32_t &word)
{
	uint32_t y = 0, z = 0;
	int h = adds(l3, s), t, u, v;
	t = subs(f0, p.go);
	if (iseg > t) z |= 1u << 4;
	iseg = subs(imax(t, iseg), p.ge);
	if (iseg > h) y = 1;
	h = imax(h, iseg);
	u = subs(h3, p.go), v = d3;
	if (v > u) z |= 1u << 5;
	t = subs(imax(u, v), r.gei);
	d_new = t;
	if (t > h) y = 2;
	h = imax(h, t);
	t = subs(h1, r.cA), v = a;
	if (v > t) z |= 1u << 6;
	a = imax(t, v);
	t = subs(a, r.aA);
	if (t > h) y = 3;
	h = imax(h, t);
	t = subs(l1, r.cB), v = b;
	if (v > t) z |= 1u << 7;
	b = imax(t, v);
	t = subs(b, r.aB);
	if (t > h) y = 4;
	h = imax(h, t);
	t = subs(l1, r.cC), v = c;
	if (v > t) z |= 1u << 8;
	c = imax(t, v);
	t = subs(c, r.aC);
	if (t > h) y = 5;
	h = imax(h, t);
	t = subs(h1, p.fs); if (t > h) y = 6; h = imax(h, t);
	t = subs(h2, p.fs); if (t > h) y = 7; h = imax(h, t);
	t = subs(l1, p.fs); if (t > h) y = 8; h = imax(h, t);
	t = subs(l2, p.fs); if (t > h) y = 9; h = imax(h, t);
	hfirst = h;
	it = subs(imax(subs(l0, p.go), it), p.ge);
	if (it > h) z |= 1u << 9, h = it;
	word = z | y;
	return h;
}

// nasw-sse.c:330-338 evaluated exactly like the x86-64 build does: every FP32 operation rounded on its own
// (no FMA contraction), then ie_coef*log2 + .5 truncated (nasw-sse.c:426)
NSW_HD int ext_len_penalty(float ie_coef, int x)
{
	if (x < 2) return 0;
	union { float f; uint32_t i; } z;
	z.f = (float)x;
	float lg = (float)((int)((z.i >> 23) & 255) - 128);
	z.i &= ~(255u << 23);
	z.i += 127u << 23;
#ifdef __CUDA_ARCH__
	float q = __fadd_rn(__fmul_rn(-0.34484843f, z.f), 2.02466578f);
	q = __fsub_rn(__fmul_rn(q, z.f), 0.67487759f);
	lg = __fadd_rn(lg, q);
	return (int)__fadd_rn(__fmul_rn(ie_coef, lg), .5f);
#else
	volatile float q = -0.34484843f * z.f;
	q = q + 2.02466578f;
	q = q * z.f;
	q = q - 0.67487759f;
	volatile float r = lg + q;
	volatile float m = ie_coef * r;
	return (int)(m + .5f);
#endif
}

// The length penalty pen(x) = (int)(ie_coef*log2(x) + .5) (x = i - 3*al) is a non-decreasing step function of x with a
// few dozen steps below 2^31, so the host tabulates where it steps (with the reference's exact FP32 arithmetic) and the
// device only advances an index: thr[k] = smallest x with pen(x) > pen(thr[k-1]), val[k] = pen(thr[k]).
constexpr int PEN_STEPS = 256;
struct PenTable { int32_t n; int32_t thr[PEN_STEPS]; int32_t val[PEN_STEPS]; };

inline void pen_table_build(float ie_coef, PenTable &t) // host only
{
	t.n = 0;
	int cur = 0;
	int64_t x = 2;
	while (x < 2147483647LL && t.n < PEN_STEPS) {
		// pen is monotone (the polynomial is increasing on [1,2) and steps up at octave boundaries): gallop + bisect
		int64_t lo = x, hi = x;
		while (hi < 2147483647LL && ext_len_penalty(ie_coef, (int)hi) <= cur) lo = hi, hi = hi * 2 < 2147483647LL ? hi * 2 : 2147483647LL;
		if (ext_len_penalty(ie_coef, (int)hi) <= cur) break;
		while (lo + 1 < hi) { const int64_t mid = (lo + hi) / 2; if (ext_len_penalty(ie_coef, (int)mid) > cur) hi = mid; else lo = mid; }
		if (ext_len_penalty(ie_coef, (int)lo) > cur) hi = lo;
		cur = ext_len_penalty(ie_coef, (int)hi);
		t.thr[t.n] = (int32_t)hi, t.val[t.n] = cur, ++t.n;
		x = hi + 1;
	}
}

// extension bookkeeping of one problem (nasw-sse.c:423-433): fed one finished row at a time
// The row maximum travels through the columns together with the column it was first seen in: best = H_adjusted * 2^cb +
// (2^cb - 1 - column), cb = code_bits(al): 12 bits as long as the columns fit (the int16 score then has 19 bits of room),
// 15 bits for extensions over up to 32767 residues (32767 * 2^15 still fits an int32).
NSW_HD int code_bits(int al) { return al <= 4095 ? 12 : 15; }
constexpr int CODE_MAX_AL = 32767;

struct ExtTracker {
	int max_sc, max_log, max_i, max_code;
	int pen, pk, next_thr; // current penalty, index of the next table step, and the x at which it applies
	int cb;                // width of the column code (code_bits)
	bool stopped;
	NSW_HD void init(int code_bits_ = 12) { max_sc = INT32_MIN, max_log = INT32_MIN, max_i = -1, max_code = 0, pen = 0, pk = 0, next_thr = 2, stopped = false, cb = code_bits_; }
	NSW_HD int aa_len(int al) const { return (max_i >= 0 && max_code != 0) ? ((1 << cb) - 1) - max_code + 1 : al + 1; }
	// best = max over columns of (H_adjusted << cb | (2^cb - 1 - column)), padding columns carry code 0
	NSW_HD void row(int i, int best, int pen_base /* 3*al */, const PenTable &pt, int xdrop)
	{
		const int x = i - pen_base;
		if (x >= next_thr) { // a few dozen times per problem
			while (pk < pt.n && x >= pt.thr[pk]) pen = pt.val[pk], ++pk;
			next_thr = pk < pt.n ? pt.thr[pk] : INT32_MAX;
		}
		const int tsc = best >> cb, tlog = tsc - pen;
		const bool better = !stopped && tlog > max_log;
		max_sc = better ? tsc : max_sc, max_i = better ? i : max_i, max_code = better ? (best & ((1 << cb) - 1)) : max_code;
		max_log = better ? tlog : max_log;
		stopped = stopped || max_log - tlog > xdrop;
	}
};

// ------------------------------------------------------------------------------------------------
// One lane of the wavefront.  The kernels keep one of these per thread (all arrays live in registers) and
// call step() once per wavefront step after fetching the left lane's outputs with warp shuffles.
//   Env must provide:  RowRec row_rec(int i)               record of row i (index clamped to [0,nl]); prefetched a step ahead
//                      const int *profile(int nas)          profile row of amino acid `nas` for THIS lane's columns
//                      carry load/store for multi-pass problems (see kernels)
// MULTI = the problem is wider than one pass (32*C columns): lane 0 of pass > 0 reads the previous pass's last
// column from the carry array and lane 31 of every pass but the last writes it.  Single-pass instantiations carry
// none of that code.  Lane-0 boundary handling and the extension tracker are branch-free on purpose: a warp pays
// for every divergent side path on every step, and one step is the unit of the critical path (nl steps).
// ------------------------------------------------------------------------------------------------
struct LaneGeom {            // where this lane sits in the problem
	int lane, pass, n_pass, nl, al, W8, col0;
	bool live;               // owns real or padding columns (col0 < W8)
};

// The rolling rows (H of rows i-1..i-3, D of row i-3, the left column's history) and the prefetched row records are
// kept in small arrays indexed by the step PHASE P = t mod 6 known at compile time (the kernels unroll the step loop by
// six), so "rotating" them is register renaming instead of ~25 moves per step.
template <int C, bool MULTI>
struct ExtLane {
	int Hr[3][C], Dr[3][C], A[C], B[C], Cc[C];
	int code[C], bonus[C], cmul;
	int Lr[3];
	int outH, outI, outB;    // what the lane to the right receives next step
	RowRec rec[2];           // records of the rows of this step (phase parity) and the next

	template <class Env>
	NSW_HD void init(const LaneGeom &g, int end_bonus, int fs, const Env &env)
	{
#pragma unroll
		for (int k = 0; k < C; ++k) {
			const int jg = g.col0 + k;
			Hr[0][k] = Hr[1][k] = Hr[2][k] = Dr[0][k] = Dr[1][k] = Dr[2][k] = A[k] = B[k] = Cc[k] = NEG;
			code[k] = jg < g.al ? ((1 << code_bits(g.al)) - 1) - jg : 0;
			bonus[k] = jg == g.al - 1 ? end_bonus : 0;
		}
		cmul = 1 << code_bits(g.al);
		Lr[0] = Lr[1] = Lr[2] = NEG;
		if (g.lane == 0 && g.pass == 0) Lr[0] = 0, Lr[1] = Lr[2] = -fs; // H(-1,-1) = 0, H(0,-1) = H(1,-1) = -fs, seen by row 2 only
		outH = outI = NEG, outB = INT32_MIN;
		rec[0] = env.row_rec(2 - g.lane), rec[1] = env.row_rec(3 - g.lane);
	}
	// the boundary column is -32768 for every row after the first (nasw-sse.c:266-270): call once after step t = 0
	NSW_HD void after_first_step(const LaneGeom &g) { if (g.lane == 0 && g.pass == 0) Lr[0] = Lr[1] = Lr[2] = NEG; }

	// P = t mod 6.  rH/rI/rB: outputs of the left lane from the previous step (ignored by lane 0 of pass 0).
	// Returns the row this lane just finished (or -1); outB then holds the best value of that row over all columns up
	// to and including this lane's -- complete in lane 31 of the last pass, which feeds the ExtTracker.
	template <int P, class Env>
	NSW_HD int step(const LaneGeom &g, const Par &par, int t, int rH, int rI, int rB, Env &env)
	{
		constexpr int h3 = P % 3, h2 = (P + 1) % 3, h1 = (P + 2) % 3, rp = P % 2; // slot h3 holds row i-3 and receives row i
		const int i = t - g.lane + 2;
		const RowRec rc = rec[rp];
		rec[rp] = env.row_rec(i + 2);
		if (g.lane == 0) env.prefetch_row(i + 24); // lane 0 is the first to touch a row: pull its line towards L1 early
		const bool row_ok = i >= 2 && i < g.nl;
		if (g.lane == 0) { // boundary column -1 (nasw-sse.c:253-271); its row history is preset by init()/after_first_step()
			if (!MULTI || g.pass == 0) rH = NEG, rI = NEG, rB = INT32_MIN;
			else if (row_ok) env.carry_load3(i, rH, rI, rB);
		}
		if (!row_ok) return -1;
		if (g.live) {
			const int *ps = env.profile(rc.nas);
			int l0 = rH, l1 = Lr[h1], l2 = Lr[h2], l3 = Lr[h3], it = rI, best = rB;
			int hn[C], dn[C];
#pragma unroll
			for (int k = 0; k < C; ++k) {
				hn[k] = cell_score(par, rc, ps[k], Hr[h1][k], Hr[h2][k], Hr[h3][k], Dr[h3][k], dn[k], A[k], B[k], Cc[k], l0, l1, l2, l3, it);
				best = imax(best, (hn[k] + bonus[k]) * cmul + code[k]);
				l0 = hn[k], l1 = Hr[h1][k], l2 = Hr[h2][k], l3 = Hr[h3][k];
			}
#pragma unroll
			for (int k = 0; k < C; ++k) Hr[h3][k] = hn[k], Dr[h3][k] = dn[k];
			outH = hn[C - 1], outI = it, outB = best;
		} else outB = rB;
		if (g.lane != 0 || (MULTI && g.pass > 0)) Lr[h3] = rH;
		if (MULTI && g.lane == 31 && g.pass < g.n_pass - 1) env.carry_store3(i, outH, outI, outB);
		return i;
	}
};

template <int C, bool MULTI>
struct TbLane {
	int Hr[3][C], Dr[3][C], A[C], B[C], Cc[C];
	int Lr[3];
	int outH, outF, outS, outI;
	RowRec rec[2];
	uint32_t seg_start;      // bit k: column col0+k starts a stripe segment of the reference layout
	int k_end;               // which of my columns is al-1 (or -1)
	int score;               // H(nl-1, al-1) once seen

	template <class Env>
	NSW_HD void init(const LaneGeom &g, int fs, const Env &env)
	{
		const int slen = g.W8 / 8;
		seg_start = 0, k_end = -1, score = NEG;
#pragma unroll
		for (int k = 0; k < C; ++k) {
			Hr[0][k] = Hr[1][k] = Hr[2][k] = Dr[0][k] = Dr[1][k] = Dr[2][k] = A[k] = B[k] = Cc[k] = NEG;
			if (slen > 0 && (g.col0 + k) % slen == 0) seg_start |= 1u << k;
			if (g.col0 + k == g.al - 1) k_end = k;
		}
		Lr[0] = Lr[1] = Lr[2] = NEG;
		if (g.lane == 0 && g.pass == 0) Lr[0] = 0, Lr[1] = Lr[2] = -fs;
		outH = outF = outS = outI = NEG;
		rec[0] = env.row_rec(2 - g.lane), rec[1] = env.row_rec(3 - g.lane);
	}
	NSW_HD void after_first_step(const LaneGeom &g) { if (g.lane == 0 && g.pass == 0) Lr[0] = Lr[1] = Lr[2] = NEG; }

	// wd[] receives the C traceback words when the function returns true
	template <int P, class Env>
	NSW_HD bool step(const LaneGeom &g, const Par &par, int t, int rH, int rF, int rS, int rI, Env &env, uint32_t *wd)
	{
		constexpr int h3 = P % 3, h2 = (P + 1) % 3, h1 = (P + 2) % 3, rp = P % 2;
		const int i = t - g.lane + 2;
		const RowRec rc = rec[rp];
		rec[rp] = env.row_rec(i + 2);
		if (g.lane == 0) env.prefetch_row(i + 24);
		const bool row_ok = i >= 2 && i < g.nl;
		if (g.lane == 0) {
			if (!MULTI || g.pass == 0) rH = rF = rS = rI = NEG;
			else if (row_ok) env.carry_load4(i, rH, rF, rS, rI);
		}
		if (!row_ok) return false;
		bool wrote = false;
		if (g.live) {
			const int *ps = env.profile(rc.nas);
			int l0 = rH, f0 = rF, l1 = Lr[h1], l2 = Lr[h2], l3 = Lr[h3], iseg = rS, it = rI;
			int hn[C], dn[C];
#pragma unroll
			for (int k = 0; k < C; ++k) {
				if (seg_start >> k & 1) f0 = NEG, iseg = NEG;
				int hf;
				hn[k] = cell_trace(par, rc, ps[k], Hr[h1][k], Hr[h2][k], Hr[h3][k], Dr[h3][k], dn[k], A[k], B[k], Cc[k], l0, f0, l1, l2, l3, iseg, it, hf, wd[k]);
				l0 = hn[k], f0 = hf, l1 = Hr[h1][k], l2 = Hr[h2][k], l3 = Hr[h3][k];
			}
			if (i == g.nl - 1 && k_end >= 0) {
#pragma unroll
				for (int k = 0; k < C; ++k) if (k == k_end) score = hn[k];
			}
#pragma unroll
			for (int k = 0; k < C; ++k) Hr[h3][k] = hn[k], Dr[h3][k] = dn[k];
			outH = hn[C - 1], outF = f0, outS = iseg, outI = it;
			wrote = true;
		}
		if (g.lane != 0 || (MULTI && g.pass > 0)) Lr[h3] = rH;
		if (MULTI && g.lane == 31 && g.pass < g.n_pass - 1) env.carry_store4(i, outH, outF, outS, outI);
		return wrote;
	}
};

// ------------------------------------------------------------------------------------------------
// Block-wide wavefront ("v3"): one THREAD per protein column, three nucleotide rows per step.
//
// The number of wavefront steps of a problem is fixed by its nucleotide length (100 k rows for a default extension
// window), and a step costs a few hundred cycles of mostly fixed overhead (neighbour exchange, row-record fetch, loop,
// tracker) however few columns a lane owns.  So the critical path is shortest when every thread owns ONE column and
// amortises the overhead over several rows: thread x of a CTA of NW warps handles column x and, at macro-step T, the
// rows 3(T-x)+2 .. 3(T-x)+4.  Three is the period of the H/D row rotation, so row r of a macro-step always lives in
// slot r (no moves).  Thread x receives what thread x-1 produced for the same three rows one macro-step earlier: by
// warp shuffle inside a warp, through a double-buffered shared-memory slot across warps (one __syncthreads per
// macro-step).  Problems of up to 32*NW (<= 256) padded columns run this way; wider ones keep the column-pass kernels.
// ------------------------------------------------------------------------------------------------
// x = column within the pass (sets the skew of the wavefront), col = column of the problem (pass * Wp + x; problems wider than
// one block run in passes of Wp columns); live = col < W8; first = col == 0 (the constant boundary is to its left)
struct Geo3 { int x, col, nl, al, W8; bool live, first; };

// Row records of a block-wide problem are stored per TRIPLE of rows (triple m = rows 3m+2 .. 3m+4, the unit of a macro-step)
// and field-major: six arrays of M 16-byte fields (record halves a/b of the three rows).  Column x works on triple T - x at
// macro-step T, so the 32 lanes of a warp read 32 CONSECUTIVE fields with each load (512 contiguous bytes) instead of 32
// records 96 bytes apart.  M covers every real row (2 .. nl-1) plus the look-ahead of the pipeline.
NSW_HD int v3_triples(int nl) { return (nl > 2 ? (nl - 2 + 2) / 3 : 0) + 4; }

template <bool TB>
struct Lane3 {
	int H[3], D[3], A, B, Cc, L[3];
	RowRec rec[6];
	int oH[3], oI[3], oX[3], oS[3]; // per row of the macro-step: final H, insertion chain, (ext) running best / (tb) first-pass H, (tb) segment chain
	int code, cmul, score; // code = bonus * cmul + column code: row maximum candidate = H * cmul + code
	bool seg_start, end_col;

	NSW_HD static int row_of(const Geo3 &g, int T, int r) { return 3 * (T - g.x) + 2 + r; }

	template <class Env>
	NSW_HD void init(const Geo3 &g, int end_bonus, int fs, const Env &env)
	{
		const int slen = g.W8 / 8;
		for (int k = 0; k < 3; ++k) H[k] = D[k] = L[k] = NEG, oH[k] = oI[k] = oS[k] = NEG, oX[k] = TB ? NEG : INT32_MIN;
		A = B = Cc = NEG;
		if (g.first) L[0] = 0, L[1] = L[2] = -fs; // H(-1,-1), H(0,-1), H(1,-1): seen by row 2 only (nasw-sse.c:253-258)
		cmul = 1 << code_bits(g.al);
		code = (g.col < g.al ? (cmul - 1) - g.col : 0) + (g.col == g.al - 1 ? end_bonus : 0) * cmul;
		seg_start = slen > 0 && g.col % slen == 0, end_col = g.col == g.al - 1, score = NEG;
		env.rec3(-g.x, rec[0], rec[1], rec[2]), env.rec3(1 - g.x, rec[3], rec[4], rec[5]);
	}

	// One real row of a column, straight-line (no validity checks): rows 0, 1, 2 of a steady-state macro-step in this order.
	// l0 = H(i, j-1); l1..l3 = H(i-1..i-3, j-1); ri / rx / rs = insertion chain, running row maximum (first-pass H in
	// traceback mode) and segment chain of the column to the left for this row (the first column gets the boundary values).
	// Dead columns (x >= W8) run the same arithmetic on their all-NEG profile columns -- nothing to their right is real, so
	// whatever they produce is never used -- and only hand the row maxima on: no divergent branch inside the warp.
	template <int R>
	NSW_HD void row(const Par &par, const RowRec &rc, bool live, int l0, int l1, int l2, int l3, int ri, int rx, int rs, const int *ps, int W, uint32_t &wd)
	{
		constexpr int h3 = R, h2 = (R + 1) % 3, h1 = (R + 2) % 3; // R = 0: (2,1,0), 1: (0,2,1), 2: (1,0,2)
		const int s = ps[rc.nas * W];
		int d_new, it = ri;
		if (TB) {
			int f0 = seg_start ? NEG : rx, iseg = seg_start ? NEG : rs, hf;
			const int h = cell_trace(par, rc, s, H[h1], H[h2], H[h3], D[h3], d_new, A, B, Cc, l0, f0, l1, l2, l3, iseg, it, hf, wd);
			H[h3] = h, D[h3] = d_new, oH[R] = h, oI[R] = it, oX[R] = hf, oS[R] = iseg;
		} else {
			const int h = cell_score(par, rc, s, H[h1], H[h2], H[h3], D[h3], d_new, A, B, Cc, l0, l1, l2, l3, it);
			H[h3] = h, D[h3] = d_new, oH[R] = h, oI[R] = it;
			const int m = imax(rx, h * cmul + code);
			oX[R] = live ? m : rx;
		}
	}
	// after the three rows: fetch the records of step T+2 through the environment's running cursor (AFTER the rows that used
	// the old ones: no register copies)
	template <int PH, class Env>
	NSW_HD void steady_tail(const Geo3 &g, Env &env)
	{
		env.next3(rec[3 * PH], rec[3 * PH + 1], rec[3 * PH + 2]);
		if (g.x < 6) env.prefetch_ahead(g.x);
	}

	// Steady-state macro-step: EVERY thread of the block has three real rows (the kernel guarantees T is in that range), so
	// there is nothing to check.  pH[r] / rH[r]: H of the column to the left for row r of the previous / of this macro-step
	// (two alternating buffers of the caller, which replace L[]).  The kernel calls the three rows itself so that it can
	// send each row's outputs to the right as soon as they exist; this form is the reference sequence (and what the CPU
	// emulation steps through).
	template <int PH, class Env>
	NSW_HD void macro_steady(const Geo3 &g, const Par &par, const int *pH, const int *rH, const int *rI, const int *rX, const int *rS, Env &env, uint32_t *wd)
	{
		const int *ps = env.profile(0);
		const int W = env.profile_stride();
		row<0>(par, rec[3 * PH], g.live, rH[0], pH[2], pH[1], pH[0], rI[0], rX[0], rS[0], ps, W, wd[0]);
		row<1>(par, rec[3 * PH + 1], g.live, rH[1], rH[0], pH[2], pH[1], rI[1], rX[1], rS[1], ps, W, wd[1]);
		row<2>(par, rec[3 * PH + 2], g.live, rH[2], rH[1], rH[0], pH[2], rI[2], rX[2], rS[2], ps, W, wd[2]);
		steady_tail<PH>(g, env);
	}
	// entering the steady range at macro-step T: hand L[] over as the first "previous" buffer and point the record cursor
	// at the rows step T will fetch
	template <class Env>
	NSW_HD void steady_enter(const Geo3 &g, int T, int *pH, Env &env) const
	{
		pH[0] = L[0], pH[1] = L[1], pH[2] = L[2];
		env.seek3(T - g.x + 2);
	}
	NSW_HD void steady_leave(const int *pH) { L[0] = pH[0], L[1] = pH[1], L[2] = pH[2]; }
	// macro-steps [lo, hi) are steady for a block of Wp columns: every column has three rows inside [3, nl - 1) (lo, hi even)
	NSW_HD static void steady_range(int nl, int Wp, int &lo, int &hi)
	{
		lo = Wp;
		hi = nl >= 6 ? (nl - 6) / 3 + 1 : 0; // ... and below nl - 1: the last row (where the global score is read) stays in the general step
		if (hi < lo) hi = lo;
		hi = lo + ((hi - lo) & ~1);
	}

	// General macro-step (ramp-up, ramp-down, tiny problems).  PH = T mod 2 selects the half of the record buffer.
	// rH/rI/rX/rS[r]: outputs of thread x-1 for row r.  wd[r] receives the traceback word of row r (TB); returns a bit mask
	// of the rows that were real rows of this thread.
	template <int PH, class Env>
	NSW_HD uint32_t macro(const Geo3 &g, const Par &par, int T, const int *rH, const int *rI, const int *rX, const int *rS, Env &env, uint32_t *wd)
	{
		const int i0 = row_of(g, T, 0);
		const RowRec rc0 = rec[3 * PH], rc1 = rec[3 * PH + 1], rc2 = rec[3 * PH + 2];
		env.rec3(T - g.x + 2, rec[3 * PH], rec[3 * PH + 1], rec[3 * PH + 2]);
		uint32_t done = 0;
#pragma unroll
		for (int r = 0; r < 3; ++r) {
			const int h3 = r, h2 = (r + 1) % 3, h1 = (r + 2) % 3;
			const int i = i0 + r;
			const RowRec &rc = r == 0 ? rc0 : r == 1 ? rc1 : rc2;
			const bool row_ok = i >= 2 && i < g.nl;
			if (!row_ok) continue;
			int l0 = rH[r], it = rI[r], lx = rX[r], ls = TB ? rS[r] : 0;
			if (g.first) l0 = NEG, it = NEG, lx = TB ? NEG : INT32_MIN, ls = NEG;
			if (g.live) {
				const int s = env.profile(rc.nas)[0];
				int d_new;
				if (TB) {
					int f0 = lx, iseg = ls, hf;
					if (seg_start) f0 = NEG, iseg = NEG;
					const int h = cell_trace(par, rc, s, H[h1], H[h2], H[h3], D[h3], d_new, A, B, Cc, l0, f0, L[h1], L[h2], L[h3], iseg, it, hf, wd[r]);
					H[h3] = h, D[h3] = d_new;
					oH[r] = h, oI[r] = it, oX[r] = hf, oS[r] = iseg;
					if (end_col && i == g.nl - 1) score = h;
				} else {
					const int h = cell_score(par, rc, s, H[h1], H[h2], H[h3], D[h3], d_new, A, B, Cc, l0, L[h1], L[h2], L[h3], it);
					H[h3] = h, D[h3] = d_new;
					oH[r] = h, oI[r] = it, oX[r] = imax(lx, h * cmul + code);
				}
				done |= 1u << r;
			} else if (!TB) oX[r] = lx, done |= 1u << r; // dead columns only hand the row maximum on
			if (!g.first) L[h3] = rH[r];
			else if (i == 2) L[0] = L[1] = L[2] = NEG; // the boundary column is -32768 for every later row (nasw-sse.c:266-270)
		}
		return done;
	}
};

// ------------------------------------------------------------------------------------------------
// sequence preparation (nasw-sse.c:91-210) for one row, from a code accessor c(k) (k in [0,nl), values 0..4).
// FORWARD orientation is used for global alignment and right extension; the LEFT variant sees the slice
// already reversed (c(k) = original[nl-1-k], not complemented) and applies the mirrored rules.
// ------------------------------------------------------------------------------------------------
// --spsc (nasw-sse.c:138-152,189-203): s = splice byte of one nucleotide (ntseq.c:130-156; 0xff = no score, else
// (score + 64) << 1 | is_acceptor), applied to the int8 donor / acceptor entries of one row with the reference's int8 wrap.
// acc_if_odd: in the forward orientation an odd byte adjusts the acceptor, in the reversed slice of a left extension the donor.
struct SpscPar { int max_spsc, null_bonus; }; // (io + 1) / 2 - 1 of the problem, ns_opt_t::sp_null_bonus
NSW_HD void spsc_adjust(int s, const SpscPar &q, bool acc_if_odd, int &don, int &acc)
{
	if (s == 0xff) { don = (int)(int8_t)(don - q.null_bonus), acc = (int)(int8_t)(acc - q.null_bonus); return; }
	int v = (s >> 1) - 64;
	if (v > q.max_spsc) v = q.max_spsc;
	if (((s & 1) != 0) == acc_if_odd) acc = (int)(int8_t)(acc - v); else don = (int)(int8_t)(don - v);
}
struct NoSpsc { NSW_HD int operator()(int) const { return -1; } }; // byte of DP row k's nucleotide, -1 = the problem has no splice bytes

template <class Code, class Spsc>
NSW_HD uint32_t prep_row_forward(const Code &c, int nl, int i, const int *sp /*[6]*/, const uint8_t *codon_tab, int aa_x, const Spsc &ss, const SpscPar &sq)
{
	int don = sp[3], acc = sp[3], nas = aa_x;
	if (i < nl) {
		if (i < nl - 3) { // donor[i]
			int t = 3;
			const int c0 = c(i), c1 = c(i + 1), c2 = c(i + 2);
			if (c1 == 2 && c2 == 3) { const int c3 = c(i + 3); t = (c3 == 0 || c3 == 2) ? (c0 == 2 ? -1 : 4) : 0; } // i+3 < nl holds here
			else if (c1 == 2 && c2 == 1 && c0 == 2) t = 1;
			else if (c1 == 0 && c2 == 3) t = 2;
			don = (int)(int8_t)(t < 0 ? 0 : sp[t]);
		} else don = (int)(int8_t)sp[3];
		if (i >= 1) { // acceptor[i]
			int t = 3, pen = 0;
			const int cm1 = c(i - 1), c0 = c(i);
			if (cm1 == 0 && c0 == 2) {
				t = (i >= 2 && (c(i - 2) == 1 || c(i - 2) == 3)) ? -1 : 0;
				for (int j = i - 4; j >= 0 && j > i - 7; --j) { const int x = c(j); if (x != 1 && x != 3) pen += sp[5]; }
			} else if (cm1 == 0 && c0 == 1) t = 2;
			acc = (int)(int8_t)(t < 0 ? 0 : sp[t]);
			if (t == -1 || t == 0) acc = (int)(int8_t)(acc + pen);
		} else acc = (int)(int8_t)sp[3];
		if (i >= 2) {
			const int a = c(i - 2), b = c(i - 1), d = c(i);
			if (a < 4 && b < 4 && d < 4) nas = codon_tab[a << 4 | b << 2 | d];
		}
	} else don = (int)(int8_t)sp[3], acc = (int)(int8_t)sp[3];
	if (i + 1 < nl) { // nasw-sse.c:140-151: ss[i+1] belongs to donor[i] / acceptor[i]
		const int s = ss(i + 1);
		if (s >= 0) spsc_adjust(s, sq, true, don, acc);
	}
	return row_pack(nas, don, acc);
}
template <class Code>
NSW_HD uint32_t prep_row_forward(const Code &c, int nl, int i, const int *sp, const uint8_t *codon_tab, int aa_x)
{
	return prep_row_forward(c, nl, i, sp, codon_tab, aa_x, NoSpsc(), SpscPar{ 0, 0 });
}

template <class Code, class Spsc>
NSW_HD uint32_t prep_row_left(const Code &c, int nl, int i, const int *sp, const uint8_t *codon_tab, int aa_x, const Spsc &ss, const SpscPar &sq)
{
	int don = (int)(int8_t)sp[3], acc = (int)(int8_t)sp[3], nas = aa_x;
	if (i < nl) {
		if (i < nl - 3) { // "donor" of the reversed string = mirrored acceptor
			int t = 3, pen = 0;
			const int c1 = c(i + 1), c2 = c(i + 2);
			if (c1 == 2 && c2 == 0) {
				const int c3 = c(i + 3);
				t = (c3 == 1 || c3 == 3) ? -1 : 0;
				for (int j = i + 5; j < nl && j < i + 8; ++j) { const int x = c(j); if (x != 1 && x != 3) pen += sp[5]; }
			} else if (c1 == 1 && c2 == 0) t = 2;
			don = (int)(int8_t)(t < 0 ? 0 : sp[t]);
			if (t == -1 || t == 0) don = (int)(int8_t)(don + pen);
		}
		if (i >= 1) { // "acceptor" of the reversed string = mirrored donor
			int t = 3;
			const int cm1 = c(i - 1), c0 = c(i);
			if (cm1 == 3 && c0 == 2) t = (i >= 2 && (c(i - 2) == 0 || c(i - 2) == 2)) ? ((i + 1 < nl && c(i + 1) == 2) ? -1 : 4) : 0;
			else if (cm1 == 1 && c0 == 2 && i + 1 < nl && c(i + 1) == 1) t = 1;
			else if (cm1 == 3 && c0 == 0) t = 2;
			acc = (int)(int8_t)(t < 0 ? 0 : sp[t]);
		}
		if (i >= 2) { // codon read in the original direction: original (p-2,p-1,p) = reversed (i, i-1, i-2)
			const int a = c(i), b = c(i - 1), d = c(i - 2);
			if (a < 4 && b < 4 && d < 4) nas = codon_tab[a << 4 | b << 2 | d];
		}
		const int s = ss(i); // nasw-sse.c:191-202: ss[x] of the original slice belongs to row nl - 1 - x, i.e. to the row of that nucleotide
		if (s >= 0) spsc_adjust(s, sq, false, don, acc);
	}
	return row_pack(nas, don, acc);
}
template <class Code>
NSW_HD uint32_t prep_row_left(const Code &c, int nl, int i, const int *sp, const uint8_t *codon_tab, int aa_x)
{
	return prep_row_left(c, nl, i, sp, codon_tab, aa_x, NoSpsc(), SpscPar{ 0, 0 });
}

// ------------------------------------------------------------------------------------------------
// backtrack over the traceback words (nasw-sse.c:40-89), run at a time.
//
// The reference walks one cell per iteration.  Consecutive cells very often repeat the same move (match runs along the
// diagonal, introns thousands of rows long, gap runs), and a run is exactly "the cells along one direction whose word
// keeps satisfying one condition", so a warp can test 32 cells of the run at once.  `Scan` supplies that test:
//     int lead(kind, i, j, &n_valid)   number of LEADING cells k = 0..31 along the direction of `kind` that are inside
//                                      the matrix and satisfy its condition; n_valid = cells of the 32 that are inside
//       kind 0  cells (i-3k, j-k)  condition: the cell decodes to state 0 (match)            [M run]
//       kind 1  cells (i,   j-k)   condition: bit 4 (insertion extended)                      [I run]
//       kind 2  cells (i-3k, j)    condition: bit 5 (deletion extended)                       [D run]
//       kind 3/4/5  cells (i-k, j) condition: bit 6/7/8 (intron phase 0/1/2 extended)         [N/U/V run]
//     uint32_t word(i, j)          one traceback word
// "inside" means i >= 2 and j >= 0, the loop condition of the reference.  The GPU kernel implements lead() with one
// load per lane and a ballot; the CPU emulation loops over k.  Everything else -- the state machine, CIGAR run-length
// merging (F and G never merge, nasw.h:141-151), the leftover rules and the tiny-U/V fix -- is this one function.
// Operations are written into out[0..cap) growing DOWN from out[cap-1]; on return out[cap-n .. cap) is the CIGAR in
// forward order.  Only the caller with `writer` set stores (lane 0 on the GPU).
// ------------------------------------------------------------------------------------------------
template <class Scan>
NSW_HD int backtrack_runs(Scan &sc, int nl, int al, uint32_t *out, int cap, bool writer)
{
	int i = nl - 1, j = al - 1, last = 0, n = 0;
	uint32_t cur_op = 0xffffffffu;
	int cur_len = 0;
	auto flush = [&]() {
		if (cur_len > 0 && n < cap) {
			++n;
			uint32_t op = cur_op;
			if ((op == 12 || op == 13) && cur_len < 3) op = 11; // nasw-sse.c:30-38: tiny U/V become G
			if (writer) out[cap - n] = (uint32_t)cur_len << 4 | op;
		}
		cur_len = 0;
	};
	auto push = [&](uint32_t op, int len) {
		if (cur_len > 0 && op == cur_op && op != 10 && op != 11) cur_len += len;
		else { flush(); cur_op = op, cur_len = len; }
	};
	while (i >= 2 && j >= 0) {
		int state = last;
		if (state == 0) {
			uint32_t x = sc.word(i, j);
			state = (x >> 9 & 1) ? 1 : (int)(x & 0xf);
		}
		if (state == 0) { // match run along the diagonal
			int nv;
			const int c = sc.lead(0, i, j, nv);
			push(0, c), i -= 3 * c, j -= c; // c >= 1: cell 0 is a match
		} else if (state <= 5) { // runs that continue while the extension bit of the current cell is set
			const int di = state == 1 ? 0 : state == 2 ? 3 : 1, dj = state == 1 ? 1 : 0;
			const uint32_t op = state <= 3 ? (uint32_t)state : (uint32_t)(state + 8); // I, D, N, U(12), V(13)
			int nv;
			const int c = sc.lead(state, i, j, nv);
			int cells;
			if (c == 32) cells = 32, last = state;          // still extending after 32 cells
			else if (c < nv) cells = c + 1, last = 0;       // cell c is inside and ends the run (its bit is clear)
			else cells = c, last = state;                   // ran out of the matrix while extending
			push(op, cells), i -= di * cells, j -= dj * cells;
			if ((state == 4 || state == 5) && c < 32 && c < nv) --j; // the closing cell of a phase-1/2 intron consumes the residue
		} else {
			switch (state) {
			case 6: push(10, 1), --i; break;
			case 7: push(10, 2), i -= 2; break;
			case 8: push(11, 1), --i, --j; break;
			case 9: push(11, 2), i -= 2, --j; break;
			default: i = -1000000; break; // states 10..15 never occur
			}
			last = 0;
		}
	}
	if (i > -1000000) {
		if (j > 0) push(1, j);
		if (i >= 0) {
			const int l = (i + 1) / 3 * 3, t = (i + 1) % 3;
			if (l > 0) push(2, l);
			if (t != 0) push(10, t);
		}
	}
	flush();
	return n;
}

// ------------------------------------------------------------------------------------------------
// backtrack over the traceback words (nasw-sse.c:40-89).  `tb(i, j)` returns the 10-bit word of cell (i,j).
// Operations are produced from the alignment end backwards, run-length merged except F and G
// (nasw.h:141-151), and written into out[0..cap) growing DOWN from out[cap-1], so that on return
// out[cap-n .. cap) is the CIGAR in forward order.  Returns n.
// ------------------------------------------------------------------------------------------------
template <class TbAt>
NSW_HD int backtrack(const TbAt &tb, int nl, int al, uint32_t *out, int cap)
{
	int i = nl - 1, j = al - 1, last = 0, n = 0;
	auto push = [&](uint32_t op, int len) {
		if (n == 0 || op != (out[cap - n] & 0xf) || op == 10 || op == 11) { if (n < cap) { ++n; out[cap - n] = (uint32_t)len << 4 | op; } }
		else out[cap - n] += (uint32_t)len << 4;
	};
	while (i >= 2 && j >= 0) {
		uint32_t x = tb(i, j);
		if (x >> 9 & 1) x = 1 | (x >> 4 << 4);
		const int state = last == 0 ? (int)(x & 0xf) : last;
		const int ext = (state >= 1 && state <= 5) ? (int)(x >> (state + 3) & 1) : 0;
		switch (state) {
		case 0: push(0, 1), i -= 3, --j; break;
		case 1: push(1, 1), --j; break;
		case 2: push(2, 1), i -= 3; break;
		case 3: push(3, 1), --i; break;
		case 4: push(12, 1), --i; if (!ext) --j; break;
		case 5: push(13, 1), --i; if (!ext) --j; break;
		case 6: push(10, 1), --i; break;
		case 7: push(10, 2), i -= 2; break;
		case 8: push(11, 1), --i, --j; break;
		case 9: push(11, 2), i -= 2, --j; break;
		default: i = -1000000; break; // states 10..15 never occur
		}
		last = (state >= 1 && state <= 5 && ext) ? state : 0;
	}
	if (i > -1000000) {
		if (j > 0) push(1, j);
		if (i >= 0) {
			const int l = (i + 1) / 3 * 3, t = (i + 1) % 3;
			if (l > 0) push(2, l);
			if (t != 0) push(10, t);
		}
	}
	for (int k = cap - n; k < cap; ++k) { // nasw-sse.c:30-38: tiny U/V become G
		const uint32_t op = out[k] & 0xf;
		if ((op == 12 || op == 13) && out[k] >> 4 < 3) out[k] = out[k] >> 4 << 4 | 11;
	}
	return n;
}

} // namespace nsw
