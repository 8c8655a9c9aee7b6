// tests/hostcheck/emu_nasw.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Lock-step CPU emulation of the nasw wavefront kernels (miniprot_b200/csrc/cuda/nasw_kernels.cu): the per-lane
// logic is the SAME header the kernels compile (nasw_core.cuh); only the warp shuffles, the shared-memory profile
// and the carry arrays are replaced by plain arrays here.  Lets the CPU test-suite compare the device arithmetic
// with the oracle for every column-per-lane variant and for multi-pass problems, without a GPU.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "cuda/nasw_core.cuh"
#include "cuda/nasw_pair.cuh"

using namespace nsw;

namespace {

struct EmuEnv {
	const RowRec *rec;
	int nl;
	const int *prof; // offset to the lane's first column
	int Wp;
	int *cy;
	RowRec row_rec(int i) const { i = i < 0 ? 0 : (i > nl ? nl : i); return rec[i]; }
	void prefetch_row(int) const {}
	// block-wide kernels: records by triple of rows (triple m = rows 3m+2..3m+4); the device stores them field-major, which
	// changes the addresses, not the values
	int cur, M;
	void rec3(int m, RowRec &r0, RowRec &r1, RowRec &r2) const { m = m < 0 ? 0 : (m >= M ? M - 1 : m); r0 = rec[3 * m + 2], r1 = rec[3 * m + 3], r2 = rec[3 * m + 4]; }
	void seek3(int m) { cur = m; }
	void prefetch_ahead(int) const {}
	void next3(RowRec &r0, RowRec &r1, RowRec &r2) { r0 = rec[3 * cur + 2], r1 = rec[3 * cur + 3], r2 = rec[3 * cur + 4], ++cur; }
	int profile_stride() const { return Wp; }
	const int *profile(int nas) const { return prof + nas * Wp; }
	void carry_load3(int i, int &a, int &b, int &c) const { a = cy[(int64_t)i * 3], b = cy[(int64_t)i * 3 + 1], c = cy[(int64_t)i * 3 + 2]; }
	void carry_store3(int i, int a, int b, int c) const { cy[(int64_t)i * 3] = a, cy[(int64_t)i * 3 + 1] = b, cy[(int64_t)i * 3 + 2] = c; }
	void carry_load4(int i, int &a, int &b, int &c, int &d) const { a = cy[(int64_t)i * 4], b = cy[(int64_t)i * 4 + 1], c = cy[(int64_t)i * 4 + 2], d = cy[(int64_t)i * 4 + 3]; }
	void carry_store4(int i, int a, int b, int c, int d) const { cy[(int64_t)i * 4] = a, cy[(int64_t)i * 4 + 1] = b, cy[(int64_t)i * 4 + 2] = c, cy[(int64_t)i * 4 + 3] = d; }
};

struct Problem {
	std::vector<uint32_t> w; // row words (nasw_core.cuh row_pack): slot x <-> row x - 2, clamped like the prep kernels
	int io;
	std::vector<RowRec> rec;
	std::vector<int> aas;
	int nl, al, W8;
	Par par;
	const int8_t *mat;
	int end_bonus, xdrop;
	float ie_coef;
};

template <int C, bool MULTI>
void run_ext(const Problem &P, int *score, int *nt_len, int *aa_len)
{
	const int Wp = 32 * C, n_pass = (P.W8 + Wp - 1) / Wp, T = P.nl > 2 ? P.nl - 2 + 32 : 0;
	std::vector<int> prof(22 * Wp), cy((size_t)(P.nl + 1) * 3);
	ExtTracker trk;
	trk.init(code_bits(P.al));
	PenTable pt;
	pen_table_build(P.ie_coef, pt);
	for (int pass = 0; pass < n_pass; ++pass) {
		for (int j = 0; j < Wp; ++j)
			for (int a = 0; a < 22; ++a) prof[a * Wp + j] = pass * Wp + j < P.al ? P.mat[a * 22 + P.aas[pass * Wp + j]] : NEG;
		ExtLane<C, MULTI> L[32];
		ExtTracker trks[32];
		for (int l = 0; l < 32; ++l) trks[l] = trk;
		LaneGeom g[32];
		EmuEnv env[32];
		for (int l = 0; l < 32; ++l) {
			g[l].lane = l, g[l].pass = pass, g[l].n_pass = n_pass, g[l].nl = P.nl, g[l].al = P.al, g[l].W8 = P.W8, g[l].col0 = pass * Wp + l * C, g[l].live = g[l].col0 < P.W8;
			env[l].rec = P.rec.data(), env[l].nl = P.nl, env[l].prof = prof.data() + l * C, env[l].Wp = Wp, env[l].cy = cy.data();
			L[l].init(g[l], P.end_bonus, P.par.fs, env[l]);
		}
		for (int t = 0; t < T; ++t) {
			int rH[32], rI[32], rB[32];
			for (int l = 0; l < 32; ++l) { const int s = l ? l - 1 : 0; rH[l] = L[s].outH, rI[l] = L[s].outI, rB[l] = L[s].outB; } // __shfl_up_sync(.., 1)
			for (int l = 0; l < 32; ++l) {
				int ri;
				switch (t % 6) { // the kernels unroll the loop by the phase period
				case 0: ri = L[l].template step<0>(g[l], P.par, t, rH[l], rI[l], rB[l], env[l]); break;
				case 1: ri = L[l].template step<1>(g[l], P.par, t, rH[l], rI[l], rB[l], env[l]); break;
				case 2: ri = L[l].template step<2>(g[l], P.par, t, rH[l], rI[l], rB[l], env[l]); break;
				case 3: ri = L[l].template step<3>(g[l], P.par, t, rH[l], rI[l], rB[l], env[l]); break;
				case 4: ri = L[l].template step<4>(g[l], P.par, t, rH[l], rI[l], rB[l], env[l]); break;
				default: ri = L[l].template step<5>(g[l], P.par, t, rH[l], rI[l], rB[l], env[l]); break;
				}
				if (ri >= 0 && pass == n_pass - 1) trks[l].row(ri, L[l].outB, P.al * 3, pt, P.xdrop); // every lane tracks; lane 31 counts
			}
			if (t == 0) for (int l = 0; l < 32; ++l) L[l].after_first_step(g[l]);
			if (pass == n_pass - 1 && (t & 15) == 15 && trks[31].stopped) break;
		}
		trk = trks[31];
	}
	*score = trk.max_sc, *nt_len = trk.max_i + 1;
	*aa_len = trk.aa_len(P.al);
}

template <int C, bool MULTI>
void run_tb(const Problem &P, int *score, std::vector<uint32_t> &cigar)
{
	const int Wp = 32 * C, n_pass = (P.W8 + Wp - 1) / Wp, T = P.nl > 2 ? P.nl - 2 + 32 : 0;
	std::vector<int> prof(22 * Wp), cy((size_t)(P.nl + 1) * 4);
	std::vector<uint16_t> tb((size_t)n_pass * (T ? T : 1) * Wp, 0xffff);
	int sc = NEG;
	for (int pass = 0; pass < n_pass; ++pass) {
		for (int j = 0; j < Wp; ++j)
			for (int a = 0; a < 22; ++a) prof[a * Wp + j] = pass * Wp + j < P.al ? P.mat[a * 22 + P.aas[pass * Wp + j]] : NEG;
		TbLane<C, MULTI> L[32];
		LaneGeom g[32];
		EmuEnv env[32];
		for (int l = 0; l < 32; ++l) {
			g[l].lane = l, g[l].pass = pass, g[l].n_pass = n_pass, g[l].nl = P.nl, g[l].al = P.al, g[l].W8 = P.W8, g[l].col0 = pass * Wp + l * C, g[l].live = g[l].col0 < P.W8;
			env[l].rec = P.rec.data(), env[l].nl = P.nl, env[l].prof = prof.data() + l * C, env[l].Wp = Wp, env[l].cy = cy.data();
			L[l].init(g[l], P.par.fs, env[l]);
		}
		for (int t = 0; t < T; ++t) {
			int rH[32], rF[32], rS[32], rI[32];
			for (int l = 0; l < 32; ++l) { const int s = l ? l - 1 : 0; rH[l] = L[s].outH, rF[l] = L[s].outF, rS[l] = L[s].outS, rI[l] = L[s].outI; }
			for (int l = 0; l < 32; ++l) {
				uint32_t wd[C];
				bool w;
				switch (t % 6) {
				case 0: w = L[l].template step<0>(g[l], P.par, t, rH[l], rF[l], rS[l], rI[l], env[l], wd); break;
				case 1: w = L[l].template step<1>(g[l], P.par, t, rH[l], rF[l], rS[l], rI[l], env[l], wd); break;
				case 2: w = L[l].template step<2>(g[l], P.par, t, rH[l], rF[l], rS[l], rI[l], env[l], wd); break;
				case 3: w = L[l].template step<3>(g[l], P.par, t, rH[l], rF[l], rS[l], rI[l], env[l], wd); break;
				case 4: w = L[l].template step<4>(g[l], P.par, t, rH[l], rF[l], rS[l], rI[l], env[l], wd); break;
				default: w = L[l].template step<5>(g[l], P.par, t, rH[l], rF[l], rS[l], rI[l], env[l], wd); break;
				}
				if (w)
					for (int k = 0; k < C; ++k) tb[((size_t)pass * T + t) * Wp + l * C + k] = (uint16_t)wd[k];
			}
			if (t == 0) for (int l = 0; l < 32; ++l) L[l].after_first_step(g[l]);
		}
		for (int l = 0; l < 32; ++l) if (L[l].k_end >= 0) sc = L[l].score;
	}
	*score = sc;
	auto at = [&](int i, int j) -> uint32_t {
		const int pass = j / Wp, jc = j - pass * Wp, lane = jc / C;
		return tb[((size_t)pass * T + (i - 2 + lane)) * Wp + jc];
	};
	const int cap = P.nl + P.al + 8;
	std::vector<uint32_t> buf((size_t)cap), buf2((size_t)cap);
	const int n = backtrack(at, P.nl, P.al, buf.data(), cap); // cell-at-a-time walk (the reference's own formulation)
	struct CpuScan { // what the GPU does with one load per lane and a ballot
		decltype(at) &tb;
		uint32_t word(int i, int j) const { return tb(i, j); }
		int lead(int kind, int i, int j, int &n_valid) const
		{
			const int di = kind == 0 ? 3 : kind == 1 ? 0 : kind == 2 ? 3 : 1, dj = kind <= 1 ? 1 : 0;
			int c = 0;
			bool open = true;
			n_valid = 0;
			for (int k = 0; k < 32; ++k) {
				const int ii = i - di * k, jj = j - dj * k;
				if (ii < 2 || jj < 0) break;
				++n_valid;
				const uint32_t x = tb(ii, jj);
				const bool ok = kind == 0 ? ((x >> 9 & 1) ? false : (x & 0xf) == 0) : (x >> (kind + 3) & 1);
				if (open && ok) ++c; else open = false;
			}
			return c;
		}
	} scan{at};
	const int n2 = backtrack_runs(scan, P.nl, P.al, buf2.data(), cap, true);
	if (n2 != n || memcmp(buf.data() + (cap - n), buf2.data() + (cap - n2), sizeof(uint32_t) * (size_t)n) != 0) { fprintf(stderr, "[emu] run-based backtrack differs from the cell walk\n"); abort(); }
	cigar.assign(buf2.begin() + (cap - n2), buf2.end());
}

// block-wide wavefront (Lane3): NW*32 threads in lockstep, thread x <- thread x-1 from the previous macro-step
template <bool TB>
void run_v3(const Problem &P, int *score, int *nt_len, int *aa_len, std::vector<uint32_t> &cigar)
{
	int NW = (P.W8 + 31) / 32;
	NW = NW <= 1 ? 1 : NW <= 2 ? 2 : NW <= 4 ? 4 : 8;
	const int Wp = 32 * NW, n_macro = P.nl > 2 ? (P.nl - 2 + 2) / 3 + Wp : 0, n_pass = (P.W8 + Wp - 1) / Wp, Trows = 3 * (n_macro + 2);
	std::vector<int> prof(22 * Wp);
	std::vector<Lane3<TB>> L((size_t)Wp);
	std::vector<Geo3> g((size_t)Wp);
	std::vector<EmuEnv> env((size_t)Wp);
	ExtTracker trk; // fed by the last column of the last pass
	trk.init(code_bits(P.al));
	std::vector<uint16_t> tb(TB ? (size_t)n_pass * Trows * Wp : 1, 0xffff);
	struct Carry { int h, i, x, s; };
	std::vector<Carry> carry((size_t)P.nl + 2, Carry{ NEG, NEG, NEG, NEG }); // what the last column of a pass leaves for the next pass, per row
	PenTable pt;
	pen_table_build(P.ie_coef, pt);
	int tb_score = NEG;
	bool stop = false;
	for (int pass = 0; pass < n_pass && !stop; ++pass) {
		const bool last_pass = pass == n_pass - 1;
		for (int j = 0; j < Wp; ++j)
			for (int a = 0; a < 22; ++a) prof[a * Wp + j] = pass * Wp + j < P.al ? P.mat[a * 22 + P.aas[pass * Wp + j]] : NEG;
		for (int x = 0; x < Wp; ++x) {
			g[x].x = x, g[x].col = pass * Wp + x, g[x].nl = P.nl, g[x].al = P.al, g[x].W8 = P.W8, g[x].live = g[x].col < P.W8, g[x].first = g[x].col == 0;
			env[x].rec = P.rec.data(), env[x].nl = P.nl, env[x].prof = prof.data() + x, env[x].Wp = Wp, env[x].cy = 0, env[x].M = v3_triples(P.nl), env[x].cur = 0;
			L[x].init(g[x], P.end_bonus, P.par.fs, env[x]);
		}
		std::vector<int> sH((size_t)Wp * 3), sI((size_t)Wp * 3), sX((size_t)Wp * 3), sS((size_t)Wp * 3);
		std::vector<Carry> carry_next = carry; // rows are read (by column 0) long before the last column rewrites them
		int t_lo, t_hi;
		Lane3<TB>::steady_range(P.nl, Wp, t_lo, t_hi);
		for (int T = 0; T < n_macro + (n_macro & 1); ++T) {
			const bool steady = T >= t_lo && T < t_hi; // the kernel's middle loop
			for (int x = 0; x < Wp; ++x)
				for (int r = 0; r < 3; ++r) sH[x * 3 + r] = L[x].oH[r], sI[x * 3 + r] = L[x].oI[r], sX[x * 3 + r] = L[x].oX[r], sS[x * 3 + r] = L[x].oS[r];
			for (int x = 0; x < Wp; ++x) {
				const int s = x ? x - 1 : 0; // what __shfl_up_sync(..., 1) / the shared-memory slot delivers
				uint32_t wd[3], done;
				// inputs from the left: the column x-1 of this pass, the carry of the previous pass, or the constant boundary
				int rH[3], rI[3], rX[3], rS[3];
				for (int r = 0; r < 3; ++r) {
					if (x) rH[r] = sH[s * 3 + r], rI[r] = sI[s * 3 + r], rX[r] = sX[s * 3 + r], rS[r] = sS[s * 3 + r];
					else if (pass > 0) {
						int i = 3 * T + 2 + r;
						i = i < P.nl ? i : P.nl;
						rH[r] = carry[(size_t)i].h, rI[r] = carry[(size_t)i].i, rX[r] = carry[(size_t)i].x, rS[r] = carry[(size_t)i].s;
					} else rH[r] = NEG, rI[r] = NEG, rS[r] = NEG, rX[r] = TB ? NEG : INT32_MIN;
				}
				if (steady) {
					// the kernel's alternating buffers: "previous" = what L[] would hold, "this" = the values just received
					int pH[3];
					if (T == t_lo) L[x].steady_enter(g[x], T, pH, env[x]);
					else for (int r = 0; r < 3; ++r) pH[r] = L[x].L[r];
					if (T & 1) L[x].template macro_steady<1>(g[x], P.par, pH, rH, rI, rX, rS, env[x], wd);
					else L[x].template macro_steady<0>(g[x], P.par, pH, rH, rI, rX, rS, env[x], wd);
					L[x].steady_leave(rH);
					done = TB ? (g[x].live ? 7u : 0u) : 7u;
				} else if (T & 1) done = L[x].template macro<1>(g[x], P.par, T, rH, rI, rX, rS, env[x], wd);
				else done = L[x].template macro<0>(g[x], P.par, T, rH, rI, rX, rS, env[x], wd);
				for (int r = 0; r < 3; ++r) {
					if (!(done >> r & 1)) continue;
					const int i = Lane3<TB>::row_of(g[x], T, r);
					if (TB) tb[((size_t)pass * Trows + (size_t)(3 * T + r)) * Wp + x] = (uint16_t)wd[r];
					else if (x == Wp - 1 && last_pass) trk.row(i, L[x].oX[r], P.al * 3, pt, P.xdrop);
					if (x == Wp - 1 && !last_pass) carry_next[(size_t)i] = Carry{ L[x].oH[r], L[x].oI[r], L[x].oX[r], L[x].oS[r] };
				}
			}
			if (!TB && (T & 1) && trk.stopped) { stop = true; break; }
		}
		carry.swap(carry_next);
		const int c_end = P.al > 0 ? P.al - 1 : 0;
		if (TB && c_end / Wp == pass) tb_score = L[(size_t)(c_end % Wp)].score;
	}
	if (TB) {
		*score = tb_score;
		auto at = [&](int i, int j) -> uint32_t { const int ps = j / Wp, jc = j - ps * Wp; return tb[((size_t)ps * Trows + (size_t)(i - 2 + 3 * jc)) * Wp + jc]; };
		struct CpuScan {
			decltype(at) &tbf;
			uint32_t word(int i, int j) const { return tbf(i, j); }
			int lead(int kind, int i, int j, int &n_valid) const
			{
				const int di = kind == 0 ? 3 : kind == 1 ? 0 : kind == 2 ? 3 : 1, dj = kind <= 1 ? 1 : 0;
				int c = 0;
				bool open = true;
				n_valid = 0;
				for (int k = 0; k < 32; ++k) {
					const int ii = i - di * k, jj = j - dj * k;
					if (ii < 2 || jj < 0) break;
					++n_valid;
					const uint32_t x = tbf(ii, jj);
					const bool ok = kind == 0 ? ((x >> 9 & 1) ? false : (x & 0xf) == 0) : (x >> (kind + 3) & 1);
					if (open && ok) ++c; else open = false;
				}
				return c;
			}
		} scan{at};
		const int cap = P.nl + P.al + 8;
		std::vector<uint32_t> buf((size_t)cap);
		const int n = backtrack_runs(scan, P.nl, P.al, buf.data(), cap, true);
		cigar.assign(buf.begin() + (cap - n), buf.end());
	} else {
		const ExtTracker &t = trk;
		*score = t.max_sc, *nt_len = t.max_i + 1;
		*aa_len = t.aa_len(P.al);
	}
}

// pair-lane wavefront (nasw_pair.cuh): W8/2 threads in lockstep, two columns each; thread x <- thread x-1 from the previous
// macro-step.  Rows in which both halves of every thread are real go through the straight-line row(), the others through
// row_masked(), like the kernels' steady and general loops.
struct PairEnv {
	const uint32_t *lo, *hi; // [22][NT] packed profile words of this thread (value in the low resp. high half)
	int nt;
	uint32_t prof(uint32_t off) const { const uint32_t a = off / 128u; return a < 22 ? lo[a * (uint32_t)nt] : hi[(a - 22) * (uint32_t)nt]; }
};

template <bool TB>
void run_pair(const Problem &P, int *score, int *nt_len, int *aa_len, std::vector<uint32_t> &cigar)
{
	const int NT = P.W8 / 2, M = P.nl > 2 ? (P.nl - 2 + 2) / 3 : 0, n_macro = M > 0 ? M + P.W8 + 1 : 0, Wp = (P.W8 + 63) / 64 * 64, Trows = 3 * (n_macro + 2);
	PairPar pp;
	pp.go = P.par.go, pp.ge = P.par.ge, pp.fs = P.par.fs, pp.end_bonus = P.end_bonus, pp.ngo = pk2(-P.par.go), pp.nfs = pk2(-P.par.fs);
	std::vector<uint32_t> plo((size_t)22 * NT), phi((size_t)22 * NT);
	for (int x = 0; x < NT; ++x)
		for (int a = 0; a < 22; ++a) {
			const int j0 = 2 * x, j1 = 2 * x + 1;
			plo[(size_t)a * NT + x] = pk(j0 < P.al ? P.mat[a * 22 + P.aas[j0]] : PAIR_DEAD, 0);
			phi[(size_t)a * NT + x] = pk(0, j1 < P.al ? P.mat[a * 22 + P.aas[j1]] : PAIR_DEAD);
		}
	auto rw = [&](int k) { k = k < 0 ? 0 : (k > P.nl ? P.nl : k); return P.w[(size_t)k + 2]; };
	std::vector<PairGeo> g((size_t)NT);
	std::vector<PairEnv> env((size_t)NT);
	std::vector<PairLane> LE((size_t)(TB ? 0 : NT));
	std::vector<PairLaneTb> LT((size_t)(TB ? NT : 0));
	std::vector<uint32_t> pH((size_t)NT * 3, 0);
	for (int x = 0; x < NT; ++x) {
		g[x].x = x, g[x].col = 2 * x, g[x].nl = P.nl, g[x].al = P.al, g[x].W8 = P.W8, g[x].first = x == 0;
		env[x].lo = plo.data() + x, env[x].hi = phi.data() + x, env[x].nt = NT;
		if (TB) LT[x].init(g[x], pp); else LE[x].init(g[x], pp);
	}
	ExtTracker trk;
	trk.init(PAIR_CB);
	PenTable pt;
	pen_table_build(P.ie_coef, pt);
	std::vector<uint16_t> tb(TB ? (size_t)Trows * Wp : 1, 0xffff);
	int tb_score = 0;
	bool stop = false;
	std::vector<uint32_t> sH((size_t)NT * 3), sQ((size_t)NT * 3), sF((size_t)NT * 3), sS((size_t)NT * 3);
	std::vector<int> sXlo((size_t)NT * 3), sXhi((size_t)NT * 3);
	for (int T = 0; T < n_macro && !stop; ++T) {
		for (int x = 0; x < NT; ++x)
			for (int r = 0; r < 3; ++r) {
				if (TB) sH[x * 3 + r] = LT[x].oH[r], sQ[x * 3 + r] = LT[x].oQ[r], sF[x * 3 + r] = LT[x].oF[r], sS[x * 3 + r] = LT[x].oS[r];
				else sH[x * 3 + r] = LE[x].oH[r], sQ[x * 3 + r] = LE[x].oQ[r], sXlo[x * 3 + r] = LE[x].oXlo[r], sXhi[x * 3 + r] = LE[x].oXhi[r];
			}
		for (int x = 0; x < NT; ++x) {
			const int s = x ? x - 1 : 0; // __shfl_up_sync(.., 1): thread 0 gets its own register back
			const int m = T - 2 * x;
			const PairRec rec = make_pair_rec(rw, m, P.io, P.par.ge, P.par.fs, 128, 22 * 128);
			uint32_t rH[3], rQ[3], rF[3], rS[3], keep[3];
			int lx[3], xp[3];
			bool bnd[3], all = true;
			for (int r = 0; r < 3; ++r) {
				const int i_lo = 3 * m + 2 + r, i_hi = i_lo - 3;
				keep[r] = (i_lo >= 2 && i_lo < P.nl ? 0xffffu : 0u) | (i_hi >= 2 && i_hi < P.nl ? 0xffff0000u : 0u);
				bnd[r] = x == 0 && i_lo == 2;
				all = all && keep[r] == 0xffffffffu && !bnd[r];
				if (TB) {
					rH[r] = LT[x].left_of(sH[s * 3 + r], sH[x * 3 + r]), rQ[r] = LT[x].left_of(sQ[s * 3 + r], sQ[x * 3 + r]);
					rF[r] = LT[x].left_of(sF[s * 3 + r], sF[x * 3 + r]), rS[r] = LT[x].left_of(sS[s * 3 + r], sS[x * 3 + r]);
				} else {
					rH[r] = LE[x].left_of(sH[s * 3 + r], sH[x * 3 + r]), rQ[r] = LE[x].left_of(sQ[s * 3 + r], sQ[x * 3 + r]);
					lx[r] = (int)((uint32_t)sXhi[s * 3 + r] & LE[x].xmask), xp[r] = sXlo[x * 3 + r];
				}
			}
			uint32_t *p = &pH[(size_t)x * 3];
			uint32_t wd[3] = { 0, 0, 0 };
			if (TB) {
				PairLaneTb &L = LT[x];
				if (all) {
					wd[0] = L.template row<0>(pp, rec, env[x], rH[0], p[2], p[1], p[0], rQ[0], rF[0], rS[0]);
					wd[1] = L.template row<1>(pp, rec, env[x], rH[1], rH[0], p[2], p[1], rQ[1], rF[1], rS[1]);
					wd[2] = L.template row<2>(pp, rec, env[x], rH[2], rH[1], rH[0], p[2], rQ[2], rF[2], rS[2]);
				} else {
					wd[0] = L.template row_masked<0>(pp, rec, env[x], rH[0], p[2], p[1], p[0], rQ[0], rF[0], rS[0], keep[0], bnd[0]);
					wd[1] = L.template row_masked<1>(pp, rec, env[x], rH[1], rH[0], p[2], p[1], rQ[1], rF[1], rS[1], keep[1], bnd[1]);
					wd[2] = L.template row_masked<2>(pp, rec, env[x], rH[2], rH[1], rH[0], p[2], rQ[2], rF[2], rS[2], keep[2], bnd[2]);
				}
				for (int r = 0; r < 3; ++r) {
					const int i_lo = 3 * m + 2 + r, i_hi = i_lo - 3;
					if (keep[r] & 0xffffu) {
						tb[(size_t)(3 * T + r) * Wp + 2 * x] = (uint16_t)(wd[r] & 0xffff);
						if (i_lo == P.nl - 1 && 2 * x == P.al - 1) tb_score = lo16(L.oH[r]) - PAIR_BIAS;
					}
					if (keep[r] >> 16) {
						tb[(size_t)(3 * T + r) * Wp + 2 * x + 1] = (uint16_t)(wd[r] >> 16);
						if (i_hi == P.nl - 1 && 2 * x + 1 == P.al - 1) tb_score = hi16(L.oH[r]) - PAIR_BIAS;
					}
				}
			} else {
				PairLane &L = LE[x];
				if (all) {
					L.template row<0>(pp, rec, env[x], rH[0], p[2], p[1], p[0], rQ[0], lx[0], xp[0]);
					L.template row<1>(pp, rec, env[x], rH[1], rH[0], p[2], p[1], rQ[1], lx[1], xp[1]);
					L.template row<2>(pp, rec, env[x], rH[2], rH[1], rH[0], p[2], rQ[2], lx[2], xp[2]);
				} else {
					L.template row_masked<0>(pp, rec, env[x], rH[0], p[2], p[1], p[0], rQ[0], lx[0], xp[0], keep[0], bnd[0]);
					L.template row_masked<1>(pp, rec, env[x], rH[1], rH[0], p[2], p[1], rQ[1], lx[1], xp[1], keep[1], bnd[1]);
					L.template row_masked<2>(pp, rec, env[x], rH[2], rH[1], rH[0], p[2], rQ[2], lx[2], xp[2], keep[2], bnd[2]);
				}
				if (x == NT - 1)
					for (int r = 0; r < 3; ++r) {
						const int i_hi = 3 * m + 2 + r - 3;
						if (keep[r] >> 16) trk.row(i_hi, L.oXhi[r], P.al * 3, pt, P.xdrop);
					}
			}
			p[0] = rH[0], p[1] = rH[1], p[2] = rH[2];
		}
		if (!TB && trk.stopped) stop = true;
	}
	if (TB) {
		*score = tb_score;
		auto at = [&](int i, int j) -> uint32_t { return tb[(size_t)(i - 2 + 3 * j) * Wp + j]; };
		struct CpuScan {
			decltype(at) &tbf;
			uint32_t word(int i, int j) const { return tbf(i, j); }
			int lead(int kind, int i, int j, int &n_valid) const
			{
				const int di = kind == 0 ? 3 : kind == 1 ? 0 : kind == 2 ? 3 : 1, dj = kind <= 1 ? 1 : 0;
				int c = 0;
				bool open = true;
				n_valid = 0;
				for (int k = 0; k < 32; ++k) {
					const int ii = i - di * k, jj = j - dj * k;
					if (ii < 2 || jj < 0) break;
					++n_valid;
					const uint32_t x = tbf(ii, jj);
					const bool ok = kind == 0 ? ((x >> 9 & 1) ? false : (x & 0xf) == 0) : (x >> (kind + 3) & 1);
					if (open && ok) ++c; else open = false;
				}
				return c;
			}
		} scan{at};
		const int cap = P.nl + P.al + 8;
		std::vector<uint32_t> buf((size_t)cap);
		const int n = backtrack_runs(scan, P.nl, P.al, buf.data(), cap, true);
		cigar.assign(buf.begin() + (cap - n), buf.end());
	} else {
		*score = trk.max_i >= 0 ? trk.max_sc - PAIR_BIAS : INT32_MIN, *nt_len = trk.max_i + 1;
		*aa_len = trk.aa_len(P.al);
	}
}

} // namespace

// ss = --spsc bytes of the slice (NULL: none), null_bonus = ns_opt_t::sp_null_bonus; problems with ss never run on the pair-lane kernels
extern "C" int emu_nasw_ss(const uint8_t *nt4, const uint8_t *aa20, const uint8_t *codon, const int8_t *mat, const int32_t *sp, int go, int ge, int io,
                           int fs, int xdrop, int end_bonus, float ie_coef, int flag, int C, const uint8_t *ns, int nl, const char *as, int al,
                           int *score, int *nt_len, int *aa_len, uint32_t *cigar, int cigar_cap, const uint8_t *ss, int null_bonus)
{
	Problem P;
	P.nl = nl, P.al = al, P.W8 = (al + 7) / 8 * 8, P.mat = mat, P.end_bonus = end_bonus, P.xdrop = xdrop, P.ie_coef = ie_coef;
	P.par.go = go, P.par.ge = ge, P.par.io = io, P.par.fs = fs, P.par.gei_stop = fs;
	const bool left = flag & 2;
	std::vector<int> code((size_t)nl);
	for (int k = 0; k < nl; ++k) code[(size_t)k] = nt4[ns[left ? nl - 1 - k : k]];
	auto c = [&](int k) { return code[(size_t)k]; };
	auto sbyte = [&](int k) { return ss ? (int)ss[left ? nl - 1 - k : k] : -1; }; // byte of the nucleotide of DP row k
	const SpscPar sq = { (io + 1) / 2 - 1, null_bonus };
	const int n_rec = std::max(nl + 1, 2 + 3 * v3_triples(nl)); // rows past nl repeat the clamped rules, like the prep kernel
	std::vector<uint32_t> w((size_t)n_rec + 3);
	for (int x = 0; x < n_rec + 3; ++x) { // slot x <-> row x - 2, clamped like the prep kernel
		int r = x - 2;
		r = r < 0 ? 0 : (r > nl ? nl : r);
		w[(size_t)x] = left ? prep_row_left(c, nl, r, sp, codon, aa20['X'], sbyte, sq) : prep_row_forward(c, nl, r, sp, codon, aa20['X'], sbyte, sq);
	}
	P.w = w, P.io = io;
	P.rec.resize((size_t)n_rec);
	for (int r = 0; r < n_rec; ++r) P.rec[(size_t)r] = make_row_rec(P.par, w[(size_t)r], w[(size_t)r + 1], w[(size_t)r + 2], w[(size_t)r + 3]);
	P.aas.resize((size_t)al);
	for (int j = 0; j < al; ++j) P.aas[(size_t)j] = aa20[(uint8_t)as[left ? al - 1 - j : j]];
	*nt_len = nl, *aa_len = al;
	int n_cig = 0;
	if (C < 0) { // pair-lane kernels (two columns per thread, int16x2)
		std::vector<uint32_t> cg;
		if (nl <= 2) { // nothing to compute: the kernels report what the block-wide kernels report
			if (flag & 6) run_v3<false>(P, score, nt_len, aa_len, cg);
			else run_v3<true>(P, score, nt_len, aa_len, cg);
		} else if (flag & 6) run_pair<false>(P, score, nt_len, aa_len, cg);
		else run_pair<true>(P, score, nt_len, aa_len, cg);
		n_cig = (int)cg.size();
		for (int k = 0; k < n_cig && k < cigar_cap; ++k) cigar[k] = cg[(size_t)k];
	} else if (C == 0) { // block-wide wavefront kernels
		std::vector<uint32_t> cg;
		if (flag & 6) run_v3<false>(P, score, nt_len, aa_len, cg);
		else run_v3<true>(P, score, nt_len, aa_len, cg);
		n_cig = (int)cg.size();
		for (int k = 0; k < n_cig && k < cigar_cap; ++k) cigar[k] = cg[(size_t)k];
	} else if (flag & 6) {
		switch (C) {
		case 1: P.W8 <= 32 ? run_ext<1, false>(P, score, nt_len, aa_len) : run_ext<1, true>(P, score, nt_len, aa_len); break;
		case 2: P.W8 <= 64 ? run_ext<2, false>(P, score, nt_len, aa_len) : run_ext<2, true>(P, score, nt_len, aa_len); break;
		case 4: P.W8 <= 128 ? run_ext<4, false>(P, score, nt_len, aa_len) : run_ext<4, true>(P, score, nt_len, aa_len); break;
		default: P.W8 <= 256 ? run_ext<8, false>(P, score, nt_len, aa_len) : run_ext<8, true>(P, score, nt_len, aa_len); break;
		}
	} else {
		std::vector<uint32_t> cg;
		switch (C) {
		case 1: P.W8 <= 32 ? run_tb<1, false>(P, score, cg) : run_tb<1, true>(P, score, cg); break;
		case 2: P.W8 <= 64 ? run_tb<2, false>(P, score, cg) : run_tb<2, true>(P, score, cg); break;
		case 4: P.W8 <= 128 ? run_tb<4, false>(P, score, cg) : run_tb<4, true>(P, score, cg); break;
		default: P.W8 <= 256 ? run_tb<8, false>(P, score, cg) : run_tb<8, true>(P, score, cg); break;
		}
		n_cig = (int)cg.size();
		for (int k = 0; k < n_cig && k < cigar_cap; ++k) cigar[k] = cg[(size_t)k];
	}
	return n_cig;
}

extern "C" int emu_nasw(const uint8_t *nt4, const uint8_t *aa20, const uint8_t *codon, const int8_t *mat, const int32_t *sp, int go, int ge, int io,
                        int fs, int xdrop, int end_bonus, float ie_coef, int flag, int C, const uint8_t *ns, int nl, const char *as, int al,
                        int *score, int *nt_len, int *aa_len, uint32_t *cigar, int cigar_cap)
{
	return emu_nasw_ss(nt4, aa20, codon, mat, sp, go, ge, io, fs, xdrop, end_bonus, ie_coef, flag, C, ns, nl, as, al, score, nt_len, aa_len, cigar, cigar_cap, 0, 0);
}

// step-table form of the extension length penalty vs the direct FP32 formula, for x in [0, xmax]
extern "C" int emu_pen_check(float ie_coef, int xmax)
{
	PenTable pt;
	pen_table_build(ie_coef, pt);
	int bad = 0, pen = 0, pk = 0;
	for (int x = 0; x <= xmax; ++x) {
		while (pk < pt.n && x >= pt.thr[pk]) pen = pt.val[pk], ++pk;
		bad += pen != ext_len_penalty(ie_coef, x);
	}
	return bad;
}

// would the dispatcher send this problem to the pair-lane kernels? (nasw_host.cu use_pair: value-domain check + width; no --spsc)
extern "C" int emu_pair_eligible(const int8_t *mat, const int32_t *sp, int go, int ge, int io, int fs, int end_bonus, int nl, int al)
{
	const PairLimits l = pair_limits(mat, sp);
	if ((al + 7) / 8 * 8 > PAIR_MAX_W8 || nl < 3) return 0;
	return pair_eligible(al, go, ge, io, fs, end_bonus, l.smin, l.smax, l.dmax, l.amax) ? 1 : 0;
}
