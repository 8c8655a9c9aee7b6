// nasw_pair_kernels.cu -- the PAIR-LANE nasw kernels ("v4") for sm_100a: two protein columns per thread as packed int16x2
// (VIADDMNMX.S16x2 / VIMNMX3.S16x2 / VIMNMX.S16x2 with predicates), see nasw_pair.cuh for the per-thread logic, the value
// domain and the exactness argument.
//
//   nasw_prep_pair_kernel   packed genome -> pair records (one 96-byte record per triple of rows, entries paired with the row
//                           three above: what a thread needs when its low column is on triple m and its high column on m-1)
//   nasw_pair_kernel<TB>    one WARP per problem of up to 64 padded columns: the wavefront moves by shuffles only (no shared-memory
//                           exchange, no barrier).  Wider problems stay on the block-wide kernels of nasw_kernels.cu: a form of this
//                           kernel with several warps per problem (neighbouring warps linked by a ring of tagged slots in shared
//                           memory instead of a block-wide barrier) was built and measured in this round and lost to them -- 326 against
//                           181 cycles per row for 120 columns, independent of the load (profiles/README.md) -- so it is not kept.
//                           TB = false: score-only extension with the warp-parallel x-drop tracker (nasw_warp.cuh);
//                           TB = true : global alignment, two traceback words per thread and row in one 32-bit store, in the
//                           wavefront-major layout nasw_bt_kernel walks.
#include <algorithm>
#include <cuda_runtime.h>
#include <stdint.h>
#include "nasw_pair.cuh"
#include "nasw_dev.hpp"
#include "nasw_warp.cuh"

namespace mpb {
namespace cuda {

using namespace nsw;

constexpr int PAIR_TRI = 1024;  // triples per prep CTA
constexpr int PAIR_PROF_HI = 22 * 128; // byte offset of the high-half profile table of a warp

__device__ __forceinline__ int pair_job_code(const uint8_t *packed, const DpDev &j, int k)
{
	const int64_t g = j.g_start + (int64_t)j.dir * k;
	int b = packed[g >> 1] >> ((g & 1) * 4) & 0xf;
	if (j.comp) b = b < 4 ? 3 - b : b;
	return b;
}

// One CTA per chunk of <= PAIR_TRI triples of one problem: phase 1 evaluates the per-row splice / codon rules (nasw-sse.c:91-210)
// into shared memory with a halo, phase 2 combines them into the pair records, stored per parity of the triple index and
// field-major (six arrays of 16-byte fields) so that the 32 lanes of a warp -- which are on triples two apart -- read 32
// consecutive fields with each load.
__global__ void __launch_bounds__(256) nasw_prep_pair_kernel(const DpDev *jobs, const PrepChunk *chunks, int n_chunks, const uint8_t *packed, const uint8_t *ss, NaswConst cst, uint4 *rec)
{
	__shared__ uint32_t w[3 * PAIR_TRI + 8];
	const int ck = blockIdx.x;
	if (ck >= n_chunks) return;
	const PrepChunk c = chunks[ck];
	const DpDev job = jobs[c.job];
	const int m0 = c.row0, nt = c.n_rows, r0 = 3 * m0 - 3; // smem slot s <-> row r0 + s
	auto code = [&](int k) { return pair_job_code(packed, job, k); };
	auto sbyte = [&](int k) { // --spsc byte of row k's nucleotide (nasw_kernels.cu job_spsc)
		if (job.ss_off < 0) return -1;
		const int64_t g = job.ss_off + job.g_start + (int64_t)job.dir * k;
		return g == job.ss_excl ? 0xff : (int)ss[g];
	};
	const SpscPar sq = { (job.io + 1) / 2 - 1, cst.sp_null_bonus };
	for (int s = threadIdx.x; s < 3 * nt + 6; s += blockDim.x) {
		int r = r0 + s;
		r = r < 0 ? 0 : (r > job.nl ? job.nl : r);
		w[s] = (job.flag & NS_F_EXT_LEFT) ? prep_row_left(code, job.nl, r, cst.sp, cst.codon, cst.aa_x, sbyte, sq) : prep_row_forward(code, job.nl, r, cst.sp, cst.codon, cst.aa_x, sbyte, sq);
	}
	__syncthreads();
	const int nb = pair_rec_blocks(job.nl);
	uint4 *base = rec + job.rw_off * 2;
	for (int t = threadIdx.x; t < nt; t += blockDim.x) {
		const int m = m0 + t;
		const PairRec r = make_pair_rec([&](int k) { return w[k - r0]; }, m, job.io, cst.ge, cst.fs, 128, PAIR_PROF_HI);
		uint4 *dst = base + pair_rec_index(nb, m);
#pragma unroll
		for (int f = 0; f < 6; ++f) dst[f * 32] = make_uint4(r.w[4 * f], r.w[4 * f + 1], r.w[4 * f + 2], r.w[4 * f + 3]);
	}
}

void nasw_launch_prep_pair(cudaStream_t st, const DpDev *jobs, const PrepChunk *chunks, int n_chunks, const uint8_t *packed, const uint8_t *ss, const NaswConst &cst, int4 *rec)
{
	if (n_chunks > 0) nasw_prep_pair_kernel<<<n_chunks, 256, 0, st>>>(jobs, chunks, n_chunks, packed, ss, cst, (uint4*)rec);
}

struct PairEnvDev {
	uint32_t prof_base; // shared-window address of this thread's word in the low-half table of its warp
	__device__ __forceinline__ uint32_t prof(uint32_t off) const { return (uint32_t)lds32(prof_base + off); }
};

// record of triple T - 2p (clamped into the stored range: what a clamped index delivers belongs to rows that are masked out)
__device__ __forceinline__ void pair_load_rec(const uint4 *base, int nb, int m_max, int T, int p, PairRec &r)
{
	int m = T - 2 * p;
	m = m < 0 ? (m & 1) : (m > m_max ? m_max - ((m ^ m_max) & 1) : m); // keep the parity: the other lanes of the warp are on that parity's array
	const uint4 *q = base + pair_rec_index(nb, m);
#pragma unroll
	for (int f = 0; f < 6; ++f) {
		const uint4 v = __ldg(q + f * 32);
		r.w[4 * f] = v.x, r.w[4 * f + 1] = v.y, r.w[4 * f + 2] = v.z, r.w[4 * f + 3] = v.w;
	}
}

template <bool TB>
__global__ void __launch_bounds__(32) nasw_pair_kernel(const DpDev *jobs, const int *order, int n_jobs, const uint4 *rec_all, const char *aa, NaswConst cst, int4 *out,
                                                       uint16_t *tb)
{
	__shared__ uint32_t prof[2 * 22 * 32];       // low-half table, high-half table (22 amino acids x 32 threads each)
	__shared__ int ring[TB ? 32 : 32 * 32];      // [slot][lane] row maxima waiting for the warp tracker
	if ((int)blockIdx.x >= n_jobs) return;
	const int jid = order[blockIdx.x];
	const DpDev job = jobs[jid];
	const int lane = threadIdx.x;
	const int W8 = (job.al + 7) / 8 * 8, nl = job.nl, al = job.al, Wp = job.pad_;
	const int p = lane, col = 2 * p;             // this thread's column pair
	const bool live = col < W8;
	{ // profile of this thread's two columns
		int r0 = -1, r1 = -1;
		if (col < al) r0 = cst.aa20[(uint8_t)aa[job.aa_off + ((job.flag & NS_F_EXT_LEFT) ? al - 1 - col : col)]];
		if (col + 1 < al) r1 = cst.aa20[(uint8_t)aa[job.aa_off + ((job.flag & NS_F_EXT_LEFT) ? al - 2 - col : col + 1)]];
		for (int a = 0; a < 22; ++a) {
			prof[a * 32 + lane] = pk(r0 >= 0 ? cst.mat[a * 22 + r0] : PAIR_DEAD, 0);
			prof[22 * 32 + a * 32 + lane] = pk(0, r1 >= 0 ? cst.mat[a * 22 + r1] : PAIR_DEAD);
		}
	}
	__syncwarp();
	const int n_macro = pair_n_macro(nl, W8);
	PairPar pp;
	pp.go = cst.go, pp.ge = cst.ge, pp.fs = cst.fs, pp.end_bonus = cst.end_bonus, pp.ngo = pk2(-cst.go), pp.nfs = pk2(-cst.fs);
	PairGeo g;
	g.x = p, g.col = col, g.nl = nl, g.al = al, g.W8 = W8, g.first = p == 0;
	typename std::conditional<TB, PairLaneTb, PairLane>::type L;
	L.init(g, pp);
	PairEnvDev env;
	env.prof_base = smem_addr(&prof[lane]);
	const uint4 *rec = rec_all + job.rw_off * 2;
	const int nb = pair_rec_blocks(nl), m_max = 2 * pair_rec_slots(nl) - 1; // records 0 .. m_max exist
	const int p_end = W8 / 2 - 1;                // the pair (= lane) that owns the problem's last column
	WarpTracker trk;
	trk.init(PAIR_CB);
	const uint32_t ring_w = smem_addr(ring) + lane * 4, ring_r = smem_addr(ring) + lane * 128 + (uint32_t)p_end * 4;
	(void)ring_w, (void)ring_r;
	int tb_score = 0;
	bool have_score = false;
	const bool has_end_lo = live && col == al - 1, has_end_hi = live && col + 1 == al - 1;
	uint16_t *tbp = TB ? tb + job.tb_off + col : 0;

	uint32_t hb[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } }; // H of the columns to the left: this macro-step's (hb[PH]) and the previous one's
	PairRec rcs[2];                                   // row records of the next even / odd macro-step, fetched two steps ahead
	pair_load_rec(rec, nb, m_max, 0, p, rcs[0]);
	pair_load_rec(rec, nb, m_max, 1, p, rcs[1]);
#define NSW_PAIR_RECV(PH) \
		uint32_t rQ[3]; \
		_Pragma("unroll") for (int r = 0; r < 3; ++r) { \
			hb[PH][r] = L.left_of(__shfl_up_sync(0xffffffffu, L.oH[r], 1), L.oH[r]); \
			rQ[r] = L.left_of(__shfl_up_sync(0xffffffffu, L.oQ[r], 1), L.oQ[r]); \
		}
	// ---- general macro-step (ramp-up, ramp-down, tiny problems): every row checked, results committed per half
	auto step = [&](int T, auto ph_tag) {
		constexpr int PH = decltype(ph_tag)::value;
		const PairRec rc = rcs[PH];
		pair_load_rec(rec, nb, m_max, T + 2, p, rcs[PH]);
		NSW_PAIR_RECV(PH)
		const uint32_t *pv = hb[PH ^ 1], *cu = hb[PH];
		const int m = T - 2 * p;
		if constexpr (TB) {
			uint32_t rF[3], rS[3], wd[3];
#pragma unroll
			for (int r = 0; r < 3; ++r) {
				rF[r] = L.left_of(__shfl_up_sync(0xffffffffu, L.oF[r], 1), L.oF[r]);
				rS[r] = L.left_of(__shfl_up_sync(0xffffffffu, L.oS[r], 1), L.oS[r]);
			}
#pragma unroll
			for (int r = 0; r < 3; ++r) {
				const int i_lo = 3 * m + 2 + r, i_hi = i_lo - 3;
				const bool vlo = live && i_lo >= 2 && i_lo < nl, vhi = live && i_hi >= 2 && i_hi < nl;
				const uint32_t keep = (vlo ? 0xffffu : 0u) | (vhi ? 0xffff0000u : 0u);
				const bool bnd = g.first && i_lo == 2;
				const uint32_t l0 = cu[r], l1 = r == 0 ? pv[2] : cu[r - 1], l2 = r == 0 ? pv[1] : r == 1 ? pv[2] : cu[0], l3 = r == 0 ? pv[0] : r == 1 ? pv[1] : pv[2];
				if (r == 0) wd[0] = L.template row_masked<0>(pp, rc, env, l0, l1, l2, l3, rQ[0], rF[0], rS[0], keep, bnd);
				else if (r == 1) wd[1] = L.template row_masked<1>(pp, rc, env, l0, l1, l2, l3, rQ[1], rF[1], rS[1], keep, bnd);
				else wd[2] = L.template row_masked<2>(pp, rc, env, l0, l1, l2, l3, rQ[2], rF[2], rS[2], keep, bnd);
				uint16_t *q = tbp + (int64_t)(3 * T + r) * Wp;
				if (vlo && vhi) *reinterpret_cast<uint32_t*>(q) = wd[r];
				else if (vlo) q[0] = (uint16_t)(wd[r] & 0xffff);
				else if (vhi) q[1] = (uint16_t)(wd[r] >> 16);
				if (vlo && has_end_lo && i_lo == nl - 1) tb_score = lo16(L.oH[r]) - PAIR_BIAS, have_score = true;
				if (vhi && has_end_hi && i_hi == nl - 1) tb_score = hi16(L.oH[r]) - PAIR_BIAS, have_score = true;
			}
		} else {
			int lx[3], xp[3];
#pragma unroll
			for (int r = 0; r < 3; ++r) lx[r] = (int)((uint32_t)__shfl_up_sync(0xffffffffu, L.oXhi[r], 1) & L.xmask), xp[r] = L.oXlo[r];
#pragma unroll
			for (int r = 0; r < 3; ++r) {
				const int i_lo = 3 * m + 2 + r, i_hi = i_lo - 3;
				const bool vlo = live && i_lo >= 2 && i_lo < nl, vhi = live && i_hi >= 2 && i_hi < nl;
				const uint32_t keep = (vlo ? 0xffffu : 0u) | (vhi ? 0xffff0000u : 0u);
				const bool bnd = g.first && i_lo == 2;
				const uint32_t l0 = cu[r], l1 = r == 0 ? pv[2] : cu[r - 1], l2 = r == 0 ? pv[1] : r == 1 ? pv[2] : cu[0], l3 = r == 0 ? pv[0] : r == 1 ? pv[1] : pv[2];
				if (r == 0) L.template row_masked<0>(pp, rc, env, l0, l1, l2, l3, rQ[0], lx[0], xp[0], keep, bnd);
				else if (r == 1) L.template row_masked<1>(pp, rc, env, l0, l1, l2, l3, rQ[1], lx[1], xp[1], keep, bnd);
				else L.template row_masked<2>(pp, rc, env, l0, l1, l2, l3, rQ[2], lx[2], xp[2], keep, bnd);
			}
#pragma unroll
			for (int r = 0; r < 3; ++r) { // the last column's rows of this step, if real, go to the tracker
				const int i_hi = 3 * (T - 2 * p_end) - 1 + r;
				if (i_hi >= 2 && i_hi < nl) trk.push(ring_w, L.oXhi[r]);
			}
			if (trk.n_ring >= 30) trk.flush(ring_r, lane, al * 3, cst.pen, cst.xdrop);
		}
	};
	// steady macro-steps: both halves of every live thread are on real rows strictly above the last row
	int t_lo = 2 * p_end + 2, t_hi = nl >= 6 ? (nl - 6) / 3 + 1 : 0;
	if (t_lo > n_macro) t_lo = n_macro; // (both even)
	if (t_hi > n_macro) t_hi = n_macro;
	if (t_hi < t_lo) t_hi = t_lo;
	t_hi = t_lo + ((t_hi - t_lo) & ~1);
	int T = 0;
	for (; T < t_lo; T += 2) {
		step(T, std::integral_constant<int, 0>());
		step(T + 1, std::integral_constant<int, 1>());
		if (!TB && trk.stopped) break;
	}
	// The steady loop, straight-line: no row checks; the row records of step T + 2 are fetched AFTER the rows of step T used the
	// old ones (same registers, a whole macro-step to arrive).  Record m of a thread sits at slot k = m >> 1 of its parity, i.e.
	// at 16-byte index k + (k >> 5) * 160 of that parity's array (nasw_pair.cuh pair_rec_index); idle threads to the right of the
	// last column read the records of the last live pair.
#define NSW_PAIR_STEADY(PH) { \
		NSW_PAIR_RECV(PH) \
		const uint32_t *pv = hb[PH ^ 1], *cu = hb[PH]; \
		if constexpr (TB) { \
			uint32_t rF[3], rS[3]; \
			_Pragma("unroll") for (int r = 0; r < 3; ++r) { \
				rF[r] = L.left_of(__shfl_up_sync(0xffffffffu, L.oF[r], 1), L.oF[r]); \
				rS[r] = L.left_of(__shfl_up_sync(0xffffffffu, L.oS[r], 1), L.oS[r]); \
			} \
			const uint32_t w0 = L.template row<0>(pp, rcs[PH], env, cu[0], pv[2], pv[1], pv[0], rQ[0], rF[0], rS[0]); \
			const uint32_t w1 = L.template row<1>(pp, rcs[PH], env, cu[1], cu[0], pv[2], pv[1], rQ[1], rF[1], rS[1]); \
			const uint32_t w2 = L.template row<2>(pp, rcs[PH], env, cu[2], cu[1], cu[0], pv[2], rQ[2], rF[2], rS[2]); \
			if (live) { \
				uint32_t *q = reinterpret_cast<uint32_t*>(tbs); \
				q[0] = w0, q[Wp / 2] = w1, q[Wp] = w2; \
			} \
			tbs += 3 * Wp; \
		} else { \
			int lx[3], xp[3]; \
			_Pragma("unroll") for (int r = 0; r < 3; ++r) lx[r] = (int)((uint32_t)__shfl_up_sync(0xffffffffu, L.oXhi[r], 1) & L.xmask), xp[r] = L.oXlo[r]; \
			L.template row<0>(pp, rcs[PH], env, cu[0], pv[2], pv[1], pv[0], rQ[0], lx[0], xp[0]); \
			L.template row<1>(pp, rcs[PH], env, cu[1], cu[0], pv[2], pv[1], rQ[1], lx[1], xp[1]); \
			L.template row<2>(pp, rcs[PH], env, cu[2], cu[1], cu[0], pv[2], rQ[2], lx[2], xp[2]); \
			_Pragma("unroll") for (int r = 0; r < 3; ++r) sts32(ring_w + (trk.n_ring + r) * 128, L.oXhi[r]); \
			trk.n_ring += 3; \
			if (trk.n_ring >= 30) trk.flush(ring_r, lane, al * 3, cst.pen, cst.xdrop); \
		} \
		{ \
			const uint4 *q = rq[PH] + (ks + 1 + ((ks + 1) >> 5) * 160); \
			_Pragma("unroll") for (int f = 0; f < 6; ++f) { \
				const uint4 v = __ldg(q + f * 32); \
				rcs[PH].w[4 * f] = v.x, rcs[PH].w[4 * f + 1] = v.y, rcs[PH].w[4 * f + 2] = v.z, rcs[PH].w[4 * f + 3] = v.w; \
			} \
		} }
	if (T < t_hi && !(!TB && trk.stopped)) {
		uint16_t *tbs = TB ? tbp + (int64_t)3 * T * Wp : 0;
		(void)tbs;
		const int ps = p > p_end ? p_end : p;
		const uint4 *rq[2] = { rec, rec + (size_t)nb * 192 }; // even / odd triples (T is even at the loop head: step T + PH has parity PH)
		int ks = (T - 2 * ps) >> 1;                           // slot of step T (and of step T + 1)
		pair_load_rec(rec, nb, m_max, T, ps, rcs[0]);         // the two records in flight, for THIS mapping of threads to records
		pair_load_rec(rec, nb, m_max, T + 1, ps, rcs[1]);
		for (; T < t_hi; T += 2, ++ks) {
			if ((T & 14) == 0) { // every eighth iteration: a lane moves on by one 16-byte record per iteration, so its 128-byte lines 32 records ahead are due
				const uint4 *pf = rq[0] + (ks + 40 + ((ks + 40) >> 5) * 160);
#pragma unroll
				for (int f = 0; f < 6; ++f) {
					asm volatile("prefetch.global.L1 [%0];" :: "l"(pf + f * 32));
					asm volatile("prefetch.global.L1 [%0];" :: "l"(pf + (size_t)nb * 192 + f * 32));
				}
			}
			NSW_PAIR_STEADY(0)
			NSW_PAIR_STEADY(1)
			if (!TB && trk.stopped) break;
		}
		if (T < n_macro) { // back to the general loop: its records (clamped indices, own pair) for the next two steps
			pair_load_rec(rec, nb, m_max, T, p, rcs[0]);
			pair_load_rec(rec, nb, m_max, T + 1, p, rcs[1]);
		}
	}
#undef NSW_PAIR_STEADY
	if (!(!TB && trk.stopped)) {
		for (; T < n_macro; T += 2) {
			step(T, std::integral_constant<int, 0>());
			step(T + 1, std::integral_constant<int, 1>());
			if (!TB && trk.stopped) break;
		}
	}
#undef NSW_PAIR_RECV
	if (TB) {
		if (have_score) out[jid] = make_int4(tb_score, nl, al, 0);
	} else {
		if (trk.n_ring > 0 && !trk.stopped) trk.flush(ring_r, lane, al * 3, cst.pen, cst.xdrop);
		if (lane == 0) {
			int4 r;
			r.x = trk.max_i >= 0 ? trk.max_sc - PAIR_BIAS : INT32_MIN, r.y = trk.max_i + 1, r.z = trk.aa_len(al), r.w = 0;
			out[jid] = r;
		}
	}
}

// pair-lane kernels: one warp per problem of up to 64 padded columns
void nasw_launch_pair(cudaStream_t st, bool is_tb, const DpDev *jobs, const int *order, int n, const int4 *rec, const char *aa, const NaswConst &cst, int4 *out, uint16_t *tb)
{
	if (n <= 0) return;
	if (is_tb) nasw_pair_kernel<true><<<n, 32, 0, st>>>(jobs, order, n, (const uint4*)rec, aa, cst, out, tb);
	else nasw_pair_kernel<false><<<n, 32, 0, st>>>(jobs, order, n, (const uint4*)rec, aa, cst, out, tb);
}

} // namespace cuda
} // namespace mpb
