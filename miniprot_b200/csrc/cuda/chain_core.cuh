// chain_core.cuh -- anchor chaining (reference chain.c) as host+device building blocks.
//
// The chaining kernels (chain_kernels.cu) use one warp per chaining problem for the score fill and one thread
// per problem for the inherently sequential backtrack; everything below is the arithmetic and the sequential
// logic they share with the CPU lock-step emulation in tests/hostcheck (test infrastructure only).
//
//   pair_score()      chain.c:112-151  gap-aware score of linking anchor j -> i (FP32 penalties, truncated)
//   resolve_chunk()   chain.c:191-205  the order-dependent part of the predecessor scan for 32 candidates at once:
//                                      records (new maxima), the max_skip counter and the early break
//   backtrack_compact() chain.c:26-110 ends sorted by score with the reference's tie order, best-first peeling,
//                                      chains reversed and ordered by target start
#pragma once
#include <stdint.h>
#include "../flagsort.hpp"

#ifdef __CUDACC__
#define CHN_HD __host__ __device__ __forceinline__
#else
#define CHN_HD inline
#endif

namespace chn {

struct Par { int32_t max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc; float chn_coef_log; int32_t is_spliced, kmer, bbit; };

// chain.c:160-174: parameter fix-ups done once per mp_chain call
CHN_HD Par normalise(Par p)
{
	if (p.max_dist_x < p.bw) p.max_dist_x = p.bw;
	if (p.max_dist_y < p.bw && !p.is_spliced) p.max_dist_y = p.bw;
	return p;
}

// mppriv.h:91-99, every FP32 operation rounded separately (the reference is built without FMA contraction)
CHN_HD float log2_approx(float x)
{
	union { float f; uint32_t i; } z;
	z.f = x;
	float r = (float)((int)((z.i >> 23) & 255) - 128);
	z.i &= ~(255u << 23);
	z.i += 127u << 23;
#ifdef __CUDA_ARCH__
	float q = __fadd_rn(__fmul_rn(-0.34484843f, z.f), 2.02466578f);
	q = __fsub_rn(__fmul_rn(q, z.f), 0.67487759f);
	return __fadd_rn(r, q);
#else
	volatile float q = -0.34484843f * z.f;
	q = q + 2.02466578f;
	q = q * z.f;
	q = q - 0.67487759f;
	volatile float s = r + q;
	return s;
#endif
}

CHN_HD int32_t pair_score(const Par &p, uint64_t ai, uint64_t aj)
{
	const int32_t dq = (int32_t)ai - (int32_t)aj, dq3 = dq * 3;
	int32_t dr3, dd, dds = 0, sc;
	if (dq <= 0 || dq3 > p.max_dist_x || dq > p.max_dist_y) return INT32_MIN;
	if (p.bbit > 0) {
		const int32_t bs = 1 << p.bbit;
		dr3 = (int32_t)(((ai >> 32) - (aj >> 32)) << p.bbit);
		if (dq3 < dr3 - bs) dd = dr3 - bs - dq3, dds = -dd;
		else if (dq3 > dr3 + bs) dd = dq3 - (dr3 + bs), dds = dd;
		else dd = 0;
	} else {
		dr3 = (int32_t)((ai >> 32) - (aj >> 32));
		if (dr3 == 0) return INT32_MIN;
		dd = dr3 > dq3 ? dr3 - dq3 : dq3 - dr3;
		dds = dq3 - dr3;
	}
	if (dd > p.bw) return INT32_MIN;
	if (p.bbit > 0) sc = p.kmer < dq ? p.kmer : dq;
	else if (p.kmer <= dq && p.kmer * 3 <= dr3) sc = p.kmer;
	else {
		const int32_t dr = dr3 / 3, g = dr < dq ? dr : dq;
		sc = g < p.kmer ? g : p.kmer;
		if (dr3 - dr * 3 != 0) --sc;
	}
	if (dd > 0) {
#ifdef __CUDA_ARCH__
		const float lin = __fmul_rn((float)dd, .33334f);
		const float lg = dd >= 2 ? __fadd_rn(__fmul_rn(p.chn_coef_log, __fsub_rn(log2_approx((float)(dd + 1)), 1.0f)), 1.0f) : (float)dd;
		if (p.is_spliced && dds < 0) sc -= (int32_t)(lin < lg ? lin : lg);
		else sc -= (int32_t)__fadd_rn(lin, lg);
#else
		volatile float lin = (float)dd * .33334f;
		volatile float lg = (float)dd;
		if (dd >= 2) { volatile float t = log2_approx((float)(dd + 1)) - 1.0f; t = p.chn_coef_log * t; lg = t + 1.0f; }
		if (p.is_spliced && dds < 0) sc -= (int32_t)(lin < lg ? lin : lg);
		else { volatile float s = lin + lg; sc -= (int32_t)s; }
#endif
	}
	if (p.bbit > 0 && ai >> 32 == aj >> 32 && dd == 0) sc += 2;
	return sc;
}

// 32 candidates j = jb, jb-1, ... (bit L = lane L = candidate jb-L) scanned in that order.
//   rec  : candidates whose score beats everything before them (running maximum, incl. the incoming max_f)
//   skip : non-record candidates already marked as "predecessor of something seen" (t[j] == i)
// Applies chain.c:196-203 and returns the lane at which the scan breaks (32 = no break); n_skip is updated.
CHN_HD int resolve_chunk(uint32_t rec, uint32_t skip, int32_t max_skip, int32_t &n_skip)
{
	if (skip == 0) { // only decrements
		const int32_t d = (int32_t)
#ifdef __CUDA_ARCH__
			__popc(rec);
#else
			__builtin_popcount(rec);
#endif
		n_skip = n_skip > d ? n_skip - d : 0;
		return 32;
	}
	uint32_t m = rec | skip;
	while (m) {
#ifdef __CUDA_ARCH__
		const int b = __ffs(m) - 1;
#else
		const int b = __builtin_ctz(m);
#endif
		if (rec >> b & 1) { if (n_skip > 0) --n_skip; }
		else if (++n_skip > max_skip) return b;
		m &= m - 1;
	}
	return 32;
}

// Sort records are packed into 8 bytes: key (score, or target coordinate) in the high 32 bits, anchor / chain index in
// the low 32.  The reference sorts 16-byte {x, y} records by x alone (radix_sort_mp128x); the resulting permutation
// depends only on the keys, so the packed form leaves ties in exactly the same places at half the traffic.
CHN_HD uint64_t rec_key(uint64_t r) { return r >> 32; }

// chain.c:8-24.  T is the mark array type (int32 in global memory, or int8 when it lives in shared memory).
template <class T, class FT, class PT>
CHN_HD int64_t bk_end(int32_t max_drop, uint64_t z, const FT *f, const PT *p, T *t)
{
	const int32_t zx = (int32_t)(z >> 32);
	int64_t i = (int64_t)(uint32_t)z, end_i = -1, max_i = i;
	int32_t max_s = 0;
	if (i < 0 || t[i] != 0) return i;
	do {
		t[i] = 2;
		end_i = i = p[i];
		const int32_t s = i < 0 ? zx : zx - (int32_t)f[i];
		if (s > max_s) max_s = s, max_i = i;
		else if (max_s - s > max_drop) break;
	} while (i >= 0 && t[i] == 0);
	for (i = (int64_t)(uint32_t)z; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
	return max_i;
}

// Sequential tail of mp_chain for ONE problem, after z[0..n_z) has been filled with (f<<32 | index) records in index
// order and t[0..n) cleared, in two parts.
//
// peel_chains (chain.c:26-75): best-first peeling, inherently sequential -- three dependent accesses per visited anchor, so
//   the kernels keep z/f/p/t in shared memory for it.  Scratch: t[n] marks, v[n], z[>=n], stack[>=CHAIN_STACK].
//   Output: u[0..n_u) = score<<32|cnt and v[0..n_v) = the chains' anchor indices, chain after chain.  Returns n_u.
//   PRESORTED: z[] has already been sorted by the caller (the warp-cooperative sort of the shared-memory kernels).
//   FT / PT: element types of the score and predecessor arrays (int32 in global memory; narrower copies in shared memory).
// compact_chains (chain.c:77-110): chains reversed to ascending order and ordered by first target coordinate.  The copy is
//   data parallel: `lane` of `n_lanes` threads share it (1 on the CPU and in the single-thread kernel, 32 in the warp
//   kernels), sync() is their barrier.  f[] is OVERWRITTEN (dead after the peeling, it serves as the chain-offset table).
//   Output: b[0..n_b) = compacted anchors, u[] permuted accordingly.
template <class T, bool PRESORTED = false, class FT = int32_t, class PT = int32_t>
CHN_HD int32_t peel_chains(const Par &p, int32_t n_z, const FT *f, const PT *pp, T *t, int32_t *v, uint64_t *z, mpb::FlagRange<uint64_t> *stack, uint64_t *u,
                           int32_t *n_v_out = 0)
{
	const int32_t max_drop = p.is_spliced ? INT32_MAX : p.bw;
	int32_t n_u = 0, n_v = 0;
	if (n_v_out) *n_v_out = 0;
	if (n_z == 0) return 0;
	if (!PRESORTED) mpb::flag_sort_by(z, z + n_z, [](const uint64_t &e) { return rec_key(e); }, stack);
	for (int32_t k = n_z - 1; k >= 0; --k) {
		const int32_t zi = (int32_t)(uint32_t)z[k], zx = (int32_t)(z[k] >> 32);
		// A score of kmer (the floor of the fill) means "no predecessor": such an end yields a one-anchor chain, which
		// min_cnt >= 2 rejects, and its mark can matter to nobody -- whoever chains through it scores higher and has been
		// peeled already.  The records are sorted, so everything from here down is of that kind (the bulk of a pre-chain
		// problem: anchors with nothing else in their block).
		if (zx <= p.kmer && p.min_cnt >= 2) break;
		if (t[zi] != 0) continue;
		const int32_t n_v0 = n_v;
		const int64_t end_i = bk_end(max_drop, z[k], f, pp, t);
		int64_t i;
		for (i = zi; i != end_i; i = pp[i]) v[n_v++] = (int32_t)i, t[i] = 1;
		const int32_t sc = i < 0 ? zx : zx - (int32_t)f[i];
		if (sc >= p.min_sc && n_v > n_v0 && n_v - n_v0 >= p.min_cnt) u[n_u++] = (uint64_t)sc << 32 | (uint64_t)(n_v - n_v0);
		else n_v = n_v0;
	}
	if (n_v_out) *n_v_out = n_v;
	return n_u;
}

template <class FT, class Sync>
CHN_HD void compact_chains(int32_t n_u, const uint64_t *a, FT *f, const int32_t *v, uint64_t *z, mpb::FlagRange<uint64_t> *stack, uint64_t *u, uint64_t *b,
                           int32_t *n_b_out, int lane, int n_lanes, Sync sync)
{
	if (n_u == 0) {
		if (lane == 0) *n_b_out = 0;
		return;
	}
	if (lane == 0) {
		int32_t k = 0;
		for (int32_t i = 0; i < n_u; ++i) {
			const int32_t ni = (int32_t)(uint32_t)u[i];
			z[i] = (a[v[k + ni - 1]] >> 32) << 32 | (uint32_t)i; // first anchor of the chain after reversal
			f[i] = (FT)k;
			k += ni;
		}
		mpb::flag_sort_by(z, z + n_u, [](const uint64_t &e) { return rec_key(e); }, stack);
	}
	sync();
	int32_t o = 0;
	for (int32_t i = 0; i < n_u; ++i) {
		const int32_t src = (int32_t)(uint32_t)z[i], k0 = (int32_t)f[src], ni = (int32_t)(uint32_t)u[src];
		for (int32_t j = lane; j < ni; j += n_lanes) b[o + j] = a[v[k0 + (ni - j - 1)]];
		o += ni;
	}
	sync();
	for (int32_t i = lane; i < n_u; i += n_lanes) z[i] = u[(uint32_t)z[i]]; // u2[i] = u[perm[i]]
	sync();
	for (int32_t i = lane; i < n_u; i += n_lanes) u[i] = z[i];
	if (lane == 0) *n_b_out = o;
}

// both parts by one thread
template <class T, bool PRESORTED = false, class FT = int32_t, class PT = int32_t>
CHN_HD int32_t peel_and_compact(const Par &p, int32_t n_z, const uint64_t *a, FT *f, const PT *pp, T *t, int32_t *v, uint64_t *z,
                                mpb::FlagRange<uint64_t> *stack, uint64_t *u, uint64_t *b, int32_t *n_b_out)
{
	const int32_t n_u = peel_chains<T, PRESORTED, FT, PT>(p, n_z, f, pp, t, v, z, stack, u);
	compact_chains(n_u, a, f, v, z, stack, u, b, n_b_out, 0, 1, [] {});
	return n_u;
}

// single-threaded form used by the CPU emulation and by problems too large for shared memory
template <class T>
CHN_HD int32_t backtrack_compact(const Par &p, int32_t n, const uint64_t *a, int32_t *f, const int32_t *pp, T *t, int32_t *v, uint64_t *z,
                                 mpb::FlagRange<uint64_t> *stack, uint64_t *u, uint64_t *b, int32_t *n_b_out)
{
	int32_t n_z = 0;
	for (int32_t i = 0; i < n; ++i) {
		t[i] = 0;
		if (f[i] >= p.min_sc) z[n_z++] = (uint64_t)(uint32_t)f[i] << 32 | (uint32_t)i;
	}
	return peel_and_compact(p, n_z, a, f, pp, t, v, z, stack, u, b, n_b_out);
}

} // namespace chn
