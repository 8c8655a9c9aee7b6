// tests/hostcheck/emu_chain.cpp -- TEST INFRASTRUCTURE ONLY.
// CPU emulation of the chaining kernels: the warp-parallel predecessor scan of chain_fill_kernel is replayed with
// explicit 32-lane arrays around the SAME building blocks (chain_core.cuh: pair_score, resolve_chunk,
// backtrack_compact) the kernels compile, and compared with the oracle by tests/test_emu_chain.py.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "cuda/chain_core.cuh"

using namespace chn;

extern "C" int emu_chain(const Par *par_, int32_t n, const uint64_t *a, int resort, uint64_t *u_out, uint64_t *b_out, int32_t *n_b_out)
{
	const Par par = normalise(*par_);
	std::vector<int32_t> f((size_t)n), p((size_t)n), t((size_t)n, 0), v((size_t)n);
	int32_t st = 0, hi = -1, hf = 0;
	for (int32_t i = 0; i < n; ++i) {
		const uint64_t ai = a[i];
		const int64_t xi = (int64_t)(ai >> 32);
		while (st < i && ((xi - (int64_t)(a[st] >> 32)) << par.bbit) > par.max_dist_x) ++st;
		int32_t max_f = par.kmer, max_j = -1, n_skip = 0;
		if (hi >= 0 && hi >= st) {
			const int32_t sc = hf + pair_score(par, ai, a[hi]);
			if (sc > max_f) max_f = sc, max_j = hi;
		} else hf = 0, hi = -1;
		if (i - st > par.max_iter) st = i - par.max_iter;
		for (int32_t jb = i - 1; jb >= st; jb -= 32) {
			int32_t sc[32], pj[32];
			bool ok[32], marked[32];
			for (int l = 0; l < 32; ++l) {
				const int32_t j = jb - l;
				ok[l] = j >= st, sc[l] = INT32_MIN, pj[l] = -1;
				if (ok[l]) {
					sc[l] = pair_score(par, ai, a[j]);
					ok[l] = sc[l] != INT32_MIN;
					if (ok[l]) sc[l] += f[(size_t)j], pj[l] = p[(size_t)j];
				}
			}
			for (int l = 0; l < 32; ++l) if (ok[l] && pj[l] >= 0) t[(size_t)pj[l]] = i; // all marks first ...
			for (int l = 0; l < 32; ++l) marked[l] = ok[l] && t[(size_t)(jb - l)] == i;  // ... then all reads (__syncwarp between)
			uint32_t R = 0, S = 0;
			int32_t run = max_f;
			for (int l = 0; l < 32; ++l) { // prefix maximum in scan order
				const bool rec = ok[l] && sc[l] > run;
				if (rec) R |= 1u << l;
				else if (ok[l] && marked[l]) S |= 1u << l;
				if (ok[l] && sc[l] > run) run = sc[l];
			}
			const int brk = resolve_chunk(R, S, par.max_skip, n_skip);
			const uint32_t Rb = brk >= 32 ? R : (R & ((1u << brk) - 1u));
			if (Rb) {
				const int top = 31 - __builtin_clz(Rb);
				max_f = sc[top], max_j = jb - top;
			}
			if (brk < 32) break;
		}
		f[(size_t)i] = max_f, p[(size_t)i] = max_j;
		if (hf < max_f) hf = max_f, hi = i;
	}
	std::vector<uint64_t> z((size_t)n + 1);
	std::vector<mpb::FlagRange<uint64_t>> stack(8 * 256 + 8);
	int32_t n_b = 0, n_u = 0;
	if (n > 0) n_u = backtrack_compact(par, n, a, f.data(), p.data(), t.data(), v.data(), z.data(), stack.data(), u_out, b_out, &n_b);
	if (resort && n_b > 1) mpb::flag_sort_by(b_out, b_out + n_b, [](const uint64_t &x) { return x; }, stack.data());
	*n_b_out = n_b;
	return n_u;
}
