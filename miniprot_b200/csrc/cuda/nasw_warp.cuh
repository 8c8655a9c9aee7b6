// nasw_warp.cuh -- device-only helpers shared by the nasw kernel files: shared-memory access by 32-bit window address and the
// warp-parallel extension bookkeeping (x-drop tracker).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nasw_core.cuh"

namespace mpb {
namespace cuda {

using namespace nsw;

// shared-memory accesses through 32-bit shared-window addresses computed once (the generic form makes the compiler rebuild
// the window base -- S2R + LEA -- next to every access of the loop)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds32(uint32_t a)
{
	int v;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ void sts128(uint32_t a, int4 v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ int4 lds128(uint32_t a)
{
	int4 v;
	asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
	return v;
}

// Predicated 16-byte accesses: one thread of a block has something to do, nobody branches (a branch that one lane takes costs the
// whole warp a divergence region on the critical path of every macro-step; the column passes of a wide problem use these).
__device__ __forceinline__ void stg128_if(bool p, const void *a, int4 v)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.global.v4.b32 [%1], {%2, %3, %4, %5};\n\t}" :: "r"((int)p), "l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void ldcg128_if(bool p, const void *a, int4 &v) // v keeps its value when p is false
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t@q ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%5];\n\t}" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"((int)p), "l"(a));
}
__device__ __forceinline__ void sts128_if(bool p, uint32_t a, int4 v)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.shared.v4.b32 [%1], {%2, %3, %4, %5};\n\t}" :: "r"((int)p), "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 16 bytes global -> shared, asynchronously (LDGSTS through L2), predicated; groups are committed and awaited per thread
__device__ __forceinline__ void cp_async16_if(bool p, uint32_t smem, const void *g)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q cp.async.cg.shared.global [%1], [%2], 16;\n\t}" :: "r"((int)p), "r"(smem), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Extension bookkeeping of the block-wide kernel, 30 rows at a time.  ExtTracker::row (nasw_core.cuh) is the specification: a
// running maximum of (row best - length penalty) with first-occurrence ties, and a stop at the first row that falls more than
// xdrop below it.  Fed row by row it costs the warp that owns the last column ~15 divergent instructions per row on the
// critical path; here the last column only drops its row maxima into a ring and, once 30 or more wait there, the 32 lanes
// of that warp evaluate the rows together (prefix maximum by shuffles, ballots for the stop row and the winner).  The stop
// is noticed up to ten macro-steps late, which is harmless: rows after the stop row are never looked at.
struct WarpTracker {
	int max_sc, max_log, max_i, max_code; // warp-uniform
	bool stopped;
	int n_ring, i_base;                   // rows [i_base, i_base + n_ring) wait in the ring
	int pen, pk, next_thr;                // per lane: lane r follows the rows i_base + r of successive batches
	int cb;                               // width of the column code in a row maximum (nasw_core.cuh code_bits)
	__device__ __forceinline__ void init(int code_bits_) { max_sc = INT32_MIN, max_log = INT32_MIN, max_i = -1, max_code = 0, stopped = false, n_ring = 0, i_base = 2, pen = 0, pk = 0, next_thr = 2, cb = code_bits_; }
	__device__ __forceinline__ int aa_len(int al) const { return (max_i >= 0 && max_code != 0) ? ((1 << cb) - 1) - max_code + 1 : al + 1; }
	// the ring is [slot][lane]: every lane of the warp stores its own value (no divergent branch on the critical path) and
	// only lane 31's column -- the last column of the problem -- is read back.  ring_w = address of (slot 0, this lane),
	// ring_r = address of (slot = this lane, lane 31)
	__device__ __forceinline__ void push(uint32_t ring_w, int v) { sts32(ring_w + n_ring * 128, v), ++n_ring; }
	// ring_r = address of (slot = this lane, column of the thread that owns the problem's last column)
	__device__ __forceinline__ void flush(uint32_t ring_r, int lane, int pen_base, const PenTable &pt, int xdrop)
	{
		__syncwarp();
		if (!stopped) {
			const bool valid = lane < n_ring;
			const int i = i_base + lane, best = valid ? lds32(ring_r) : 0, x = i - pen_base;
			if (valid && x >= next_thr) {
				while (pk < pt.n && x >= pt.thr[pk]) pen = pt.val[pk], ++pk;
				next_thr = pk < pt.n ? pt.thr[pk] : INT32_MAX;
			}
			const int tsc = best >> cb, tlog = valid ? tsc - pen : INT32_MIN;
			int pm = tlog; // inclusive prefix maximum over the batch
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(0xffffffffu, pm, d);
				if (lane >= d) pm = max(pm, o);
			}
			const int run = max(pm, max_log);
			const unsigned sm = __ballot_sync(0xffffffffu, valid && run - tlog > xdrop);
			const int last = sm ? __ffs(sm) - 1 : n_ring - 1; // the last row that is still looked at
			const int bm = __shfl_sync(0xffffffffu, pm, last);
			if (bm > max_log) {
				const int w = __ffs(__ballot_sync(0xffffffffu, lane <= last && tlog == bm)) - 1;
				max_log = bm, max_sc = __shfl_sync(0xffffffffu, tsc, w), max_code = __shfl_sync(0xffffffffu, best & ((1 << cb) - 1), w), max_i = i_base + w;
			}
			stopped = sm != 0;
		}
		i_base += n_ring, n_ring = 0;
		__syncwarp();
	}
};

} // namespace cuda
} // namespace mpb
