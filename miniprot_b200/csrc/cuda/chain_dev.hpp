// chain_dev.hpp -- launcher prototypes of the chaining kernels (chain_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "chain_core.cuh"

namespace mpb {
namespace cuda {

// pending ranges of the flag sort per problem: the sort works depth first, every pass pops one range and pushes at most 255
// (its buckets of more than 64 items), and a 64-bit key has 8 digits: 8 * 255 + 1 is the worst case for any input size
constexpr int CHAIN_STACK = 2048;

// score fill of the problems list[0..n_prob) (list == NULL: problems 0..n_prob-1), per-anchor state in global memory
void chain_launch_fill(cudaStream_t st, const int32_t *list, const int64_t *a_off, const int32_t *cnt, const uint64_t *a, int n_prob, const chn::Par &par, int32_t *f,
                       int32_t *p, int32_t *t);
// fill + backtrack + compaction of problems of at most `cap` (< 32768) anchors, everything in shared memory
void chain_launch_smem(cudaStream_t st, const int32_t *list, int n_list, int cap, const int64_t *a_off, const int32_t *cnt, const uint64_t *a, const chn::Par &par,
                       int32_t *v, void *stack, uint64_t *u, uint64_t *b, int32_t *n_u, int32_t *n_b, int resort);
// backtrack + compaction after a global-memory fill: one warp per problem, records / marks / 16-bit copies of f and p in
// shared memory (13 B per anchor)
void chain_launch_bt_smem(cudaStream_t st, const int32_t *list, int n_list, int cap, const int64_t *a_off, const int32_t *cnt, const uint64_t *a, const chn::Par &par,
                          int32_t *f, const int32_t *p, int32_t *v, void *stack, uint64_t *u, uint64_t *b, int32_t *n_u, int32_t *n_b, int resort);
// backtrack + compaction for the largest problems: one thread each, global memory
void chain_launch_bt(cudaStream_t st, const int32_t *list, int n_list, const int64_t *a_off, const int32_t *cnt, const uint64_t *a, const chn::Par &par, int32_t *f,
                     const int32_t *p, int32_t *t, int32_t *v, void *z, void *stack, uint64_t *u, uint64_t *b, int32_t *n_u, int32_t *n_b, int resort);

} // namespace cuda
} // namespace mpb
