// seed_dev.hpp -- descriptors and launcher prototypes of the seeding / window-join kernels (seed_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct mpb_ctx_s;

namespace mpb {
namespace cuda {

constexpr int SEED_THREADS = 128;
constexpr int WIN_TILE = 2048;                 // window positions per shared-memory tile
constexpr int WIN_SMEM_SPAN = WIN_TILE + 256;  // + halos of 3*(min_aa_len+1) on the left, 3*min_aa_len on the right
constexpr int WIN_MAX_MIN_AA = 40;             // largest min_aa_len the halos cover

struct SeedConst {          // passed by value (constant bank)
	uint8_t aa13[256];      // residue char -> 4-bit reduced alphabet (>= 14: stop / unknown)
	uint8_t codon[64];      // codon -> amino acid (20 = stop)
	uint8_t codon13[64];    // codon -> reduced alphabet
	int32_t kmer, mod_bit, max_occ;
	int64_t n_kb;
};

struct WinJob {             // one refinement window
	int64_t g_start;        // nibble index of window position 0
	int32_t dir, comp;
	int64_t len;
	int32_t qid, pad_;
	int64_t grp_off;        // per-job group counters live at grp[grp_off .. grp_off + n_pk[qid])
};

void seed_launch_sketch(cudaStream_t st, const char *aa, const int32_t *aa_off, int n_q, const SeedConst &cst, const int64_t *ki, uint32_t *sd_hash, int32_t *sd_pos,
                        int64_t *sd_cnt, int64_t *sd_aoff, int32_t *n_sd, int64_t *tot);
void seed_launch_expand(cudaStream_t st, const int32_t *aa_off, int n_q, const int64_t *ki, const uint32_t *kb, const uint32_t *sd_hash, const int32_t *sd_pos,
                        const int64_t *sd_cnt, const int64_t *sd_aoff, const int32_t *n_sd, const int64_t *a_off, uint64_t *a);
void seed_launch_prot_kmer(cudaStream_t st, const char *aa, const int32_t *aa_off, int n_q, const SeedConst &cst, int kmer, uint64_t *keys, int32_t *n_out);
void win_launch_count(cudaStream_t st, const WinJob *jobs, int n_jobs, const uint8_t *packed, const SeedConst &cst, int kmer, int min_aa_len, int max_ava,
                      const uint64_t *pk, const int32_t *aa_off, const int32_t *n_pk, int32_t *grp, int64_t *n_a);
void win_launch_emit(cudaStream_t st, const WinJob *jobs, int n_jobs, const uint8_t *packed, const SeedConst &cst, int kmer, int min_aa_len, const uint64_t *pk,
                     const int32_t *aa_off, const int32_t *n_pk, const int32_t *grp, const int64_t *a_off, uint64_t *a);

// segmented ascending sort of 64-bit keys (device-wide primitive; see seg_sort.cu)
// segmented ascending sort of 64-bit keys, in place (seg_sort.cu); the segment bounds are host arrays
void seg_sort_u64(mpb_ctx_s *ctx, cudaStream_t st, uint64_t *keys, uint64_t *tmp, int n_seg, const int64_t *h_begin, const int64_t *h_end);

} // namespace cuda
} // namespace mpb
