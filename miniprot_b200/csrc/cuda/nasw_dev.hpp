// nasw_dev.hpp -- device-side job descriptors of the nasw stage and the launcher prototypes (nasw_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nasw_core.cuh"

#ifndef NS_F_CIGAR
#define NS_F_CIGAR 0x1
#define NS_F_EXT_LEFT 0x2
#define NS_F_EXT_RIGHT 0x4
#endif

namespace mpb {
namespace cuda {

constexpr int NASW_WARPS = 4;   // problems per CTA (one warp each)
constexpr int NASW_CMAX = 8;
constexpr int PREP_ROWS = 4096; // rows per prep CTA    // columns per lane in the widest instantiation (32*8 = 256 columns per pass)

struct DpDev {                  // one DP problem, resident in HBM for the duration of a wave
	int64_t g_start;            // nibble index (packed genome) of DP row 0
	int32_t dir, comp;          // +1/-1 walk direction; 1 = complement bases (minus strand)
	int32_t nl, al;
	int32_t aa_off;             // first residue of the protein slice in the batch residue buffer
	int32_t flag, io;
	int32_t C;                  // column-pass kernels: columns per lane (1, 2, 4, 8); 0 = block-wide wavefront kernel
	int64_t rw_off;             // row records (32 B each): nl + 1 entries, offset in rows
	int64_t tb_off;             // traceback words (uint16 units); tb problems only
	int64_t cig_off;            // CIGAR slot
	int32_t cig_cap;
	int32_t pad_;               // block-wide wavefront: traceback row width (32 * warps per problem)
	int64_t carry_off;          // per-row carry between column passes (int units)
	int64_t ss_off;             // --spsc: byte of DP row k = ss[ss_off + g_start + dir * k]; < 0 = the problem has no splice bytes
	int64_t ss_excl;            // ... except at this index, which reads 0xff (the first position of the region's window, ntseq.c:130-156); -1 = none
};

struct PrepChunk { int32_t job, row0, n_rows, pad_; };

struct NaswConst {              // problem-independent parameters, passed by value (constant bank)
	int8_t mat[484];            // 22 x 22 substitution matrix
	uint8_t aa20[256];
	uint8_t codon[64];
	int32_t sp[6];
	int32_t go, ge, fs, xdrop, end_bonus;
	float ie_coef;
	int32_t aa_x;               // code of 'X'
	int32_t sp_null_bonus;      // --spsc0 (nasw-sse.c:143-145)
	nsw::PenTable pen;          // extension length penalty as a step table (nasw-sse.c:426, FP32 done on the host)
};

// pair-lane family (nasw_pair_kernels.cu): chunks carry {job, first triple, number of triples}
void nasw_launch_prep_pair(cudaStream_t st, const DpDev *jobs, const PrepChunk *chunks, int n_chunks, const uint8_t *packed, const uint8_t *ss, const NaswConst &cst, int4 *rec);
void nasw_launch_pair(cudaStream_t st, bool is_tb, const DpDev *jobs, const int *order, int n, const int4 *rec, const char *aa, const NaswConst &cst, int4 *out,
                      uint16_t *tb);
void nasw_launch_prep(cudaStream_t st, const DpDev *jobs, const PrepChunk *chunks, int n_chunks, const uint8_t *packed, const uint8_t *ss, const NaswConst &cst, int4 *rec);
void nasw_launch_ext(cudaStream_t st, int C, const DpDev *jobs, const int *order, int n, const int4 *rec, const char *aa, const NaswConst &cst, int4 *out, int *carry);
void nasw_launch_tb(cudaStream_t st, int C, const DpDev *jobs, const int *order, int n, const int4 *rec, const char *aa, const NaswConst &cst, int4 *out, int *carry,
                    uint16_t *tb);
void nasw_launch_v3(cudaStream_t st, int nw, bool is_tb, const DpDev *jobs, const int *order, int n, const int4 *rec, const char *aa, const NaswConst &cst, int4 *out,
                    uint16_t *tb, int warps_per_sm = 0, int *carry = 0, bool multi = false, const int2 *units = 0, int *progress = 0);
// packs the CIGARs of a wave (each written at the end of its own worst-case slot) back to back in job order, so that the
// device-to-host copy moves what was produced instead of the slots; offs[] (n + 1 entries) is scratch
void nasw_launch_pack(cudaStream_t st, const DpDev *jobs, int n, const int4 *out, const uint32_t *cigar, int64_t *offs, uint32_t *packed);
// keeps `st` busy for about `us` microseconds (a head start for whatever does not wait behind it)
void nasw_launch_spacer(cudaStream_t st, int us);
void nasw_launch_bt(cudaStream_t st, const DpDev *jobs, const int *order, int n, const uint16_t *tb, uint32_t *cigar, int4 *out);

} // namespace cuda
} // namespace mpb
