/*
 * nasw_b200.h -- C ABI of the splice/frameshift-aware protein-to-DNA DP ("nasw") as served
 * by libminiprot_b200.so.
 *
 * Every entry point below replaces the reference symbol of the same name, so a program
 * compiled against the reference's nasw.h links against this library unchanged.  Struct
 * layouts are part of the ABI and mirror reference nasw.h:61-78 field for field.
 *
 *   ns_make_tables      <- reference nasw.h:94   (nasw-tab.c:85)
 *   ns_opt_init         <- reference nasw.h:101  (nasw-tab.c:131)
 *   ns_opt_set_sp       <- reference nasw.h:134  (nasw-tab.c:124)
 *   ns_set_stop_sc      <- reference nasw.h:139  (nasw-tab.c:149)
 *   ns_global_gs16      <- reference nasw.h:128  (nasw-sse.c:553)
 *   ns_global_gs16b     <- reference nasw.h:131  (nasw-sse.c:340)   ** the DP hot kernel **
 *
 * The DP itself runs on the GPU (sm_100a kernels in miniprot_b200/csrc/cuda/nasw_*.cu).
 * There is no CPU fallback: without a CUDA device the call aborts with a message.
 */
#ifndef NASW_B200_H
#define NASW_B200_H

#include <stdint.h>

/* CIGAR operators: low 4 bits of a CIGAR word, length in the upper 28 (reference nasw.h:33-44) */
enum {
	NS_CIGAR_M = 0,  /* codon vs residue                         */
	NS_CIGAR_I = 1,  /* residue(s) without codon                 */
	NS_CIGAR_D = 2,  /* codon(s) without residue; length in aa   */
	NS_CIGAR_N = 3,  /* phase-0 intron                           */
	NS_CIGAR_F = 10, /* frameshift: 1-2 nt consumed, no residue  */
	NS_CIGAR_G = 11, /* frameshift: 1-2 nt consumed, one residue */
	NS_CIGAR_U = 12, /* phase-1 intron                           */
	NS_CIGAR_V = 13  /* phase-2 intron                           */
};
#define NS_CIGAR_STR "MIDNSHP=XBFGUVE"

/* ns_opt_t::flag (reference nasw.h:46-48) */
#define NS_F_CIGAR     0x1 /* global alignment with traceback */
#define NS_F_EXT_LEFT  0x2 /* score-only extension to the left (sequences reversed internally) */
#define NS_F_EXT_RIGHT 0x4 /* score-only extension to the right */

/* splice models for ns_opt_set_sp (reference nasw.h:50-52) */
#define NS_S_NONE    0
#define NS_S_GENERIC 1
#define NS_S_MAMMAL  2

#define NS_SPSC_OFFSET 64

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int32_t flag;
	int32_t go, ge, io, fs;   /* gap open / extend, intron open, frameshift (all positive penalties) */
	int32_t xdrop, end_bonus; /* extension only */
	int32_t asize;            /* alphabet size, 22 */
	int32_t sp[6];            /* splice penalties: 0 GT+!R / AG+!Y, 1 GC-AG, 2 AT-AC, 3 other, 4 !G before GTR, 5 poly-Y */
	int32_t sp_null_bonus;
	float ie_coef;            /* extension length penalty: ie_coef * log2(nt - 3*aa) */
	const int8_t *sc;         /* asize x asize substitution matrix */
	uint8_t *nt4, *aa20, *codon;
} ns_opt_t;

typedef struct {
	int32_t n_cigar, m_cigar;
	int32_t nt_len, aa_len;
	int32_t score;
	uint32_t *cigar; /* malloc'ed (the km argument is accepted for ABI compatibility and ignored) */
} ns_rst_t;

extern char *ns_tab_nt_i2c, *ns_tab_aa_i2c;
extern uint8_t ns_tab_a2r[22], ns_tab_nt4[256], ns_tab_aa20[256], ns_tab_aa13[256];
extern uint8_t ns_tab_codon[64], ns_tab_codon13[64];
extern int8_t ns_mat_blosum62[484];

/* Build the char->code and codon tables for an NCBI genetic code; 0 on success, <0 if undefined. */
int ns_make_tables(int codon_type);
void ns_opt_init(ns_opt_t *opt);
void ns_opt_set_sp(ns_opt_t *opt, int32_t model);
void ns_set_stop_sc(int32_t asize, int8_t *mat, int8_t score);

/* One DP problem; dispatched to the GPU as a batch of one (use mpb_nasw_batch for real batches).
 * Memory contract: the reference allocates r->cigar from the kalloc arena `km` and its callers kfree(km, r->cigar) (nasw.h:77,
 * align.c:77).  Here `km` is accepted and IGNORED: r->cigar is always malloc'ed and is released with free() (or mpb_free()).
 * That is what kfree(0, ptr) does in the reference, so callers that pass km = NULL -- example code, tests -- need no change;
 * a caller that passes a real arena must switch the release to free().
 * ss: per-base splice-score bytes of the slice (--spsc, ntseq.c:130-156; NULL = none), applied like nasw-sse.c:138-152,189-203.
 * Preconditions: opt->go >= 1 (with a gap open penalty of 0 the reference's lazy-F loop, nasw-sse.c:408-422 / 521-537, stops at
 * once and its result depends on the SSE stripe layout) and an ie_coef whose length penalty fits the kernels'
 * step table (<= ~5).  A call that violates them prints a message and aborts (the function has no way to report an error;
 * mpb_nasw_batch() returns -3 instead).
 * ns_global_gs32 / ns_global_gs32b (reference nasw.h:129,132; never called by miniprot) are NOT exported: see DESIGN.md section 8. */
void ns_global_gs16(void *km, const char *ns, int32_t nl, const char *as, int32_t al, const ns_opt_t *opt, ns_rst_t *r);
void ns_global_gs16b(void *km, const char *ns, int32_t nl, const char *as, int32_t al, const ns_opt_t *opt, const uint8_t *ss, ns_rst_t *r);

#ifdef __cplusplus
}
#endif
#endif
