/*
 * oracle/ora.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar restatement of the reference's hot-path algorithms (lh3/miniprot 0.18-r281).
 * It exists to CHECK the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load liboracle.so.  The product library never links or calls it.
 *
 * Parity pinning: every function here is compared against the compiled reference
 * (oracle/_ref/libref.so, built by oracle/Makefile from /root/reference) on seeded random
 * inputs by tests/test_oracle_pin.py; the reference ships no golden vectors of its own
 * (SURVEY.md section 4), so the reference binary is the pin.
 *
 * The functions take the character/codon tables as arguments instead of owning a copy, so the
 * oracle carries no data of its own: tests pass the reference's tables (ref_ns_tab_*) or the
 * product's (which are themselves checked against the reference's).
 */
#ifndef ORA_H
#define ORA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	const uint8_t *nt4;     /* [256] char/code -> 0..4                       (nasw-tab.c:93-95)  */
	const uint8_t *aa20;    /* [256] char/code -> 0..21                      (nasw-tab.c:96-98)  */
	const uint8_t *aa13;    /* [256] char/code -> reduced 4-bit alphabet     (nasw-tab.c:99-101) */
	const uint8_t *codon;   /* [64]  codon -> aa20 code (20 = stop)          (nasw-tab.c:102-105)*/
	const uint8_t *codon13; /* [64]  codon -> reduced alphabet                                    */
} ora_tab_t;

/* ---- nasw (nasw-sse.c:340 ns_global_gs16b) ------------------------------------------- */
typedef struct {
	int32_t flag;            /* 1 CIGAR, 2 EXT_LEFT, 4 EXT_RIGHT (nasw.h:46-48) */
	int32_t go, ge, io, fs, xdrop, end_bonus;
	int32_t sp[6], sp_null_bonus;
	float ie_coef;
	const int8_t *mat;       /* 22x22 */
} ora_nasw_par_t;

typedef struct {
	int32_t score, nt_len, aa_len;
	int32_t n_cigar, m_cigar;
	uint32_t *cigar;         /* malloc'ed; caller frees */
} ora_nasw_rst_t;

void ora_nasw(const ora_tab_t *tab, const ora_nasw_par_t *par, const uint8_t *ns, int32_t nl,
              const char *as, int32_t al, const uint8_t *ss, ora_nasw_rst_t *r);
/* debugging aid: when non-NULL receives the final H rows / traceback words, row-major [nl][8*slen] */
extern int16_t *ora_nasw_dbg_H;
extern uint16_t *ora_nasw_dbg_tb;

/* ---- sorts (ksort.h:112-162 via misc.c:4-8) -------------------------------------------- */
typedef struct { uint64_t x, y; } ora128_t;
void ora_sort64(uint64_t *beg, uint64_t *end);           /* radix_sort_mp64: any correct sort is equivalent */
void ora_sort128x(ora128_t *beg, ora128_t *end);         /* radix_sort_mp128x: key .x only, UNSTABLE, order restated exactly */

/* ---- sketch (sketch.c) ------------------------------------------------------------------ */
uint32_t ora_hash32_mask(uint32_t key, uint32_t mask);   /* sketch.c:7 */
/* sketch.c:18; returns number of seeds written to out (capacity >= len) */
int32_t ora_sketch_prot(const ora_tab_t *tab, const char *seq, int32_t len, int32_t kmer, int32_t mod_bit, uint64_t *out);
/* sketch.c:62; out must hold up to len entries; returns count after sort+dedup */
int64_t ora_sketch_nt4(const ora_tab_t *tab, const uint8_t *seq, int64_t len, int32_t min_aa_len, int32_t kmer,
                       int32_t mod_bit, int32_t bbit, int64_t boff, uint64_t *out);

/* ---- seed lookup (map.c:126-177) ---------------------------------------------------------- */
/* returns malloc'ed sorted anchors (block<<32|qpos) for one protein; *n_a receives the count */
uint64_t *ora_seed_anchors(const ora_tab_t *tab, const int64_t *ki, int64_t n_kb, const uint32_t *kb, int32_t kmer,
                           int32_t mod_bit, int32_t max_occ_cap, const char *seq, int32_t len, int64_t *n_a);

/* ---- chaining (chain.c:160 mp_chain) ------------------------------------------------------- */
typedef struct {
	int32_t max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc;
	float chn_coef_log;
	int32_t is_spliced, kmer, bbit;
} ora_chain_par_t;
/* a[n] sorted anchors (not modified).  Returns malloc'ed compacted anchors (NULL if no chain);
 * *u_ receives malloc'ed score<<32|cnt per chain, *n_u_ the number of chains. */
uint64_t *ora_chain(const ora_chain_par_t *par, int64_t n, const uint64_t *a, int32_t *n_u_, uint64_t **u_);
int32_t ora_comput_sc(const ora_chain_par_t *par, uint64_t ai, uint64_t aj); /* chain.c:112 */

/* ---- second-round refinement core (map.c:41-97) ------------------------------------------- */
/* window nt codes (already strand-oriented) + protein -> best chain.  Returns malloc'ed anchors of
 * the best chain (ntpos<<32|aapos, window-relative), *n_best its length, *sc_best its score;
 * NULL when no chain. */
uint64_t *ora_refine(const ora_tab_t *tab, const ora_chain_par_t *par, int32_t min_aa_len, int32_t max_ava,
                     const uint8_t *nt, int64_t l_nt, const char *aa, int32_t l_aa, int32_t *n_best, int32_t *sc_best);

#ifdef __cplusplus
}
#endif
#endif
