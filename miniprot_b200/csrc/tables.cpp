// tables.cpp -- character / codon / scoring tables, option defaults and small runtime helpers.
//
// Replaces (same exported symbols, same values): reference nasw-tab.c (tables, ns_make_tables,
// ns_opt_init, ns_opt_set_sp, ns_set_stop_sc), options.c (mp_idxopt_init, mp_mapopt_init,
// mp_mapopt_set_fs, mp_mapopt_set_max_intron, mp_mapopt_check), misc.c:10-16 (mp_verbose,
// mp_dbg_flag, mp_start) and sys.c:93-127 (timers).  tests/test_tables.py compares every table and
// default with the compiled reference.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <sys/resource.h>
#include <sys/time.h>
#include "internal.hpp"

extern "C" {

int32_t mp_verbose = 3, mp_dbg_flag = 0;

static char g_nt_i2c[] = "ACGTN";
static char g_aa_i2c[] = "ARNDCQEGHILKMFPSTWYV*X";
char *ns_tab_nt_i2c = g_nt_i2c, *ns_tab_aa_i2c = g_aa_i2c;

// 20 amino acids + stop + X -> 4-bit reduced alphabet used for seeding (nasw-tab.c:12):
// A0 ST1 RK2 H3 ND4 EQ5 C6 P7 G8 IV10 LM11 FY12 W13, '*' = 14, X = 15
uint8_t ns_tab_a2r[22] = { 0, 2, 4, 4, 6, 5, 5, 8, 3, 10, 11, 2, 11, 12, 7, 1, 1, 13, 12, 10, 14, 15 };
uint8_t ns_tab_nt4[256], ns_tab_aa20[256], ns_tab_aa13[256], ns_tab_codon[64], ns_tab_codon13[64];
int8_t ns_mat_blosum62[484];

} // extern "C"

namespace {

// NCBI genetic codes, in NCBI's own published layout (first/second/third base in T,C,A,G order).
// Ids without an entry are undefined (ns_make_tables returns -2), as in the reference (nasw-tab.c:16-55).
struct GenCode { int id; const char *tcag; };
const GenCode kGenCodes[] = {
	{ 1, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{ 2, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG"},
	{ 3, "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{ 4, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{ 5, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG"},
	{ 6, "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{ 9, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
	{10, "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{11, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{12, "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{13, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG"},
	{14, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
	{15, "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{16, "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{21, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
	{22, "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{23, "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{24, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG"},
	{25, "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{26, "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{27, "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{28, "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{29, "FFLLSSSSYYYYCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{30, "FFLLSSSSYYEECC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{31, "FFLLSSSSYYEECCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{32, "FFLLSSSSYY*WCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
	{33, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG"},
};

// BLOSUM62 in the order ARNDCQEGHILKMFPSTWYV, lower triangle row by row (diagonal last in each row),
// then the X row; '*' scores -4 against everything and +1 against itself.
const int8_t kB62Lower[] = {
	 4,
	-1, 5,
	-2, 0, 6,
	-2,-2, 1, 6,
	 0,-3,-3,-3, 9,
	-1, 1, 0, 0,-3, 5,
	-1, 0, 0, 2,-4, 2, 5,
	 0,-2, 0,-1,-3,-2,-2, 6,
	-2, 0, 1,-1,-3, 0, 0,-2, 8,
	-1,-3,-3,-3,-1,-3,-3,-4,-3, 4,
	-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,
	-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,
	-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5,
	-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,
	-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,
	 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4,
	 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,
	-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11,
	-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,
	 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,
};
const int8_t kB62X[20] = { 0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1 };

struct MatInit {
	MatInit() {
		int k = 0;
		for (int i = 0; i < 20; ++i)
			for (int j = 0; j <= i; ++j, ++k)
				ns_mat_blosum62[i * 22 + j] = ns_mat_blosum62[j * 22 + i] = kB62Lower[k];
		for (int i = 0; i < 22; ++i) ns_mat_blosum62[20 * 22 + i] = ns_mat_blosum62[i * 22 + 20] = -4;
		ns_mat_blosum62[20 * 22 + 20] = 1;
		for (int i = 0; i < 20; ++i) ns_mat_blosum62[21 * 22 + i] = ns_mat_blosum62[i * 22 + 21] = kB62X[i];
		ns_mat_blosum62[21 * 22 + 21] = -1;
	}
} g_mat_init;

void fill_char_table(uint8_t *tab, uint8_t dflt, const char *alphabet, const uint8_t *value /* per alphabet index, or NULL = index */)
{
	memset(tab, dflt, 256);
	for (int i = 0; alphabet[i]; ++i) {
		uint8_t v = value ? value[i] : (uint8_t)i;
		tab[i] = v; // codes map to themselves, so pre-encoded input is accepted too
		tab[(uint8_t)toupper(alphabet[i])] = tab[(uint8_t)tolower(alphabet[i])] = v;
	}
}

} // namespace

extern "C" {

int ns_make_tables(int codon_type) // nasw-tab.c:85-107
{
	const char *tcag = 0;
	if (codon_type < 0 || codon_type > 33) return -1;
	for (const GenCode &g : kGenCodes) if (g.id == codon_type) tcag = g.tcag;
	if (!tcag) return -2;
	fill_char_table(ns_tab_nt4, 4, g_nt_i2c, 0);
	fill_char_table(ns_tab_aa20, 21, g_aa_i2c, 0);
	fill_char_table(ns_tab_aa13, 15, g_aa_i2c, ns_tab_a2r);
	// our codon index is base1*16 + base2*4 + base3 with A,C,G,T = 0..3; NCBI strings use T,C,A,G
	static const int acgt2tcag[4] = { 2, 1, 3, 0 };
	for (int c = 0; c < 64; ++c) {
		int t = acgt2tcag[c >> 4] * 16 + acgt2tcag[(c >> 2) & 3] * 4 + acgt2tcag[c & 3];
		ns_tab_codon[c] = ns_tab_aa20[(uint8_t)tcag[t]];
		ns_tab_codon13[c] = ns_tab_a2r[ns_tab_codon[c]];
	}
	return 0;
}

void ns_opt_set_sp(ns_opt_t *opt, int32_t model) // nasw-tab.c:124-129
{
	static const int32_t generic[6] = { 8, 15, 21, 30, 0, 0 }, mammal[6] = { 8, 15, 21, 30, 4, 4 };
	for (int i = 0; i < 6; ++i)
		opt->sp[i] = model == NS_S_MAMMAL ? mammal[i] : model == NS_S_GENERIC ? generic[i] : 0;
}

void ns_opt_init(ns_opt_t *opt) // nasw-tab.c:131-147
{
	memset(opt, 0, sizeof(*opt));
	opt->go = 11, opt->ge = 1, opt->io = 29, opt->fs = 17;
	opt->xdrop = 100, opt->end_bonus = 5;
	ns_opt_set_sp(opt, NS_S_MAMMAL);
	opt->sp_null_bonus = -7;
	opt->asize = 22;
	opt->ie_coef = .5f;
	opt->sc = ns_mat_blosum62;
	opt->nt4 = ns_tab_nt4, opt->aa20 = ns_tab_aa20, opt->codon = ns_tab_codon;
}

void ns_set_stop_sc(int32_t asize, int8_t *mat, int8_t pen) // nasw-tab.c:149-156
{
	const int32_t stop = ns_tab_aa20[(uint8_t)'*'];
	const int8_t keep = mat[stop * asize + stop];
	for (int32_t i = 0; i < asize; ++i) mat[stop * asize + i] = mat[i * asize + stop] = (int8_t)-pen;
	mat[stop * asize + stop] = keep;
}

void mp_start(void) // misc.c:12-16
{
	ns_make_tables(1);
	mp_realtime();
}

void mp_idxopt_init(mp_idxopt_t *io) // options.c:10-22
{
	memset(io, 0, sizeof(*io));
	io->trans_code = 1;
	io->bbit = 8;
	io->min_aa_len = 30;
	io->kmer = 6;
	io->mod_bit = 1;
}

void mp_mapopt_set_fs(mp_mapopt_t *mo, int32_t fs) // options.c:24-29
{
	mo->fs = fs;
	ns_set_stop_sc(mo->asize, mo->mat, (int8_t)mo->fs);
}

void mp_mapopt_set_max_intron(mp_mapopt_t *mo, int64_t gsize) // options.c:31-40
{
	int64_t x = (int64_t)(sqrt((double)gsize) * 3.6 + 1.);
	if (x < mo->min_max_intron) x = mo->min_max_intron;
	if (x > mo->max_max_intron) x = mo->max_max_intron;
	mo->bw = mo->max_intron = (int32_t)x;
	if (mp_verbose >= 3) fprintf(stderr, "[M::%s] set max intron size to %d\n", __func__, mo->max_intron);
}

void mp_mapopt_init(mp_mapopt_t *mo) // options.c:42-90
{
	memset(mo, 0, sizeof(*mo));
	mo->mini_batch_size = 2000000;
	mo->max_occ = 20000;
	mo->max_gap = 1000;
	mo->max_intron = 200000;
	mo->min_max_intron = 10000, mo->max_max_intron = 300000;
	mo->bw = mo->max_intron;
	mo->max_ext = 10000;
	mo->max_ava = 1000;
	mo->min_chn_cnt = 3;
	mo->max_chn_max_skip = 25;
	mo->max_chn_iter = 1000000;
	mo->min_chn_sc = 0;
	mo->chn_coef_log = 0.75f;
	mo->mask_level = 0.5f;
	mo->mask_len = INT32_MAX;
	mo->pri_ratio = 0.7f;
	mo->out_sim = 0.99f, mo->out_cov = 0.1f;
	mo->best_n = 30, mo->out_n = 1000;
	mo->kmer2 = 5;
	mo->go = 11, mo->ge = 1, mo->io = 29, mo->fs = 23;
	mo->io_end = 19;
	mo->ie_coef = .5f;
	mo->sp_model = NS_S_GENERIC;
	mo->sp_null_bonus = -7, mo->sp_max_bonus = 14;
	mo->sp_scale = 1.0f;
	mo->xdrop = 100;
	mo->end_bonus = 5;
	mo->asize = 22;
	mo->gff_delim = -1;
	mo->max_intron_flank = 200;
	mo->gff_prefix = "MP";
	memcpy(mo->mat, ns_mat_blosum62, 484);
	ns_set_stop_sc(mo->asize, mo->mat, (int8_t)mo->fs);
}

int32_t mp_mapopt_check(const mp_mapopt_t *mo) // options.c:92-99
{
	if (mo->sp_model < 0 || mo->sp_model > 2) {
		fprintf(stderr, "[ERROR] option -j should be between 0 and 2\n");
		return -1;
	}
	return 0;
}

double mp_realtime(void) // sys.c:93-105: seconds since the first call
{
	static double t0 = -1.0;
	struct timeval tv;
	gettimeofday(&tv, 0);
	double t = (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
	if (t0 < 0) t0 = t;
	return t - t0;
}

double mp_cputime(void) // sys.c:107-114
{
	struct rusage r;
	getrusage(RUSAGE_SELF, &r);
	return (double)(r.ru_utime.tv_sec + r.ru_stime.tv_sec) + 1e-6 * (double)(r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

long mp_peakrss(void) // sys.c:116-127
{
	struct rusage r;
	getrusage(RUSAGE_SELF, &r);
	return r.ru_maxrss * 1024L;
}

mp_tbuf_t *mp_tbuf_init(void) { return (mp_tbuf_t*)calloc(1, 16); } // scratch lives in the GPU context; kept for ABI (map.c:16)
void mp_tbuf_destroy(mp_tbuf_t *b) { free(b); }

int32_t mp_ntseq_read_spsc(mp_ntdb_t *nt, const char *fn, int32_t max_sc) { return mpb::ntdb_read_spsc(nt, fn, max_sc); } // ntseq.c:234

void mp_set_spsc(const char *fn, mp_idx_t *mi, mp_mapopt_t *mo, int32_t keep_io) // index.c:239-248
{
	if (fn == 0) return;
	if (!keep_io) mo->io += 10, mo->io_end += 10;
	int32_t max_sc = (mo->io + 1) / 2 - 1;
	if (max_sc > mo->io - mo->go) max_sc = mo->io - mo->go;
	if (max_sc > mo->sp_max_bonus) max_sc = mo->sp_max_bonus;
	mp_ntseq_read_spsc(mi->nt, fn, max_sc);
}

} // extern "C"
