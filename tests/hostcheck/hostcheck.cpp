// tests/hostcheck/hostcheck.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Plugs the C oracle (oracle/*.c) into the product's stage interface (mpb::Stages) so that the HOST side of
// the batch dispatcher -- region bookkeeping, alignment planning, statistics, PAF writer, index builder --
// can be checked for byte-identical PAF against the compiled reference on a machine without a GPU.
// This library is built under tests/_build/ by tests/build_hostcheck.py and is never shipped or linked
// into libminiprot_b200.so; the product's stages are CUDA kernels only.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "internal.hpp"
extern "C" {
#include "ora.h"
}

using namespace mpb;

namespace {

ora_tab_t product_tables()
{
	ora_tab_t t;
	t.nt4 = ns_tab_nt4, t.aa20 = ns_tab_aa20, t.aa13 = ns_tab_aa13, t.codon = ns_tab_codon, t.codon13 = ns_tab_codon13;
	return t;
}

ora_chain_par_t chain_par(int32_t mdx, int32_t mdy, int32_t bw, const mp_mapopt_t *o, int32_t min_cnt, int32_t min_sc, int32_t kmer, int32_t bbit)
{
	ora_chain_par_t p;
	p.max_dist_x = mdx, p.max_dist_y = mdy, p.bw = bw, p.max_skip = o->max_chn_max_skip, p.max_iter = o->max_chn_iter;
	p.min_cnt = min_cnt, p.min_sc = min_sc, p.chn_coef_log = o->chn_coef_log, p.is_spliced = !(o->flag & MP_F_NO_SPLICE);
	p.kmer = kmer, p.bbit = bbit;
	return p;
}

struct OracleStages : Stages {
	int64_t cells_ext = 0, cells_tb = 0;
	void seed_chain(const mp_idx_t *mi, const mp_mapopt_t *opt, const Batch &b, ChainSet &out) override
	{
		ora_tab_t tab = product_tables();
		const int32_t w = 1 << mi->opt.bbit, spl = !(opt->flag & MP_F_NO_SPLICE);
		out.u_off.assign(1, 0), out.a_off.assign(1, 0);
		for (int32_t q = 0; q < b.n; ++q) {
			int64_t n_a = 0;
			uint64_t *a = ora_seed_anchors(&tab, mi->ki, mi->n_kb, mi->kb, mi->opt.kmer, mi->opt.mod_bit, opt->max_occ, b.seq[q], b.len[q], &n_a);
			int32_t n_u = 0;
			uint64_t *u = 0;
			if (!(opt->flag & MP_F_NO_PRE_CHAIN) && spl) { // map.c:186-192
				ora_chain_par_t p = chain_par(w, w, w, opt, 2, 0, mi->opt.kmer, mi->opt.bbit);
				uint64_t *a2 = ora_chain(&p, n_a, a, &n_u, &u);
				free(a);
				a = a2, n_a = 0;
				for (int32_t i = 0; i < n_u; ++i) n_a += (uint32_t)u[i];
				free(u);
				u = 0;
				if (a) ora_sort64(a, a + n_a);
			}
			ora_chain_par_t p = chain_par(opt->max_intron, opt->max_gap, opt->bw, opt, opt->min_chn_cnt, opt->min_chn_sc, mi->opt.kmer, mi->opt.bbit);
			uint64_t *c = ora_chain(&p, n_a, a, &n_u, &u);
			free(a);
			int64_t nc = 0;
			for (int32_t i = 0; i < n_u; ++i) nc += (uint32_t)u[i];
			out.u.insert(out.u.end(), u, u + n_u);
			if (c) out.a.insert(out.a.end(), c, c + nc);
			out.u_off.push_back((int64_t)out.u.size()), out.a_off.push_back((int64_t)out.a.size());
			free(u); free(c);
		}
	}
	void refine(const mp_idx_t *mi, const mp_mapopt_t *opt, const Batch &b, const std::vector<RefineJob> &jobs, RefineSet &out) override
	{
		ora_tab_t tab = product_tables();
		ora_chain_par_t p = chain_par(opt->max_intron, opt->max_gap, opt->bw, opt, opt->min_chn_cnt, opt->min_chn_sc, opt->kmer2, 0);
		std::vector<uint8_t> nt;
		out.off.assign(1, 0), out.sc.clear(), out.a.clear();
		for (const RefineJob &j : jobs) {
			nt.resize((size_t)(j.ae - j.as) + 1);
			int64_t l = nt_fetch_v(mi->nt, j.vid, j.as, j.ae, nt.data());
			int32_t nb = 0, sc = 0;
			uint64_t *a = ora_refine(&tab, &p, mi->opt.min_aa_len, opt->max_ava, nt.data(), l, b.seq[j.qid], b.len[j.qid], &nb, &sc);
			if (a) out.a.insert(out.a.end(), a, a + nb);
			out.off.push_back((int64_t)out.a.size()), out.sc.push_back(sc);
			free(a);
		}
	}
	void nasw(const mp_idx_t *mi, const ns_opt_t *base, const Batch &b, const std::vector<DpJob> &jobs, DpSet &out) override
	{
		ora_tab_t tab = product_tables();
		std::vector<uint8_t> nt, ss;
		out.score.clear(), out.nt_len.clear(), out.aa_len.clear(), out.cig.clear(), out.cig_off.assign(1, 0);
		for (const DpJob &j : jobs) {
			ora_nasw_par_t p;
			p.flag = j.flag, p.go = base->go, p.ge = base->ge, p.io = j.io, p.fs = base->fs, p.xdrop = base->xdrop, p.end_bonus = base->end_bonus;
			memcpy(p.sp, base->sp, sizeof(p.sp));
			p.sp_null_bonus = base->sp_null_bonus, p.ie_coef = base->ie_coef, p.mat = base->sc;
			nt.resize((size_t)j.nl + 1);
			int64_t l = j.nl > 0 ? nt_fetch_v(mi->nt, j.vid, j.nt_st, j.nt_st + j.nl, nt.data()) : 0;
			if (l != j.nl) { fprintf(stderr, "[hostcheck] bad slice %ld != %d\n", (long)l, j.nl); abort(); }
			ora_nasw_rst_t r;
			memset(&r, 0, sizeof(r));
			// --spsc: the bytes of the slice as mp_ntseq_spsc_get() fills them for the region's window (ntseq.c:130-156): the largest
			// byte of a position, 0xff where there is none -- and none at the window's first position, which the reference's
			// interval search always skips
			const uint8_t *ssp = 0;
			if (mi->nt->spsc) {
				ss.assign((size_t)j.nl + 1, 0xff);
				const mp_spsc_t *sv = &mi->nt->spsc[j.vid];
				const uint64_t *lo = std::lower_bound(sv->a, sv->a + sv->n, (uint64_t)j.nt_st << 8);
				for (const uint64_t *e = lo; e < sv->a + sv->n && (int64_t)(*e >> 8) < j.nt_st + j.nl; ++e) {
					const int64_t pos = (int64_t)(*e >> 8);
					const uint8_t sc = (uint8_t)(*e & 0xff);
					if (pos == j.win_st) continue;
					uint8_t &x = ss[(size_t)(pos - j.nt_st)];
					if (x == 0xff || x < sc) x = sc;
				}
				ssp = ss.data();
			}
			ora_nasw(&tab, &p, nt.data(), j.nl, b.seq[j.qid] + j.aa_st, j.al, ssp, &r);
			out.score.push_back(r.score), out.nt_len.push_back(r.nt_len), out.aa_len.push_back(r.aa_len);
			if (r.n_cigar) out.cig.insert(out.cig.end(), r.cigar, r.cigar + r.n_cigar);
			out.cig_off.push_back((int64_t)out.cig.size());
			if (FILE *df = getenv("HC_DUMP") ? fopen(getenv("HC_DUMP"), "a") : 0) {
				fprintf(df, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t", j.flag, j.io, j.nl, j.al, r.score, r.nt_len, r.aa_len);
				for (int32_t k = 0; k < j.nl; ++k) fputc("ACGTN"[nt[k]], df);
				fputc('\t', df);
				fwrite(b.seq[j.qid] + j.aa_st, 1, j.al, df);
				fputc('\n', df);
				fclose(df);
			}
			free(r.cigar);
			((j.flag & NS_F_CIGAR) ? cells_tb : cells_ext) += (int64_t)j.nl * j.al;
		}
	}
};

} // namespace

extern "C" {

// map a protein FASTA against a genome FASTA / .mpi index and write PAF to out_path; returns 0 on success
int hc_map_file(const char *genome, const char *prot, const char *out_path, uint32_t flag, int32_t max_intron, int32_t auto_intron,
                int32_t sp_model, int32_t n_threads, int64_t mini_batch)
{
	mp_idxopt_t io;
	mp_mapopt_t mo;
	mp_verbose = 1;
	mp_start();
	mp_idxopt_init(&io);
	mp_mapopt_init(&mo);
	mo.flag |= flag;
	if (max_intron > 0) mo.max_intron = mo.bw = max_intron;
	if (sp_model >= 0) mo.sp_model = sp_model;
	if (mini_batch > 0) mo.mini_batch_size = mini_batch;
	if (getenv("HC_GFF_DELIM")) mo.gff_delim = getenv("HC_GFF_DELIM")[0]; // --gff-delim (test hook)
	mp_idx_t *mi = mp_idx_load(genome, &io, n_threads);
	if (!mi) return -1;
	if (auto_intron) mp_mapopt_set_max_intron(&mo, mi->nt->l_seq);
	FILE *fp = fopen(out_path, "wb");
	if (!fp) return -2;
	OracleStages st;
	int32_t rc = map_file(&st, mi, prot, &mo, fp);
	fclose(fp);
	mp_idx_destroy(mi);
	return rc;
}

// The reference's mp_map_file() entry with the oracle stages behind it: lets the UNMODIFIED reference CLI (main.c) be linked
// against this test library (tests/test_host_cli.py), so that the host side of the product -- option handling, index builder
// and .mpi I/O, hit bookkeeping, alignment planner, statistics, every output format -- is compared with the reference binary
// under arbitrary command lines, without a GPU.
int32_t mp_map_file(const mp_idx_t *mi, const char *fn, const mp_mapopt_t *opt, int)
{
	OracleStages st;
	return map_file(&st, mi, fn, opt, stdout);
}

int hc_idx_dump(const char *genome, const char *out_mpi, int32_t n_threads)
{
	mp_idxopt_t io;
	mp_verbose = 1;
	mp_start();
	mp_idxopt_init(&io);
	mp_idx_t *mi = mp_idx_load(genome, &io, n_threads);
	if (!mi) return -1;
	int rc = mp_idx_dump(out_mpi, mi);
	mp_idx_destroy(mi);
	return rc;
}

} // extern "C"
