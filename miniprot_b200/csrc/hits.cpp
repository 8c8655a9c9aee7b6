// hits.cpp -- host bookkeeping between the GPU stages: chains -> regions, ordering, primary/secondary
// classification, selection, extension budgets.  Small, branchy, float-compare code that SURVEY 8(a8) keeps
// on the host; it has to reproduce reference hit.c decision for decision (including FP32/FP64 compares and the
// unstable tie order of the sorts), otherwise the PAF diverges.
#include <assert.h>
#include "internal.hpp"

namespace mpb {

// hit.c:6-16: chain score without gap costs, block resolution
static int32_t chain_score_ungapped_blocks(const mp_reg1_t *r, const uint64_t *a, int32_t kmer)
{
	int32_t sc = kmer;
	for (int32_t i = 1; i < r->cnt; ++i) {
		const uint64_t p = a[r->off + i - 1], q = a[r->off + i];
		const int32_t dq = (int32_t)q - (int32_t)p;
		sc += dq < kmer ? dq : kmer;
		if (q >> 32 == p >> 32) sc += 2; // same-block bonus (miniprot.h:23)
	}
	return sc;
}

// hit.c:18-30: same at base resolution (after refinement)
int32_t chain_score_ungapped(int32_t n_a, const uint64_t *a, int32_t kmer)
{
	int32_t sc = kmer;
	for (int32_t i = 1; i < n_a; ++i) {
		const int32_t dq = (int32_t)a[i] - (int32_t)a[i - 1], dr3 = (int32_t)((a[i] >> 32) - (a[i - 1] >> 32));
		const int32_t dr = dr3 / 3, rem = dr3 - dr * 3, dg = dq < dr ? dq : dr;
		if (dq >= dr && rem != 0) --sc;
		else sc += dg < kmer ? dg : kmer;
	}
	return sc;
}

// hit.c:32-76: one region per chain; a chain that crosses a contig/strand boundary keeps its larger side
mp_reg1_t *regs_from_chains(const mp_idx_t *mi, int32_t n_u, const uint64_t *u, const uint64_t *a, int32_t *n_reg)
{
	mp_reg1_t *reg = (mp_reg1_t*)calloc((size_t)(n_u > 0 ? n_u : 1), sizeof(mp_reg1_t));
	const int32_t bbit = mi->opt.bbit;
	int32_t k = 0;
	for (int32_t i = 0; i < n_u; ++i) {
		const int32_t n = (int32_t)(uint32_t)u[i];
		mp_reg1_t *r = &reg[i];
		int32_t first = k, last = k + n - 1;
		const int32_t vs_id = idx_block2vid(mi, (uint32_t)(a[first] >> 32)), ve_id = idx_block2vid(mi, (uint32_t)(a[last] >> 32));
		r->off = k, r->cnt = n;
		if (vs_id == ve_id) r->vid = (uint32_t)vs_id;
		else {
			int32_t j, head_end, tail_beg;
			for (j = k; j < k + n; ++j) if (a[j] >> 32 >= mi->bo[vs_id + 1]) break;
			head_end = j; // anchors [k, head_end) lie on the first strand
			for (j = k + n - 1; j >= head_end; --j) if (a[j] >> 32 < mi->bo[ve_id]) break;
			tail_beg = j + 1; // anchors [tail_beg, k+n) lie on the last strand
			if (head_end - k > k + n - tail_beg) r->vid = (uint32_t)vs_id, last = head_end - 1;
			else r->vid = (uint32_t)ve_id, first = tail_beg;
		}
		r->vs = (int64_t)((a[first] >> 32) - mi->bo[r->vid]) << bbit;
		r->ve = (int64_t)((a[last] >> 32) - mi->bo[r->vid] + 1) << bbit;
		r->qs = (int32_t)(uint32_t)a[first];
		r->qe = (int32_t)(uint32_t)a[last];
		r->chn_sc = vs_id == ve_id ? (int32_t)(u[i] >> 32) : (int32_t)(uint32_t)((double)(u[i] >> 32) * (last - first + 1) / n + .499);
		r->chn_sc_ungap = chain_score_ungapped_blocks(r, a, mi->opt.kmer);
		k += n;
	}
	*n_reg = n_u;
	return reg;
}

// hit.c:97-126: descending by DP score (or chain score before alignment); soft-deleted entries squeezed out
void regs_sort(int32_t *n_regs, mp_reg1_t *r)
{
	const int32_t n = *n_regs;
	if (n <= 1) return;
	std::vector<mp128_t> key;
	key.reserve((size_t)n);
	for (int32_t i = 0; i < n; ++i) {
		if (r[i].cnt > 0) {
			const int32_t sc = r[i].p ? r[i].p->dp_max : r[i].chn_sc;
			key.push_back(mp128_t{ (uint64_t)sc << 32 | r[i].hash, (uint64_t)i });
		} else if (r[i].p) {
			free(r[i].p); free(r[i].feat);
			r[i].p = 0, r[i].feat = 0;
		}
	}
	sort_128x(key.data(), key.data() + key.size());
	std::vector<mp_reg1_t> tmp(key.size());
	for (size_t i = 0; i < key.size(); ++i) tmp[i] = r[key[key.size() - 1 - i].y];
	if (!tmp.empty()) memcpy(r, tmp.data(), sizeof(mp_reg1_t) * tmp.size());
	*n_regs = (int32_t)tmp.size();
}

// hit.c:128-191: walk hits best-first; a hit overlapping an existing primary on the QUERY by more than
// mask_level becomes its secondary.  All arithmetic in the tests is FP32 exactly as in the reference.
void regs_set_parent(float mask_level, int32_t mask_len, int32_t n, mp_reg1_t *r, int32_t sub_diff, int32_t hard_mask_level)
{
	if (n <= 0) return;
	for (int32_t i = 0; i < n; ++i) r[i].id = i;
	std::vector<int32_t> prim(1, 0);
	std::vector<uint64_t> cov;
	r[0].parent = 0;
	for (int32_t i = 1; i < n; ++i) {
		mp_reg1_t *ri = &r[i];
		const int32_t si = ri->qs, ei = ri->qe;
		int32_t uncov = 0;
		bool secondary = false;
		cov.clear();
		if (!hard_mask_level) {
			for (int32_t w : prim) {
				int32_t sj = r[w].qs, ej = r[w].qe;
				if (ej <= si || sj >= ei) continue;
				cov.push_back((uint64_t)(sj < si ? si : sj) << 32 | (uint32_t)(ej > ei ? ei : ej));
			}
			if (cov.empty()) goto new_primary; // overlaps nothing
			sort_u64(cov.data(), cov.data() + cov.size());
			int32_t x = si;
			for (uint64_t c : cov) {
				if ((int32_t)(c >> 32) > x) uncov += (int32_t)(c >> 32) - x;
				x = (int32_t)c > x ? (int32_t)c : x;
			}
			if (ei > x) uncov += ei - x;
		}
		for (int32_t w : prim) {
			mp_reg1_t *rp = &r[w];
			const int32_t sj = rp->qs, ej = rp->qe;
			if (ej <= si || sj >= ei) continue;
			const int32_t lmin = ej - sj < ei - si ? ej - sj : ei - si, lmax = ej - sj > ei - si ? ej - sj : ei - si;
			const int32_t ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
			if ((float)ol / lmin - (float)uncov / lmax > mask_level && uncov <= mask_len) {
				int32_t counts = 0, sci = ri->chn_sc;
				ri->parent = rp->parent;
				rp->subsc = rp->subsc > sci ? rp->subsc : sci;
				if (ri->cnt >= rp->cnt) counts = 1;
				if (rp->p && ri->p && (rp->vid != ri->vid || rp->vs != ri->vs || rp->ve != ri->ve || ol != lmin)) {
					sci = ri->p->dp_max;
					rp->p->dp_max2 = rp->p->dp_max2 > sci ? rp->p->dp_max2 : sci;
					if (rp->p->dp_max - ri->p->dp_max <= sub_diff) counts = 1;
				}
				if (counts) ++rp->n_sub;
				secondary = true;
				break;
			}
		}
		if (secondary) continue;
new_primary:
		prim.push_back(i), ri->parent = i, ri->n_sub = 0;
	}
}

// hit.c:193-210: after deletions, renumber ids and translate parent links
static void regs_sync(int32_t n, mp_reg1_t *r)
{
	if (n <= 0) return;
	int32_t max_id = -1;
	for (int32_t i = 0; i < n; ++i) max_id = max_id > r[i].id ? max_id : r[i].id;
	std::vector<int32_t> where((size_t)(max_id + 1), -1);
	for (int32_t i = 0; i < n; ++i) if (r[i].id >= 0) where[(size_t)r[i].id] = i;
	for (int32_t i = 0; i < n; ++i) {
		r[i].id = i;
		if (r[i].parent == -2) r[i].parent = i;
		else if (r[i].parent >= 0 && where[(size_t)r[i].parent] >= 0) r[i].parent = where[(size_t)r[i].parent];
		else r[i].parent = -1;
	}
}

static inline bool same_hit(const mp_reg1_t &x, const mp_reg1_t &y)
{
	return x.qs == y.qs && x.qe == y.qe && x.vid == y.vid && x.vs == y.vs && x.ve == y.ve;
}

// hit.c:212-236: keep primaries, and up to best_n secondaries that score within pri_ratio of their parent
void regs_select_sub(float pri_ratio, int32_t min_diff, int32_t best_n, int32_t *n_, mp_reg1_t *r)
{
	if (!(pri_ratio > 0.0f) || *n_ <= 0) return;
	const int32_t n = *n_;
	int32_t k = 0, n_2nd = 0, top_ungap = -1;
	for (int32_t i = 0; i < n; ++i) top_ungap = top_ungap > r[i].chn_sc_ungap ? top_ungap : r[i].chn_sc_ungap;
	for (int32_t i = 0; i < n; ++i) {
		const int32_t p = r[i].parent;
		const int32_t sci = r[i].p ? r[i].p->dp_max : r[i].chn_sc, scp = r[p].p ? r[p].p->dp_max : r[p].chn_sc;
		if (p == i) {
			r[k++] = r[i];
		} else if ((sci >= scp * pri_ratio || sci + min_diff >= scp) && n_2nd < best_n) {
			if (!same_hit(r[i], r[p])) r[k++] = r[i], ++n_2nd;
			else if (r[i].p) { free(r[i].p); free(r[i].feat); }
		} else if (r[i].p == 0 && r[p].p == 0 && top_ungap > 0 && r[i].chn_sc_ungap >= top_ungap * pri_ratio && n_2nd < best_n) {
			if (!same_hit(r[i], r[p])) r[k++] = r[i], ++n_2nd;
		} else if (r[i].p) { free(r[i].p); free(r[i].feat); }
	}
	if (k != n) regs_sync(k, r);
	*n_ = k;
}

// hit.c:238-250: prefer a multi-exon hit over a marginally better single-exon top hit
void regs_select_multi_exon(int32_t n, mp_reg1_t *r, int32_t single_penalty)
{
	if (n < 2 || r[0].n_exon != 1) return;
	int32_t i = 1;
	while (i < n && r[i].n_exon < 2) ++i;
	if (i == n || r[0].p == 0 || r[i].p == 0) return;
	if (r[0].p->dp_max < r[i].p->dp_max + single_penalty) { mp_reg1_t t = r[0]; r[0] = r[i]; r[i] = t; }
}

// hit.c:252-287: how far each region may be extended before it runs into its neighbour on the same strand
void regs_max_ext(const mp_ntdb_t *nt, int32_t n_reg, mp_reg1_t *reg, const uint64_t *a, int32_t min_ext, int32_t max_ext,
                  std::vector<uint64_t> &ext)
{
	ext.assign((size_t)(n_reg > 0 ? n_reg : 0), 0);
	if (n_reg <= 0) return;
	std::vector<mp128_t> ord((size_t)n_reg);
	for (int32_t i = 0; i < n_reg; ++i) {
		const mp_reg1_t *r = &reg[i];
		ord[(size_t)i].x = nt ? (uint64_t)(r->vs + nt->ctg[r->vid >> 1].off + ((r->vid & 1) ? nt->ctg[r->vid >> 1].len : 0)) : a[r->off] >> 32;
		ord[(size_t)i].y = (uint64_t)i;
	}
	sort_128x(ord.data(), ord.data() + n_reg);
	for (int32_t i = 0; i < n_reg; ++i) {
		const int32_t j = (int32_t)ord[(size_t)i].y;
		const mp_reg1_t *r = &reg[j];
		int32_t left = max_ext, right = max_ext;
		if (i > 0) {
			const mp_reg1_t *q = &reg[ord[(size_t)i - 1].y];
			if (q->vid == r->vid && q->qe >= r->qs) {
				left = r->vs - q->ve < max_ext ? (int32_t)(r->vs - q->ve) : max_ext;
				left = left > min_ext ? left : min_ext;
			}
		}
		if (i < n_reg - 1) {
			const mp_reg1_t *q = &reg[ord[(size_t)i + 1].y];
			if (q->vid == r->vid && r->qe >= q->qs) {
				right = q->vs - r->ve < max_ext ? (int32_t)(q->vs - r->ve) : max_ext;
				right = right > min_ext ? right : min_ext;
			}
		}
		ext[(size_t)j] = (uint64_t)left << 32 | (uint32_t)right;
	}
}

} // namespace mpb
