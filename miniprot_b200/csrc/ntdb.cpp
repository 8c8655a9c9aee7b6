// ntdb.cpp -- 4-bit packed genome store: FASTA ingest, window fetch, (de)serialisation.
//
// Layout contract (reference ntseq.c:29-77, miniprot.h:89-98): all contigs are concatenated into one
// nibble array, base i of the concatenation sits in seq[i>>1] >> ((i&1)*4) & 15 with codes A,C,G,T = 0..3
// and everything else 4; contig c covers [ctg[c].off, ctg[c].off + ctg[c].len).  The same array is what
// gets uploaded to HBM, so the device kernels unpack windows with the identical rule.
#include <stdio.h>
#include <algorithm>
#include <string>
#include <unordered_map>
#include "internal.hpp"
#include "fastx.hpp"

namespace mpb {

mp_ntdb_t *ntdb_read_fasta(const char *fn)
{
	FastxReader rd(fn);
	if (!rd.fp) return 0;
	mp_ntdb_t *db = (mp_ntdb_t*)calloc(1, sizeof(mp_ntdb_t));
	std::vector<uint8_t> packed;
	std::vector<std::string> names;
	std::vector<int64_t> lens;
	std::string name, seq;
	int64_t off = 0;
	while (rd.next(name, seq)) {
		names.push_back(name);
		lens.push_back((int64_t)seq.size());
		packed.resize((size_t)((off + (int64_t)seq.size() + 1) >> 1), 0);
		for (size_t i = 0; i < seq.size(); ++i, ++off)
			packed[(size_t)(off >> 1)] |= (uint8_t)(ns_tab_nt4[(uint8_t)seq[i]] << ((off & 1) * 4));
	}
	db->n_ctg = db->m_ctg = (int32_t)names.size();
	db->l_seq = off;
	db->m_seq = (off + 1) >> 1 << 1;
	db->seq = (uint8_t*)malloc(packed.size() ? packed.size() : 1);
	memcpy(db->seq, packed.data(), packed.size());
	db->ctg = (mp_ctg_t*)calloc((size_t)(db->n_ctg ? db->n_ctg : 1), sizeof(mp_ctg_t));
	for (size_t i = 0; i < names.size(); ++i) db->l_name += (int32_t)names[i].size() + 1;
	db->name = (char*)malloc((size_t)(db->l_name ? db->l_name : 1));
	char *p = db->name;
	off = 0;
	for (size_t i = 0; i < names.size(); ++i) {
		memcpy(p, names[i].c_str(), names[i].size() + 1);
		db->ctg[i].name = p, db->ctg[i].off = off, db->ctg[i].len = lens[i];
		p += names[i].size() + 1, off += lens[i];
	}
	if (mp_verbose >= 3)
		fprintf(stderr, "[M::%s@%.3f] read %ld bases in %d contigs\n", __func__, mp_realtime(), (long)db->l_seq, db->n_ctg);
	return db;
}

void ntdb_destroy(mp_ntdb_t *db)
{
	if (!db) return;
	if (db->spsc) for (int32_t i = 0; i < db->n_ctg * 2; ++i) free(db->spsc[i].a);
	free(db->seq); free(db->ctg); free(db->name); free(db->spsc);
	free(db);
}

// ntseq.c:234-296: "ctg offset +|- D|A score" lines -> per (contig, strand) a sorted array of pos << 8 | (score + 64) << 1 | is_acceptor,
// positions counted on that strand.  The quirks of the reference reader are kept: any strand character but '+' means '-', lines with
// fewer than five fields or an unknown contig / type are skipped, scores are clamped to +-max_sc, sites at either contig end ignored.
int32_t ntdb_read_spsc(mp_ntdb_t *nt, const char *fn, int32_t max_sc)
{
	gzFile fp = fn && strcmp(fn, "-") != 0 ? gzopen(fn, "rb") : gzdopen(0, "rb");
	if (fp == 0) return -1;
	if (max_sc > 63) max_sc = 63;
	std::unordered_map<std::string, int32_t> name2id;
	for (int32_t i = 0; i < nt->n_ctg; ++i)
		if (!name2id.emplace(nt->ctg[i].name, i).second) fprintf(stderr, "ERROR: duplicated contig name!\n");
	if (nt->spsc) for (int32_t i = 0; i < nt->n_ctg * 2; ++i) free(nt->spsc[i].a);
	free(nt->spsc);
	nt->spsc = (mp_spsc_t*)calloc((size_t)nt->n_ctg * 2, sizeof(mp_spsc_t));
	std::vector<std::vector<uint64_t>> per((size_t)nt->n_ctg * 2);
	std::string line;
	std::vector<char> buf(1 << 16);
	int64_t n_read = 0;
	auto take_line = [&]() {
		char *f[5];
		int n_f = 0;
		char *q = &line[0];
		for (char *p = q;; ++p) {
			if (*p == '\t' || *p == 0) {
				const char c = *p;
				*p = 0;
				f[n_f++] = q;
				if (n_f == 5 || c == 0) break;
				q = p + 1;
			}
		}
		if (n_f < 5) return;
		const int64_t pos0 = atol(f[1]);
		const int strand = *f[2] == '+' ? 1 : -1;
		const int type = *f[3] == 'D' ? 0 : *f[3] == 'A' ? 1 : -1;
		int score = atoi(f[4]);
		if (score > max_sc) score = max_sc;
		if (score < -max_sc) score = -max_sc;
		const auto it = name2id.find(f[0]);
		if (it == name2id.end() || type < 0 || pos0 < 0) return;
		const int32_t cid = it->second;
		const int64_t pos = strand < 0 ? nt->ctg[cid].len - pos0 : pos0;
		if (pos > 0 && pos < nt->ctg[cid].len) {
			per[(size_t)cid << 1 | (strand > 0 ? 0 : 1)].push_back((uint64_t)pos << 8 | (uint64_t)((score + 64) << 1 | type));
			++n_read;
		}
	};
	for (bool more = true; more;) {
		line.clear();
		while (line.empty() || line.back() != '\n') {
			if (gzgets(fp, buf.data(), (int)buf.size()) == 0) { more = false; break; }
			line += buf.data();
		}
		while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
		if (!line.empty()) take_line();
	}
	gzclose(fp);
	for (size_t j = 0; j < per.size(); ++j) {
		std::vector<uint64_t> &v = per[j];
		if (v.empty()) continue;
		std::sort(v.begin(), v.end());
		mp_spsc_t *sp = &nt->spsc[j];
		sp->n = sp->m = (uint32_t)v.size();
		sp->a = (uint64_t*)malloc(sizeof(uint64_t) * v.size());
		memcpy(sp->a, v.data(), sizeof(uint64_t) * v.size());
	}
	if (mp_verbose >= 3) fprintf(stderr, "[M::%s] read %ld splice scores\n", "mp_ntseq_read_spsc", (long)n_read);
	return 0;
}

int64_t nt_fetch(const mp_ntdb_t *db, int32_t cid, int64_t st, int64_t en, int32_t rev, uint8_t *out)
{
	if (cid < 0 || cid >= db->n_ctg) return -1;
	const mp_ctg_t *c = &db->ctg[cid];
	if (en < 0 || en > c->len) en = c->len;
	int64_t k = 0;
	if (!rev) {
		for (int64_t g = c->off + st; g < c->off + en; ++g) out[k++] = db->seq[g >> 1] >> ((g & 1) * 4) & 0xf;
	} else {
		for (int64_t g = c->off + en - 1; g >= c->off + st; --g) {
			uint8_t b = db->seq[g >> 1] >> ((g & 1) * 4) & 0xf;
			out[k++] = b >= 4 ? b : (uint8_t)(3 - b);
		}
	}
	return k;
}

int64_t nt_fetch_v(const mp_ntdb_t *db, uint32_t vid, int64_t st, int64_t en, uint8_t *out)
{
	const int64_t L = db->ctg[vid >> 1].len;
	if (st < 0 || en < 0 || st >= L) return -1;
	if (en > L) en = L;
	return (vid & 1) ? nt_fetch(db, (int32_t)(vid >> 1), L - en, L - st, 1, out) : nt_fetch(db, (int32_t)(vid >> 1), st, en, 0, out);
}

// .mpi genome section (ntseq.c:163-205): n_ctg, l_name (int32 each), l_seq (int64), len[n_ctg] (int64),
// packed bases ((l_seq+1)/2 bytes), NUL-separated names (l_name bytes)
void ntdb_dump(FILE *fp, const mp_ntdb_t *db)
{
	int32_t hdr[2] = { db->n_ctg, db->l_name };
	fwrite(hdr, 4, 2, fp);
	fwrite(&db->l_seq, 8, 1, fp);
	for (int32_t i = 0; i < db->n_ctg; ++i) fwrite(&db->ctg[i].len, 8, 1, fp);
	fwrite(db->seq, 1, (size_t)((db->l_seq + 1) >> 1), fp);
	fwrite(db->name, 1, (size_t)db->l_name, fp);
}

mp_ntdb_t *ntdb_restore(FILE *fp)
{
	int32_t hdr[2];
	mp_ntdb_t *db = (mp_ntdb_t*)calloc(1, sizeof(mp_ntdb_t));
	if (fread(hdr, 4, 2, fp) != 2 || fread(&db->l_seq, 8, 1, fp) != 1) { free(db); return 0; }
	db->n_ctg = db->m_ctg = hdr[0], db->l_name = hdr[1], db->m_seq = db->l_seq;
	db->ctg = (mp_ctg_t*)calloc((size_t)(db->n_ctg ? db->n_ctg : 1), sizeof(mp_ctg_t));
	int64_t off = 0;
	for (int32_t i = 0; i < db->n_ctg; ++i) {
		if (fread(&db->ctg[i].len, 8, 1, fp) != 1) break;
		db->ctg[i].off = off, off += db->ctg[i].len;
	}
	size_t nb = (size_t)((db->l_seq + 1) >> 1);
	db->seq = (uint8_t*)malloc(nb ? nb : 1);
	db->name = (char*)malloc((size_t)(db->l_name ? db->l_name : 1));
	if (fread(db->seq, 1, nb, fp) != nb || fread(db->name, 1, (size_t)db->l_name, fp) != (size_t)db->l_name) {
		ntdb_destroy(db);
		return 0;
	}
	char *p = db->name;
	for (int32_t i = 0; i < db->n_ctg; ++i) db->ctg[i].name = p, p += strlen(p) + 1;
	return db;
}

} // namespace mpb
