// ntdb.cpp -- 4-bit packed genome store: FASTA ingest, window fetch, (de)serialisation.
//
// Layout contract (reference ntseq.c:29-77, miniprot.h:89-98): all contigs are concatenated into one
// nibble array, base i of the concatenation sits in seq[i>>1] >> ((i&1)*4) & 15 with codes A,C,G,T = 0..3
// and everything else 4; contig c covers [ctg[c].off, ctg[c].off + ctg[c].len).  The same array is what
// gets uploaded to HBM, so the device kernels unpack windows with the identical rule.
#include <stdio.h>
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
	free(db->seq); free(db->ctg); free(db->name); free(db->spsc);
	free(db);
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
