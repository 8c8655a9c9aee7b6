"""Deterministic synthetic genome / proteome generator.

Genome : i.i.d. uniform ACGT, split into contigs of ``ctg_len`` bp named chr1..chrN.
Genes  : one per equal-width slot, random strand.  Protein = 'M' + uniform over the 20 amino
         acids, length U[300,500]; CDS by uniform synonymous-codon choice (standard code) +
         stop codon; 3-8 exons (cut points >= 30 nt from the CDS ends); introns
         ``GT[AG]...[CT]AG`` of length U[60, Lmax].
Queries: the planted protein with each residue replaced by a uniform random residue with
         probability 1-identity.  Frameshifts: per-base deletion/insertion in the planted CDS
         with probability ``fs_per_base`` (the query keeps the un-shifted protein).

Everything is a pure function of the arguments (numpy Generator seeded with ``seed``), so the
CPU reference and the GPU path always see byte-identical inputs.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass

import numpy as np

AA20 = "ARNDCQEGHILKMFPSTWYV"
_BASES = "ACGT"
# standard genetic code in TCAG order (NCBI table 1)
_STD_TCAG = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"


def _codon_map():
    order = "TCAG"
    aa2codons = {}
    for i, a in enumerate(_STD_TCAG):
        c = order[i >> 4] + order[(i >> 2) & 3] + order[i & 3]
        aa2codons.setdefault(a, []).append(c)
    return aa2codons


_AA2CODONS = _codon_map()
_COMP = bytes.maketrans(b"ACGT", b"TGCA")


@dataclass
class SynthSpec:
    genome_len: int = 100_000_000
    n_genes: int = 1000
    identity: float = 0.8
    lmax: int = 5000
    seed: int = 12
    ctg_len: int = 50_000_000
    fs_per_base: float = 0.0
    min_exons: int = 3
    max_exons: int = 8
    long_intron: tuple | None = None  # (lo, hi): one intron per gene drawn from U[lo,hi] (C4)

    def tag(self) -> str:
        t = f"g{self.genome_len}_n{self.n_genes}_id{int(self.identity * 100)}_L{self.lmax}_s{self.seed}"
        if self.fs_per_base > 0:
            t += f"_fs{int(self.fs_per_base * 1e5)}"
        if self.long_intron:
            t += f"_li{self.long_intron[0]}-{self.long_intron[1]}"
        return t


def _rand_dna(rng: np.random.Generator, n: int) -> bytearray:
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    return bytearray(lut[rng.integers(0, 4, size=n, dtype=np.uint8)].tobytes())


def _make_gene(rng: np.random.Generator, spec: SynthSpec, slot: int):
    plen = int(rng.integers(300, 501))
    prot = "M" + "".join(AA20[i] for i in rng.integers(0, 20, size=plen - 1))
    cds = []
    for a in prot:
        cs = _AA2CODONS[a]
        cds.append(cs[int(rng.integers(0, len(cs)))])
    stops = _AA2CODONS["*"]
    cds.append(stops[int(rng.integers(0, len(stops)))])
    cds = "".join(cds)
    # planted frameshifts: per-base delete / insert
    if spec.fs_per_base > 0:
        out = []
        hit = rng.random(len(cds)) < spec.fs_per_base
        kind = rng.integers(0, 2, size=len(cds))
        ins = rng.integers(0, 4, size=len(cds))
        for k, ch in enumerate(cds):
            if hit[k] and 30 < k < len(cds) - 30:
                if kind[k] == 0:
                    continue  # deletion
                out.append(ch)
                out.append(_BASES[ins[k]])
            else:
                out.append(ch)
        cds = "".join(out)
    n_exon = int(rng.integers(spec.min_exons, spec.max_exons + 1))
    lo, hi = 30, len(cds) - 30
    cuts = sorted(set(int(x) for x in rng.integers(lo, hi, size=n_exon - 1)))
    ilens = [int(x) for x in rng.integers(60, spec.lmax + 1, size=len(cuts))]
    if spec.long_intron and ilens:
        ilens[int(rng.integers(0, len(ilens)))] = int(rng.integers(spec.long_intron[0], spec.long_intron[1] + 1))
    budget = int(slot * 0.9) - len(cds)
    tot = sum(ilens)
    if tot > budget > 0:
        ilens = [max(60, int(x * budget / tot)) for x in ilens]
    pieces = []
    prev = 0
    for c, il in zip(cuts, ilens):
        pieces.append(cds[prev:c])
        body = _rand_dna(rng, il)
        body[0:2] = b"GT"
        body[2] = ord("AG"[int(rng.integers(0, 2))])
        body[-3] = ord("CT"[int(rng.integers(0, 2))])
        body[-2:] = b"AG"
        pieces.append(body.decode())
        prev = c
    pieces.append(cds[prev:])
    gene = "".join(pieces)
    # query protein
    mut = rng.random(plen) >= spec.identity
    sub = rng.integers(0, 20, size=plen)
    q = "".join(AA20[sub[i]] if mut[i] else prot[i] for i in range(plen))
    return gene, q, prot


def generate(spec: SynthSpec, outdir: str, gz: bool = False):
    """Write <outdir>/<tag>.fa and <tag>.faa (deterministic); return (genome_path, protein_path)."""
    os.makedirs(outdir, exist_ok=True)
    tag = spec.tag()
    gpath = os.path.join(outdir, tag + (".fa.gz" if gz else ".fa"))
    ppath = os.path.join(outdir, tag + ".faa")
    if os.path.exists(gpath) and os.path.exists(ppath) and os.path.exists(os.path.join(outdir, tag + ".planted.faa")):
        return gpath, ppath
    rng = np.random.default_rng(spec.seed)
    genome = _rand_dna(rng, spec.genome_len)
    slot = spec.genome_len // max(spec.n_genes, 1)
    prots = []
    plants = []
    for g in range(spec.n_genes):
        gene, q, planted = _make_gene(rng, spec, slot)
        plants.append((f"p{g}", planted))
        if len(gene) > slot - 200:
            prots.append((f"p{g}", q))  # does not fit: query kept (unmappable), nothing planted
            continue
        off = g * slot + int(rng.integers(100, slot - len(gene) - 99))
        # never straddle a contig boundary
        c0, c1 = off // spec.ctg_len, (off + len(gene) - 1) // spec.ctg_len
        if c0 != c1:
            off = c1 * spec.ctg_len + 100
            if off + len(gene) > min((g + 1) * slot + slot, spec.genome_len):
                prots.append((f"p{g}", q))
                continue
        seq = gene.encode()
        if rng.integers(0, 2):
            seq = seq.translate(_COMP)[::-1]
        genome[off:off + len(seq)] = seq
        prots.append((f"p{g}", q))
    opener = (lambda p: gzip.open(p, "wb", compresslevel=1)) if gz else (lambda p: open(p, "wb"))
    tmp = gpath + ".tmp"
    with opener(tmp) as f:
        n_ctg = (spec.genome_len + spec.ctg_len - 1) // spec.ctg_len
        mv = memoryview(genome)
        for c in range(n_ctg):
            f.write(f">chr{c + 1}\n".encode())
            f.write(mv[c * spec.ctg_len:min((c + 1) * spec.ctg_len, spec.genome_len)])
            f.write(b"\n")
    os.replace(tmp, gpath)
    with open(ppath + ".tmp", "w") as f:
        for name, q in prots:
            f.write(f">{name}\n{q}\n")
    os.replace(ppath + ".tmp", ppath)
    with open(os.path.join(outdir, tag + ".planted.faa"), "w") as f:
        for name, q in plants:
            f.write(f">{name}\n{q}\n")
    return gpath, ppath


def shard_queries(spec: SynthSpec, outdir: str, shard: int) -> str:
    """Query set of shard `shard` (weak scaling): the same planted proteins, mutated with a shard-specific seed.
    Shard 0 is the default query file written by generate()."""
    tag = spec.tag()
    if shard == 0:
        return os.path.join(outdir, tag + ".faa")
    out = os.path.join(outdir, f"{tag}.shard{shard}.faa")
    if os.path.exists(out):
        return out
    rng = np.random.default_rng(spec.seed * 1000003 + shard)
    with open(os.path.join(outdir, tag + ".planted.faa")) as f, open(out + ".tmp", "w") as o:
        for line in f:
            if line.startswith(">"):
                o.write(f">{line[1:].strip()}s{shard}\n")
                continue
            prot = line.strip()
            mut = rng.random(len(prot)) >= spec.identity
            sub = rng.integers(0, 20, size=len(prot))
            o.write("".join(AA20[sub[i]] if mut[i] else prot[i] for i in range(len(prot))) + "\n")
    os.replace(out + ".tmp", out)
    return out


# Named workloads of BASELINE.json (configs[1..4]) plus small test-sized variants.
CONFIGS = {
    "C2": SynthSpec(100_000_000, 1000, 0.8, 5000, 12),
    "C3": SynthSpec(3_000_000_000, 20000, 0.8, 50000, 13),
    "C4": SynthSpec(2_000_000_000, 5000, 0.8, 20000, 14, min_exons=4, max_exons=8, long_intron=(50_000, 150_000)),
    "C5": SynthSpec(100_000_000, 1000, 0.6, 5000, 15, fs_per_base=0.0067),
    # cut-down shapes of C3 / C4 for the parity tests (same generator, same options as the full configs):
    #   C3s: 1 Gbp, so that every protein has >= 16384 anchors (the large chaining class) and -I gives bw ~ 1.2e5
    #   C4s: 400 kbp slots, one 50-150 kbp intron per gene, for the -G x -e sweep
    "C3s": SynthSpec(1_000_000_000, 2000, 0.8, 50000, 13),
    "C4s": SynthSpec(120_000_000, 300, 0.8, 20000, 14, ctg_len=40_000_000, min_exons=4, max_exons=8, long_intron=(50_000, 150_000)),
    "tiny": SynthSpec(2_000_000, 40, 0.8, 2000, 7, ctg_len=700_000),
    "tiny5": SynthSpec(2_000_000, 40, 0.6, 2000, 8, ctg_len=700_000, fs_per_base=0.0067),
    "small": SynthSpec(20_000_000, 300, 0.8, 5000, 9, ctg_len=7_000_000),
    "small5": SynthSpec(20_000_000, 300, 0.6, 5000, 10, ctg_len=7_000_000, fs_per_base=0.0067),
}


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("outdir")
    a = ap.parse_args()
    print(*generate(CONFIGS[a.config], a.outdir))


def make_spsc(genome_fa: str, out_path: str, seed: int = 1, p_site: float = 0.6, p_noise: float = 0.002) -> str:
    """A splice-score file for --spsc ("ctg offset +|- D|A score", reference ntseq.c:234-296) over a FASTA genome: scores at a share of
    the GT / AG dinucleotides of both strands (donor: offset of the G; acceptor: offset just past the AG -- the offsets at which
    nasw-sse.c:138-152 applies them to those sites) plus a sprinkle of sites elsewhere, a few of them duplicated with another score."""
    rng = np.random.default_rng(seed)
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    ctgs, name, seq = [], None, []
    with open(genome_fa, "rb") as f:
        for ln in f:
            if ln.startswith(b">"):
                if name is not None:
                    ctgs.append((name, b"".join(seq)))
                name, seq = ln[1:].split()[0].decode(), []
            else:
                seq.append(ln.strip())
    if name is not None:
        ctgs.append((name, b"".join(seq)))
    with open(out_path, "w") as out:
        for name, s in ctgs:
            n = len(s)
            for strand, t in (("+", s.upper()), ("-", s.translate(comp)[::-1].upper())):
                a = np.frombuffer(t, dtype=np.uint8)
                gt = np.flatnonzero((a[:-1] == ord("G")) & (a[1:] == ord("T")))
                ag = np.flatnonzero((a[:-1] == ord("A")) & (a[1:] == ord("G"))) + 2
                for typ, pos in (("D", gt), ("A", ag)):
                    pos = pos[rng.random(len(pos)) < p_site]
                    sc = rng.integers(-8, 16, len(pos))
                    for p, v in zip(pos, sc):
                        out.write(f"{name}\t{p if strand == '+' else n - p}\t{strand}\t{typ}\t{v}\n")
                        if rng.random() < 0.01:  # the same site again with another score: the larger byte wins (ntseq.c:146-152)
                            out.write(f"{name}\t{p if strand == '+' else n - p}\t{strand}\t{typ}\t{int(v) - 3}\n")
                noise = np.flatnonzero(rng.random(n) < p_noise)
                for p in noise:
                    out.write(f"{name}\t{p if strand == '+' else n - p}\t{strand}\t{'DA'[int(rng.integers(0, 2))]}\t{int(rng.integers(-20, 21))}\n")
        out.write("no_such_contig\t10\t+\tD\t5\n" + f"{ctgs[0][0]}\t12\t+\n")  # lines the reader skips
    return out_path
