"""Drop-in boundary on the GPU: the reference's UNMODIFIED callers linked against libminiprot_b200.so.

  oracle/_ref/miniprot_b200_cli = /root/reference/main.c   + libminiprot_b200.so   (oracle/Makefile)
  oracle/_ref/example_b200      = /root/reference/example.c + libminiprot_b200.so  (mp_map one query at a time, reads r->p->cigar)
Their stdout must equal that of the same two programs built from the reference's own objects (oracle/_ref/miniprot,
oracle/_ref/example_ref).  Plus the reference-named single-problem entry points called directly: ns_global_gs16[b], mp_map."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import miniprot_b200 as mp
import oracle_lib as ol
from miniprot_b200 import synth

pytestmark = pytest.mark.gpu
REFD = os.path.join(ol.ORA_DIR, "_ref")
CLI, EX, EX_REF = (os.path.join(REFD, x) for x in ("miniprot_b200_cli", "example_b200", "example_ref"))
DATA = os.path.join(REFD, "data")
need_bins = pytest.mark.skipif(not all(os.path.exists(x) for x in (CLI, EX, EX_REF, ol.REF_BIN)), reason="drop-in binaries not built (oracle/Makefile)")


def out(cmd):
    return subprocess.run(cmd, check=True, capture_output=True, timeout=300).stdout


@need_bins
def test_reference_cli_on_our_library(tmp_path):
    g, p = os.path.join(DATA, "DPP3-hs.gen.fa.gz"), os.path.join(DATA, "DPP3-mm.pep.fa.gz")
    for args in ((), ("-j2",), ("-G", "2k"), ("-u", "--outn=3")):
        assert out([CLI, "-t4", *args, g, p]) == out([ol.REF_BIN, "-t4", *args, g, p]), args
    g, p = synth.generate(synth.CONFIGS["small"], str(tmp_path))
    mpi = str(tmp_path / "ours.mpi")
    subprocess.run([CLI, "-t8", "-d", mpi, g], check=True, capture_output=True, timeout=300)  # index built and dumped by our library ...
    assert out([ol.REF_BIN, "-t8", mpi, p]) == out([CLI, "-t8", mpi, p])        # ... restored by the reference, and by us
    assert out([CLI, "-t8", "-I", g, p]) == out([ol.REF_BIN, "-t8", "-I", g, p])


@need_bins
def test_index_built_on_device(tmp_path):
    """mp_idx_load on a FASTA builds ki / kb on the GPU (cuda/idx_build.cu).  The dumped .mpi must equal, byte for byte,
    the one of the reference and the one of our host builder -- on a synthetic genome, on DPP3, and on a FASTA of awkward contigs
    (shorter than a codon, shorter than an ORF, runs of N, lower case, a 300 kb ORF-rich stretch without stop codons)."""
    rng = np.random.default_rng(3)
    odd = str(tmp_path / "odd.fa")
    with open(odd, "w") as f:
        for i, n in enumerate((1, 2, 3, 5, 89, 90, 91, 92, 93, 2047, 2048, 2049, 4096 + 91, 70000)):
            s = "".join("ACGT"[x] for x in rng.integers(0, 4, n))
            if n > 100:
                k = int(rng.integers(0, n - 50))
                s = s[:k] + "N" * int(rng.integers(1, 40)) + s[k:].lower()
            f.write(f">c{i} some text\n{s}\n")
        no_stop = [c for c in (a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT") if c not in ("TAA", "TAG", "TGA")]
        f.write(">orf\n" + "".join(no_stop[x] for x in rng.integers(0, len(no_stop), 100000)) + "\n")
        f.write(">polyA\n" + "A" * 50000 + "\n")
    gs, _ = synth.generate(synth.CONFIGS["small"], str(tmp_path))
    for tag, g in (("odd", odd), ("dpp3", os.path.join(DATA, "DPP3-hs.gen.fa.gz")), ("small", gs)):
        dev, host, ref = (str(tmp_path / f"{tag}.{x}.mpi") for x in ("dev", "host", "ref"))
        r = subprocess.run([CLI, "-t8", "-d", dev, g], check=True, capture_output=True, timeout=300)
        assert b"built the k-mer tables on the device" in r.stderr, r.stderr[-500:]
        r = subprocess.run([CLI, "-t8", "-d", host, g], check=True, capture_output=True, timeout=300, env=dict(os.environ, MPB_IDX_BUILD="host"))
        assert b"on the device" not in r.stderr
        subprocess.run([ol.REF_BIN, "-t8", "-d", ref, g], check=True, capture_output=True, timeout=300)
        a, b, c = (open(x, "rb").read() for x in (dev, host, ref))
        assert len(a) == len(c) and a == c, tag
        assert b == c, tag


@need_bins
def test_splice_score_file(tmp_path):
    """--spsc: the reference's own CLI reads the score file through our mp_set_spsc / mp_ntseq_read_spsc, the scores
    reach the DP as a dense byte table in HBM.  Output identical to the reference's, and different from a run without the file."""
    g, p = synth.generate(synth.CONFIGS["small"], str(tmp_path))
    sp = synth.make_spsc(g, str(tmp_path / "small.spsc"), seed=5)
    plain = out([CLI, "-t8", g, p])
    for args in (("--spsc", sp), ("--spsc", sp, "--spsc0=-3", "--spsc-max=10"), ("--spsc", sp, "--gff", "-j2")):
        a, b = out([CLI, "-t8", *args, g, p]), out([ol.REF_BIN, "-t8", *args, g, p])
        assert a == b, args
        assert len(a) > 1000
    assert out([CLI, "-t8", "--spsc", sp, g, p]) != plain
    g, p = os.path.join(DATA, "DPP3-hs.gen.fa.gz"), os.path.join(DATA, "DPP3-mm.pep.fa.gz")
    import gzip
    fa = str(tmp_path / "dpp3.fa")
    with gzip.open(g, "rb") as f, open(fa, "wb") as o:
        o.write(f.read())
    sp = synth.make_spsc(fa, str(tmp_path / "dpp3.spsc"), seed=6)
    assert out([CLI, "-t4", "--spsc", sp, g, p]) == out([ol.REF_BIN, "-t4", "--spsc", sp, g, p])


@need_bins
def test_reference_example_on_our_library(tmp_path):
    g, p = synth.generate(synth.CONFIGS["tiny"], str(tmp_path))
    assert out([EX, g, p]) == out([EX_REF, g, p])
    g, p = os.path.join(DATA, "DPP3-hs.gen.fa.gz"), os.path.join(DATA, "DPP3-mm.pep.fa.gz")
    assert out([EX, g, p]) == out([EX_REF, g, p])


def test_ns_global_gs16b_with_splice_bytes():
    """ns_global_gs16b(.., ss, ..) (nasw.h:131): the per-base splice bytes of one problem travel with the batch of one."""
    L = mp.lib()
    rng = np.random.default_rng(12)
    opt = mp.nsopt()
    opt.io = 39
    tab = ol.OraTab()
    for f, sym in (("nt4", "ns_tab_nt4"), ("aa20", "ns_tab_aa20"), ("aa13", "ns_tab_aa13"), ("codon", "ns_tab_codon"), ("codon13", "ns_tab_codon13")):
        setattr(tab, f, C.addressof(C.c_uint8.in_dll(L, sym)))
    par = dict(go=opt.go, ge=opt.ge, io=opt.io, fs=opt.fs, xdrop=opt.xdrop, end_bonus=opt.end_bonus, sp=tuple(opt.sp), sp_null_bonus=opt.sp_null_bonus, ie_coef=opt.ie_coef)
    f = L.ns_global_gs16b
    f.restype = None
    for it in range(12):
        for flag in (1, 2, 4):
            nt, aa = ol.random_dp_problem(rng, al_max=(70, 300)[it & 1], flank=50)
            while len(nt) < 3:
                nt, aa = ol.random_dp_problem(rng, al_max=70, flank=50)
            ss = np.full(len(nt), 0xff, dtype=np.uint8)
            k = rng.random(len(nt)) < 0.4
            ss[k] = ((rng.integers(-14, 15, int(k.sum())) + 64) << 1 | rng.integers(0, 2, int(k.sum()))).astype(np.uint8)
            opt.flag = flag
            r = NsRst()
            f(None, nt.ctypes.data_as(C.c_char_p), C.c_int32(len(nt)), aa, C.c_int32(len(aa)), C.byref(opt), ss.ctypes.data_as(C.c_void_p), C.byref(r))
            w = ol.ora_nasw(tab, nt, aa, flag, opt._mat_keepalive, par, ss)
            if flag == 1:
                assert w[0] == r.score and w[3] == [r.cigar[k] for k in range(r.n_cigar)], (it, flag)
                L.mpb_free(r.cigar)
            else:
                assert w[:3] == (r.score, r.nt_len, r.aa_len), (it, flag)


class NsRst(C.Structure):  # ns_rst_t (nasw.h:73-78)
    _fields_ = [("n_cigar", C.c_int32), ("m_cigar", C.c_int32), ("nt_len", C.c_int32), ("aa_len", C.c_int32), ("score", C.c_int32), ("cigar", C.POINTER(C.c_uint32))]


def test_ns_global_gs16b_single_calls():
    """The reference-named DP entry points, one problem per call (a GPU batch of one), against the oracle."""
    L = mp.lib()
    rng = np.random.default_rng(11)
    opt = mp.nsopt()
    tab_syms = (("nt4", "ns_tab_nt4"), ("aa20", "ns_tab_aa20"), ("aa13", "ns_tab_aa13"), ("codon", "ns_tab_codon"), ("codon13", "ns_tab_codon13"))
    tab = ol.OraTab()
    for f, sym in tab_syms:
        setattr(tab, f, C.addressof(C.c_uint8.in_dll(L, sym)))
    par = dict(go=opt.go, ge=opt.ge, io=opt.io, fs=opt.fs, xdrop=opt.xdrop, end_bonus=opt.end_bonus, sp=tuple(opt.sp), sp_null_bonus=opt.sp_null_bonus, ie_coef=opt.ie_coef)
    for fn in ("ns_global_gs16b", "ns_global_gs16"):
        f = getattr(L, fn)
        f.restype = None
        for flag in (1, 2, 4):
            nt, aa = ol.random_dp_problem(rng, al_max=70, flank=50)
            while len(nt) < 3:
                nt, aa = ol.random_dp_problem(rng, al_max=70, flank=50)
            opt.flag = flag
            r = NsRst()
            args = [None, nt.ctypes.data_as(C.c_char_p), C.c_int32(len(nt)), aa, C.c_int32(len(aa)), C.byref(opt)]
            if fn.endswith("b"):
                args.append(None)
            f(*args, C.byref(r))
            w = ol.ora_nasw(tab, nt, aa, flag, opt._mat_keepalive, par)
            if flag == 1:
                assert w[0] == r.score and w[3] == [r.cigar[k] for k in range(r.n_cigar)] and (r.nt_len, r.aa_len) == (len(nt), len(aa)), fn
                L.mpb_free(r.cigar)
            else:
                assert w[:3] == (r.score, r.nt_len, r.aa_len), (fn, flag)
    opt.flag = 0
