"""CPU: the product library loads, exports every symbol of include/*.h, and its tables / defaults equal the reference's."""
import ctypes as C
import os
import re

import pytest

import miniprot_b200 as mp
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(mp.LIB_PATH):
        mp.build()
    return mp.lib()


def test_exports_every_declared_symbol(L):
    names = set()
    for h in ("miniprot_b200.h", "nasw_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:mp|mpb|ns)_[a-z0-9_]+)\s*\(", src))
    names -= {"mp_tbuf_s"}
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    for g in ("mp_verbose", "mp_dbg_flag", "ns_tab_nt4", "ns_tab_aa20", "ns_tab_aa13", "ns_tab_codon", "ns_tab_codon13", "ns_tab_a2r",
              "ns_mat_blosum62", "ns_tab_nt_i2c", "ns_tab_aa_i2c"):
        C.c_uint8.in_dll(L, g)


@pytest.mark.skipif(not ol.have_ref(), reason="needs oracle/_ref")
def test_tables_and_defaults_equal_reference(L):
    r = ol.ref()
    for code in (1, 2, 4, 11, 33):
        assert L.ns_make_tables(code) == 0 and r.ref_ns_make_tables(code) == 0
        for name, n in (("ns_tab_nt4", 256), ("ns_tab_aa20", 256), ("ns_tab_aa13", 256), ("ns_tab_codon", 64), ("ns_tab_codon13", 64), ("ns_tab_a2r", 22)):
            a = bytes((C.c_uint8 * n).in_dll(L, name))
            b = bytes((C.c_uint8 * n).in_dll(r, "ref_" + name))
            assert a == b, (code, name)
    assert L.ns_make_tables(7) == r.ref_ns_make_tables(7) == -2
    L.ns_make_tables(1), r.ref_ns_make_tables(1)
    assert bytes((C.c_int8 * 484).in_dll(L, "ns_mat_blosum62")) == bytes((C.c_int8 * 484).in_dll(r, "ref_ns_mat_blosum62"))
    mo, ro = mp.MapOpt(), mp.MapOpt()
    L.mp_mapopt_init(C.byref(mo)), r.ref_mp_mapopt_init(C.byref(ro))
    for f, _ in mp.MapOpt._fields_:
        a, b = getattr(mo, f), getattr(ro, f)
        if f == "mat":
            assert bytes(a) == bytes(b)
        else:
            assert a == b, f
    io, rio = mp.IdxOpt(), mp.IdxOpt()
    L.mp_idxopt_init(C.byref(io)), r.ref_mp_idxopt_init(C.byref(rio))
    assert bytes(io) == bytes(rio)
    no, rno = mp.NsOpt(), mp.NsOpt()
    L.ns_opt_init(C.byref(no)), r.ref_ns_opt_init(C.byref(rno))
    for f in ("go", "ge", "io", "fs", "xdrop", "end_bonus", "asize", "sp_null_bonus", "ie_coef"):
        assert getattr(no, f) == getattr(rno, f), f
    assert list(no.sp) == list(rno.sp)
    r.ref_mp_mapopt_set_max_intron.argtypes = [C.c_void_p, C.c_int64]
    C.c_int32.in_dll(r, "ref_mp_verbose").value = 1
    for gsize in (1000, 10**8, 3 * 10**9, 10**11):
        L.mp_mapopt_set_max_intron(C.byref(mo), gsize), r.ref_mp_mapopt_set_max_intron(C.addressof(ro), gsize)
        assert (mo.max_intron, mo.bw) == (ro.max_intron, ro.bw)


@pytest.mark.skipif(not ol.have_ref(), reason="needs oracle/_ref")
def test_index_file_is_byte_identical(L, tmp_path):
    """mp_idx_build + mp_idx_dump vs the reference's -d output on a small synthetic genome (ki/kb order is part of the contract)."""
    import subprocess
    from miniprot_b200 import synth

    g, _ = synth.generate(synth.CONFIGS["tiny"], str(tmp_path))
    ours, ref = str(tmp_path / "ours.mpi"), str(tmp_path / "ref.mpi")
    mi = mp.idx_load(g, 4)
    assert L.mp_idx_dump(ours.encode(), mi) == 0
    L.mp_idx_destroy(mi)
    subprocess.run([ol.REF_BIN, "-t4", "-d", ref, g], check=True, stderr=subprocess.DEVNULL)
    assert open(ours, "rb").read() == open(ref, "rb").read()
    mi2 = L.mp_idx_restore(ref.encode())
    assert mi2 and mi2.contents.n_kb > 0
    L.mp_idx_destroy(mi2)


def test_no_gpu_means_loud_failure(L):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        mp.Context(0)


@pytest.mark.skipif(not ol.have_ref(), reason="needs oracle/_ref")
def test_splice_score_reader_equals_reference(L, tmp_path):
    """mp_set_spsc / mp_ntseq_read_spsc (index.c:239-248, ntseq.c:234-296): option side effects and the sorted per-strand arrays."""
    from miniprot_b200 import synth
    r = ol.ref()
    g, _ = synth.generate(synth.CONFIGS["tiny"], str(tmp_path))
    sp = synth.make_spsc(g, str(tmp_path / "t.spsc"), seed=3, p_site=0.2, p_noise=0.001)

    class Spsc(C.Structure):
        _fields_ = [("n", C.c_uint32), ("m", C.c_uint32), ("a", C.POINTER(C.c_uint64))]

    io = mp.idxopt()
    for fn in (L.mp_idx_load, r.ref_mp_idx_load):
        fn.restype = C.POINTER(mp.Idx)
    a, b = L.mp_idx_load(g.encode(), C.byref(io), 4), r.ref_mp_idx_load(g.encode(), C.byref(io), 4)
    for keep_io in (0, 1):
        mo, ro = mp.mapopt(), mp.mapopt()
        L.mp_set_spsc(sp.encode(), a, C.byref(mo), keep_io), r.ref_mp_set_spsc(sp.encode(), b, C.byref(ro), keep_io)
        assert (mo.io, mo.io_end) == (ro.io, ro.io_end) == ((29, 19) if keep_io else (39, 29))
        n_ctg = a.contents.nt.contents.n_ctg
        sa, sb = C.cast(a.contents.nt.contents.spsc, C.POINTER(Spsc)), C.cast(b.contents.nt.contents.spsc, C.POINTER(Spsc))
        tot = 0
        for j in range(2 * n_ctg):
            assert sa[j].n == sb[j].n, j
            assert [sa[j].a[k] for k in range(sa[j].n)] == [sb[j].a[k] for k in range(sb[j].n)], j
            tot += sa[j].n
        assert tot > 1000
    L.mp_idx_destroy(a)


def test_build_record_matches_the_loaded_library():
    """build() records what the library was compiled from (BUILD_INFO.json); bench.py and smoke() report it, so a stale or foreign
    .so would show (so_is_that_build / sources_unchanged_since false)."""
    import os

    import miniprot_b200 as mp

    if not os.path.exists(mp.BUILD_INFO):
        import pytest

        pytest.skip("library was built outside miniprot_b200.build()")
    bi = mp.build_info()
    assert "error" not in bi and bi["so_is_that_build"] and bi["sources_unchanged_since"], bi
    assert "compute_100a" in bi["arch"] and len(bi["source_sha256"]) == 64
