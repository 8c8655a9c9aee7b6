"""GPU parity tests, stage by stage, through the C ABI of libminiprot_b200.so (mpb_nasw_batch / mpb_chain_batch):
the CUDA kernels against the C oracle on the same seeded inputs.  Bit-exact: integer scores, lengths, CIGAR words."""
import ctypes as C

import numpy as np
import pytest

import miniprot_b200 as mp
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mp.Context(0)
    yield c
    c.close()


def product_tables():
    L = mp.lib()
    t = ol.OraTab()
    for f, sym in (("nt4", "ns_tab_nt4"), ("aa20", "ns_tab_aa20"), ("aa13", "ns_tab_aa13"), ("codon", "ns_tab_codon"), ("codon13", "ns_tab_codon13")):
        setattr(t, f, C.addressof(C.c_uint8.in_dll(L, sym)))
    return t


def _par(opt):
    return dict(go=opt.go, ge=opt.ge, io=opt.io, fs=opt.fs, xdrop=opt.xdrop, end_bonus=opt.end_bonus, sp=tuple(opt.sp),
                sp_null_bonus=opt.sp_null_bonus, ie_coef=opt.ie_coef)


@pytest.mark.parametrize("model,family", [(1, "auto"), (2, "pair"), (1, "pair"), (2, "v3"), (1, "v3"), (2, "cols")])
def test_nasw_batch_matches_oracle(ctx, model, family, monkeypatch):
    """family: which kernel family serves the problems -- the pair-lane kernels for everything they are eligible for (incl. their
    multi-warp form), the block-wide wavefront (v3), the column-pass kernels, or the production choice (auto)."""
    if family != "auto":
        monkeypatch.setenv("MPB_NASW_KERNEL", family)
    rng = np.random.default_rng(77 + model)
    opt = mp.nsopt()
    mp.lib().ns_opt_set_sp(C.byref(opt), model)
    mat = opt._mat_keepalive
    tab = product_tables()
    probs = []
    for it in range(700):
        al_max = (30, 60, 120, 250, 600)[it % 5]
        nt, aa = ol.random_dp_problem(rng, al_max=al_max, flank=80, intron_max=600 if it % 3 else 3000)
        if len(nt) < 3 and it % 50:
            continue
        flag = (1, 4, 2)[it % 3]
        io = 19 if (flag != 1 and it % 7 == 0) else opt.io
        if flag != 1 and len(nt) < 3:
            continue
        probs.append((nt, aa, flag, io))
    for nl in (0, 1, 2, 3, 4):  # degenerate global problems
        probs.append((rng.integers(0, 4, size=nl).astype(np.uint8), b"MKV", 1, opt.io))
    got = mp.nasw_batch(ctx, opt, probs)
    bad = 0
    for (nt, aa, flag, io), g in zip(probs, got):
        par = _par(opt)
        par["io"] = io
        w = ol.ora_nasw(tab, nt, aa, flag, mat, par)
        ok = (w[0] == g[0] and w[3] == g[3]) if flag == 1 else (w[:3] == g[:3])
        if not ok:
            bad += 1
            if bad < 5:
                print("MISMATCH", flag, len(nt), len(aa), w[:3], g[:3], w[3][:8], g[3][:8])
    assert bad == 0
    st = ctx.stats()
    assert st.kernel_launches > 0 and st.dp_cells_ext > 0 and st.dp_cells_tb > 0


def test_nasw_xdrop_long_tail(ctx):
    rng = np.random.default_rng(5)
    opt = mp.nsopt(xdrop=30)
    tab, mat = product_tables(), opt._mat_keepalive
    probs = []
    for it in range(60):
        nt, aa = ol.random_dp_problem(rng, al_max=40, intron_max=0, flank=0)
        probs.append((np.concatenate([nt, np.full(900, 4, np.uint8)]), aa, 4, opt.io))
    got = mp.nasw_batch(ctx, opt, probs)
    for (nt, aa, flag, io), g in zip(probs, got):
        assert ol.ora_nasw(tab, nt, aa, flag, mat, _par(opt))[:3] == g[:3]


def test_nasw_extension_wider_than_4095_columns(ctx):
    """Extensions over more than 4095 residues (a protein whose first pinned anchor lies deep inside): the row maximum
    carries a 15-bit column code there (nasw_core.cuh code_bits).  Reference behaviour: nasw-sse.c:435-443 has no limit."""
    rng = np.random.default_rng(4242)
    opt = mp.nsopt()
    tab, mat = product_tables(), opt._mat_keepalive
    probs = []
    for flag, al in ((4, 4300), (2, 4500), (4, 4096), (2, 4095)):
        nt, aa = ol.random_dp_problem(rng, al_max=al, flank=40, intron_max=300, p_sub=0.25)
        while len(aa) < al - 200:  # random_dp_problem draws the length from [1, al_max]
            nt, aa = ol.random_dp_problem(rng, al_max=al, flank=40, intron_max=300, p_sub=0.25)
        probs.append((nt, aa, flag, opt.io))
    got = mp.nasw_batch(ctx, opt, probs)
    for (nt, aa, flag, io), g in zip(probs, got):
        assert ol.ora_nasw(tab, nt, aa, flag, mat, _par(opt))[:3] == g[:3], (flag, len(nt), len(aa))


@pytest.mark.parametrize("split", ["0", "1"])
def test_nasw_long_wide_extension(ctx, split, monkeypatch):
    """Extensions of 129..256 columns over >= 32768 rows: on one 8-warp CTA (default) and as two concurrent 4-warp column passes
    linked by the carry row (MPB_NASW_SPLIT=1, nasw_host.cu class 10).  A protein tail that never aligns keeps the extension going
    over the whole window, like the 100 k-row end extensions of a real run."""
    monkeypatch.setenv("MPB_NASW_SPLIT", split)
    rng = np.random.default_rng(99)
    opt = mp.nsopt()
    tab, mat = product_tables(), opt._mat_keepalive
    probs = []
    for flag, al, tail in ((4, 140, 34000), (2, 200, 33000), (4, 256, 40000), (2, 129, 36000), (4, 136, 32768 - 400)):
        nt, aa = ol.random_dp_problem(rng, al_max=al, flank=40, intron_max=300, p_sub=0.25)
        while len(aa) < al - 8 or len(aa) > al:
            nt, aa = ol.random_dp_problem(rng, al_max=al, flank=40, intron_max=300, p_sub=0.25)
        junk = rng.integers(0, 4, size=tail).astype(np.uint8)
        nt = np.concatenate([junk, nt]) if flag == 2 else np.concatenate([nt, junk])
        probs.append((nt, aa, flag, opt.io))
    got = mp.nasw_batch(ctx, opt, probs)
    for (nt, aa, flag, io), g in zip(probs, got):
        assert ol.ora_nasw(tab, nt, aa, flag, mat, _par(opt))[:3] == g[:3], (flag, len(nt), len(aa))


@pytest.mark.parametrize("mode", ["pre", "main", "refine"])
def test_chain_batch_matches_oracle(ctx, mode):
    rng = np.random.default_rng({"pre": 31, "main": 32, "refine": 33}[mode])
    for over in ({}, dict(max_skip=2), dict(is_spliced=0, bw=500, max_dist_x=500), dict(max_iter=50)):
        par = ol.chain_par(mode, **over)
        lists = [ol.random_chain_problem(rng, int(rng.integers(1, 80 if i % 5 else 3000)), mode) for i in range(120)]
        mpar = mp.ChainPar(**{f: getattr(par, f) for f, _ in mp.ChainPar._fields_})
        got = mp.chain_batch(ctx, mpar, lists)
        for a, (u, b) in zip(lists, got):
            wu, wb = ol.ora_chain(par, a)
            assert len(wu) == len(u) and (wu == u).all() and len(wb) == len(b) and (wb == b).all(), (mode, over, len(a))


def test_chain_batch_large_problems(ctx):
    """Problems that land in the larger shared-memory classes (<= 8192, <= 13312 anchors) and in the global-memory path."""
    rng = np.random.default_rng(99)
    for mode in ("pre", "main"):
        par = ol.chain_par(mode)
        lists = [ol.random_chain_problem(rng, n, mode) for n in (5000, 7000, 9000, 12000, 15000, 22000, 300)]
        mpar = mp.ChainPar(**{f: getattr(par, f) for f, _ in mp.ChainPar._fields_})
        got = mp.chain_batch(ctx, mpar, lists)
        for a, (u, b) in zip(lists, got):
            wu, wb = ol.ora_chain(par, a)
            assert len(wu) == len(u) and (wu == u).all() and len(wb) == len(b) and (wb == b).all(), (mode, len(a))


def test_chain_batch_gigabase_prechain(ctx):
    """Pre-chain problems of the size a gigabase genome produces (every one above the shared-memory classes: the
    warp-cooperative global-memory backtrack): dense block ids and the sparse shape (most anchors alone in their block)."""
    rng = np.random.default_rng(2024)
    par = ol.chain_par("pre")
    lists = [ol.random_chain_problem(rng, n, "pre") for n in (50000, 150000, 300000)]
    for n in (60000, 200000):  # sparse: block ids spread over 40 n blocks, a few planted collinear runs
        x = np.sort(rng.integers(1000, 1000 + 40 * n, size=n)).astype(np.uint64)
        runs = rng.integers(0, n - 40, size=n // 200)
        for r in runs:
            x[r:r + 30] = x[r] + np.arange(30, dtype=np.uint64) // 3
        x = np.sort(x)
        y = ((x.astype(np.int64) * 85) % 380 + rng.integers(0, 6, size=n) + 5).astype(np.uint64)
        lists.append(np.ascontiguousarray(np.unique((x << np.uint64(32)) | y), dtype=np.uint64))
    mpar = mp.ChainPar(**{f: getattr(par, f) for f, _ in mp.ChainPar._fields_})
    got = mp.chain_batch(ctx, mpar, lists)
    for a, (u, b) in zip(lists, got):
        wu, wb = ol.ora_chain(par, a)
        assert len(wu) == len(u) and (wu == u).all() and len(wb) == len(b) and (wb == b).all(), len(a)


def test_seed_batch_matches_oracle(ctx, tmp_path):
    """mpb_seed_batch (sketch, adaptive occupancy cut-off, bucket expansion, sort) against the oracle's restatement of
    map.c:126-177 on the same index, protein by protein."""
    from miniprot_b200 import synth

    g, p = synth.generate(synth.CONFIGS["small"], str(tmp_path))
    mi = mp.idx_load(g, 8)
    idx = mi.contents
    seqs = []
    with open(p) as fh:
        cur = []
        for line in fh:
            if line.startswith(">"):
                if cur:
                    seqs.append("".join(cur).encode())
                cur = []
            else:
                cur.append(line.strip())
        if cur:
            seqs.append("".join(cur).encode())
    seqs += [b"", b"MKV", b"M" * 40, b"ACDEFGHIKLMNPQRSTVWY" * 3 + b"XX*" + b"WWHHKK" * 5]  # degenerate and low-complexity queries
    tab = product_tables()
    ora = ol.ora()
    for max_occ in (20000, 50):
        got = mp.seed_batch(ctx, mi, max_occ, seqs)
        assert len(got) == len(seqs)
        for s, a in zip(seqs, got):
            n_a = C.c_int64(0)
            ptr = ora.ora_seed_anchors(C.byref(tab), C.c_void_p(idx.ki), C.c_int64(idx.n_kb), C.c_void_p(idx.kb), C.c_int32(idx.opt.kmer), C.c_int32(idx.opt.mod_bit),
                                       C.c_int32(max_occ), C.c_char_p(s), C.c_int32(len(s)), C.byref(n_a))
            want = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(max(n_a.value, 1),)).copy()[:n_a.value] if ptr else np.zeros(0, np.uint64)
            if ptr:
                ol._libc.free(C.c_void_p(ptr))
            assert np.array_equal(a, want), (len(s), len(a), n_a.value)
    mp.lib().mp_idx_destroy(mi)


def test_refine_batch_matches_oracle(ctx, tmp_path):
    """mpb_refine_batch (window ORF 5-mers x protein 5-mers -> anchors -> base-level chain, best chain kept) against the
    oracle's restatement of map.c:41-97, window by window: both strands of every protein's own slot, plus a foreign slot."""
    from miniprot_b200 import synth

    spec = synth.CONFIGS["small"]
    g, p = synth.generate(spec, str(tmp_path))
    mi = mp.idx_load(g, 8)
    mo = mp.mapopt()
    seqs, ctgs = [], []
    for path, out in ((p, seqs), (g, ctgs)):
        cur = []
        with open(path) as fh:
            for line in fh:
                if line.startswith(">"):
                    if cur:
                        out.append("".join(cur).encode())
                    cur = []
                else:
                    cur.append(line.strip())
        if cur:
            out.append("".join(cur).encode())
    seqs = seqs[:80]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    nt4 = np.full(256, 4, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        nt4[ch] = i
    slot = spec.genome_len // spec.n_genes
    windows, slices = [], []
    for q in range(len(seqs)):
        for gslot in (q, (q + 37) % spec.n_genes):
            lo, hi = gslot * slot, (gslot + 1) * slot
            c = lo // spec.ctg_len
            c_len = len(ctgs[c])
            lo -= c * spec.ctg_len
            hi = min(hi - c * spec.ctg_len, c_len)
            for rev in (0, 1):
                as_, ae = (lo, hi) if not rev else (c_len - hi, c_len - lo)
                sl = ctgs[c][lo:hi] if not rev else ctgs[c][lo:hi].translate(comp)[::-1]
                windows.append((q, c << 1 | rev, as_, ae))
                slices.append(nt4[np.frombuffer(sl, np.uint8)])
    got = mp.refine_batch(ctx, mi, mo, seqs, windows)
    tab = product_tables()
    ora = ol.ora()
    par = mp.ChainPar(mo.max_intron, mo.max_gap, mo.bw, mo.max_chn_max_skip, mo.max_chn_iter, mo.min_chn_cnt, mo.min_chn_sc, mo.chn_coef_log,
                      0 if (mo.flag & 0x1) else 1, mo.kmer2, 0)
    n_hit = 0
    for (q, vid, as_, ae), nt, (a, sc) in zip(windows, slices, got):
        nb, sb = C.c_int32(0), C.c_int32(0)
        ptr = ora.ora_refine(C.byref(tab), C.byref(par), C.c_int32(mi.contents.opt.min_aa_len), C.c_int32(mo.max_ava), C.c_void_p(nt.ctypes.data), C.c_int64(len(nt)),
                             C.c_char_p(seqs[q]), C.c_int32(len(seqs[q])), C.byref(nb), C.byref(sb))
        want = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(max(nb.value, 1),)).copy()[:nb.value] if ptr else np.zeros(0, np.uint64)
        if ptr:
            ol._libc.free(C.c_void_p(ptr))
        assert np.array_equal(a, want), (q, vid, len(a), nb.value)
        if len(want):
            assert sc == sb.value
            n_hit += 1
    assert n_hit >= len(seqs) // 2  # the planted genes are found
    mp.lib().mp_idx_destroy(mi)


def test_nasw_global_score_end_column_first_of_pass(ctx):
    """al = 1, 257, 513: the end column is the first column of a block / pass, its last row must not be handled by the
    check-free steady loop (see tests/test_emu_nasw.py, same name)."""
    rng = np.random.default_rng(4242)
    opt = mp.nsopt()
    mat = opt._mat_keepalive
    tab = product_tables()
    probs = []
    for al, nls in ((1, (104, 107, 110, 113, 116)), (257, (803, 806, 809, 812, 815)), (513, (806, 809, 812, 1211))):
        for nl in nls:
            nt0, aa = ol.random_dp_problem(rng, al_max=al, flank=10, intron_max=100)
            aa = (aa + bytes(b"ARNDCQEGHILKMFPSTWYV"[i] for i in rng.integers(0, 20, size=al)))[:al]
            nt = np.concatenate([nt0, rng.integers(0, 4, size=nl).astype(np.uint8)])[:nl]
            probs.append((nt, aa, 1, opt.io))
    got = mp.nasw_batch(ctx, opt, probs)
    for (nt, aa, flag, io), g in zip(probs, got):
        w = ol.ora_nasw(tab, nt, aa, flag, mat, _par(opt))
        assert w[0] == g[0] and list(w[3]) == list(g[3]), (len(nt), len(aa), w[0], g[0])


def test_segmented_sort(ctx):
    """seg_sort.cu through mpb_sort_segments: empty / tiny segments, the small-tile class, one large tile, and several tiles with
    one to six merge passes, against numpy's sort."""
    rng = np.random.default_rng(8)
    sizes = [0, 1, 2, 3, 31, 33, 1000, 1024, 1025, 5000, 8191, 8192, 8193, 16384, 20000, 47000, 100000, 300001, 7, 0, 9000]
    off = np.zeros(len(sizes) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    keys = rng.integers(0, 1 << 62, size=int(off[-1]), dtype=np.int64).astype(np.uint64)
    keys[off[9]:off[9] + 100] = keys[off[9]]  # duplicates are legal input
    want = keys.copy()
    for s in range(len(sizes)):
        want[off[s]:off[s + 1]].sort()
    L = mp.lib()
    L.mpb_sort_segments.restype = C.c_int
    L.mpb_sort_segments.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    assert L.mpb_sort_segments(ctx.h, len(sizes), off.ctypes.data, keys.ctypes.data) == 0
    assert (keys == want).all()


def test_nasw_batch_random_scoring_parameters(ctx):
    """Ten batches, each with its own random scoring parameters (-O >= 1, -E, -J, -F, -B, splice penalties, x-drop, ie_coef) and
    stop-codon score: every kernel family the dispatcher picks (incl. the pair-lane kernels where the value-domain check admits
    the problem) against the oracle.  The CPU twin of this test is tests/test_emu_nasw.py::test_emu_random_scoring_parameters."""
    rng = np.random.default_rng(31337)
    tab = product_tables()
    for batch in range(10):
        over = dict(go=int(rng.integers(1, 31)), ge=int(rng.integers(0, 6)), io=int(rng.integers(3, 61)), fs=int(rng.integers(1, 61)),
                    end_bonus=int(rng.integers(0, 21)), xdrop=int(rng.choice([5, 30, 100, 400])), ie_coef=float(rng.choice([0.0, 0.25, 0.5, 1.0, 2.5])))
        if batch % 2:
            over["sp"] = tuple(int(x) for x in rng.integers(0, 40, size=4)) + (int(rng.integers(0, 8)), int(rng.integers(0, 8)))
        opt = mp.nsopt(**over)
        mat = opt._mat_keepalive
        if batch % 3 == 0:
            mp.lib().ns_set_stop_sc(22, mat.ctypes.data_as(C.c_void_p), int(rng.integers(1, 60)))
        probs = []
        for it in range(120):
            nt, aa = ol.random_dp_problem(rng, al_max=(12, 64, 140, 300)[it % 4], flank=(0, 3, 60)[it % 3], intron_max=(0, 60, 400)[it % 3],
                                          p_sub=(0.05, 0.2, 0.5)[it % 3])
            if len(nt) >= 3:
                probs.append((nt, aa, (1, 4, 2)[it % 3], opt.io))
        got = mp.nasw_batch(ctx, opt, probs)
        par = _par(opt)
        for (nt, aa, flag, io), g in zip(probs, got):
            w = ol.ora_nasw(tab, nt, aa, flag, mat, par)
            assert (w[0] == g[0] and w[3] == g[3]) if flag == 1 else (w[:3] == g[:3]), (batch, over, flag, len(nt), len(aa), w[:3], g[:3])


def test_unsupported_scoring_is_refused(ctx):
    """Gap open 0 (the reference's result then depends on its SSE stripe layout, backend.cu bad_scoring) and an ie_coef beyond the
    penalty table are refused with -3 instead of being mapped approximately."""
    nt, aa = ol.random_dp_problem(np.random.default_rng(5), al_max=40)
    for over in (dict(go=0), dict(ie_coef=40.0)):
        with pytest.raises(RuntimeError, match="-3"):
            mp.nasw_batch(ctx, mp.nsopt(**over), [(nt, aa, 1, 29)])
