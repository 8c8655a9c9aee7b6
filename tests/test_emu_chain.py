"""CPU emulation of the chaining kernels (same building blocks the kernels compile) vs the oracle."""
import ctypes as C

import numpy as np
import pytest

import build_hostcheck
import oracle_lib as ol


@pytest.fixture(scope="module")
def hc():
    lib = C.CDLL(build_hostcheck.build())
    lib.emu_chain.restype = C.c_int
    lib.emu_chain.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def emu_chain(hc, par, a, resort=0):
    n = len(a)
    u = np.zeros(n + 1, np.uint64)
    b = np.zeros(n + 1, np.uint64)
    nb = C.c_int32(0)
    nu = hc.emu_chain(C.addressof(par), n, a.ctypes.data, resort, u.ctypes.data, b.ctypes.data, C.addressof(nb))
    return u[:nu].copy(), b[:nb.value].copy()


@pytest.mark.parametrize("mode", ["pre", "main", "refine"])
def test_emu_chain(hc, mode):
    rng = np.random.default_rng({"pre": 21, "main": 22, "refine": 23}[mode])
    nonempty = 0
    for it in range(150):
        n = int(rng.integers(1, 70 if it % 4 else 1500))
        a = ol.random_chain_problem(rng, n, mode)
        over = {}
        if it % 10 == 0:
            over = dict(is_spliced=0, bw=500, max_dist_x=500)
        if it % 7 == 0:
            over["max_skip"] = int(rng.integers(0, 4))
        if it % 11 == 0:
            over["max_iter"] = 40
        par = ol.chain_par(mode, **over)
        ua, ba = ol.ora_chain(par, a)
        ub, bb = emu_chain(hc, par, a)
        assert len(ua) == len(ub) and (ua == ub).all() and len(ba) == len(bb) and (ba == bb).all(), (mode, it, n)
        nonempty += len(ua) > 0
    assert nonempty > 30


def test_emu_chain_long_skip_runs(hc):
    """Dense collinear anchors: many candidates per row, so max_skip breaks happen inside and across 32-wide chunks."""
    rng = np.random.default_rng(5)
    for it in range(30):
        n = int(rng.integers(200, 1200))
        x = np.sort(rng.integers(14, 14 + 4 * n, size=n)).astype(np.uint64)
        y = (x.astype(np.int64) // 3 + rng.integers(-1, 2, size=n)).clip(4, 5000).astype(np.uint64)
        a = np.unique((x << np.uint64(32)) | y)
        par = ol.chain_par("refine", max_skip=int(rng.integers(1, 30)))
        ua, ba = ol.ora_chain(par, a)
        ub, bb = emu_chain(hc, par, a)
        assert (ua == ub).all() and (ba == bb).all(), it
