"""Differential fuzzing of the chaining kernels' building blocks on the CPU (tests/hostcheck/emu_chain.cpp: the same
chain_core.cuh the kernels compile) against the reference's own mp_chain (oracle/_ref/libref.so) and the C oracle, with random
anchors AND random chaining parameters (band width up to the -I value of a 3 Gbp genome, max_skip, max_iter, min_cnt, ...).
Test infrastructure, not part of the product.   usage: python tools/fuzz_emu_chain.py [seed] [n_iterations]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C  # noqa: E402

import build_hostcheck  # noqa: E402
import oracle_lib as ol  # noqa: E402
from test_emu_chain import emu_chain  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    hc = C.CDLL(build_hostcheck.build())
    hc.emu_chain.restype = C.c_int
    hc.emu_chain.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(n_it):
        if it and it % 2000 == 0:
            print(f"progress seed {seed}: {it} problems, {bad} mismatches", flush=True)
        mode = ("pre", "main", "refine")[it % 3]
        n = int(rng.integers(1, 80)) if rng.random() < 0.5 else int(rng.integers(80, 4000))
        a = ol.random_chain_problem(rng, n, mode)
        if rng.random() < 0.2 and mode != "refine":  # clusters of anchors in the same block / neighbouring blocks (repeats)
            x = np.sort(rng.integers(1000, 1000 + max(2, n // 40), size=n)).astype(np.uint64)
            y = rng.integers(5, 60, size=n).astype(np.uint64)
            a = np.unique((x << np.uint64(32)) | y)
        over = {}
        if rng.random() < 0.6:
            over = dict(max_skip=int(rng.choice([0, 1, 3, 25, 100])), max_iter=int(rng.choice([5, 40, 1000000])), min_cnt=int(rng.integers(1, 5)),
                        min_sc=int(rng.choice([0, 10, 40])), chn_coef_log=float(rng.choice([0.0, 0.25, 0.75, 2.0])), is_spliced=int(rng.integers(0, 2)))
            bw = int(rng.choice([16, 256, 2000, 200000, 197181 // (256 if mode != "refine" else 1), 1000000]))
            over["bw"] = bw
            over["max_dist_x"] = bw if rng.random() < 0.7 else int(rng.choice([100, 5000, 500000]))
            over["max_dist_y"] = int(rng.choice([50, 256, 1000, 5000]))
        par = ol.chain_par(mode, **over)
        ur, br = ol.ref_chain(par, a)
        uo, bo = ol.ora_chain(par, a)
        ue, be = emu_chain(hc, par, a)
        ok = len(ur) == len(uo) == len(ue) and (ur == uo).all() and (ur == ue).all() and len(br) == len(bo) == len(be) and (br == bo).all() and (br == be).all()
        if not ok:
            bad += 1
            print(f"MISMATCH seed={seed} it={it} mode={mode} n={len(a)} over={over}: chains ref {len(ur)} ora {len(uo)} emu {len(ue)}", flush=True)
    print(f"seed {seed}: {n_it} problems, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
