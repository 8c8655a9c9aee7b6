"""DP-only micro-benchmark: synthetic extension / traceback problems of a given shape through mpb_nasw_batch.
Prints Gcell/s from the CUDA-event stage timers.  Usage: python tools/dp_bench.py [n_ext] [nl] [al] [n_tb] [nl_tb] [al_tb]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniprot_b200 as mp  # noqa: E402

a = [int(x) for x in sys.argv[1:]] + [0] * 6
n_ext, nl, al, n_tb, nl_tb, al_tb = a[0] or 4000, a[1] or 10000, a[2] or 24, a[3] or 8000, a[4] or 600, a[5] or 40
rng = np.random.default_rng(3)
ctx = mp.Context(0)
opt = mp.nsopt()
AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)


def mk(n, nl, al, flag):
    out = []
    for i in range(n):
        nt = rng.integers(0, 4, size=nl).astype(np.uint8)
        aa = AA[rng.integers(0, 20, size=al)].tobytes()
        out.append((nt, aa, flag, opt.io))
    return out


for name, probs in (("ext", mk(n_ext, nl, al, 4)), ("tb", mk(n_tb, nl_tb, al_tb, 1))):
    mp.nasw_batch(ctx, opt, probs[:64])  # warm-up
    ctx.reset_stats()
    t = time.time()
    mp.nasw_batch(ctx, opt, probs)
    wall = time.time() - t
    st = ctx.stats()
    cells = st.dp_cells_ext + st.dp_cells_tb
    ms = st.ms_dp_ext + st.ms_dp_tb
    print(f"{name}: n={len(probs)} nl={probs[0][0].size} al={len(probs[0][1])} cells={cells:.3e} kernel_ms={ms:.2f} "
          f"Gcell/s(kernel)={cells / ms / 1e6:.1f} wall_s={wall:.2f} Gcell/s(e2e)={cells / wall / 1e9:.2f}")
ctx.close()
