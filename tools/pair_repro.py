"""Repro helper: the stage test's problem mix through mpb_nasw_batch (optionally a slice of it)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miniprot_b200 as mp, oracle_lib as ol
model = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
rng = np.random.default_rng(77 + model)
ctx = mp.Context(0)
opt = mp.nsopt()
mp.lib().ns_opt_set_sp(C.byref(opt), model)
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
probs = probs[lo:hi]
print("problems", len(probs), flush=True)
got = mp.nasw_batch(ctx, opt, probs)
print("done", len(got))
