"""Debug / measurement helper: random DP problems of one width class through mpb_nasw_batch against the oracle, per kernel family.
   python tools/pair_debug.py AL_LO AL_HI N [NL_FLANK]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MPB_NASW_KERNEL", "pair")
import miniprot_b200 as mp  # noqa: E402
import oracle_lib as ol  # noqa: E402

al_lo, al_hi, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
flank = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rng = np.random.default_rng(al_lo * 1000 + al_hi)
ctx = mp.Context(0)
opt = mp.nsopt()
L = mp.lib()
tab = ol.OraTab()
for f, sym in (("nt4", "ns_tab_nt4"), ("aa20", "ns_tab_aa20"), ("aa13", "ns_tab_aa13"), ("codon", "ns_tab_codon"), ("codon13", "ns_tab_codon13")):
    setattr(tab, f, C.addressof(C.c_uint8.in_dll(L, sym)))
par = dict(go=opt.go, ge=opt.ge, io=opt.io, fs=opt.fs, xdrop=opt.xdrop, end_bonus=opt.end_bonus, sp=tuple(opt.sp), sp_null_bonus=opt.sp_null_bonus, ie_coef=opt.ie_coef)
probs = []
while len(probs) < n:
    nt, aa = ol.random_dp_problem(rng, al_max=al_hi, flank=flank, intron_max=600)
    if len(aa) < al_lo or len(nt) < 3:
        continue
    probs.append((nt, aa, (1, 4, 2)[len(probs) % 3], opt.io))
for flag in (4, 2, 1):
    sub = [p for p in probs if p[2] == flag]
    t0 = time.time()
    got = mp.nasw_batch(ctx, opt, sub)
    dt = time.time() - t0
    bad = 0
    for (nt, aa, fl, io), g in zip(sub, got):
        w = ol.ora_nasw(tab, nt, aa, fl, opt._mat_keepalive, par)
        ok = (w[0] == g[0] and w[3] == g[3]) if fl == 1 else (w[:3] == g[:3])
        if not ok:
            bad += 1
            if bad <= 3:
                print("  MISMATCH flag", fl, "nl", len(nt), "al", len(aa), "want", w[:3], "got", g[:3], "cig", w[3][:6], g[3][:6])
    print(f"al {al_lo}-{al_hi} flag {flag}: {len(sub)} problems, {bad} mismatches, {dt:.2f} s", flush=True)
ctx.close()
