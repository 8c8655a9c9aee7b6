"""Measured integer peak of the GPU for the nasw kernels' operations (mpb_int_peak): prints one JSON object.
   python tools/int_peak.py [out.json]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniprot_b200 as mp  # noqa: E402

NAMES = ["viaddmnmx_s32", "vimnmx3_s32", "viaddmnmx_s16x2", "vimnmx3_s16x2", "viaddmnmx_s16x2_relu"]


def measure(ctx):
    L = mp.lib()
    L.mpb_int_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    out = {}
    for v, name in enumerate(NAMES):
        a, b = C.c_double(), C.c_double()
        assert L.mpb_int_peak(ctx.h, v, C.byref(a), C.byref(b)) == 0
        out[name] = {"thread_instr_per_s": a.value, "int_ops_per_s": b.value}
    return out


if __name__ == "__main__":
    ctx = mp.Context(0)
    r = measure(ctx)
    ctx.close()
    print(json.dumps(r, indent=1))
    if len(sys.argv) > 1:
        json.dump(r, open(sys.argv[1], "w"), indent=1)
