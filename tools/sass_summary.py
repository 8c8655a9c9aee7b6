"""Per-kernel SASS mnemonic counts of libminiprot_b200.so (cuobjdump -sass; runs without a GPU): which Blackwell instructions the
hot kernels are made of -- fused add/min/max (VIADDMNMX, VIMNMX3) in 32-bit and packed .S16x2 form, asynchronous global->shared
copies (LDGSTS = cp.async; UBLKCP / UTMALDG would be TMA bulk copies), shuffles, barriers.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "miniprot_b200", "libminiprot_b200.so")
KEYS = ["VIADDMNMX.S16x2", "VIADDMNMX", "VIMNMX3.S16x2", "VIMNMX3", "VIMNMX.S16x2", "VIMNMX", "IADD3", "LOP3", "PRMT", "SHFL", "LDGSTS", "UBLKCP", "UTMALDG", "LDG", "STG", "LDS", "STS",
        "BAR.SYNC", "VOTE", "ATOM", "RED", "BRA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], check=True, capture_output=True, text=True).stdout
    names = {}
    cur = None
    counts = collections.OrderedDict()
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.x]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_n"] += 1
            for k in KEYS:  # the most specific key that prefixes the opcode
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
                    break
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    for mangled, d in zip(counts, dem):
        names[mangled] = re.sub(r"\(.*", "", d).replace("mpb::cuda::", "").replace("void ", "")
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# {os.path.relpath(SO, ROOT)}: {len(counts)} kernels, arch {sorted(set(arch))}; columns = SASS instructions per kernel (static counts)")
    cols = ["_n"] + KEYS
    print("kernel".ljust(58) + "".join(c.replace("_n", "total").rjust(9 if len(c) < 9 else len(c) + 1) for c in cols))
    for mangled, c in sorted(counts.items(), key=lambda kv: -kv[1]["_n"]):
        total.update(c)
        print(names[mangled][:57].ljust(58) + "".join(str(c[k]).rjust(9 if len(k) < 9 else len(k) + 1) for k in cols))
    print("ALL".ljust(58) + "".join(str(total[k]).rjust(9 if len(k) < 9 else len(k) + 1) for k in cols))


if __name__ == "__main__":
    sys.exit(main())
