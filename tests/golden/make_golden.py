"""Regenerate the committed golden PAFs from the compiled reference (run in the build container only).

    python tests/golden/make_golden.py

Writes tests/golden/DPP3_default.paf (+ option variants) and synthetic 'tiny'/'tiny5' PAFs: the reference ships no
expected outputs, so these files ARE the golden vectors; md5 of the default DPP3 PAF is
74fd00200bda6c03380bb3062fb5178b.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from miniprot_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "miniprot")
DATA = os.path.join(ROOT, "oracle", "_ref", "data")


def run(args, out):
    with open(out, "wb") as f:
        subprocess.run([REF, "-t4"] + args, check=True, stdout=f, stderr=subprocess.DEVNULL)


if __name__ == "__main__":
    g, p = os.path.join(DATA, "DPP3-hs.gen.fa.gz"), os.path.join(DATA, "DPP3-mm.pep.fa.gz")
    run([g, p], os.path.join(HERE, "DPP3_default.paf"))
    run(["-j2", g, p], os.path.join(HERE, "DPP3_j2.paf"))
    run(["-G", "2k", g, p], os.path.join(HERE, "DPP3_G2k.paf"))
    for cfg in ("tiny", "tiny5"):
        gg, pp = synth.generate(synth.CONFIGS[cfg], "/tmp/mpb_golden")
        run([gg, pp], os.path.join(HERE, cfg + ".paf"))
    # the other output formats (format.c:189-452), on the set that exercises every CIGAR operation, and on DPP3
    gg, pp = synth.generate(synth.CONFIGS["tiny5"], "/tmp/mpb_golden")
    for name, args in (("gff", ["--gff"]), ("gtf", ["--gtf"]), ("aln", ["--aln"]), ("trans", ["--trans", "-u"]), ("gff_only", ["--gff-only", "--gff-delim", "#"])):
        run(args + [gg, pp], os.path.join(HERE, "tiny5_" + name + ".txt"))
        run(args + [g, p], os.path.join(HERE, "DPP3_" + name + ".txt"))
