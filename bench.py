#!/usr/bin/env python
"""bench.py -- the protein-to-genome mapping hot path on N B200s (BASELINE.json metric: proteins/s, DP Gcell/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|small|tiny]

One "step" = one pass of the hot path (seed + chain + refine + nasw DP waves + host bookkeeping) over one batch of
synthetic proteins of the named workload.  N>1 is launched by torchrun, one rank per GPU: rank 0 builds the index,
ONE NCCL broadcast puts ki / kb / packed genome into every GPU's HBM, every rank then maps its own shard of
queries (weak scaling: fixed proteins per GPU) with no further communication.

JSON line (rank 0): metric proteins/s; `value` = mapping with the index resident in HBM and the parsed proteins in
host memory (mpb_map_batch); `e2e` = the reference-facing call on files (mpb_map_file: FASTA in, PAF out, every
host<->device copy inside); `roofline` for the dominant kernel; `cpu_baseline` = the reference CLI on this box.
`--impl reference` times the unmodified reference (oracle/_ref/miniprot -t<all cores>) on the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from miniprot_b200 import synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "miniprot")
CACHE = os.environ.get("MPB_BENCH_DIR", "/tmp/mpb_bench")


def read_fasta(path):
    names, seqs = [], []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                names.append(line[1:].split()[0])
                seqs.append([])
            else:
                seqs[-1].append(line.strip())
    return names, ["".join(s) for s in seqs]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows if len(r) > 3 + k)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def prepare(workload, rank, world):
    """Synthetic inputs + prebuilt .mpi index (rank 0 writes, others wait at the barrier outside)."""
    spec = synth.CONFIGS[workload]
    d = os.path.join(CACHE, workload)
    if rank == 0:
        synth.generate(spec, d)
        for r in range(world):
            synth.shard_queries(spec, d, r)
    return spec, d


def run_reference(args, spec, d, mpi):
    """--impl reference: the unmodified reference CLI on this box's host cores, same workload, same metric."""
    prot = synth.shard_queries(spec, d, 0)
    n_prot = len(read_fasta(prot)[0])
    cores = os.cpu_count() or 1
    cmd = [REF_BIN, f"-t{cores}", mpi, prot]
    for _ in range(args.warmup):
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    dt = (time.perf_counter() - t0) / args.steps
    v = n_prot / dt
    print(json.dumps({
        "impl": "reference", "metric": "proteins/s", "value": v, "unit": "proteins/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec.genome_len} bp genome, {n_prot} proteins, defaults", "index": "prebuilt .mpi (load time inside the step)"},
        "cpu_baseline": {"value": v, "unit": "proteins/s", "cores": cores, "kind": "reference", "sample": f"all {n_prot} proteins, {args.steps} runs of miniprot -t{cores}"},
        "e2e": {"value": v, "unit": "proteins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return
        spec, d = prepare(args.workload, 0, 1)
        g = os.path.join(d, spec.tag() + ".fa")
        mpi = os.path.join(d, spec.tag() + ".ref.mpi")
        if not os.path.exists(mpi):
            subprocess.run([REF_BIN, f"-t{os.cpu_count()}", "-d", mpi, g], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        run_reference(args, spec, d, mpi)
        return

    import numpy as np
    import torch
    import miniprot_b200 as mp

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = mp.lib()
    L.mpb_event_end_ms.restype = C.c_double
    L.mpb_event_begin.argtypes = [C.c_void_p]
    L.mpb_event_end_ms.argtypes = [C.c_void_p]
    L.mpb_map_batch.argtypes = [C.c_void_p, C.POINTER(mp.Idx), C.POINTER(mp.MapOpt), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mpb_regs_free.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]

    spec, d = prepare(args.workload, rank, world)
    g = os.path.join(d, spec.tag() + ".fa")
    mpi = os.path.join(d, spec.tag() + ".mpi")
    t_idx = time.perf_counter()
    if rank == 0 and not os.path.exists(mpi):
        mi0 = mp.idx_load(g, os.cpu_count() or 8)
        L.mp_idx_dump((mpi + ".tmp").encode(), mi0)
        os.replace(mpi + ".tmp", mpi)
        L.mp_idx_destroy(mi0)
    if dist:
        dist.barrier()
    mi = mp.idx_load(mpi)  # host copy: contig table, block offsets, packed genome for the host-side statistics
    t_idx = time.perf_counter() - t_idx
    ctx = mp.Context(local)
    nb = mp.n_bucket(mi.contents.opt)
    n_kb, l_seq = mi.contents.n_kb, mi.contents.nt.contents.l_seq
    bcast_ms = None
    if dist:
        # the ONE collective of the path: NCCL broadcast of the read-only index from rank 0 over NVLink
        ki = torch.empty(nb + 1, dtype=torch.int64, device=dev)
        kb = torch.empty(max(n_kb, 1), dtype=torch.int32, device=dev)
        sq = torch.empty((l_seq + 1) // 2 + 16, dtype=torch.uint8, device=dev)
        if rank == 0:
            h_ki = np.ctypeslib.as_array(C.cast(mi.contents.ki, C.POINTER(C.c_int64)), shape=(nb,))
            ki[:nb].copy_(torch.from_numpy(h_ki))
            ki[nb] = n_kb
            kb.copy_(torch.from_numpy(np.ctypeslib.as_array(C.cast(mi.contents.kb, C.POINTER(C.c_int32)), shape=(max(n_kb, 1),))))
            sq[:(l_seq + 1) // 2].copy_(torch.from_numpy(np.ctypeslib.as_array(C.cast(mi.contents.nt.contents.seq, C.POINTER(C.c_uint8)), shape=((l_seq + 1) // 2,))))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in (ki, kb, sq):
            dist.broadcast(t, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        assert L.mpb_idx_attach_device(ctx.h, mi, ki.data_ptr(), kb.data_ptr(), sq.data_ptr()) == 0
    else:
        assert L.mpb_idx_upload(ctx.h, mi) == 0

    prot = synth.shard_queries(spec, d, rank)
    names, seqs = read_fasta(prot)
    n = len(seqs)
    mo = mp.mapopt()
    c_seqs = (C.c_char_p * n)(*[s.encode() for s in seqs])
    c_names = (C.c_char_p * n)(*[s.encode() for s in names])
    c_lens = (C.c_int32 * n)(*[len(s) for s in seqs])
    n_reg = (C.c_int32 * n)()
    regs = (C.c_void_p * n)()
    out_paf = os.path.join(d, f"ours.rank{rank}.paf")

    def step_resident():
        assert L.mpb_map_batch(ctx.h, mi, C.byref(mo), n, c_seqs, c_lens, c_names, n_reg, regs) == 0
        L.mpb_regs_free(n, n_reg, regs)

    def step_e2e():
        mp.map_file(ctx, mi, prot, out_paf, mo)

    def timed(fn, steps):
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        L.mpb_event_begin(ctx.h)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = L.mpb_event_end_ms(ctx.h)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(ms, wall)  # a step ends with host bookkeeping after the last kernel: the wall clock closes the bracket
        if dist:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    st = ctx.stats()
    step_e2e()
    ctx.reset_stats()
    ms_e2e = timed(step_e2e, args.steps)
    st2 = ctx.stats()
    sampler.stop_flag = True

    tot_prot = n * world
    value = tot_prot / (ms / args.steps / 1e3)
    e2e = tot_prot / (ms_e2e / args.steps / 1e3)
    cells = (st.dp_cells_ext + st.dp_cells_tb) / args.steps
    if rank != 0:
        ctx.close()
        if dist:
            dist.destroy_process_group()
        return

    # parity gate + CPU baseline (reference CLI on this box, bounded sample)
    cores = os.cpu_count() or 1
    cpu = {"value": None, "unit": "proteins/s", "cores": cores, "kind": "reference", "sample": "reference binary not present"}
    parity = None
    if os.path.exists(REF_BIN):
        ref_mpi = os.path.join(d, spec.tag() + ".ref.mpi")
        if not os.path.exists(ref_mpi):
            subprocess.run([REF_BIN, f"-t{cores}", "-d", ref_mpi, g], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t0 = time.perf_counter()
        ref = subprocess.run([REF_BIN, f"-t{cores}", ref_mpi, prot], check=True, capture_output=True).stdout
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "proteins/s", "cores": cores, "kind": "reference",
               "sample": f"all {n} proteins of rank 0's shard, miniprot -t{cores}, prebuilt .mpi, {dt:.2f} s wall (index load included)"}
        parity = open(out_paf, "rb").read() == ref

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # Dominant kernel by time: the score-only nasw extension kernel (nasw_v3_kernel<NW,false>).  Its algorithmic HBM traffic is
    # ~0, so the HBM roofline is reported for the traceback kernels (nasw_v3_kernel<NW,true>: 2 B per DP cell written
    # once), and the integer rate for both.  achieved = algorithmic bytes of one step / sum of the traceback launches' CUDA-event
    # durations of that step (events on the launching side streams; each bracket also holds the CIGAR backtrack kernel that
    # follows); traffic = DRAM bytes of the same launches from the committed ncu capture, both divided by the launches per step.
    tb_gbs = (st.dp_cells_tb * 2 / 1e9) / (st.ms_dp_tb / 1e3) if st.ms_dp_tb > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_v6.json")))
        if args.workload == "C2":
            traffic = tr["tb_traffic_bytes_per_step"] / tr["tb_launches_per_step"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "nasw_v3_kernel<NW,true> (traceback; 8 launches per step)", "achieved": tb_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": tb_gbs / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": st.dp_cells_tb * 2 / args.steps / 8,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                "note": "nasw is integer-issue / latency bound, not HBM bound (SURVEY 8d tension): see int_rate, DESIGN.md 4 and profiles/README.md"}
    int_rate = {
        "ext_gcell_per_s": st.dp_cells_ext / st.ms_dp_ext / 1e6 if st.ms_dp_ext > 0 else None,
        "tb_gcell_per_s": st.dp_cells_tb / st.ms_dp_tb / 1e6 if st.ms_dp_tb > 0 else None,
        "ext_int_ops_per_cell": 32, "tb_int_ops_per_cell": 46}
    print(json.dumps({
        "metric": "proteins/s", "value": value, "unit": "proteins/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16 scores (int32 lanes)",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec.genome_len} bp genome, {n} proteins per GPU ~400 aa, identity {spec.identity}, defaults",
                   "l2": "index (ki+kb+genome) and DP working set exceed the 126 MB L2", "index_build_or_load_s": round(t_idx, 2),
                   "nccl_index_broadcast_ms": bcast_ms, "paf_identical_to_reference": parity},
        "dp_gcell_per_s": cells / (ms / args.steps / 1e3) / 1e9 * world,
        "stage_ms_per_step": {"seed": st.ms_seed / args.steps, "chain": st.ms_chain / args.steps, "refine": st.ms_refine / args.steps,
                              "dp_ext": st.ms_dp_ext / args.steps, "dp_tb": st.ms_dp_tb / args.steps},
        "wall_ms_per_step": dict(zip(["S1_seed_chain", "H1_regions", "S2_refine", "H2_plan", "S3_dp_waves", "H3_finish"], [w / args.steps for w in st.ms_wall])),
        "e2e": {"value": e2e, "unit": "proteins/s", "h2d_bytes_per_step": st2.h2d_bytes // args.steps, "d2h_bytes_per_step": st2.d2h_bytes // args.steps,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(st.kernel_launches), "roofline": roofline, "int_rate": int_rate, "cpu_baseline": cpu, "clocks": sampler.summary()}))
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
