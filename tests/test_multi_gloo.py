"""N>1 host path on CPU (gloo, world_size 2): the index broadcast protocol of bench.py (ki with its sentinel, kb, packed
genome) and the query sharding.  Mapping itself needs a GPU; what is checked here is that every rank ends up with the
same index bytes as rank 0 and with a disjoint, deterministic shard of queries."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

import miniprot_b200 as mp
from miniprot_b200 import synth


def _worker(rank, world, d, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = synth.CONFIGS["tiny"]
    g = os.path.join(d, spec.tag() + ".fa")
    mpi = os.path.join(d, spec.tag() + ".mpi")
    L = mp.lib()
    if rank == 0:
        mi0 = mp.idx_load(g, 2)
        assert L.mp_idx_dump(mpi.encode(), mi0) == 0
        L.mp_idx_destroy(mi0)
    dist.barrier()
    mi = mp.idx_load(mpi)
    nb, n_kb, l_seq = mp.n_bucket(mi.contents.opt), mi.contents.n_kb, mi.contents.nt.contents.l_seq
    ki = torch.zeros(nb + 1, dtype=torch.int64)
    kb = torch.zeros(max(n_kb, 1), dtype=torch.int32)
    sq = torch.zeros((l_seq + 1) // 2, dtype=torch.uint8)
    own_ki = np.ctypeslib.as_array(C.cast(mi.contents.ki, C.POINTER(C.c_int64)), shape=(nb,)).copy()
    own_kb = np.ctypeslib.as_array(C.cast(mi.contents.kb, C.POINTER(C.c_int32)), shape=(max(n_kb, 1),)).copy()
    own_sq = np.ctypeslib.as_array(C.cast(mi.contents.nt.contents.seq, C.POINTER(C.c_uint8)), shape=((l_seq + 1) // 2,)).copy()
    if rank == 0:
        ki[:nb] = torch.from_numpy(own_ki)
        ki[nb] = n_kb
        kb.copy_(torch.from_numpy(own_kb))
        sq.copy_(torch.from_numpy(own_sq))
    for t in (ki, kb, sq):
        dist.broadcast(t, 0)
    ok = bool((ki[:nb].numpy() == own_ki).all() and int(ki[nb]) == n_kb and (kb.numpy() == own_kb).all() and (sq.numpy() == own_sq).all())
    # what a rank other than 0 loads in bench.py: the head of the file only (no host copy of ki / kb)
    meta = L.mpb_idx_load_meta(mpi.encode())
    ok = ok and bool(meta) and not meta.contents.ki and not meta.contents.kb and meta.contents.n_kb == n_kb and meta.contents.nt.contents.l_seq == l_seq
    ok = ok and meta.contents.n_block == mi.contents.n_block and meta.contents.nt.contents.n_ctg == mi.contents.nt.contents.n_ctg
    n_bo = 2 * mi.contents.nt.contents.n_ctg + 1
    ok = ok and bytes(C.string_at(meta.contents.bo, 4 * n_bo)) == bytes(C.string_at(mi.contents.bo, 4 * n_bo))
    L.mp_idx_destroy(meta)
    shard = synth.shard_queries(spec, d, rank)
    names = [l[1:].strip() for l in open(shard) if l.startswith(">")]
    q.put((rank, ok, names[:3], len(names)))
    L.mp_idx_destroy(mi)
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_shards(tmp_path):
    d = str(tmp_path)
    spec = synth.CONFIGS["tiny"]
    synth.generate(spec, d)
    for r in range(2):
        synth.shard_queries(spec, d, r)
    if not os.path.exists(mp.LIB_PATH):
        mp.build()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, d, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "a rank received an index that differs from its own load"
    assert res[0][3] == res[1][3] == spec.n_genes
    assert set(res[0][2]).isdisjoint(res[1][2])  # shard 1 proteins carry their own names (p<i>s1)
