"""World-size-2 `gloo` test (CPU) of the N>1 host logic: read sharding (mdbg_shard_reads), the
fingerprint-range owner function (mdbg_owner_of_fingerprint / mdbg_tuple_fingerprint), serial
ordinals across ranks and the node-index rule (sum of lower_bounds over the ranks' sorted
first-sighting lists) -- the same plan graph.cu executes with NCCL.  The per-read minimizers
come from the oracle (this is a test); the merged table must equal the single-process oracle."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, k, l, d, minab, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import oracle_py
    import rust_mdbg_b200 as m
    from helpers import genome_reads, py_kminmers
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = m.ffi.lib()
    rng = np.random.default_rng(42)
    seqs = genome_reads(rng, 30000, 81, mean=4000, sd=1500, err=0.003)
    lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
    L.mdbg_shard_reads(len(seqs), world, rank, ctypes.byref(lo), ctypes.byref(hi))
    # local sightings in serial order
    local = []
    for r in range(lo.value, hi.value):
        hs, ps = oracle_py.extract(seqs[r], l, d)
        local += [node for node, _, _, _ in py_kminmers([int(x) for x in hs], [int(x) for x in ps], k, l)]
    counts = [None] * world
    dist.all_gather_object(counts, len(local))
    base = sum(counts[:rank])
    # all-to-all by fingerprint range
    send = [[] for _ in range(world)]
    for g, node in enumerate(local):
        t = np.array(node, dtype=np.uint64)
        fp = L.mdbg_tuple_fingerprint(t.ctypes.data, k, 0x6d64626700000000)
        send[L.mdbg_owner_of_fingerprint(fp, world)].append((base + g, node))
    recv = [None] * world
    dist.all_to_all_object(recv, send) if hasattr(dist, "all_to_all_object") else None
    if recv[0] is None:      # older torch: emulate with all_gather_object
        allsend = [None] * world
        dist.all_gather_object(allsend, send)
        recv = [allsend[s][rank] for s in range(world)]
    # owner: records arrive grouped by source = ascending ordinal
    table = {}
    for part in recv:
        for ordinal, node in part:
            e = table.setdefault(node, [ordinal, 0])
            e[1] += 1
    firsts = sorted(e[0] for e in table.values())
    allfirsts = [None] * world
    dist.all_gather_object(allfirsts, firsts)
    mine = {}
    for node, (first, cnt) in table.items():
        index = sum(int(np.searchsorted(np.array(f, dtype=np.int64), first, side="left")) for f in allfirsts)
        if minab == 1 or (cnt & 0xFFFF) >= minab:
            mine[node] = (index, cnt & 0xFFFF)
    allnodes = [None] * world
    dist.all_gather_object(allnodes, mine)
    if rank == 0:
        merged = {}
        for part in allnodes:
            assert not (set(part) & set(merged)), "a tuple was owned by two ranks"
            merged.update(part)
        out.put(merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k,l,d,minab", [(5, 10, 0.01, 2), (8, 10, 0.02, 1)])
def test_two_rank_merge_equals_oracle(oracle, k, l, d, minab):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import genome_reads, pack_reads
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + k
    procs = [ctx.Process(target=_worker, args=(r, 2, port, k, l, d, minab, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    seqs = genome_reads(rng, 30000, 81, mean=4000, sd=1500, err=0.003)
    bases, off = pack_reads(seqs)
    o = oracle.build_graph(bases, off, k, l, d, minab, 0.0)
    assert o.stats["n_nodes"] == len(merged) > 0
    for i in range(len(o.index)):
        assert merged[tuple(int(x) for x in o.tuple[i])] == (int(o.index[i]), int(o.abundance[i]))
