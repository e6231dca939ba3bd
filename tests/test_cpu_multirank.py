"""World-size-2 `gloo` test (CPU) of the N>1 plan graph.cu executes with NCCL: read sharding
(mdbg_shard_reads), serial ordinals across ranks (ordinal base = sightings of the ranks before), the
fingerprint-prefix owner function (mdbg_owner_of_fingerprint / mdbg_tuple_fingerprint), the all-to-all
of records bucketed by owner (stable: every bucket ascends in ordinal), and the node-index rule: every
owner marks the first sighting of each of its tuples in a bitmap of two bits per ordinal, the bitmaps
are SUMMED over the ranks (an ordinal has one owner, so the sum is the union) and the prefix popcount
below a tuple's first sighting is its node index.  The per-read minimizers come from the oracle (this
is a test); the merged table must equal the single-process oracle."""
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
    ktot = sum(counts)
    nwords = (ktot + 1 + 15) // 16
    bits = np.zeros(nwords, np.int64)
    for node, (first, cnt) in table.items():
        solid = minab == 1 or (cnt & 0xFFFF) >= minab
        bits[first >> 4] |= (3 if solid else 1) << (2 * (first & 15))
    t = torch.from_numpy(bits)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # graph.cu: ncclAllReduce(ncclUint32, ncclSum)
    bits = t.numpy().astype(np.uint32)
    pop = np.array([bin(int(w) & 0x55555555).count("1") for w in bits], np.int64)
    wscan = np.concatenate([[0], np.cumsum(pop)])
    mine = {}
    for node, (first, cnt) in table.items():
        below = int(bits[first >> 4]) & ((1 << (2 * (first & 15))) - 1)
        index = int(wscan[first >> 4]) + bin(below & 0x55555555).count("1")
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
