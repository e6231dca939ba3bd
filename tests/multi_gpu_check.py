"""Multi-GPU parity check, run as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tests/multi_gpu_check.py
Reads are sharded by record (contiguous ranges); rank 0 compares the gathered graph with the CPU
oracle on the whole read set and checks that it equals the single-GPU graph."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

import rust_mdbg_b200 as m
from helpers import genome_reads, pack_reads


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for case, (k, l, d, minab, presimp, err) in enumerate([(5, 10, 0.01, 2, 0.01, 0.002), (10, 12, 0.02, 3, 0.3, 0.001),
                                                            (4, 8, 0.02, 1, 0.5, 0.003), (21, 12, 0.05, 2, 0.01, 0.001)]):
        rng = np.random.default_rng(1000 + case)
        seqs = genome_reads(rng, 60000, 301, mean=6000, sd=2500, err=err) + [b"", b"ACGT"]
        lo = len(seqs) * rank // world
        hi = len(seqs) * (rank + 1) // world
        bases, off = pack_reads(seqs[lo:hi])
        ctx = m.Context(m.Params(k=k, l=l, density=d, min_abundance=minab, presimp=presimp, device=local))
        ids = [m.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        ctx.set_read_base(lo)
        ctx.push_reads(bases, off)
        g = ctx.finish()
        dev_stats = ctx.finish_device()
        ctx.close()
        if rank == 0:
            import oracle_py
            ab, ao = pack_reads(seqs)
            o = oracle_py.build_graph(ab, ao, k, l, d, minab, presimp)
            for key in ("n_kminmers", "n_distinct", "n_nodes", "n_edges", "presimp_removed", "n_seqlines"):
                if g.stats[key] != o.stats[key] or (key != "n_seqlines" and dev_stats[key] != o.stats[key]):
                    print("MISMATCH", case, key, g.stats[key], dev_stats[key], o.stats[key]); ok = False
            for a in ("index", "abundance", "seqlen", "shift", "tuple", "e_n1", "e_n2", "e_o1", "e_o2", "e_ov",
                      "q_index", "q_read", "q_start", "q_end", "q_rev", "q_shift"):
                if not np.array_equal(getattr(g, a), getattr(o, a)):
                    print("MISMATCH", case, a); ok = False
            print("case", case, "world", world, "nodes", g.stats["n_nodes"], "edges", g.stats["n_edges"], "ok" if ok else "FAIL")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
