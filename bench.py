#!/usr/bin/env python
"""bench.py -- reads -> mdBG throughput on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  (N > 1: launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`)

A "step" is one pass of the hot path over one batch of synthetic HiFi-shape reads: minimizer
extraction (K-A) + windowing/canonicalisation (K-B) + node table (K-C/K-D) + edges (K-E).
`value` is timed with the reads already resident in HBM (CUDA events on the launching stream);
`e2e` is the same metric through the host-facing C ABI with pinned HOST buffers: H2D of the
reads and D2H of the graph are inside the timed region.  The default workload is BASELINE
config 3 (synthetic D. melanogaster 140 Mbp, HiFi 50x, k=35 l=12 d=0.002: the largest
single-GPU configuration, 7 Gbases per step); the N=1 line also carries a short device-resident
measurement of config 2 (`ecoli50x`).  Per-GPU work is fixed as N grows (weak scaling): rank r
builds reads [r*R, (r+1)*R) of the same job, and the N>1 line carries `parity_vs_single_gpu`:
the N-GPU graph of a bounded job against the graph ONE GPU context builds from the same reads.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the Rust
reference cannot be built in this image) in the reference's thread structure on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] / configs[2] (configs[0] is the CPU-runnable example, a parity case)
    "ecoli50x": dict(genome_len=5_000_000, coverage=50.0, k=21, l=12, density=0.003,
                     desc="synthetic E. coli 5 Mbp HiFi 50x, k=21 l=12 d=0.003"),
    "dmel50x": dict(genome_len=140_000_000, coverage=50.0, k=35, l=12, density=0.002,
                    desc="synthetic D. melanogaster 140 Mbp HiFi 50x, k=35 l=12 d=0.002"),
    # BASELINE config 4 as weak-scaling shards: 6.5x of a 3 Gbp genome per GPU (19.5 Gbases, 19.5 GB of ASCII
    # bases in HBM) = the 52x / 156 Gbases job on 8 GPUs (run with --gpus 8 --workload human52x_per8).
    "human52x_per8": dict(genome_len=3_000_000_000, coverage=6.5, k=35, l=12, density=0.002,
                          desc="synthetic human 3 Gbp HiFi, 6.5x per GPU (52x on 8 GPUs), k=35 l=12 d=0.002"),
    "tiny": dict(genome_len=200_000, coverage=20.0, k=21, l=12, density=0.003, desc="smoke-size"),
}
MIN_ABUNDANCE, PRESIMP = 2, 0.01


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(wl, n_reads, steps, warmup):
    """The reference algorithm on the host cores (oracle/ in the reference's thread structure)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import rust_mdbg_b200 as m
    cores = host_threads()
    s = m.Synth(genome_len=wl["genome_len"])
    ro, total = s.plan(0, n_reads)
    host = s.fill_host(0, n_reads, ro, threads=cores)
    times = []
    stats = None
    for i in range(warmup + steps):
        t = time.perf_counter()
        g = oracle_py.build_graph(host, ro, wl["k"], wl["l"], wl["density"], MIN_ABUNDANCE, PRESIMP, threads=cores)
        dt = time.perf_counter() - t
        stats = g.stats
        g.close()
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / max(1, len(times))
    return total / sec / 1e9, cores, total, sec, stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dmel50x", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the short ecoli50x line at N=1")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the N-GPU == 1-GPU graph comparison")
    ap.add_argument("--parity-gbases", type=float, default=24.0, help="N>1: size cap of the parity job (whole job)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ka-variant", default="default", choices=["default", "classic", "bitslice"],
                    help="K-A kernel (mdbg_params.ka_variant); default = the library's choice")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sweep-k", default="", help="comma list of k: also time finish() per k over the resident "
                    "minimizers (BASELINE config 5, the multi-k idea of utils/multik)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    import rust_mdbg_b200 as m
    synth = m.Synth(genome_len=wl["genome_len"])
    reads_per_rank = synth.num_reads(wl["coverage"])
    config = {"workload": wl["desc"], "reads_per_gpu": reads_per_rank, "min_abundance": MIN_ABUNDANCE,
              "presimp": PRESIMP, "hpc": True, "input": "ASCII bases, 1 B/base",
              "l2": "inputs (>= 250 MB per GPU) larger than the 126 MB L2; no explicit flush",
              "sharding": "reads by record, contiguous ranges per rank" if world > 1 else "single GPU"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n = max(64, min(reads_per_rank, 60000))   # bounded sample of the same workload per step (~0.9 Gbases)
        v, cores, total, sec, st = cpu_reference_run(wl, n, steps, warmup)
        sample = "first %d reads (%d bases) of the workload per step" % (n, total)
        out = {"impl": "reference", "metric": "Gbases/sec reads->mdBG", "value": v, "unit": "Gbases/s",
               "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
               "data": "synthetic", "config": config,
               "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample},
               "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "note": "CPU restatement of the reference algorithm in its thread structure (oracle/); the Rust "
                       "binary cannot be built in this image (no cargo/rustc)"}
        _emit(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    P = m.Params(k=wl["k"], l=wl["l"], density=wl["density"], min_abundance=MIN_ABUNDANCE, presimp=PRESIMP,
                 device=local_rank, ka_variant=args.ka_variant)
    ctx = m.Context(P)
    if world > 1:
        ids = [m.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        ctx.set_read_base(rank * reads_per_rank)

    first = rank * reads_per_rank
    ro, total = synth.plan(first, reads_per_rank)
    d_bases = ctx.device_malloc(total + 64)
    d_off = ctx.device_malloc((reads_per_rank + 1) * 8)
    synth.fill_device(ctx, first, reads_per_rank, ro, d_bases, d_off)

    def step_device():
        ctx.reset()
        ctx.push_reads_device(d_bases, d_off, reads_per_rank, total)
        return ctx.finish_device()

    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    ctx.sync(); barrier()
    sampler.start()
    ctx.timer_start()
    t0 = time.perf_counter()
    ka_ms, launches, stats = 0.0, 0, None
    stage = {"ka": 0.0, "kb": 0.0, "kc": 0.0, "kd": 0.0, "ke": 0.0}
    kern_ms = [0.0] * 8
    for _ in range(steps):
        stats = step_device()
        tm = ctx.timings()
        ka_ms += tm["ms_ka_kernel"]
        for i_, v_ in enumerate(tm["ms_kernels"]):
            kern_ms[i_] += v_
        stage["exchange"] = stage.get("exchange", 0.0) + tm["ms_exchange"]
        stage["finish_total"] = stage.get("finish_total", 0.0) + tm["ms_total_finish"]
        launches += tm["launches_push"] + tm["launches_finish"]
        for s_ in ("ka", "kb", "kc", "kd", "ke"):
            stage[s_] += tm["ms_" + s_]
    ctx.sync()
    dev_ms = ctx.timer_stop()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop()
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    bases_all = sum_over_ranks(float(total))
    value = bases_all * steps / (dev_ms * 1e-3) / 1e9

    # roofline of the dominant kernel (K-A), algorithmic bytes per launch
    M = stats["n_minimizers"]
    alg_bytes = total + 12 * M + 16 * (reads_per_rank + 1)
    ka_avg_ms = ka_ms / steps
    achieved = alg_bytes / (ka_avg_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    # DRAM bytes of one launch from the committed `ncu --set full` capture (tools/ncu_traffic.py) of the
    # kernel variant that ran -- only while the capture belongs to THIS kernel (hash of its sources) and workload
    traffic, traffic_note = None, "no capture"
    variant_name = "bitslice" if int(tm.get("ka_variant_used", 1)) == 2 else "classic"
    tp = os.path.join(ROOT, "profiles", "ka_traffic_bitslice.json" if variant_name == "bitslice" else "ka_traffic.json")
    if os.path.exists(tp):
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from ka_hash import ka_source_hash
            tj = json.load(open(tp))
            if tj.get("workload") != args.workload:
                traffic_note = "capture is of workload %s" % tj.get("workload")
            elif tj.get("source_hash") != ka_source_hash(variant_name):
                traffic_note = "stale: the kernel sources changed since the capture"
            else:
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), "ncu dram__bytes_read.sum + dram__bytes_write.sum, " + str(tj.get("report"))
        except Exception as e:
            traffic_note = "unreadable capture: %s" % e
    ka_variant = {1: "classic", 2: "bitslice"}.get(int(tm.get("ka_variant_used", 1)), "classic")
    roofline = {"kernel": "ka_bitslice_kernel" if ka_variant == "bitslice" else "ka_minimizers_kernel",
                "ka_variant": ka_variant, "ka_dirty_tiles": int(tm.get("ka_dirty_tiles", 0)), "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ka_avg_ms,
                "share_of_step": ka_avg_ms / (dev_ms / steps)}

    # the other kernels of the step against the same HBM roofline (algorithmic bytes per SURVEY 8d; times from CUDA
    # events around each kernel inside the library): they move few bytes per item, latency- not bandwidth-bound
    Kc, Dc, Sc, kk_ = stats["n_kminmers"] / max(1, world), stats["n_distinct"] / max(1, world), stats["n_nodes"], wl["k"]
    kernel_bytes = [("kb_records_kernel", Kc * (8 + 48)), ("kc_insert_kernel", Kc * (64 + 8 + 4)),
                    ("kc_verify_kernel", Kc * 8 + max(0.0, Kc - Dc) * 2 * kk_ * 8), ("cub radix sort by slot (3 passes)", Kc * 8 * 2 * 3),
                    ("ke_join_kernel<0>", Sc / max(1, world) * (2 * (kk_ - 1) * 8 * 2 + 26)), ("ka_finalize_kernel", M * 24 * 2),
                    ("kd_nodes_kernel + kd_expand_kernel", Sc * (20 * 2 + kk_ * 8 * 2 + 14))]
    kernel_rooflines = []
    for i_, (name_, by_) in enumerate(kernel_bytes):
        ms_ = kern_ms[i_] / steps
        if ms_ > 0:
            kernel_rooflines.append({"kernel": name_, "avg_ms": ms_, "algorithmic_bytes": int(by_),
                                     "achieved_gbs": by_ / (ms_ * 1e-3) / 1e9, "frac": by_ / (ms_ * 1e-3) / 1e9 / peak,
                                     "share_of_step": ms_ / (dev_ms / steps)})

    # ------------------------------------------------------------------ multi-k sweep (config 5)
    multik = None
    if args.sweep_k:
        multik = []
        ctx.reset()
        ctx.push_reads_device(d_bases, d_off, reads_per_rank, total)     # minimizers stay resident
        for kk in [int(x) for x in args.sweep_k.split(",") if x]:
            ctx.set_k(kk)
            ctx.finish_device()                                         # warm-up for this k
            ctx.sync(); barrier()
            ctx.timer_start()
            st_k = ctx.finish_device()
            ms_k = max_over_ranks(ctx.timer_stop())
            multik.append({"k": kk, "finish_ms": ms_k, "Gbases_per_s_graph_only": bases_all / (ms_k * 1e-3) / 1e9,
                           "n_kminmers": int(st_k["n_kminmers"]), "n_nodes": int(st_k["n_nodes"]),
                           "n_edges": int(st_k["n_edges"])})
        ctx.set_k(wl["k"])

    # ------------------------------------------------------------------ e2e through the host ABI
    e2e = None
    if not args.no_e2e:
        import ctypes
        import numpy as np
        hb = ctx.host_alloc_pinned(total + 64)
        ho = ctx.host_alloc_pinned((reads_per_rank + 1) * 8)
        hb_arr = np.ctypeslib.as_array(ctypes.cast(hb, ctypes.POINTER(ctypes.c_uint8)), shape=(total,))
        ho_arr = np.ctypeslib.as_array(ctypes.cast(ho, ctypes.POINTER(ctypes.c_uint64)), shape=(reads_per_rank + 1,))
        ho_arr[:] = ro
        ctx.d2h(hb_arr, d_bases)     # same bytes as the device copy (generated on the device)
        d2h_bytes = 0

        e2e_counts = {}
        e2e_parts = {"h2d": 0.0, "push": 0.0, "ka_start": 0.0, "ka_kernels_done": 0.0, "ka_done": 0.0, "finish": 0.0, "d2h": 0.0}

        def step_e2e():
            ctx.reset()
            ctx.push_reads_ptr(hb, ho, reads_per_rank)
            cg = ctx.finish_raw(want_seqlines=False)
            tm = ctx.timings()
            e2e_parts["h2d"] += tm["ms_h2d"]; e2e_parts["push"] += tm["ms_total_push"]
            e2e_parts["finish"] += tm["ms_total_finish"]; e2e_parts["d2h"] += tm["ms_d2h"]
            e2e_parts["ka_kernels_done"] += tm["ms_ka_kernel"]; e2e_parts["ka_done"] += tm["ms_ka"]
            e2e_parts["ka_start"] += tm["ms_ka_start"]
            nb = cg.n_nodes * (4 + 2 + 4 + 4 + 8 * cg.k) + cg.n_edges * (4 + 1 + 4 + 1 + 4)
            e2e_counts.update(n_minimizers=int(cg.n_minimizers), n_kminmers=int(cg.n_kminmers),
                              n_distinct=int(cg.n_distinct), n_nodes=int(cg.n_nodes), n_edges=int(cg.n_edges))
            ctx.graph_free(cg)
            return nb

        for _ in range(max(1, warmup)):
            step_e2e()
        ctx.sync(); barrier()
        for k_ in e2e_parts:
            e2e_parts[k_] = 0.0
        t0 = time.perf_counter()
        for _ in range(steps):
            d2h_bytes = step_e2e()
        ctx.sync()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        barrier()
        tm_last = ctx.timings()
        packed_up = bool(tm_last.get("upload_packed", 0))
        try:
            counts_equal = bool(e2e_counts) and all(int(stats[k_]) == v_ for k_, v_ in e2e_counts.items())
        except Exception:
            counts_equal = None
        # bytes that crossed PCIe host->device: the bases as 2-bit planes (8 B per 32 bases) plus any 4 KiB tile
        # sent as ASCII, or the ASCII bases; plus the read offsets
        h2d_bases = int(tm_last.get("upload_h2d_bytes", total)) if packed_up else total
        e2e = {"value": bases_all * steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(h2d_bases + 8 * (reads_per_rank + 1)), "d2h_bytes_per_step": int(d2h_bytes),
               "upload": ("%s: 2-bit planes packed by %s host threads inside the call and expanded on the device; "
                          "%d of the 4 KiB tiles went unpacked (copy engine idle / bytes outside ACGT)" %
                          (os.environ.get("MDBG_UPLOAD", "hybrid"), os.environ.get("MDBG_PACK_THREADS", "min(32, nproc / local ranks)"),
                           int(tm_last.get("upload_ascii_tiles", 0)))) if packed_up else "ASCII",
               "ms_per_step": e2e_ms / steps, "timing": "host wall clock around K steps, max over ranks",
               # the graph built from the host buffers (packed / hybrid upload) against the device-resident one
               "counts_equal_device_resident": counts_equal,
               "stage_ms_per_step": {k_: v_ / steps for k_, v_ in e2e_parts.items()}}
        # the same reads handed over as the caller's own 2-bit planes (mdbg_push_reads_packed): what the path does when
        # the host keeps reads packed -- 0.25 B/base over PCIe, no packing inside the call.  Reported next to `e2e`
        # (whose input is the reference's: ASCII), never instead of it.
        if world == 1 and not args.no_extra:
            n_words = (total + 31) // 32
            hp = ctx.host_alloc_pinned(8 * n_words + 64)
            rc = m.ffi.lib().mdbg_pack_bases_host(hb, total, hp, None, min(32, os.cpu_count() or 1))
            assert rc == 0
            pk_counts = {}
            pk_parts = {"h2d": 0.0, "push": 0.0, "ka_done": 0.0, "finish": 0.0, "d2h": 0.0}

            def step_packed():
                ctx.reset()
                ctx.push_reads_packed_ptr(hp, ho, reads_per_rank)
                cg = ctx.finish_raw(want_seqlines=False)
                tm = ctx.timings()
                pk_parts["h2d"] += tm["ms_h2d"]; pk_parts["push"] += tm["ms_total_push"]; pk_parts["ka_done"] += tm["ms_ka"]
                pk_parts["finish"] += tm["ms_total_finish"]; pk_parts["d2h"] += tm["ms_d2h"]
                pk_counts.update(n_minimizers=int(cg.n_minimizers), n_kminmers=int(cg.n_kminmers),
                                 n_distinct=int(cg.n_distinct), n_nodes=int(cg.n_nodes), n_edges=int(cg.n_edges))
                ctx.graph_free(cg)

            for _ in range(max(1, warmup)):
                step_packed()
            ctx.sync()
            for k_ in pk_parts:
                pk_parts[k_] = 0.0
            t0 = time.perf_counter()
            for _ in range(steps):
                step_packed()
            ctx.sync()
            pk_ms = (time.perf_counter() - t0) * 1e3
            tm_pk = ctx.timings()
            e2e["packed_input"] = {"value": bases_all * steps / (pk_ms * 1e-3) / 1e9, "unit": "Gbases/s",
                                   "ms_per_step": pk_ms / steps, "stage_ms_per_step": {k_: v_ / steps for k_, v_ in pk_parts.items()},
                                   "h2d_bytes_per_step": int(tm_pk["upload_h2d_bytes"]) + 8 * (reads_per_rank + 1),
                                   "input": "the caller's 2-bit planes in pinned host memory (mdbg_push_reads_packed), 0.25 B/base",
                                   "counts_equal_device_resident": all(int(stats[k_]) == v_ for k_, v_ in pk_counts.items())}
            ctx.host_free_pinned(hp)
        ctx.host_free_pinned(hb); ctx.host_free_pinned(ho)

    # ------------------------------------------------------------------ N>1: the N-GPU graph == the 1-GPU graph
    parity = None
    if world > 1 and not args.no_parity:
        import numpy as np
        # a bounded job: the first P reads of every rank's shard (all of them if the whole job fits the cap)
        P = int(min(reads_per_rank, max(64, args.parity_gbases * 1e9 / 15000.0 / world)))
        nb_p = int(ro[P])
        ctx.reset()
        ctx.set_read_base(rank * P)
        ctx.push_reads_device(d_bases, d_off, P, nb_p)
        g_multi = ctx.finish(want_seqlines=False)          # rank 0: the whole graph; other ranks: counters
        ctx.reset()
        ctx.set_read_base(rank * reads_per_rank)
        ok, detail = None, ""
        if rank == 0:
            one = m.Context(m.Params(k=wl["k"], l=wl["l"], density=wl["density"], min_abundance=MIN_ABUNDANCE,
                                     presimp=PRESIMP, device=local_rank, ka_variant=args.ka_variant))
            for r_ in range(world):                        # the same reads, in rank order, through ONE context
                ro_r, tot_r = synth.plan(r_ * reads_per_rank, P)
                one_off = one.device_malloc((P + 1) * 8)
                one_b = one.device_malloc(tot_r + 64)
                synth.fill_device(one, r_ * reads_per_rank, P, ro_r, one_b, one_off)
                one.push_reads_device(one_b, one_off, P, tot_r)
                one.device_free(one_b); one.device_free(one_off)
            g_one = one.finish(want_seqlines=False)
            one.close()
            bad = [a for a in ("index", "abundance", "seqlen", "shift", "tuple", "e_n1", "e_n2", "e_o1", "e_o2", "e_ov")
                   if not np.array_equal(getattr(g_multi, a), getattr(g_one, a))]
            bad += [c_ for c_ in ("n_kminmers", "n_distinct", "n_nodes", "n_edges", "presimp_removed")
                    if g_multi.stats[c_] != g_one.stats[c_]]
            ok = not bad and g_one.stats["n_nodes"] > 0
            detail = "mismatch: " + ",".join(bad) if bad else "nodes (index, abundance, seqlen, shift, tuple) and edges bit-identical"
            parity = {"parity_vs_single_gpu": bool(ok), "detail": detail,
                      "job": "first %d reads of every rank's shard (%d reads, %.2f Gbases in all)" %
                             (P, P * world, P * world * 15000.0 / 1e9),
                      "n_nodes": int(g_one.stats["n_nodes"]), "n_edges": int(g_one.stats["n_edges"]),
                      "n_kminmers": int(g_one.stats["n_kminmers"])}
        barrier()

    # ------------------------------------------------------------------ N=1: config 2 as an extra line
    extra = None
    if world == 1 and rank == 0 and not args.no_extra and args.workload != "ecoli50x":
        w2 = WORKLOADS["ecoli50x"]
        s2 = m.Synth(genome_len=w2["genome_len"])
        n2 = s2.num_reads(w2["coverage"])
        ro2, tot2 = s2.plan(0, n2)
        c2 = m.Context(m.Params(k=w2["k"], l=w2["l"], density=w2["density"], min_abundance=MIN_ABUNDANCE, presimp=PRESIMP,
                                device=local_rank, ka_variant=args.ka_variant))
        b2 = c2.device_malloc(tot2 + 64); o2 = c2.device_malloc((n2 + 1) * 8)
        s2.fill_device(c2, 0, n2, ro2, b2, o2)

        def step2():
            c2.reset()
            c2.push_reads_device(b2, o2, n2, tot2)
            return c2.finish_device()
        for _ in range(5):
            step2()
        c2.sync(); c2.timer_start()
        ka2, st2 = 0.0, None
        for _ in range(20):
            st2 = step2()
            ka2 += c2.timings()["ms_ka_kernel"]
        ms2 = c2.timer_stop() / 20
        alg2 = tot2 + 12 * st2["n_minimizers"] + 16 * (n2 + 1)
        extra = {"ecoli50x": {"workload": w2["desc"], "value": tot2 / (ms2 * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": ms2,
                              "steps": 20, "warmup": 5, "ka_kernel_ms": ka2 / 20,
                              "ka_roofline_frac": alg2 / (ka2 / 20 * 1e-3) / 1e9 / peak,
                              "counts": {k_: int(v_) for k_, v_ in st2.items()}}}
        c2.device_free(b2); c2.device_free(o2)
        c2.close()

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = max(64, min(reads_per_rank // 2, 140000))        # <= ~2 Gbases: 10-30 s of CPU work
        v, cores, tot, sec, _ = cpu_reference_run(wl, n, 1, 0)
        cpu = {"value": v, "unit": "Gbases/s", "cores": cores, "kind": "port",
               "sample": "first %d reads (%d bases) of the workload, %.1f s" % (n, tot, sec)}

    if rank == 0:
        out = {"metric": "Gbases/sec reads->mdBG", "value": value, "unit": "Gbases/s", "n_gpus": world,
               "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
               "timing": "CUDA events on the launching stream around K steps, max over ranks",
               "wall_ms_per_step": wall_ms / steps, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roofline, "kernel_rooflines": kernel_rooflines, "cpu_baseline": cpu,
               "stage_ms_per_step": {k_: v_ / steps for k_, v_ in stage.items()},
               "nvlink_bytes_sent_per_gpu_per_step": int(tm.get("exchange_bytes", 0)),
               "record_exchange": (("kx_scatter_kernel: peer stores over NVLink (CUDA IPC inboxes)" if int(tm.get("exchange_p2p", 0))
                                    else "grouped ncclSend/ncclRecv") if world > 1 else None),
               "counts": {k_: int(v_) for k_, v_ in stats.items()}}
        if multik is not None:
            out["multik"] = multik
        if parity is not None:
            out.update(parity)
        if extra is not None:
            out["extra"] = extra
        _emit(json.dumps(out))
    ctx.device_free(d_bases); ctx.device_free(d_off)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _emit(line):
    """The one JSON line goes to the REAL stdout; see __main__."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    # Anything native libraries print on fd 1 (e.g. "NCCL version ..." under NCCL_DEBUG=VERSION)
    # is sent to stderr, so that stdout carries exactly one JSON line.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(main())
