"""Context / Graph: thin object layer over the C ABI (one Context per GPU)."""
import ctypes

import numpy as np

from . import ffi
from .ffi import CGraph, CParams, CSynth, CTimings, MdbgError, as_u8, ptr


class Params:
    """The members of the reference's `Params` (src/main.rs:92-114) the hot path reads."""

    def __init__(self, k=10, l=12, density=0.10, min_abundance=2, presimp=0.01, hpc=True,
                 device=0, debug_fp_bits=0, bf=False, ka_variant=0):
        # defaults as in src/main.rs:437-450 (k=10, l=12, density=0.10, minabund=2, presimp=0.01)
        self.k, self.l, self.density = int(k), int(l), float(density)
        self.min_abundance, self.presimp, self.hpc = int(min_abundance), float(presimp), bool(hpc)
        self.device, self.debug_fp_bits, self.bf = int(device), int(debug_fp_bits), bool(bf)
        self.ka_variant = {"default": 0, "classic": 1, "bitslice": 2}.get(ka_variant, ka_variant)

    def c(self):
        return CParams(self.k, self.l, self.density, self.min_abundance, self.presimp,
                       1 if self.hpc else 0, self.device, 0, self.debug_fp_bits, 1 if self.bf else 0,
                       int(self.ka_variant))


class Graph:
    """Host copy of a finished mdBG: nodes (ascending index), edges (sorted), .sequences lines."""

    def __init__(self, cg, copy_arrays=True):
        self.stats = {n: getattr(cg, n) for n in ("n_reads", "n_bases", "n_minimizers", "n_kminmers",
                                                  "n_distinct", "n_nodes", "n_edges", "presimp_removed",
                                                  "n_seqlines")}
        self.k, self.l = cg.k, cg.l
        S, E, Q, k = cg.n_nodes, cg.n_edges, cg.n_seqlines, cg.k

        def arr(p, n, dt, shape=None):
            if not p or not copy_arrays:
                return np.zeros(0 if shape is None else (0,) + shape[1:], dt)
            a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))),
                                      shape=(n,)).copy()
            return a if shape is None else a.reshape(shape)

        self.index = arr(cg.node_index, S, np.uint32)
        self.abundance = arr(cg.abundance, S, np.uint16)
        self.seqlen = arr(cg.seqlen, S, np.uint32)
        self.shift = arr(cg.shift, 2 * S, np.uint16, (S, 2))
        self.tuple = arr(cg.tuple, S * k, np.uint64, (S, k))
        self.e_n1 = arr(cg.e_n1, E, np.uint32); self.e_o1 = arr(cg.e_o1, E, np.uint8)
        self.e_n2 = arr(cg.e_n2, E, np.uint32); self.e_o2 = arr(cg.e_o2, E, np.uint8)
        self.e_ov = arr(cg.e_overlap, E, np.uint32)
        has_q = bool(cg.q_index)
        Q = Q if has_q else 0
        self.q_index = arr(cg.q_index, Q, np.uint32); self.q_read = arr(cg.q_read, Q, np.uint64)
        self.q_start = arr(cg.q_start, Q, np.uint64); self.q_end = arr(cg.q_end, Q, np.uint64)
        self.q_rev = arr(cg.q_reversed, Q, np.uint8)
        self.q_shift = arr(cg.q_shift, 2 * Q, np.uint64, (Q, 2))


class Context:
    """One GPU's engine.  Raises MdbgError(MDBG_ERR_NO_DEVICE) when there is no CUDA device."""

    def __init__(self, params):
        self.L = ffi.lib()
        self.params = params
        h = ffi.vp()
        cp = params.c()
        rc = self.L.mdbg_ctx_create(ctypes.byref(cp), ctypes.byref(h))
        if rc != 0:
            raise MdbgError(rc, (self.L.mdbg_last_error(None) or b"").decode())
        self.h = h
        self.n_reads = 0

    def _ck(self, rc):
        if rc != 0:
            raise MdbgError(rc, (self.L.mdbg_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mdbg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- Entry 1 ----------------------------------------------------------------------------
    def extract_minimizers(self, bases, read_off, cap=None):
        """Batch Read::extract -> (hash u64[M], pos u64[M], min_read_off u64[R+1])."""
        b = as_u8(bases)
        ro = np.ascontiguousarray(read_off, dtype=np.uint64)
        R = len(ro) - 1
        if cap is None:
            cap = int(len(b) * min(1.0, 3.0 * self.params.density + 1e-3)) + 1024
        n = ffi.u64(0)
        while True:
            h = np.zeros(cap, np.uint64); p = np.zeros(cap, np.uint64); mo = np.zeros(R + 1, np.uint64)
            rc = self.L.mdbg_extract_minimizers(self.h, ptr(b), ptr(ro), R, ptr(h), ptr(p), ptr(mo), cap,
                                                ctypes.byref(n))
            if rc == -5 and n.value > cap:   # MDBG_ERR_CAPACITY: retry with the reported size
                cap = n.value
                continue
            self._ck(rc)
            return h[:n.value].copy(), p[:n.value].copy(), mo

    def read_extract(self, seq):
        b = as_u8(seq)
        cap = len(b) + 16
        h = np.zeros(cap, np.uint64); p = np.zeros(cap, np.uint64)
        n = ffi.u64(0)
        self._ck(self.L.mdbg_read_extract(self.h, ptr(b), len(b), ptr(h), ptr(p), cap, ctypes.byref(n)))
        return h[:n.value].copy(), p[:n.value].copy()

    # ---- Entry 2 (batch) ----------------------------------------------------------------------
    def window(self, hash_, pos, min_read_off):
        h = np.ascontiguousarray(hash_, np.uint64); p = np.ascontiguousarray(pos, np.uint64)
        mo = np.ascontiguousarray(min_read_off, np.uint64)
        R = len(mo) - 1
        k = self.params.k
        m = np.diff(mo.astype(np.int64))
        K = int(np.where(m > k, m - k + 1, 0).sum())
        tup = np.zeros((K, k), np.uint64); rev = np.zeros(K, np.uint8)
        sh = np.zeros((K, 2), np.uint64); of = np.zeros((K, 3), np.uint64); ko = np.zeros(R + 1, np.uint64)
        n = ffi.u64(0)
        self._ck(self.L.mdbg_window(self.h, ptr(h), ptr(p), ptr(mo), R, ptr(tup), ptr(rev), ptr(sh), ptr(of),
                                    ptr(ko), K, ctypes.byref(n)))
        assert n.value == K
        return tup, rev, sh, of, ko

    # ---- Entry 3 ----------------------------------------------------------------------------
    def push_reads(self, bases, read_off):
        b = as_u8(bases)
        ro = np.ascontiguousarray(read_off, dtype=np.uint64)
        self._ck(self.L.mdbg_push_reads(self.h, ptr(b), ptr(ro), len(ro) - 1))
        self.n_reads += len(ro) - 1

    def push_reads_ptr(self, bases_ptr, read_off_ptr, n_reads):
        """Host pointers (e.g. pinned buffers) without numpy marshalling."""
        self._ck(self.L.mdbg_push_reads(self.h, bases_ptr, read_off_ptr, n_reads))
        self.n_reads += n_reads

    def push_reads_packed(self, planes, read_off):
        """Reads the host already holds as 2-bit planes (pack_bases): a quarter of the PCIe bytes, no packing inside."""
        pl = np.ascontiguousarray(planes, dtype=np.uint32)
        ro = np.ascontiguousarray(read_off, dtype=np.uint64)
        assert len(pl) >= 2 * ((int(ro[-1]) + 31) // 32)
        self._ck(self.L.mdbg_push_reads_packed(self.h, ptr(pl), ptr(ro), len(ro) - 1))
        self.n_reads += len(ro) - 1

    def push_reads_packed_ptr(self, planes_ptr, read_off_ptr, n_reads):
        self._ck(self.L.mdbg_push_reads_packed(self.h, planes_ptr, read_off_ptr, n_reads))
        self.n_reads += n_reads

    def push_reads_device(self, d_bases, d_read_off, n_reads, n_bases):
        self._ck(self.L.mdbg_push_reads_device(self.h, d_bases, d_read_off, n_reads, n_bases))
        self.n_reads += n_reads

    def reset(self):
        self._ck(self.L.mdbg_reset(self.h))
        self.n_reads = 0

    def set_k(self, k, min_abundance=None, presimp=None):
        if min_abundance is None:
            min_abundance = self.params.min_abundance
        if presimp is None:
            presimp = self.params.presimp
        self._ck(self.L.mdbg_ctx_set_k(self.h, k, min_abundance, presimp))
        self.params.k, self.params.min_abundance, self.params.presimp = k, min_abundance, presimp

    def finish(self, want_seqlines=True):
        cg = CGraph()
        self._ck(self.L.mdbg_finish(self.h, 1 if want_seqlines else 0, ctypes.byref(cg)))
        g = Graph(cg)
        g._c = cg
        self.L.mdbg_graph_free(ctypes.byref(cg))
        return g

    def finish_raw(self, want_seqlines=True):
        """Like finish() but returns the C struct (for the file writers); free with graph_free."""
        cg = CGraph()
        self._ck(self.L.mdbg_finish(self.h, 1 if want_seqlines else 0, ctypes.byref(cg)))
        return cg

    def graph_free(self, cg):
        self.L.mdbg_graph_free(ctypes.byref(cg))

    def finish_device(self):
        cg = CGraph()
        self._ck(self.L.mdbg_finish_device(self.h, ctypes.byref(cg)))
        return Graph(cg, copy_arrays=False).stats

    def read_stats(self, bases, read_off):
        """--read-stats (main.rs:939-975) against the graph of the last finish():
        -> (counts u32[K], first count of every read u64[R+1])."""
        b = as_u8(bases)
        ro = np.ascontiguousarray(read_off, dtype=np.uint64)
        R = len(ro) - 1
        off = np.zeros(R + 1, np.uint64)
        n = ffi.u64(0)
        rc = self.L.mdbg_read_stats(self.h, ptr(b), ptr(ro), R, None, ptr(off), 0, ctypes.byref(n))
        if rc not in (0, -5):
            self._ck(rc)
        cnt = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.L.mdbg_read_stats(self.h, ptr(b), ptr(ro), R, ptr(cnt), ptr(off), n.value, ctypes.byref(n)))
        return cnt, off

    def get_minimizers(self):
        n = ffi.u64(0)
        self._ck(self.L.mdbg_get_minimizers(self.h, None, None, None, 0, ctypes.byref(n)))
        M = n.value
        h = np.zeros(M, np.uint64); p = np.zeros(M, np.uint64)
        ro = np.zeros(self.n_reads + 1, np.uint64)
        self._ck(self.L.mdbg_get_minimizers(self.h, ptr(h), ptr(p), ptr(ro), M, ctypes.byref(n)))
        return h, p, ro

    def timings(self):
        t = CTimings()
        self._ck(self.L.mdbg_get_timings(self.h, ctypes.byref(t)))
        return {n: (list(getattr(t, n)) if n == "ms_kernels" else getattr(t, n)) for n, _ in CTimings._fields_}

    # ---- memory / sync ------------------------------------------------------------------------
    def device_malloc(self, nbytes):
        p = ffi.vp()
        self._ck(self.L.mdbg_device_malloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def device_free(self, p):
        self._ck(self.L.mdbg_device_free(self.h, p))

    def sync(self):
        self._ck(self.L.mdbg_sync(self.h))

    def timer_start(self):
        self._ck(self.L.mdbg_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        self._ck(self.L.mdbg_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def host_alloc_pinned(self, nbytes):
        p = ffi.vp()
        rc = self.L.mdbg_host_alloc_pinned(nbytes, ctypes.byref(p))
        if rc != 0:
            raise MdbgError(rc, "cudaMallocHost failed")
        return p.value

    def host_free_pinned(self, p):
        self.L.mdbg_host_free_pinned(p)

    def flush_l2(self):
        self._ck(self.L.mdbg_flush_l2(self.h))

    def h2d(self, dst, src_arr):
        self._ck(self.L.mdbg_memcpy_h2d(self.h, dst, ptr(src_arr), src_arr.nbytes))

    def d2h(self, dst_arr, src):
        self._ck(self.L.mdbg_memcpy_d2h(self.h, ptr(dst_arr), src, dst_arr.nbytes))

    # ---- multi-GPU ------------------------------------------------------------------------------
    def comm_init(self, unique_id, rank, world):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.mdbg_comm_init(self.h, buf, rank, world))

    def set_read_base(self, first_read):
        self._ck(self.L.mdbg_comm_set_read_base(self.h, first_read))


def pack_bases(bases, threads=8):
    """2-bit planes of a buffer of bases (mdbg_pack_bases_host): (planes u32[2 * ceil(n / 32)], bad_tiles u8[ceil(n / 4096)]);
    bad_tiles[t] = 1 when tile t holds a byte outside ACGT (such a batch cannot be pushed packed)."""
    b = as_u8(bases)
    n = len(b)
    planes = np.zeros(2 * ((n + 31) // 32), np.uint32)
    bad = np.zeros((n + 4095) // 4096 + 1, np.uint8)
    if n:
        rc = ffi.lib().mdbg_pack_bases_host(ptr(b), n, ptr(planes), ptr(bad), int(threads))
        if rc != 0:
            raise MdbgError(rc, "mdbg_pack_bases_host")
    return planes, bad[:(n + 4095) // 4096]


def nccl_unique_id():
    L = ffi.lib()
    buf = (ctypes.c_uint8 * 128)()
    rc = L.mdbg_nccl_unique_id(buf)
    if rc != 0:
        raise MdbgError(rc, "ncclGetUniqueId failed")
    return bytes(buf)


class Synth:
    """Synthetic HiFi-shape reads: counter-based, identical on host and device."""

    def __init__(self, genome_len, mean_len=15000.0, sd_len=4000.0, min_len=1000, max_len=60000,
                 error_rate=0.001, seed=0x6d646267):
        self.c = CSynth(int(genome_len), float(mean_len), float(sd_len), int(min_len), int(max_len),
                        float(error_rate), int(seed))
        self.L = ffi.lib()

    def num_reads(self, coverage):
        return int(self.L.mdbg_synth_num_reads(ctypes.byref(self.c), float(coverage)))

    def plan(self, first_read, n_reads):
        ro = np.zeros(n_reads + 1, np.uint64)
        total = self.L.mdbg_synth_plan(ctypes.byref(self.c), first_read, n_reads, ptr(ro), None, None)
        return ro, int(total)

    def fill_host(self, first_read, n_reads, read_off, out=None, threads=8):
        if out is None:
            out = np.zeros(int(read_off[-1]), np.uint8)
        self.L.mdbg_synth_fill_host(ctypes.byref(self.c), first_read, n_reads, ptr(read_off), ptr(out), threads)
        return out

    def fill_device(self, ctx, first_read, n_reads, read_off, d_bases, d_read_off):
        ctx._ck(self.L.mdbg_synth_fill_device(ctx.h, ctypes.byref(self.c), first_read, n_reads, ptr(read_off),
                                              d_bases, d_read_off))
