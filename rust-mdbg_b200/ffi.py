"""ctypes binding of libmdbg_b200.so (the C ABI in include/mdbg.h).

The library is the product: hand-written sm_100a kernels behind plain-C entry points.  This
module only marshals numpy arrays to pointers.  There is NO fallback: if the shared library
is missing the import fails, and without a CUDA device `Context()` raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDBG_LIB", os.path.join(_HERE, "libmdbg_b200.so"))   # MDBG_LIB: tuning variants

MDBG_OK = 0
ERR_NAMES = {0: "MDBG_OK", -1: "MDBG_ERR_NO_DEVICE", -2: "MDBG_ERR_CUDA", -3: "MDBG_ERR_BAD_ARG",
             -4: "MDBG_ERR_ALPHABET", -5: "MDBG_ERR_CAPACITY", -6: "MDBG_ERR_RANGE",
             -7: "MDBG_ERR_NCCL", -8: "MDBG_ERR_IO", -9: "MDBG_ERR_UNSUPPORTED"}

vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32


class MdbgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, str(code)), msg))
        self.code = code


class CParams(ctypes.Structure):
    _fields_ = [("k", u32), ("l", u32), ("density", ctypes.c_double), ("min_abundance", u32),
                ("presimp", ctypes.c_float), ("hpc", i32), ("device", i32), ("keep_bases", i32),
                ("debug_fp_bits", u32), ("bf", u32), ("ka_variant", u32), ("reserved", u32 * 5)]


class CGraph(ctypes.Structure):
    _fields_ = [(n, u64) for n in ("n_reads", "n_bases", "n_minimizers", "n_kminmers", "n_distinct",
                                   "n_nodes", "n_edges", "presimp_removed", "n_seqlines")] + \
               [("k", u32), ("l", u32)] + \
               [(n, vp) for n in ("node_index", "abundance", "seqlen", "shift", "tuple",
                                  "e_n1", "e_o1", "e_n2", "e_o2", "e_overlap",
                                  "q_index", "q_read", "q_start", "q_end", "q_reversed", "q_shift",
                                  "_owner")]


class CTimings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("ms_h2d", "ms_ka", "ms_kb", "ms_kc", "ms_kd", "ms_ke",
                                              "ms_d2h", "ms_total_push", "ms_total_finish")] + \
               [("launches_push", u64), ("launches_finish", u64), ("ka_launches", u64),
                ("ka_ms_sum", ctypes.c_float), ("table_attempts", u32), ("ka_dense_tiles", u32),
                ("ms_ka_kernel", ctypes.c_float), ("ms_ka_start", ctypes.c_float),
                ("ka_variant_used", u32), ("ka_dirty_tiles", u32), ("upload_packed", u32),
                ("upload_ascii_tiles", u32), ("upload_h2d_bytes", u64), ("ms_exchange", ctypes.c_float),
                ("exchange_bytes", u64), ("exchange_p2p", u32), ("ms_kernels", ctypes.c_float * 8)]


class CSynth(ctypes.Structure):
    _fields_ = [("genome_len", u64), ("mean_len", ctypes.c_double), ("sd_len", ctypes.c_double),
                ("min_len", u64), ("max_len", u64), ("error_rate", ctypes.c_double), ("seed", u64)]


# every symbol include/mdbg.h declares: name -> (restype, argtypes)
PP, GP = ctypes.POINTER(CParams), ctypes.POINTER(CGraph)
SYMBOLS = {
    "mdbg_version": (ctypes.c_char_p, []),
    "mdbg_device_count": (ctypes.c_int, []),
    "mdbg_ctx_create": (ctypes.c_int, [PP, ctypes.POINTER(vp)]),
    "mdbg_ctx_destroy": (None, [vp]),
    "mdbg_last_error": (ctypes.c_char_p, [vp]),
    "mdbg_ctx_set_k": (ctypes.c_int, [vp, u32, u32, ctypes.c_float]),
    "mdbg_hash_bound": (u64, [ctypes.c_double]),
    "mdbg_extract_minimizers": (ctypes.c_int, [vp, vp, vp, u64, vp, vp, vp, u64, ctypes.POINTER(u64)]),
    "mdbg_read_extract": (ctypes.c_int, [vp, vp, u64, vp, vp, u64, ctypes.POINTER(u64)]),
    "mdbg_kminmer_normalize": (None, [vp, u32, vp, ctypes.POINTER(ctypes.c_int)]),
    "mdbg_kminmer_reverse": (None, [vp, u32, vp]),
    "mdbg_kminmer_prefix": (None, [vp, u32, vp]),
    "mdbg_kminmer_suffix": (None, [vp, u32, vp]),
    "mdbg_kminmer_cmp": (ctypes.c_int, [vp, vp, u32]),
    "mdbg_window": (ctypes.c_int, [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, u64, ctypes.POINTER(u64)]),
    "mdbg_push_reads": (ctypes.c_int, [vp, vp, vp, u64]),
    "mdbg_push_reads_packed": (ctypes.c_int, [vp, vp, vp, u64]),
    "mdbg_push_reads_device": (ctypes.c_int, [vp, vp, vp, u64, u64]),
    "mdbg_reset": (ctypes.c_int, [vp]),
    "mdbg_finish": (ctypes.c_int, [vp, ctypes.c_int, GP]),
    "mdbg_finish_device": (ctypes.c_int, [vp, GP]),
    "mdbg_graph_free": (None, [GP]),
    "mdbg_get_minimizers": (ctypes.c_int, [vp, vp, vp, vp, u64, ctypes.POINTER(u64)]),
    "mdbg_get_timings": (ctypes.c_int, [vp, ctypes.POINTER(CTimings)]),
    "mdbg_stream": (vp, [vp]),
    "mdbg_timer_start": (ctypes.c_int, [vp]),
    "mdbg_timer_stop": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_float)]),
    "mdbg_synth_num_reads": (u64, [ctypes.POINTER(CSynth), ctypes.c_double]),
    "mdbg_synth_plan": (u64, [ctypes.POINTER(CSynth), u64, u64, vp, vp, vp]),
    "mdbg_synth_fill_device": (ctypes.c_int, [vp, ctypes.POINTER(CSynth), u64, u64, vp, vp, vp]),
    "mdbg_synth_fill_host": (None, [ctypes.POINTER(CSynth), u64, u64, vp, vp, ctypes.c_int]),
    "mdbg_device_malloc": (ctypes.c_int, [vp, u64, ctypes.POINTER(vp)]),
    "mdbg_device_free": (ctypes.c_int, [vp, vp]),
    "mdbg_host_alloc_pinned": (ctypes.c_int, [u64, ctypes.POINTER(vp)]),
    "mdbg_host_free_pinned": (ctypes.c_int, [vp]),
    "mdbg_memcpy_h2d": (ctypes.c_int, [vp, vp, vp, u64]),
    "mdbg_memcpy_d2h": (ctypes.c_int, [vp, vp, vp, u64]),
    "mdbg_sync": (ctypes.c_int, [vp]),
    "mdbg_flush_l2": (ctypes.c_int, [vp]),
    "mdbg_nccl_unique_id": (ctypes.c_int, [vp]),
    "mdbg_comm_init": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int]),
    "mdbg_comm_set_read_base": (ctypes.c_int, [vp, u64]),
    "mdbg_shard_reads": (None, [u64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "mdbg_owner_of_fingerprint": (u32, [u64, ctypes.c_int]),
    "mdbg_tuple_fingerprint": (u64, [vp, u32, u64]),
    "mdbg_write_gfa": (ctypes.c_int, [GP, ctypes.c_char_p]),
    "mdbg_write_sequences": (ctypes.c_int, [GP, vp, vp, ctypes.c_char_p, ctypes.c_int]),
    "mdbg_seq_writer_open": (ctypes.c_int, [GP, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]),
    "mdbg_seq_writer_open_part": (ctypes.c_int, [GP, ctypes.c_char_p, ctypes.c_int, u32, u32, ctypes.POINTER(vp)]),
    "mdbg_seq_writer_next_read": (u64, [vp]),
    "mdbg_seq_writer_read": (ctypes.c_int, [vp, u64, vp, u64]),
    "mdbg_seq_writer_close": (ctypes.c_int, [vp]),
    "mdbg_pack_bases_host": (ctypes.c_int, [vp, u64, vp, vp, ctypes.c_int]),
    "mdbg_read_stats": (ctypes.c_int, [vp, vp, vp, u64, vp, vp, u64, ctypes.POINTER(u64)]),
}

_lib = None


def lib():
    """Load libmdbg_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmdbg_b200.so is missing (%s): run `make` or "
                              "`python -c 'import __graft_entry__ as g; g.build()'`; "
                              "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)   # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data


def as_u8(b):
    if isinstance(b, str):
        b = b.encode()
    if isinstance(b, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(b), dtype=np.uint8)
    return np.ascontiguousarray(b, dtype=np.uint8)
