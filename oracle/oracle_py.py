"""ctypes binding of the CPU ORACLE (oracle/libmdbg_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg; never from the product package.
See oracle/mdbg_oracle.h for the reference file:line each function follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmdbg_oracle.so")


def build(force=False):
    """Compile the oracle with g++ (seconds)."""
    src = os.path.join(_HERE, "mdbg_oracle.cpp")
    hdr = os.path.join(_HERE, "mdbg_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libmdbg_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OrcParams(ctypes.Structure):
    _fields_ = [("k", ctypes.c_uint32), ("l", ctypes.c_uint32), ("density", ctypes.c_double),
                ("min_abundance", ctypes.c_uint32), ("presimp", ctypes.c_float),
                ("hpc", ctypes.c_int32), ("bf", ctypes.c_int32)]


class OrcStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("n_reads", "n_bases", "n_hpc_bases", "n_minimizers", "n_kminmers", "n_distinct",
                 "n_nodes", "n_edges", "presimp_removed", "n_seqlines")] + \
               [("error", ctypes.c_int64), ("error_read", ctypes.c_uint64),
                ("error_offset", ctypes.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_LIB_PATH)
    vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
    PP = ctypes.POINTER(OrcParams)
    L.orc_ntf64.argtypes = [vp, u64, u32, vp]
    L.orc_ntr64.argtypes = [vp, u64, u32, vp]
    L.orc_ntc64.argtypes = [vp, u64, u32, vp]
    L.orc_nthash_iter.restype = ctypes.c_int64
    L.orc_nthash_iter.argtypes = [vp, u64, u32, vp]
    L.orc_hash_bound.restype = u64
    L.orc_hash_bound.argtypes = [ctypes.c_double]
    L.orc_encode_rle.restype = u64
    L.orc_encode_rle.argtypes = [vp, u64, vp, vp]
    L.orc_extract.restype = ctypes.c_int64
    L.orc_extract.argtypes = [vp, u64, PP, vp, vp, u64, vp]
    L.orc_normalize.argtypes = [vp, u32, vp, vp]
    L.orc_revcomp.argtypes = [vp, u64, vp]
    L.orc_build.restype = vp
    L.orc_build.argtypes = [vp, vp, u64, PP]
    L.orc_build_mt.restype = vp
    L.orc_build_mt.argtypes = [vp, vp, u64, PP, ctypes.c_int]
    L.orc_graph_free.argtypes = [vp]
    L.orc_graph_stats.argtypes = [vp, ctypes.POINTER(OrcStats)]
    L.orc_graph_nodes.argtypes = [vp] * 6
    L.orc_graph_edges.argtypes = [vp] * 6
    L.orc_graph_seqlines.argtypes = [vp] * 7
    L.orc_graph_minimizers.restype = u64
    L.orc_graph_minimizers.argtypes = [vp] * 4
    L.orc_write_gfa.argtypes = [vp, ctypes.c_char_p]
    L.orc_read_stats.restype = ctypes.c_int64
    L.orc_read_stats.argtypes = [vp, vp, vp, u64, vp, vp, u64]
    L.orc_write_sequences.argtypes = [vp, vp, vp, ctypes.c_char_p]
    _lib = L
    return L


def _buf(b):
    if isinstance(b, str):
        b = b.encode()
    if isinstance(b, (bytes, bytearray)):
        return np.frombuffer(bytes(b), dtype=np.uint8)
    return np.ascontiguousarray(b, dtype=np.uint8)


def params(k, l, density, min_abundance=2, presimp=0.01, hpc=True, bf=False):
    return OrcParams(k, l, float(density), int(min_abundance), float(presimp), 1 if hpc else 0, 1 if bf else 0)


def hash_bound(density):
    return int(lib().orc_hash_bound(float(density)))


def _nt(fn, s, i, k):
    a = _buf(s)
    out = ctypes.c_uint64(0)
    rc = fn(a.ctypes.data, i, k, ctypes.byref(out))
    if rc != 0:
        raise ValueError("Non-ACGTN nucleotide encountered!")
    return out.value


def ntf64(s, i, k): return _nt(lib().orc_ntf64, s, i, k)
def ntr64(s, i, k): return _nt(lib().orc_ntr64, s, i, k)
def ntc64(s, i, k): return _nt(lib().orc_ntc64, s, i, k)


def nthash_iter(s, k):
    a = _buf(s)
    out = np.zeros(max(1, len(a)), dtype=np.uint64)
    n = lib().orc_nthash_iter(a.ctypes.data, len(a), k, out.ctypes.data)
    if n == -1:
        raise ValueError("Non-ACGTN nucleotide encountered!")
    if n == -2:
        raise ValueError("k out of range")
    return out[:n].copy()


def encode_rle(s):
    a = _buf(s)
    h = np.zeros(max(1, len(a)), dtype=np.uint8)
    p = np.zeros(max(1, len(a)), dtype=np.uint64)
    n = lib().orc_encode_rle(a.ctypes.data, len(a), h.ctypes.data, p.ctypes.data)
    return h[:n].copy(), p[:n].copy()


def extract(s, l, density, hpc=True):
    """Read::extract_density -> (hashes u64[], raw positions u64[])."""
    a = _buf(s)
    p = params(1, l, density, hpc=hpc)
    cap = len(a) + 1
    h = np.zeros(cap, dtype=np.uint64)
    ps = np.zeros(cap, dtype=np.uint64)
    bad = ctypes.c_uint64(0)
    n = lib().orc_extract(a.ctypes.data, len(a), ctypes.byref(p), h.ctypes.data, ps.ctypes.data,
                          cap, ctypes.byref(bad))
    if n < 0:
        raise ValueError("Non-ACGTN nucleotide encountered at %d" % bad.value)
    return h[:n].copy(), ps[:n].copy()


def normalize(t):
    a = np.ascontiguousarray(t, dtype=np.uint64)
    out = np.zeros_like(a)
    rev = ctypes.c_int(0)
    lib().orc_normalize(a.ctypes.data, len(a), out.ctypes.data, ctypes.byref(rev))
    return out, bool(rev.value)


def revcomp(s):
    a = _buf(s)
    out = np.zeros(len(a), dtype=np.uint8)
    lib().orc_revcomp(a.ctypes.data, len(a), out.ctypes.data)
    return out.tobytes()


class Graph:
    """Result of the serial-order oracle build, copied into numpy arrays."""

    def __init__(self, handle, k, bases, read_off):
        L = lib()
        self._h = handle
        self.k = k
        st = OrcStats()
        L.orc_graph_stats(handle, ctypes.byref(st))
        self.stats = {n: getattr(st, n) for n, _ in OrcStats._fields_}
        S, E, Q = st.n_nodes, st.n_edges, st.n_seqlines
        self.index = np.zeros(S, np.uint32)
        self.abundance = np.zeros(S, np.uint16)
        self.seqlen = np.zeros(S, np.uint32)
        self.shift = np.zeros((S, 2), np.uint16)
        self.tuple = np.zeros((S, k), np.uint64)
        self.e_n1 = np.zeros(E, np.uint32); self.e_o1 = np.zeros(E, np.uint8)
        self.e_n2 = np.zeros(E, np.uint32); self.e_o2 = np.zeros(E, np.uint8)
        self.e_ov = np.zeros(E, np.uint32)
        self.q_index = np.zeros(Q, np.uint32); self.q_read = np.zeros(Q, np.uint64)
        self.q_start = np.zeros(Q, np.uint64); self.q_end = np.zeros(Q, np.uint64)
        self.q_rev = np.zeros(Q, np.uint8); self.q_shift = np.zeros((Q, 2), np.uint64)
        if st.error == 0:
            L.orc_graph_nodes(handle, self.index.ctypes.data, self.abundance.ctypes.data,
                              self.seqlen.ctypes.data, self.shift.ctypes.data, self.tuple.ctypes.data)
            L.orc_graph_edges(handle, self.e_n1.ctypes.data, self.e_o1.ctypes.data,
                              self.e_n2.ctypes.data, self.e_o2.ctypes.data, self.e_ov.ctypes.data)
            L.orc_graph_seqlines(handle, self.q_index.ctypes.data, self.q_read.ctypes.data,
                                 self.q_start.ctypes.data, self.q_end.ctypes.data,
                                 self.q_rev.ctypes.data, self.q_shift.ctypes.data)
        M = st.n_minimizers
        self.m_hash = np.zeros(M, np.uint64); self.m_pos = np.zeros(M, np.uint64)
        self.m_off = np.zeros(st.n_reads + 1, np.uint64)
        if st.error == 0 and st.n_reads + 1 == len(read_off):
            L.orc_graph_minimizers(handle, self.m_hash.ctypes.data, self.m_pos.ctypes.data,
                                   self.m_off.ctypes.data)
        self._bases, self._read_off = bases, read_off

    def read_stats(self, bases, read_off):
        """--read-stats (main.rs:939-975) -> (counts u32[K], first count of every read u64[R+1])."""
        b = _buf(bases)
        ro = np.ascontiguousarray(read_off, dtype=np.uint64)
        R = len(ro) - 1
        off = np.zeros(R + 1, np.uint64)
        n = lib().orc_read_stats(self._h, b.ctypes.data, ro.ctypes.data, R, None, off.ctypes.data, 0)
        if n < 0:
            raise ValueError("Non-ACGTN nucleotide encountered!")
        cnt = np.zeros(n, np.uint32)
        lib().orc_read_stats(self._h, b.ctypes.data, ro.ctypes.data, R, cnt.ctypes.data, off.ctypes.data, n)
        return cnt, off

    def write_gfa(self, path):
        return lib().orc_write_gfa(self._h, path.encode())

    def write_sequences(self, path):
        return lib().orc_write_sequences(self._h, self._bases.ctypes.data,
                                         self._read_off.ctypes.data, path.encode())

    def close(self):
        if self._h:
            lib().orc_graph_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def build_graph(bases, read_off, k, l, density, min_abundance=2, presimp=0.01, hpc=True,
                threads=0, bf=False):
    """threads == 0: serial-order oracle; threads >= 1: reference thread structure (timing)."""
    b = _buf(bases)
    ro = np.ascontiguousarray(read_off, dtype=np.uint64)
    p = params(k, l, density, min_abundance, presimp, hpc, bf)
    L = lib()
    if threads:
        h = L.orc_build_mt(b.ctypes.data, ro.ctypes.data, len(ro) - 1, ctypes.byref(p), threads)
    else:
        h = L.orc_build(b.ctypes.data, ro.ctypes.data, len(ro) - 1, ctypes.byref(p))
    return Graph(h, k, b, ro)
