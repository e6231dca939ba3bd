"""KmerVec -- mirror of the reference's k-min-mer value type (src/kmer_vec.rs).

Same method names and semantics: `make_from`, `prefix`, `suffix`, `reverse`, `normalize`
(returns (KmerVec, reversed); a palindromic tuple is reported reversed, kmer_vec.rs:37-38),
`print_as_string` (Rust `{:?}` of Vec<u64>), and Eq/Hash/Ord = lexicographic on the u64 vector
(kmer_vec.rs:54-84).  Single-tuple value ops go through the C ABI helpers; the batch
canonicalisation of every window of every read is the K-B kernel (Context.window / finish).
"""
import ctypes
import functools

import numpy as np

from . import ffi


@functools.total_ordering
class KmerVec:
    __slots__ = ("data",)

    def __init__(self, data):
        self.data = np.ascontiguousarray(data, dtype=np.uint64)

    @staticmethod
    def make_from(ar):                       # kmer_vec.rs:41
        return KmerVec(np.array(ar, dtype=np.uint64))

    def suffix(self):                        # kmer_vec.rs:16-20
        out = np.zeros(len(self.data) - 1, np.uint64)
        ffi.lib().mdbg_kminmer_suffix(ffi.ptr(self.data), len(self.data), ffi.ptr(out))
        return KmerVec(out)

    def prefix(self):                        # kmer_vec.rs:22-26
        out = np.zeros(len(self.data) - 1, np.uint64)
        ffi.lib().mdbg_kminmer_prefix(ffi.ptr(self.data), len(self.data), ffi.ptr(out))
        return KmerVec(out)

    def reverse(self):                       # kmer_vec.rs:28-32
        out = np.zeros(len(self.data), np.uint64)
        ffi.lib().mdbg_kminmer_reverse(ffi.ptr(self.data), len(self.data), ffi.ptr(out))
        return KmerVec(out)

    def normalize(self):                     # kmer_vec.rs:34-39
        out = np.zeros(len(self.data), np.uint64)
        rev = ctypes.c_int(0)
        ffi.lib().mdbg_kminmer_normalize(ffi.ptr(self.data), len(self.data), ffi.ptr(out), ctypes.byref(rev))
        return KmerVec(out), bool(rev.value)

    def print_as_string(self):               # kmer_vec.rs:45-47
        return "[" + ", ".join(str(int(x)) for x in self.data) + "]"

    def __eq__(self, o):
        return len(self.data) == len(o.data) and bool(np.array_equal(self.data, o.data))

    def __lt__(self, o):
        n = min(len(self.data), len(o.data))
        c = ffi.lib().mdbg_kminmer_cmp(ffi.ptr(self.data), ffi.ptr(o.data), n) if n else 0
        return c < 0 or (c == 0 and len(self.data) < len(o.data))

    def __hash__(self):
        return hash(self.data.tobytes())

    def __repr__(self):
        return "KmerVec { data: %s }" % self.print_as_string()
