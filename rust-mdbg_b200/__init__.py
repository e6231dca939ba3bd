"""rust-mdbg_b200: B200-native (sm_100a) reads -> minimizer-space de Bruijn graph hot path of
ekimb/rust-mdbg behind the reference's Read / minimizers / KmerVec module surface.
All compute is in libmdbg_b200.so (hand-written CUDA behind the C ABI of include/mdbg.h)."""
from . import ffi, minimizers
from .engine import Context, Graph, Params, Synth, nccl_unique_id, pack_bases
from .ffi import MdbgError
from .kmer_vec import KmerVec
from .read import Read

__all__ = ["Context", "Graph", "Params", "Synth", "KmerVec", "Read", "MdbgError", "minimizers", "ffi",
           "nccl_unique_id"]
