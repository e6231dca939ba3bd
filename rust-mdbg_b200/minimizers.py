"""minimizers -- the density-threshold arithmetic of the reference's minimizer selection.

`hash_bound` is src/read.rs:183.  The l-mer pre-tables of src/minimizers.rs
(`minimizers_preparation`, only used with --lmer-counts / --error-correct) and the UHS / LCP
schemes are outside the hot path and not provided."""
from . import ffi


def hash_bound(density):
    """(density as f64 * u64::MAX as f64) as u64"""
    return int(ffi.lib().mdbg_hash_bound(float(density)))
