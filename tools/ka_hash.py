"""Hash of the K-A kernel sources: ties an ncu DRAM-traffic capture (profiles/ka_traffic*.json) to the
kernel it was taken from.  bench.py reports `roofline.traffic` only while the hash recorded in the
capture equals the hash of the sources the running library was built from."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rust-mdbg_b200", "csrc")
FILES = {
    "bitslice": ["ka_bitslice_body.h", "ka_bitslice_math.h", "ka_bitslice.cu", "mdbg_kernels.h", "mdbg_common.cuh"],
    "classic": ["ka_minimizers.cu", "mdbg_kernels.h", "mdbg_common.cuh"],
}


def ka_source_hash(variant):
    h = hashlib.sha256()
    for name in FILES[variant]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]
