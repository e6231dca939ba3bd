"""Diagnostic: per-tile state of the bit-sliced K-A kernel on the GPU (MDBG_BS_DEBUG_DUMP) next to the
same state from the CPU emulation of the same kernel body (tests/model).  Prints the first tiles
where they differ.  Not a test; a tool for bringing the kernel up on hardware."""
import ctypes
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
import oracle_py  # noqa: E402
import rust_mdbg_b200 as M  # noqa: E402
from helpers import pack_reads, random_reads  # noqa: E402

TILE = 4096
l, d = 12, 0.003
rng = np.random.default_rng(3)
seqs = random_reads(rng, 8, mean=9000, sd=3000, hp=0.25)
bases, off = pack_reads(seqs)
B, R = int(off[-1]), len(seqs)
n_tiles = (B + TILE - 1) // TILE
bound = oracle_py.lib().orc_hash_bound(d)

# emulator
L = ctypes.CDLL(os.path.join(HERE, "model", "libka_bitslice_model.so"))
vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
L.bs_model_run.restype = ctypes.c_int
L.bs_model_run.argtypes = [vp, vp, u64, u64, u32, u64, ctypes.c_int, u32, ctypes.c_int, u64, u64,
                           vp, vp, vp, vp, u64, vp, vp, vp, vp, vp]
cap = B // 4
tc = np.zeros(n_tiles, np.uint64); ts = np.zeros(n_tiles, np.uint64)
sh, sp = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
oro = np.zeros(R + 1, np.uint64); dl = np.zeros(n_tiles, np.uint32)
dn, st = ctypes.c_uint32(0), ctypes.c_uint64(0)
edbg = np.zeros((n_tiles, 8), np.uint32)
group = int(os.environ.get("MDBG_BS_GROUP", "1"))
assert L.bs_model_run(bases.ctypes.data, off.ctypes.data, R, B, l, bound, 1, group, 1, 0, 0, tc.ctypes.data,
                      ts.ctypes.data, sh.ctypes.data, sp.ctypes.data, cap, oro.ctypes.data, dl.ctypes.data,
                      ctypes.byref(dn), ctypes.byref(st), edbg.ctypes.data) == 0

dump = os.path.join(ROOT, "gpurun_out", "bs_dbg.bin")
os.makedirs(os.path.dirname(dump), exist_ok=True)
os.environ["MDBG_BS_DEBUG_DUMP"] = dump
os.environ["MDBG_BS_GROUP"] = str(group)
with M.Context(M.Params(k=5, l=l, density=d, ka_variant=2)) as ctx:
    h, p, mo = ctx.extract_minimizers(bases, off)
    tm = ctx.timings()
raw = open(dump, "rb").read()
nt = struct.unpack("<Q", raw[:8])[0]
gdbg = np.frombuffer(raw[8:8 + nt * 32], np.uint32).reshape(nt, 8)
gcnt = np.frombuffer(raw[8 + nt * 32:], np.uint64)
print("tiles", nt, n_tiles, "variant", tm["ka_variant_used"], "dirty", tm["ka_dirty_tiles"], "gpu minimizers", len(h),
      "emu minimizers", int(st.value))
names = ["Ctile", "Ctotal", "flags", "qn", "accepted", "CA0", "CB0", "mraw0"]
shown = 0
for t in range(nt):
    if not np.array_equal(gdbg[t], edbg[t]) or int(gcnt[t]) != int(tc[t]):
        print("tile", t, "gpu", dict(zip(names, ["%x" % x for x in gdbg[t]])), "cnt", int(gcnt[t]))
        print("      emu", dict(zip(names, ["%x" % x for x in edbg[t]])), "cnt", int(tc[t]))
        shown += 1
        if shown >= 4:
            break
print("differing tiles shown:", shown)
