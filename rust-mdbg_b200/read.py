"""Read -- mirror of the reference's `Read` (src/read.rs:12-20) and `Read::extract`
(src/read.rs:85-90 -> extract_density, read.rs:176-211) on the GPU engine."""
import numpy as np

from .ffi import as_u8


class Read:
    __slots__ = ("id", "minimizers", "minimizers_pos", "transformed", "seq", "corrected")

    def __init__(self, id_, minimizers_pos, transformed, seq):
        self.id = id_
        self.minimizers = []              # stays empty on the density path (read.rs:179)
        self.minimizers_pos = minimizers_pos  # raw coordinates (read.rs:206-207)
        self.transformed = transformed        # canonical ntHash values <= hash_bound (read.rs:208)
        self.seq = seq                        # the RAW sequence (read.rs:210)
        self.corrected = False

    @staticmethod
    def extract(inp_id, inp_seq, ctx):
        """Read::extract(inp_id, inp_seq, &params, ..) with params bound in `ctx` (a Context).
        The --uhs/--lcp/--syncmers schemes of read.rs:86-88 are outside the hot path."""
        h, p = ctx.read_extract(inp_seq)
        return Read(inp_id, p, h, inp_seq if isinstance(inp_seq, (bytes, str)) else bytes(as_u8(inp_seq)))

    @staticmethod
    def extract_batch(ids, bases, read_off, ctx):
        """All reads of a batch in one kernel launch."""
        h, p, mo = ctx.extract_minimizers(bases, read_off)
        b = as_u8(bases)
        out = []
        for r, name in enumerate(ids):
            lo, hi = int(mo[r]), int(mo[r + 1])
            out.append(Read(name, p[lo:hi], h[lo:hi], b[int(read_off[r]):int(read_off[r + 1])].tobytes()))
        return out

    def read_to_kmers(self, ctx):
        """src/read.rs:358-413 (twin of main.rs:756-781): canonical k-min-mers of this read as
        (KmerVec, reversed, shift pair, read_offsets) via the K-B kernel."""
        from .kmer_vec import KmerVec
        mo = np.array([0, len(self.transformed)], np.uint64)
        tup, rev, sh, of, _ = ctx.window(self.transformed, self.minimizers_pos, mo)
        return [(KmerVec(tup[i]), bool(rev[i]), (int(sh[i, 0]), int(sh[i, 1])),
                 (int(of[i, 0]), int(of[i, 1]), int(of[i, 2]))) for i in range(len(rev))]
