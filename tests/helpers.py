"""shared test helpers: hand-made FASTQ edge cases and oracle glue (oracle/ is only ever used as the checker)."""
import numpy as np

from oracle import dsk, image as oimg


def fastq(reads, quals=None, final_newline=True, headers=None):
    out = []
    for i, r in enumerate(reads):
        h = headers[i] if headers else f"@r{i}"
        q = quals[i] if quals else "I" * len(r)
        out.append(f"{h}\n{r}\n+\n{q}\n")
    s = "".join(out).encode("ascii")
    return s if final_newline else s[:-1]


def rand_reads(rng, n, lmin, lmax, p_n=0.01, alphabet="ACGT"):
    reads = []
    for _ in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        a = rng.choice(list(alphabet), L)
        if p_n:
            a[rng.random(L) < p_n] = "N"
        reads.append("".join(a))
    return reads


def oracle_levels(buf, k, seed, levels, nsites, read_index_base=0, hist=None, calibrated=True):
    """canon_full per level with the project's selection rule, computed by the CPU oracle.  ``calibrated`` (the product's
    default, VK_SAMPLING_CALIBRATED): thresholds fitted to the base targets from the histogram of the WHOLE sample -- the
    buffer's own reads, or ``hist`` when the buffer is a shard."""
    p = dsk.parse_fastq(buf)
    if calibrated and hist is None:
        hist = dsk.prio_hist(p["lens"], seed, read_index_base)
    out = []
    for bp in levels:
        sel = dsk.select_reads(p["n_reads"], seed, bp, nsites, read_index_base, hist=hist if calibrated else None)
        fwd = dsk.count_forward(buf, p["starts"], p["lens"], k, sel)
        out.append(dsk.fold_canonical(fwd, k))
    return np.stack(out) if out else np.zeros((0, 4 ** k), dtype=np.uint64)


def oracle_images(canon_levels, lut):
    return np.stack([oimg.image_exact(c, lut) for c in canon_levels])


# ---------------------------------------------------------------------------------------------------------
# CPU stand-in for varkoder_b200.engine.Engine, built on the oracle.  TESTS ONLY: it lets the host-side logic that
# drives an engine (the read-sharded driver under gloo, the reference's own run_clean2img with the stage functions
# rebound) run on a machine without a GPU.  The product never sees it.
# ---------------------------------------------------------------------------------------------------------
class OracleEngine:
    """Engine methods the host code uses; per-segment forward histograms, lex index."""
    device = 0

    def __init__(self):
        self._own = None

    def upload(self, buf):
        self.buf = bytes(buf)
        return len(self.buf)

    def parse(self):
        self.p = dsk.parse_fastq(self.buf)
        return dict(n_bytes=len(self.buf), n_lines=self.p["n_lines"], n_reads=self.p["n_reads"],
                    nsites=self.p["nsites_ref"], nsites_true=self.p["nsites_true"])

    def _seg(self, seg_hist_ptr, nk):
        import ctypes
        from varkoder_b200 import _lib
        if seg_hist_ptr is None:
            if self._own is None or self._own.shape[1] != nk:
                self._own = np.zeros((_lib.VK_MAX_LEVELS, nk), dtype=np.uint64)
            return self._own
        return np.ctypeslib.as_array(ctypes.cast(seg_hist_ptr, ctypes.POINTER(ctypes.c_uint64)),
                                     shape=(_lib.VK_MAX_LEVELS, nk))

    def prio_hist(self, params, hist_ptr):
        import ctypes
        from varkoder_b200 import _lib
        from varkoder_b200.ladder import parse_seed
        h = np.ctypeslib.as_array(ctypes.cast(hist_ptr, ctypes.POINTER(ctypes.c_int64)), shape=(_lib.VK_PRIO_BUCKETS,))
        h += dsk.prio_hist(self.p["lens"], parse_seed(params.seed), params.read_index_base)

    def count(self, params, seg_hist_ptr=None):
        from varkoder_b200 import _lib
        from varkoder_b200.engine import Result
        from varkoder_b200.ladder import LessThanMinimumData, ladder, parse_seed
        p, k = self.p, params.k
        nk = 4 ** k
        nsites = params.nsites_override or p["nsites_ref"]
        try:
            levels = ladder(nsites, params.min_bp, params.max_bp, params.is_query)
            status = 0
        except LessThanMinimumData:
            levels, status = [], _lib.VK_LADDER_LESS_THAN_MIN
        seed = parse_seed(params.seed)
        hist = None
        if params.sampling == _lib.VK_SAMPLING_CALIBRATED:
            if params.prio_hist:
                import ctypes
                hist = np.ctypeslib.as_array(ctypes.cast(params.prio_hist, ctypes.POINTER(ctypes.c_int64)),
                                             shape=(_lib.VK_PRIO_BUCKETS,)).copy()
            else:
                hist = dsk.prio_hist(p["lens"], seed, params.read_index_base)
        member = [dsk.select_reads(p["n_reads"], seed, bp, nsites, params.read_index_base, hist=hist).astype(bool)
                  for bp in levels]
        long_enough = p["lens"] >= k
        out = self._seg(seg_hist_ptr, nk)
        out[:] = 0
        reads, bases = [], []
        for s in range(len(levels)):
            sel = member[s] & ~(member[s + 1] if s + 1 < len(levels) else np.zeros_like(member[s]))
            out[s] = dsk.count_forward(self.buf, p["starts"], p["lens"], k, sel.astype(np.uint8))
            reads.append(int((member[s] & long_enough).sum()))
            bases.append(int(p["lens"][member[s] & long_enough].sum()))
        return Result(len(self.buf), p["n_lines"], p["n_reads"], p["nsites_ref"], p["nsites_true"], status,
                      levels, reads, bases)

    def render(self, table, k, n_levels, seg_hist_ptr=None, want_canon=True):
        nk = 4 ** k
        seg = self._seg(seg_hist_ptr, nk)
        cum = np.cumsum(seg[:n_levels][::-1], axis=0, dtype=np.uint64)[::-1]
        canon = np.stack([dsk.fold_canonical(c, k) for c in cum]) if n_levels else np.zeros((0, nk), np.uint64)
        pixels = np.stack([oimg.image_exact(c, table.lut) for c in canon]) if n_levels and table is not None else None
        return (canon if want_canon else None), pixels

    def render_counts(self, table, canon):
        canon = np.ascontiguousarray(canon, dtype=np.uint64)
        if canon.ndim == 1:
            canon = canon[None, :]
        return np.stack([oimg.image_exact(c, table.lut) for c in canon])
