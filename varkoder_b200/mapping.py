"""Pixel tables: which k-mer is shown at which pixel (reference: core/utils.py:152-217, image.py:900-913).

A table is reduced to ``lut[row, col]`` = lexicographic index (first base most significant, A0 C1 G2 T3) of
one k-mer of the canonical class drawn at that pixel of the FINAL image, or -1 for pixels no k-mer maps to.
``row = H-1-y``, ``col = x`` accounts for the transpose + flip of make_image (image.py:911-913).
"""
import os
from functools import lru_cache

import numpy as np

MAPPING_CHOICES = ["varKode", "cgr"]      # core/config.py:27
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "varkode_lut.npz")
_LEX = np.full(256, -1, dtype=np.int64)
for _i, _c in enumerate("ACGT"):
    _LEX[ord(_c)] = _i


def kmer_strings_to_index(kmers):
    a = np.frombuffer("".join(kmers).encode("ascii"), dtype=np.uint8).reshape(len(kmers), -1)
    codes = _LEX[a]
    if (codes < 0).any():
        raise ValueError("pixel table holds a k-mer with a letter outside ACGT")
    k = a.shape[1]
    w = 4 ** np.arange(k - 1, -1, -1, dtype=np.int64)
    return (codes * w).sum(axis=1), k


def revcomp_index(idx, k):
    idx = np.asarray(idx, dtype=np.int64)
    r = np.zeros_like(idx)
    x = idx.copy()
    for _ in range(k):
        r = (r << 2) | (3 - (x & 3))
        x >>= 2
    return r


class PixelTable:
    """What ``get_kmer_mapping`` returns here: the LUT plus what make_image reads off the DataFrame."""

    def __init__(self, lut, k, method):
        self.lut = np.ascontiguousarray(lut, dtype=np.int32)
        self.k = int(k)
        self.method = method
        if self.lut.ndim != 2 or self.lut.shape[0] != self.lut.shape[1]:
            raise ValueError("pixel tables are square")
        self.side = int(self.lut.shape[0])

    @property
    def n_unused(self):
        return int((self.lut < 0).sum())

    def __repr__(self):
        return f"PixelTable(k={self.k}, method={self.method!r}, side={self.side}, unused={self.n_unused})"


def lut_from_dataframe(df):
    """Reference DataFrame (index = k-mer string, columns x, y) -> LUT.

    make_image left-joins this table with the canonical counts and averages per pixel (image.py:900), which
    equals "the abundance of the pixel's canonical class" exactly when all rows of a pixel belong to one class
    (true for the shipped varKode tables and for get_cgr); anything else is rejected.
    """
    idx, k = kmer_strings_to_index([str(s) for s in df.index])
    x = np.asarray(df["x"]).astype(np.int64)
    y = np.asarray(df["y"]).astype(np.int64)
    H, W = int(y.max()) + 1, int(x.max()) + 1
    if H != W:
        raise ValueError("pixel tables are square")
    pix = (H - 1 - y) * W + x
    canon = np.minimum(idx, revcomp_index(idx, k))
    order = np.argsort(pix, kind="stable")
    ps, cs = pix[order], canon[order]
    same = ps[1:] == ps[:-1]
    if np.any(cs[1:][same] != cs[:-1][same]):
        raise ValueError("pixel table draws two canonical classes on one pixel: not supported")
    lut = np.full(H * W, -1, dtype=np.int32)
    lut[pix] = idx
    return lut.reshape(H, W)


def cgr_lut(k):
    """Closed form of get_cgr (core/utils.py:174-217): corners A(0,0) C(0,1) G(1,1) T(1,0); the i-th base of the
    k-mer sets bit i of x and y; side 2^k; every pixel used."""
    n = 4 ** k
    idx = np.arange(n, dtype=np.int64)
    x = np.zeros(n, dtype=np.int64)
    y = np.zeros(n, dtype=np.int64)
    xbit = np.array([0, 0, 1, 1])
    ybit = np.array([0, 1, 1, 0])
    for i in range(k):
        d = (idx >> (2 * (k - 1 - i))) & 3
        x |= xbit[d] << i
        y |= ybit[d] << i
    side = 2 ** k
    lut = np.full(side * side, -1, dtype=np.int32)
    lut[(side - 1 - y) * side + x] = idx
    return lut.reshape(side, side)


@lru_cache(maxsize=None)
def get_kmer_mapping(kmer_size=7, method="varKode"):
    """Same arguments as the reference's get_kmer_mapping (core/utils.py:152-171); returns a PixelTable."""
    k = int(kmer_size)
    if not 5 <= k <= 9:
        raise ValueError("kmer size must be between 5 and 9")          # image.py:1209
    if method == "varKode":
        if not os.path.exists(_DATA):
            raise FileNotFoundError(_DATA + " missing: run tools/make_mapping_data.py")
        with np.load(_DATA) as z:
            return PixelTable(z[f"k{k}"], k, method)
    if method == "cgr":
        return PixelTable(cgr_lut(k), k, method)
    raise Exception('method must be "varKode" or "cgr"')               # core/utils.py:169


def as_pixel_table(kmer_mapping, mapping_code=None):
    """Accept either a PixelTable or the reference's DataFrame (drop-in callers pass the latter)."""
    if isinstance(kmer_mapping, PixelTable):
        return kmer_mapping
    key = id(kmer_mapping)
    hit = _DF_CACHE.get(key)
    if hit is not None and hit[0] is kmer_mapping:
        return hit[1]
    lut = lut_from_dataframe(kmer_mapping)
    k = len(str(kmer_mapping.index[0]))
    t = PixelTable(lut, k, mapping_code or "custom")
    _DF_CACHE[key] = (kmer_mapping, t)
    return t


_DF_CACHE = {}


# ------------------------------------------------------------------------------------------------- remap (convert)
@lru_cache(maxsize=None)
def remap_plan(k, in_mapping, out_mapping):
    """The inner join of ``convert.remap`` (varKoder/commands/convert.py:52-60) between the two pixel tables of size k,
    reduced to what the GPU needs: per output pixel (flattened, final orientation) two source pixels and how many rows
    of the join carry each.  Returns ``(src0, src1, mult[n_out, 2], out_shape)``.

    Multiplicities: the varKode table lists K and rc K on one pixel (a palindrome twice); get_cgr lists every K twice,
    at its own pixel and at the pixel of rc K (core/utils.py:199-208).  So varKode -> cgr puts old[vk(K)] on pixel
    cgr(K) once through K and once through rc K (the same source pixel: 1 + 1, a palindrome 4), and cgr -> varKode
    puts old[cgr(S)] and old[cgr(rc S)] on pixel vk(S) twice each (a palindrome 4 times its one pixel).  Without
    sum_rc the reference keeps the last row written; for images whose pixels agree under reverse complement (anything
    ``varKoder image`` wrote) every row carries the same value and src0 is used.
    """
    if in_mapping not in MAPPING_CHOICES or out_mapping not in MAPPING_CHOICES:
        raise Exception("Input and output mapping must be one of: " + str(MAPPING_CHOICES))      # convert.py:49
    if in_mapping == out_mapping:
        raise ValueError("input and output mapping are the same: nothing to remap")
    tin, tout = get_kmer_mapping(k, in_mapping), get_kmer_mapping(k, out_mapping)
    n = 4 ** k
    rc = revcomp_index(np.arange(n, dtype=np.int64), k)
    lin = tin.lut.reshape(-1).astype(np.int64)
    used_in = np.flatnonzero(lin >= 0)
    pin = np.full(n, -1, dtype=np.int64)
    pin[lin[used_in]] = used_in
    if in_mapping == "varKode":
        pin[rc[lin[used_in]]] = used_in
    lout = tout.lut.reshape(-1).astype(np.int64)
    used = np.flatnonzero(lout >= 0)
    kk = lout[used]
    pal = rc[kk] == kk
    src0 = np.full(lout.size, -1, dtype=np.int32)
    src1 = np.full(lout.size, -1, dtype=np.int32)
    mult = np.zeros((lout.size, 2), dtype=np.uint8)
    src0[used] = pin[kk]
    src1[used] = pin[rc[kk]]
    each = 1 if out_mapping == "cgr" else 2
    mult[used, 0] = np.where(pal, 4, each)
    mult[used, 1] = np.where(pal, 0, each)
    for a in (src0, src1, mult):
        a.setflags(write=False)
    return src0, src1, mult, (tout.side, tout.side)
