"""numpy restatement of the Python half of the reference hot path -- TEST INFRASTRUCTURE ONLY.

PINNED: every function here is checked against the imported, unmodified reference
(``oracle/make_golden.py`` -> ``tests/golden``; ``tests/test_oracle.py`` when
``/root/reference`` is mounted) and against the properties of the reference's shipped ``docs/*.png``.

Follows (reference file:line):
  ladder()             split_fastq           commands/image.py:669-709
  cgr_xy()/cgr_lut()   get_cgr               core/utils.py:174-217
  lut_from_table()     join/groupby/scatter  commands/image.py:900-913
  image_exact()        +1, quantile, digitize commands/image.py:911-919 (exact integer form)
  image_float()        the same lines, literally, in float64
Index convention: lexicographic k-mer index, first base most significant, A=0 C=1 G=2 T=3.
"""
import math

import numpy as np

_LEX = {"A": 0, "C": 1, "G": 2, "T": 3}


# ----------------------------------------------------------------------------------------------- R1
def ladder(nsites, min_bp, max_bp, is_query=False):
    """image.py:669-695, literally (same float log10 / float division as the reference)."""
    nsites = int(nsites)
    if max_bp is None:
        sites_per_file = [int(nsites)]
    elif is_query or int(nsites) > min_bp:
        sites_per_file = [min(int(nsites), int(max_bp))]
    else:
        raise Exception("Input file has less than minimum data.")
    if not is_query:
        while sites_per_file[-1] > min_bp:
            oneless = sites_per_file[-1] - 1
            nzeros = int(math.log10(oneless))
            first_digit = int(oneless / (10 ** nzeros))
            if first_digit in [1, 2, 5]:
                sites_per_file.append(first_digit * (10 ** nzeros))
            else:
                multiplier = max([x for x in [1, 2, 5] if x < first_digit])
                sites_per_file.append(multiplier * (10 ** nzeros))
        if sites_per_file[-1] < min_bp:
            del sites_per_file[-1]
    return sites_per_file


def level_tag(bp):
    """image.py:704-705: '%08dK' of the target (not realised) bases."""
    return str(int(bp / 1000)).rjust(8, "0") + "K"


def image_name(sample, bp, mapping_code, k):
    """image.py:699-709 + 752-758 + 843-849."""
    return f"{sample}@{level_tag(bp)}+{mapping_code}+k{k}.png"


# ------------------------------------------------------------------------------------------ R8 / R9
def lex_index(kmer):
    x = 0
    for ch in kmer:
        x = x * 4 + _LEX[ch]
    return x


def revcomp_index(x, k):
    r = 0
    for _ in range(k):
        r = (r << 2) | (3 - (x & 3))
        x >>= 2
    return r


def cgr_xy(k):
    """closed form of get_cgr (utils.py:185-215): corners A(0,0) C(0,1) G(1,1) T(1,0); the i-th base
    contributes bit i (first base = least significant).  Returns x[4^k], y[4^k] by lexicographic index."""
    n = 4 ** k
    idx = np.arange(n, dtype=np.int64)
    x = np.zeros(n, dtype=np.int64)
    y = np.zeros(n, dtype=np.int64)
    xbit = np.array([0, 0, 1, 1])      # A C G T
    ybit = np.array([0, 1, 1, 0])
    for i in range(k):                 # i-th base of the k-mer (0 = first) sits at lex digit k-1-i
        d = (idx >> (2 * (k - 1 - i))) & 3
        x |= xbit[d] << i
        y |= ybit[d] << i
    return x, y


def lut_from_xy(kmer_idx, x, y):
    """Pixel table rows (k-mer lex index, x, y) -> pix2kmer int32[H, W] in FINAL image orientation.

    image.py:906-913: H = max(y)+1, W = max(x)+1, A[x, y] = value, A = flip(A.T, 0) => value lands at
    row H-1-y, column x.  Every row of a pixel must belong to one canonical class (true for the shipped
    tables: each pixel lists K and rc(K)); the LUT stores one member, -1 marks unused pixels.
    """
    kmer_idx = np.asarray(kmer_idx, dtype=np.int64)
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    H = int(y.max()) + 1
    W = int(x.max()) + 1
    lut = np.full((H, W), -1, dtype=np.int32)
    lut[H - 1 - y, x] = kmer_idx
    return lut


def cgr_lut(k):
    x, y = cgr_xy(k)
    return lut_from_xy(np.arange(4 ** k), x, y)


def lut_from_table(df):
    """reference DataFrame (index = k-mer string, columns x, y) -> LUT, checking the one-class-per-pixel
    property the join/groupby-mean of image.py:900 relies on."""
    kmers = [str(s) for s in df.index]
    k = len(kmers[0])
    idx = np.array([lex_index(s) for s in kmers], dtype=np.int64)
    x = df["x"].to_numpy().astype(np.int64)
    y = df["y"].to_numpy().astype(np.int64)
    canon = np.minimum(idx, np.array([revcomp_index(int(i), k) for i in idx], dtype=np.int64))
    H = int(y.max()) + 1
    pix = (H - 1 - y) * (int(x.max()) + 1) + x
    order = np.argsort(pix, kind="stable")
    ps, cs = pix[order], canon[order]
    same = ps[1:] == ps[:-1]
    if np.any(cs[1:][same] != cs[:-1][same]):
        raise ValueError("pixel table maps two canonical classes onto one pixel")
    return lut_from_xy(idx, x, y)


# ------------------------------------------------------------------------------------------ R4 - R6
def pixel_values(canon_full, lut):
    """image.py:900-911: pixel = canonical abundance + 1; unused pixels stay 0; unseen k-mers give 1."""
    c = np.asarray(canon_full, dtype=np.uint64)
    v = np.zeros(lut.shape, dtype=np.uint64)
    used = lut >= 0
    v[used] = c[lut[used]] + np.uint64(1)
    return v


def rank_exact(values):
    """image.py:916-919 in exact integer arithmetic.

    bins = quantile(A, i/256) with linear interpolation: virtual index (n-1)*i/256 = p + g/256, value
    s[p] + (s[p+1]-s[p])*g/256; out(v) = #{i : bins[i] <= v} - 1.  Scaled by 256 everything is an integer.
    Python ints are used so there is no overflow at any count.
    """
    flat = [int(t) for t in np.asarray(values).ravel()]
    n = len(flat)
    s = sorted(flat)
    bins256 = []
    for i in range(256):
        p, g = divmod((n - 1) * i, 256)
        q = min(p + 1, n - 1)
        bins256.append(256 * s[p] + (s[q] - s[p]) * g)
    b = np.array(bins256, dtype=object)
    # bins256 is non-decreasing: count of bins <= 256*v by bisection
    import bisect
    out = np.array([bisect.bisect_right(bins256, 256 * t) - 1 for t in flat], dtype=np.int64)
    assert out.min() >= 0 and out.max() <= 255 and len(b) == 256
    return out.astype(np.uint8).reshape(np.asarray(values).shape)


def image_exact(canon_full, lut):
    return rank_exact(pixel_values(canon_full, lut))


def image_float(canon_full, lut):
    """the literal float64 route of image.py:910-919 (for cross-checking rank_exact)."""
    a = pixel_values(canon_full, lut).astype(np.float64)
    bins = np.quantile(a, np.arange(0, 1, 1 / 256))
    return np.uint8(np.digitize(a, bins, right=False) - 1)


# ---------------------------------------------------------------------------------------------- remap (convert)
def remap_plan(lut_in, lut_out, k, in_is_cgr, out_is_cgr):
    """Restatement of the pandas merge of ``convert.remap`` (varKoder/commands/convert.py:52-72) for one direction.

    Returns ``(src0, src1, mult)`` per OUTPUT pixel (flattened, final image orientation): the output pixel takes
    ``old[src0]`` in plain mode and ``mult * (old[src0] + old[src1]) / 2 ...`` -- precisely: in sum_rc mode it receives
    ``m0 * old[src0] + m1 * old[src1]`` with ``(m0, m1) = mult``; unused output pixels have src0 = -1.

    Row multiplicities of the merge: the varKode table lists K and rc K on one pixel (a palindrome twice), the cgr table
    lists every K twice, once at its own pixel and once at the pixel of rc K (utils.py:199-208).
    """
    n = 4 ** k
    idx = np.arange(n, dtype=np.int64)
    rc = np.array([revcomp_index(int(i), k) for i in idx], dtype=np.int64)
    lin, lout = np.asarray(lut_in).reshape(-1), np.asarray(lut_out).reshape(-1)

    def kmer_to_pixel(lut_flat, is_cgr):
        pix = np.full(n, -1, dtype=np.int64)
        used = np.flatnonzero(lut_flat >= 0)
        pix[lut_flat[used]] = used
        if not is_cgr:                                  # varKode: rc K is drawn on the same pixel
            pix[rc[lut_flat[used]]] = used
        return pix

    pin = kmer_to_pixel(lin, in_is_cgr)
    n_out = lout.size
    src0 = np.full(n_out, -1, dtype=np.int32)
    src1 = np.full(n_out, -1, dtype=np.int32)
    mult = np.zeros((n_out, 2), dtype=np.uint8)
    used = np.flatnonzero(lout >= 0)
    K = lout[used].astype(np.int64)                     # a k-mer shown at the output pixel (its own K for cgr)
    pal = rc[K] == K
    if out_is_cgr and not in_is_cgr:                    # varKode -> cgr: pixel xy(K) <- old[vk(K)] (+ old[vk(rc K)], same pixel)
        src0[used] = pin[K]
        src1[used] = pin[rc[K]]
        mult[used, 0] = np.where(pal, 4, 1)
        mult[used, 1] = np.where(pal, 0, 1)
    elif in_is_cgr and not out_is_cgr:                  # cgr -> varKode: pixel vk(S) <- old[cgr(S)], old[cgr(rc S)]
        src0[used] = pin[K]
        src1[used] = pin[rc[K]]
        mult[used, 0] = np.where(pal, 4, 2)
        mult[used, 1] = np.where(pal, 0, 2)
    else:
        raise ValueError("remap is between varKode and cgr")
    return src0, src1, mult


def remap_exact(img, lut_in, lut_out, k, in_is_cgr, out_is_cgr, sum_rc=False):
    """pixels of ``convert.remap(img, k, in, out, sum_rc)`` for an image that is consistent under reverse complement"""
    old = np.asarray(img, dtype=np.uint8).reshape(-1)
    src0, src1, mult = remap_plan(lut_in, lut_out, k, in_is_cgr, out_is_cgr)
    used = src0 >= 0
    new = np.zeros(src0.size, dtype=np.uint8)
    if not sum_rc:
        new[used] = old[src0[used]]
        return new.reshape(np.asarray(lut_out).shape)
    acc = (mult[used, 0].astype(np.uint32) * old[src0[used]] + mult[used, 1].astype(np.uint32) * old[np.maximum(src1[used], 0)])
    new[used] = (acc & 0xFF).astype(np.uint8)           # np.add.at on a uint8 array wraps
    mn, mx = int(new.min()), int(new.max())
    with np.errstate(all="ignore"):
        out = np.uint8((new - np.uint8(mn)) / np.uint8(mx) * 255) if mx else np.zeros_like(new)
    return out.reshape(np.asarray(lut_out).shape)


# ---- quality flag ---------------------------------------------------------------------------
def base_content(buf, starts, lens, pos_begin=5, pos_end=40):
    """numpy restatement of what fastp accumulates for its content curves (per cycle: bases by ``byte & 7`` and the
    number of reads that reach the cycle), over the framed reads: -> uint64 [pos_end - pos_begin, 5] (A, T, C, G, reach).
    The reference consumes the curves in get_basefrequency_sd (varKoder/commands/image.py:49-88)."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    starts = np.asarray(starts, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int64)
    out = np.zeros((pos_end - pos_begin, 5), dtype=np.uint64)
    for i, p in enumerate(range(pos_begin, pos_end)):
        sel = lens > p
        b = a[starts[sel] + p] & 7
        out[i] = [np.count_nonzero(b == 1), np.count_nonzero(b == 4), np.count_nonzero(b == 3),
                  np.count_nonzero(b == 7), b.size]
    return out
