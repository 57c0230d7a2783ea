"""ctypes front-end of ``oracle/dsk_oracle.c`` plus a brute-force pure-Python cross-check.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All histograms use the lexicographic index
(first base most significant, A=0 C=1 G=2 T=3).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BREAKLENGTH = 500          # reformat.sh breaklength=500, image.py:586
MASK64 = (1 << 64) - 1


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libvkoracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        c = ctypes
        L.vko_prio.restype = c.c_uint64
        L.vko_prio.argtypes = [c.c_uint64, c.c_uint64]
        L.vko_threshold.restype = c.c_uint64
        L.vko_threshold.argtypes = [c.c_uint64, c.c_uint64]
        L.vko_parse_fastq.restype = c.c_int64
        L.vko_parse_fastq.argtypes = [c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p, c.c_int64,
                                      c.POINTER(c.c_int64), c.POINTER(c.c_int64), c.POINTER(c.c_int64)]
        L.vko_count_forward.restype = None
        L.vko_count_forward.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int,
                                        c.c_void_p, c.c_void_p, c.c_int]
        L.vko_fold_canonical.restype = None
        L.vko_fold_canonical.argtypes = [c.c_void_p, c.c_int, c.c_void_p]
        L.vko_revcomp.restype = c.c_uint32
        L.vko_revcomp.argtypes = [c.c_uint32, c.c_int]
        L.vko_dsk2ascii.restype = c.c_int64
        L.vko_dsk2ascii.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int64]
        L.vko_count_levels.restype = c.c_int64
        L.vko_count_levels.argtypes = [c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_uint64, c.c_uint64,
                                       c.c_void_p, c.c_void_p, c.c_int, c.c_void_p, c.c_int]
        _LIB = L
    return _LIB


def _as_u8(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def prio(seed, read_index):
    return int(lib().vko_prio(seed & MASK64, read_index & MASK64))


def prio_py(seed, read_index):
    """same hash in plain Python ints (cross-check of the C and CUDA versions)."""
    z = (seed + (read_index + 1) * 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def threshold(target, nsites):
    """floor(target * 2^64 / nsites); only meaningful for target < nsites."""
    return (int(target) << 64) // int(nsites)


def parse_fastq(buf):
    """-> dict(starts, lens, nsites_ref, nsites_true, n_lines) following split_fastq's framing."""
    a = _as_u8(buf)
    cap = a.size // 2 + 16
    starts = np.zeros(cap, dtype=np.int64)
    lens = np.zeros(cap, dtype=np.int64)
    r = ctypes.c_int64()
    t = ctypes.c_int64()
    nl = ctypes.c_int64()
    n = lib().vko_parse_fastq(a.ctypes.data, a.size, starts.ctypes.data, lens.ctypes.data, cap,
                              ctypes.byref(r), ctypes.byref(t), ctypes.byref(nl))
    return dict(starts=starts[:n].copy(), lens=lens[:n].copy(), nsites_ref=r.value, nsites_true=t.value,
                n_lines=nl.value, n_reads=int(n))


PRIO_BUCKETS = 1 << 16          # include/varkoder_b200.h VK_PRIO_BUCKETS
PRIO_SHIFT = 48


def prio_array(seed, first, count):
    """prio64(seed, r) for r = first .. first + count - 1, vectorised (uint64 arithmetic wraps)."""
    with np.errstate(over="ignore"):
        r = np.arange(first + 1, first + 1 + count, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + r * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def prio_hist(lens, seed, read_index_base=0):
    """bases of the reads (ALL records of the buffer, whatever their length) per priority bucket, prio >> 48
    (varkoder_b200/csrc/vk_sample.cuh K1h).  Shards of one sample: add their histograms."""
    lens = np.asarray(lens, dtype=np.int64)
    b = (prio_array(seed, read_index_base, lens.size) >> np.uint64(PRIO_SHIFT)).astype(np.int64)
    # float64 weights are exact here: a bucket holds far fewer than 2^53 bases
    return np.bincount(b, weights=lens.astype(np.float64), minlength=PRIO_BUCKETS).astype(np.int64)


def calibrated_thresholds(hist, levels, nsites):
    """the thresholds fitted to the base targets (vk_sample.cuh K1t), in exact integers:
    level with target T -> first bucket b whose cumulative base count reaches T, C = bases before it,
    thr = (b << 48) + floor((T - C) * 2^48 / hist[b]).  -> (thr[], take_all[])."""
    hist = np.asarray(hist, dtype=np.int64)
    cum = np.cumsum(hist)
    total = int(cum[-1])
    thr, take_all = [], []
    for T in levels:
        T = int(T)
        if T >= nsites or T >= total:
            thr.append(0)
            take_all.append(1)
            continue
        b = int(np.searchsorted(cum, T, side="left"))          # first b with cum[b] >= T
        C = int(cum[b - 1]) if b else 0
        W = int(hist[b])
        rem = T - C
        if rem >= W:
            t = (b + 1) << PRIO_SHIFT
        else:
            t = (b << PRIO_SHIFT) + (((rem << 64) // W) >> (64 - PRIO_SHIFT))
        if t > MASK64:
            thr.append(0)
            take_all.append(1)
        else:
            thr.append(t)
            take_all.append(0)
    return thr, take_all


def level_thresholds(levels, nsites, seed=0, lens=None, hist=None, read_index_base=0):
    """(thr[], take_all[]) of the ladder levels.  With ``lens`` (all records of the sample, in order) or ``hist`` (their
    prio_hist, summed over shards): the calibrated rule, the product's default (VK_SAMPLING_CALIBRATED).  With neither:
    thr = floor(bp * 2^64 / nsites) (VK_SAMPLING_EXPECTED)."""
    if hist is None and lens is not None:
        hist = prio_hist(lens, seed, read_index_base)
    if hist is not None:
        return calibrated_thresholds(hist, levels, nsites)
    return ([0 if bp >= nsites else threshold(bp, nsites) for bp in levels], [1 if bp >= nsites else 0 for bp in levels])


def select_reads(n_reads, seed, target, nsites, read_index_base=0, lens=None, hist=None):
    """the project's sub-sampling rule: read r is kept iff prio(seed, base + r) < thr(target); target >= nsites keeps
    everything.  ``lens`` (all records of the SAMPLE, so that ``n_reads`` == len(lens) for an unsharded buffer) or ``hist``
    select the calibrated threshold; without them thr = floor(target * 2^64 / nsites)."""
    (thr,), (all_,) = level_thresholds([target], nsites, seed, lens, hist, 0 if lens is None else 0)
    if all_:
        return np.ones(n_reads, dtype=np.uint8)
    return (prio_array(seed, read_index_base, n_reads) < np.uint64(thr)).astype(np.uint8)


def count_forward(buf, starts, lens, k, select=None, breaklen=BREAKLENGTH, threads=1):
    a = _as_u8(buf)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    fwd = np.zeros(4 ** k, dtype=np.uint64)
    sel = None if select is None else np.ascontiguousarray(select, dtype=np.uint8)
    lib().vko_count_forward(a.ctypes.data, starts.ctypes.data, lens.ctypes.data, starts.size, k, breaklen,
                            None if sel is None else sel.ctypes.data, fwd.ctypes.data, threads)
    return fwd


def fold_canonical(fwd, k):
    fwd = np.ascontiguousarray(fwd, dtype=np.uint64)
    out = np.zeros_like(fwd)
    lib().vko_fold_canonical(fwd.ctypes.data, k, out.ctypes.data)
    return out


def canonical_counts(buf, k, select=None, breaklen=BREAKLENGTH, threads=1):
    """FASTQ bytes -> canon_full uint64[4^k] (dsk restated)."""
    p = parse_fastq(buf)
    return fold_canonical(count_forward(buf, p["starts"], p["lens"], k, select, breaklen, threads), k)


def dsk2ascii_text(canon_full, k):
    c = np.ascontiguousarray(canon_full, dtype=np.uint64)
    cap = (k + 24) * (c.size // 2 + c.size // 64 + 64)
    out = ctypes.create_string_buffer(cap)
    n = lib().vko_dsk2ascii(c.ctypes.data, k, out, cap)
    assert n <= cap
    return out.raw[:n].decode("ascii")


def count_levels(buf, k, seed, thresholds, take_all, read_index_base=0, breaklen=BREAKLENGTH, threads=0):
    """whole native half for all ladder levels (CPU baseline leg): -> (n_reads, canon_full[L, 4^k])."""
    a = _as_u8(buf)
    thr = np.ascontiguousarray(thresholds, dtype=np.uint64)
    ta = np.ascontiguousarray(take_all, dtype=np.uint8)
    out = np.zeros((thr.size, 4 ** k), dtype=np.uint64)
    n = lib().vko_count_levels(a.ctypes.data, a.size, k, breaklen, seed & MASK64, read_index_base,
                               thr.ctypes.data, ta.ctypes.data, thr.size, out.ctypes.data, threads)
    return int(n), out


# ---------------------------------------------------------------------------------------------------
# brute force, for tiny inputs: the oracle's oracle
# ---------------------------------------------------------------------------------------------------
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
_LEX = {"A": 0, "C": 1, "G": 2, "T": 3}


def lex_index(kmer):
    x = 0
    for ch in kmer:
        x = x * 4 + _LEX[ch]
    return x


def brute_canonical_counts(fastq_bytes, k, select=None, breaklen=BREAKLENGTH):
    """pure-Python: split lines, cut reads at breaklen, slide a window, drop windows with non-ACGT."""
    lines = fastq_bytes.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    out = np.zeros(4 ** k, dtype=np.uint64)
    r = 0
    for i, l in enumerate(lines):
        if i % 4 != 1:
            continue
        keep = True if select is None else bool(select[r])
        r += 1
        if not keep:
            continue
        s = l.decode("latin-1").upper()
        pieces = [s[j:j + breaklen] for j in range(0, len(s), breaklen)] if breaklen else [s]
        for p in pieces:
            for j in range(len(p) - k + 1):
                w = p[j:j + k]
                if all(ch in _LEX for ch in w):
                    rc = "".join(_COMP[ch] for ch in reversed(w))
                    out[lex_index(w)] += 1
                    if rc != w:
                        out[lex_index(rc)] += 1
    return out
