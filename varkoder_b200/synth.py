"""Synthetic cleaned-FASTQ generator (numpy) -- byte-identical to csrc/vk_synth.cuh for fixed read length.

Shape from SURVEY.md section 8d: record = "@S%010d\\n" + L bases + "\\n+\\n" + L qualities + "\\n"; bases A,T 0.30 /
C,G 0.20, N with p ~ 0.001, 0.5 % of reads with a poly-G tail of 20..60, qualities uniform '#'..'I'.
``variable()`` adds the Bembidion-shaped case (read lengths 60..280, 1 % shorter than k including empty).
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _hash(seed, r, slot):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + r.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) \
            + slot.astype(np.uint64) * np.uint64(0xD1B54A32D192ED03)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def record_size(read_len):
    return 2 * read_len + 17


def fixed_total_bytes(n_bases, read_len):
    n_reads = (n_bases + read_len - 1) // read_len
    last = n_bases - (n_reads - 1) * read_len
    return (n_reads - 1) * record_size(read_len) + 2 * last + 17


def _bases_quals(seed, r, lens):
    """r: [n] read indices, lens: [n] lengths -> (bases[n, Lmax], quals[n, Lmax]) uint8 (padding undefined)."""
    n = r.size
    lmax = int(lens.max()) if n else 0
    i = np.arange(lmax, dtype=np.uint64)[None, :]
    rr = r.astype(np.uint64)[:, None]
    h = _hash(seed, rr, i + np.uint64(1))
    u = (h & np.uint64(0xFFFF)).astype(np.int64)
    b = np.where(u < 19661, ord("A"), np.where(u < 32768, ord("C"), np.where(u < 45875, ord("G"), ord("T"))))
    isn = ((h >> np.uint64(16)) & np.uint64(0xFFFFF)) < np.uint64(1049)
    b = np.where(isn, ord("N"), b)
    hr = _hash(seed, r.astype(np.uint64), np.zeros(n, dtype=np.uint64))
    tail = np.where((hr & np.uint64(0xFFFF)) < np.uint64(328),
                    20 + ((hr >> np.uint64(16)) % np.uint64(41)).astype(np.int64), 0)
    polyg = (np.arange(lmax, dtype=np.int64)[None, :] + tail[:, None]) >= lens.astype(np.int64)[:, None]
    b = np.where(polyg, ord("G"), b).astype(np.uint8)
    q = (35 + ((h >> np.uint64(40)) % np.uint64(39)).astype(np.int64)).astype(np.uint8)
    return b, q


def _assemble(seed, r, lens):
    n = r.size
    rec = 2 * lens.astype(np.int64) + 17
    off = np.concatenate([[0], np.cumsum(rec)])
    out = np.empty(int(off[-1]), dtype=np.uint8)
    b, q = _bases_quals(seed, r, lens)
    hdr = np.array([list(b"@S%010d\n" % (int(x) % 10**10)) for x in r], dtype=np.uint8).reshape(n, 13)
    for j in range(n):
        o, L = int(off[j]), int(lens[j])
        out[o:o + 13] = hdr[j]
        out[o + 13:o + 13 + L] = b[j, :L]
        out[o + 13 + L:o + 16 + L] = (10, 43, 10)
        out[o + 16 + L:o + 16 + 2 * L] = q[j, :L]
        out[o + 16 + 2 * L] = 10
    return out


def fixed(n_bases, read_len=150, seed=0, first_read=0):
    """same bytes as vk_synth_fastq(n_bases, read_len, seed, first_read)."""
    n_reads = (n_bases + read_len - 1) // read_len
    lens = np.full(n_reads, read_len, dtype=np.int64)
    lens[-1] = n_bases - (n_reads - 1) * read_len
    r = first_read + np.arange(n_reads, dtype=np.uint64)
    if n_reads > 20000:      # vectorised assembly for equal-length records
        return _fixed_fast(seed, r, lens, read_len)
    return _assemble(seed, r, lens)


def _fixed_fast(seed, r, lens, L):
    n = r.size
    rs = record_size(L)
    body = np.empty((n - 1, rs), dtype=np.uint8)
    chunk = 50000
    for s in range(0, n - 1, chunk):
        e = min(n - 1, s + chunk)
        b, q = _bases_quals(seed, r[s:e], lens[s:e])
        body[s:e, 0] = ord("@")
        body[s:e, 1] = ord("S")
        v = (r[s:e] % np.uint64(10**10)).astype(np.int64)
        for d in range(10):
            body[s:e, 11 - d] = ord("0") + (v % 10)
            v //= 10
        body[s:e, 12] = 10
        body[s:e, 13:13 + L] = b
        body[s:e, 13 + L] = 10
        body[s:e, 14 + L] = ord("+")
        body[s:e, 15 + L] = 10
        body[s:e, 16 + L:16 + 2 * L] = q
        body[s:e, 16 + 2 * L] = 10
    tail = _assemble(seed, r[-1:], lens[-1:])
    return np.concatenate([body.reshape(-1), tail])


def variable(n_reads, seed=0, min_len=60, max_len=280, short_frac=0.01, k=7, first_read=0, final_newline=True):
    """Bembidion-shaped: lengths uniform min..max, ``short_frac`` of reads shorter than k (incl. empty)."""
    r = first_read + np.arange(n_reads, dtype=np.uint64)
    hl = _hash(seed ^ 0x5EED, r, np.zeros(n_reads, dtype=np.uint64))
    lens = (min_len + (hl % np.uint64(max_len - min_len + 1)).astype(np.int64))
    short = ((hl >> np.uint64(32)) % np.uint64(10000)).astype(np.int64) < int(short_frac * 10000)
    lens = np.where(short, ((hl >> np.uint64(48)) % np.uint64(k)).astype(np.int64), lens)
    out = _assemble(seed, r, lens)
    return out if final_newline else out[:-1]
