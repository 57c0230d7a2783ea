"""A SECOND, independently written restatement of the native half of the path (framing + sub-sample membership + dsk's
counting rules), in vectorised numpy -- TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Why it exists.  ``dsk_oracle.c`` is the checker of the CUDA path, but GATB dsk and BBTools cannot be run or built here
(conda binaries, ``conda_environments/linux.yml:10-11``), so the C restatement is "parity unpinned" at that boundary.  What
CAN be done is to restate the same published rules a second time, with a different algorithm and in a different language,
and require the two to agree on the BASELINE-size inputs, not only on the tiny ones the brute-force Python counter reaches:

* ``dsk_oracle.c`` walks every read once with a rolling 2-bit window and a run counter of valid bases;
* this file never rolls anything: it enumerates EVERY window start of a chunk of reads as a flat index array, gathers the k
  bytes of each window, drops a window if any gathered byte is not a base or if the window crosses a cut point, and
  histograms the survivors with ``np.bincount``.

Rules restated (reference call sites in ``varKoder/commands/image.py``):
  framing      lines split on ``\\n``; line index % 4 == 1 is a sequence line; an unterminated last line counts (:662-667)
  D1           k-mers never span records, nor the <= ``breaklength``-base pieces ``reformat.sh breaklength=500`` cuts (:586)
  D2 / D3      A C G T (either case) are bases; a window holding anything else is dropped (``iupacToN``, :586; dsk skips N)
  D4 / D5      abundance of a canonical class = forward count of K + forward count of rc(K) (once for a palindrome);
               ``canon_full[K] = canon_full[rc K]`` = that abundance, lexicographic index, A0 C1 G2 T3 (dsk2ascii, :875-891)
  membership   read r is in level l iff ``take_all[l]`` or ``prio64(seed, read_index_base + r) < thr[l]`` (this project's
               rule, DESIGN.md "Sub-sampling"; the thresholds come from ``dsk.level_thresholds``)
"""
import numpy as np

from . import dsk

_LEX = np.full(256, -1, dtype=np.int8)
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _LEX[ord(_ch)] = _v
    _LEX[ord(_ch.lower())] = _v


def frame(buf):
    """-> (starts, lens) of the sequence lines, as Python's binary line iterator frames them."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    n = int(a.size)
    nl = np.flatnonzero(a == 10).astype(np.int64)
    line_start = np.concatenate((np.zeros(1, dtype=np.int64), nl + 1))
    line_end = np.concatenate((nl, np.array([n], dtype=np.int64)))
    n_lines = nl.size + (1 if n > 0 and a[n - 1] != 10 else 0)
    idx = np.arange(1, n_lines, 4, dtype=np.int64)
    return line_start[idx], line_end[idx] - line_start[idx]


def revcomp_lex(k):
    """lexicographic index of the reverse complement of every k-mer."""
    x = np.arange(4 ** k, dtype=np.int64)
    r = np.zeros_like(x)
    for _ in range(k):
        r = r * 4 + (3 - (x & 3))
        x >>= 2
    return r


def count_levels(buf, k, seed, thresholds, take_all, breaklen=dsk.BREAKLENGTH, read_index_base=0, chunk_reads=150_000):
    """-> canon_full uint64 [n_levels, 4^k] of the nested levels (level 0 first)."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, dtype=np.uint8)
    starts, lens = frame(a)
    n_levels = len(thresholds)
    nk = 4 ** k
    # number of (nested) levels every read is in; segment s = n_in - 1
    prio = dsk.prio_array(seed, read_index_base, starts.size)
    n_in = np.zeros(starts.size, dtype=np.int64)
    for l in range(n_levels):
        inside = np.ones(starts.size, dtype=bool) if take_all[l] else prio < np.uint64(thresholds[l])
        assert (n_in[inside] == l).all(), "levels are nested"
        n_in += inside
    seg_hist = np.zeros((n_levels, nk), dtype=np.int64)
    for c0 in range(0, starts.size, chunk_reads):
        st, ln, ni = starts[c0:c0 + chunk_reads], lens[c0:c0 + chunk_reads], n_in[c0:c0 + chunk_reads]
        keep = (ni > 0) & (ln >= k)
        st, ln, seg = st[keep], ln[keep], ni[keep] - 1
        nwin = ln - k + 1                                            # window starts per read
        total = int(nwin.sum())
        if total == 0:
            continue
        owner = np.repeat(np.arange(st.size, dtype=np.int64), nwin)  # read of every window
        first = np.cumsum(nwin) - nwin
        off = np.arange(total, dtype=np.int64) - first[owner]        # window start inside its read
        ok = np.ones(total, dtype=bool)
        if breaklen:
            ok &= (off // breaklen) == ((off + k - 1) // breaklen)    # D1: inside one piece
        pos = st[owner] + off
        val = np.zeros(total, dtype=np.int64)
        for j in range(k):
            c = _LEX[a[pos + j]]
            ok &= c >= 0                                              # D3
            val = val * 4 + c
        flat = seg[owner][ok] * nk + val[ok]
        seg_hist += np.bincount(flat, minlength=n_levels * nk).reshape(n_levels, nk)
    fwd = np.cumsum(seg_hist[::-1], axis=0)[::-1]                     # level l = segments l, l + 1, ...
    rc = revcomp_lex(k)
    pal = rc == np.arange(nk)
    canon = fwd + fwd[:, rc]
    canon[:, pal] = fwd[:, pal]
    return canon.astype(np.uint64)
