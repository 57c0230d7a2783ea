"""Host feed of the image hot path: cleaned ``.fq.gz`` files -> pinned host buffers -> GPU, several samples in flight.

The reference reads ``<int>/clean_reads/<sample>.fq.gz`` (single-member gzip written by pigz,
varKoder/commands/image.py:529-540) with Python's gzip module, once in ``split_fastq`` (:662-667) and then L more
times through ``reformat.sh``.  Here every sample is inflated ONCE, by a worker thread (the decoder runs without the
GIL, so N threads inflate N samples in parallel -- a single gzip member cannot be split), straight into page-locked
memory, and handed to the GPU in submission order while the next samples are still inflating.  PNG encoding of
finished samples runs in the same pool, off the critical path.

The decoder is ``libvk_feed.so`` (``csrc/vk_inflate.c``: table-driven DEFLATE with a 64-bit bit buffer, output in place,
CRC-32 by carry-less multiplication), 1.5-2 x zlib's speed on FASTQ text; anything it rejects, and every file when
the library is absent, goes through zlib's streaming decoder as before.  When there are fewer samples than threads,
the spare threads split single members: pigz ends every 128 KiB chunk with an empty stored block, and
``gunzip_parallel`` decodes the pieces between such sync points concurrently (0.18 s instead of 1.9 s for a 200 Mbp
sample on 16 threads).
"""
import ctypes as C
import os
import threading
import zlib
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_CHUNK = 1 << 20


def gzip_isize(path):
    """ISIZE field of the last gzip member: uncompressed size mod 2^32 (a hint for the buffer size)."""
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        if f.tell() < 18:
            return 0
        f.seek(-4, os.SEEK_END)
        return int.from_bytes(f.read(4), "little")


class PinnedBuffer:
    """A growable page-locked byte buffer (torch pinned memory when CUDA is there, plain numpy otherwise)."""

    def __init__(self, nbytes=0, pinned=True):
        self.pinned = pinned
        self._t = None
        self.array = np.empty(0, dtype=np.uint8)
        self.reserve(max(int(nbytes), 1))

    def reserve(self, nbytes):
        if nbytes <= self.array.size:
            return
        new = int(nbytes + nbytes // 8 + 4096)
        old = self.array
        t = None
        if self.pinned:
            try:
                import torch
                if torch.cuda.is_available():
                    t = torch.empty(new, dtype=torch.uint8, pin_memory=True)
            except Exception:
                t = None
        arr = t.numpy() if t is not None else np.empty(new, dtype=np.uint8)
        if old.size:
            arr[:old.size] = old
        self._t, self.array = t, arr

    @property
    def is_pinned(self):
        return self._t is not None


_HERE = os.path.dirname(os.path.abspath(__file__))
_FEED_LIB_PATH = os.path.join(_HERE, "libvk_feed.so")
_feed_lib = None
VKF_OK, VKF_ESPACE = 0, -2


def feed_lib():
    """``libvk_feed.so`` or None (then zlib does the work)."""
    global _feed_lib
    if _feed_lib is None:
        try:
            L = C.CDLL(_FEED_LIB_PATH)
            L.vkf_gunzip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
            L.vkf_gunzip.restype = C.c_int
            L.vkf_crc32.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
            L.vkf_crc32.restype = C.c_uint32
            L.vkf_gzip_header_len.argtypes = [C.c_void_p, C.c_size_t]
            L.vkf_gzip_header_len.restype = C.c_size_t
            L.vkf_next_sync.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
            L.vkf_next_sync.restype = C.c_size_t
            L.vkf_inflate_piece.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
            L.vkf_inflate_piece.restype = C.c_int
            L.vkf_resolve16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
            L.vkf_resolve16.restype = C.c_size_t
            L.vkf_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
            L.vkf_crc32_combine.restype = C.c_uint32
            _feed_lib = L
        except OSError:
            _feed_lib = False
    return _feed_lib or None


def gunzip_into(comp, buf: "PinnedBuffer", size_hint=0, verify_crc=True):
    """Inflate a whole gzip file held in memory (``comp``: bytes / uint8 array) into ``buf`` with ``libvk_feed.so``.
    Returns the number of bytes, or None when the library is absent or does not accept the stream."""
    L = feed_lib()
    if L is None:
        return None
    a = np.frombuffer(comp, dtype=np.uint8) if not isinstance(comp, np.ndarray) else comp
    cap = max(int(size_hint), 1 << 16)
    while cap < a.size:                       # ISIZE is the size mod 2^32: text does not shrink below its gzip
        cap += 1 << 32 if size_hint else cap
    n = C.c_size_t()
    while True:
        buf.reserve(cap + 64)
        rc = L.vkf_gunzip(a.ctypes.data, a.size, buf.array.ctypes.data, buf.array.size, C.byref(n), 1 if verify_crc else 0)
        if rc == VKF_OK:
            return int(n.value)
        if rc != VKF_ESPACE:
            return None
        cap = max(2 * buf.array.size, cap + (1 << 32 if size_hint else 0))


_piece_pool = None
_piece_pool_lock = threading.Lock()
WINDOW = 32768


def piece_pool():
    """threads that decode the pieces of ONE gzip member (leaf tasks only; separate from a feeder's per-sample pool)"""
    global _piece_pool
    with _piece_pool_lock:
        if _piece_pool is None:
            _piece_pool = ThreadPoolExecutor(max_workers=max(2, len(os.sched_getaffinity(0))), thread_name_prefix="vk-piece")
    return _piece_pool


def gunzip_parallel(comp, buf: "PinnedBuffer", size_hint, threads, verify_crc=True, min_piece=1 << 20):
    """One single-member gzip file on several threads (see ``csrc/vk_inflate.c``, "one gzip member on several threads"):
    cut the stream behind the empty stored blocks pigz leaves after every chunk, decode the pieces concurrently (the
    first as bytes in place, the others as 16-bit symbols with placeholders for the 32 KiB in front of them), resolve
    the placeholders front to back, check ISIZE and the CRC-32 of the whole member.  Returns the number of bytes, or
    None when the file is not of that shape or anything does not add up (the caller then decodes it serially)."""
    L = feed_lib()
    if L is None or threads < 2 or size_hint <= 0:
        return None
    a = np.frombuffer(comp, dtype=np.uint8) if not isinstance(comp, np.ndarray) else comp
    n, base = a.size, a.ctypes.data
    hl = L.vkf_gzip_header_len(base, n)
    if hl == 0 or n - hl - 8 < 2 * min_piece or size_hint < n // 2:
        return None
    want = int(min(threads * 3, (n - hl - 8) // min_piece))
    step = (n - hl - 8) // want
    cuts = [hl]
    for k in range(1, want):
        pos = L.vkf_next_sync(base, n - 8, hl + k * step)
        if pos >= n - 8:
            break
        if pos > cuts[-1]:
            cuts.append(int(pos))
    P = len(cuts)
    if P < 2:
        return None
    ratio = size_hint / float(n - hl)
    buf.reserve(size_hint + 64)
    out_base = buf.array.ctypes.data

    def decode(i):
        start = cuts[i]
        last = i + 1 == P
        ln = (n - start) if last else (cuts[i + 1] - start)
        got, used = C.c_size_t(), C.c_size_t()
        if i == 0:
            rc = L.vkf_inflate_piece(base + start, ln, 0, 0, out_base, size_hint, C.byref(got), C.byref(used))
            return rc, None, int(got.value), int(used.value), ln
        cap = int(ln * ratio * 1.25) + (1 << 16)
        while True:
            sym = np.empty(cap, dtype=np.uint16)
            rc = L.vkf_inflate_piece(base + start, ln, 1 if last else 0, 1, sym.ctypes.data, cap, C.byref(got), C.byref(used))
            if rc != VKF_ESPACE or cap > 64 * ln + (1 << 20):
                return rc, sym, int(got.value), int(used.value), ln
            cap *= 2

    pool = piece_pool()
    res = list(pool.map(decode, range(P)))
    if any(r[0] != VKF_OK for r in res):
        return None
    if res[-1][3] != res[-1][4] - 8:                       # the final block must end right in front of the trailer
        return None
    lens = [r[2] for r in res]
    offs = [0]
    for x in lens:
        offs.append(offs[-1] + x)
    total = offs[-1]
    trailer = a[n - 8:].tobytes()
    if (total & 0xFFFFFFFF) != int.from_bytes(trailer[4:], "little") or total > size_hint:
        return None

    def window_of(off):
        if off >= WINDOW:
            return buf.array[off - WINDOW:off], 0
        w = np.zeros(WINDOW, dtype=np.uint8)
        w[WINDOW - off:] = buf.array[:off]
        return w, WINDOW - off

    # front to back: the last 32 KiB of every piece (the window of the next one), then everything else concurrently
    bad = 0
    for i in range(1, P):
        t0 = max(0, lens[i] - WINDOW)
        w, valid_from = window_of(offs[i])
        bad += L.vkf_resolve16(res[i][1].ctypes.data + 2 * t0, lens[i] - t0, w.ctypes.data, valid_from, out_base + offs[i] + t0)

    def head(i):
        t0 = max(0, lens[i] - WINDOW)
        if t0 == 0:
            return 0
        w, valid_from = window_of(offs[i])
        return L.vkf_resolve16(res[i][1].ctypes.data, t0, w.ctypes.data, valid_from, out_base + offs[i])

    bad += sum(pool.map(head, range(1, P)))
    if bad:
        return None
    if verify_crc:
        crcs = list(pool.map(lambda i: L.vkf_crc32(0, out_base + offs[i], lens[i]), range(P)))
        crc = crcs[0]
        for i in range(1, P):
            crc = L.vkf_crc32_combine(crc, crcs[i], lens[i])
        if crc != int.from_bytes(trailer[:4], "little"):
            return None
    return total


def inflate_into(path, buf: PinnedBuffer, threads=1):
    """Read a FASTQ file (gzip, possibly multi-member, or plain text) into ``buf``; returns the number of bytes.
    ``threads`` > 1 lets one pigz-written member be decoded by several threads (``gunzip_parallel``)."""
    path = str(path)
    with open(path, "rb") as f:
        magic = f.read(2)
        f.seek(0)
        if magic != b"\x1f\x8b":
            n = os.fstat(f.fileno()).st_size
            buf.reserve(n)
            got = f.readinto(memoryview(buf.array)[:n]) if n else 0
            return int(got)
        hint = gzip_isize(path)
        if feed_lib() is not None:
            comp = np.empty(os.fstat(f.fileno()).st_size, dtype=np.uint8)
            f.readinto(memoryview(comp))
            got = gunzip_parallel(comp, buf, hint, threads) if threads > 1 else None
            if got is None:
                got = gunzip_into(comp, buf, hint)
            if got is not None:
                return got
            f.seek(0)                                  # let zlib have the last word (and raise its own error)
        buf.reserve(max(hint, _CHUNK))
        n = 0
        d = zlib.decompressobj(wbits=31)
        while True:
            raw = f.read(_CHUNK)
            if not raw:
                break
            while raw:
                out = d.decompress(raw, 8 * _CHUNK)
                if out:
                    if n + len(out) > buf.array.size:
                        buf.reserve(max(2 * buf.array.size, n + len(out)))
                    buf.array[n:n + len(out)] = np.frombuffer(out, dtype=np.uint8)
                    n += len(out)
                if d.eof:                                   # next member of a multi-member file (cat of .gz files)
                    raw = d.unused_data
                    d = zlib.decompressobj(wbits=31)
                else:
                    raw = d.unconsumed_tail
        tail = d.flush()
        if tail:
            if n + len(tail) > buf.array.size:
                buf.reserve(n + len(tail))
            buf.array[n:n + len(tail)] = np.frombuffer(tail, dtype=np.uint8)
            n += len(tail)
        return n


# Page-locking memory is slow (the driver pins every page), so buffers outlive a feeder: the next batch of the
# process starts with the pinned buffers of the previous one.
_buffer_cache = {True: deque(), False: deque()}
_buffer_cache_lock = threading.Lock()
_BUFFER_CACHE_MAX = 64


def _cached_buffer(pinned):
    with _buffer_cache_lock:
        q = _buffer_cache[bool(pinned)]
        if q:
            return q.popleft()
    return PinnedBuffer(0, pinned)


def _return_buffers(bufs):
    with _buffer_cache_lock:
        for b in bufs:
            q = _buffer_cache[bool(b.pinned)]
            if len(q) < _BUFFER_CACHE_MAX:
                q.append(b)


def pigz_compress(raw, level=6, threads=None, block=128 * 1024):
    """The stream ``pigz -p N`` writes -- what the reference's ``clean_reads`` leaves as ``<sample>.fq.gz``
    (``cat ... | pigz -p``, image.py:534-540): ONE gzip member; every 128 KiB block is deflated on its own (primed with
    the previous 32 KiB as dictionary) and ends with a sync flush (an empty stored block), so the blocks can be produced --
    and, by :func:`gunzip_parallel`, decoded -- concurrently.  zlib releases the GIL, so this scales over threads.
    The producer side of the host feed: tests and benches use it to make inputs of the reference's real format."""
    raw = memoryview(raw).cast("B") if not isinstance(raw, (bytes, bytearray)) else raw
    n = len(raw)
    starts = list(range(0, max(n, 1), block))

    def one(i):
        zd = bytes(raw[max(0, i - 32768):i]) if i else b""
        c = (zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd
             else zlib.compressobj(level, zlib.DEFLATED, -15, 8))
        return c.compress(raw[i:i + block]) + c.flush(zlib.Z_FINISH if i + block >= n else zlib.Z_SYNC_FLUSH)

    threads = threads or len(os.sched_getaffinity(0))
    if threads > 1 and len(starts) > 1:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            parts = list(pool.map(one, starts, chunksize=8))
    else:
        parts = [one(i) for i in starts]
    crc = 0
    for i in range(0, n, 1 << 26):
        crc = zlib.crc32(raw[i:i + (1 << 26)], crc)
    return b"".join([b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"] + parts
                    + [crc.to_bytes(4, "little") + (n & 0xFFFFFFFF).to_bytes(4, "little")])


class SampleFeeder:
    """Iterate ``(index, item, PinnedBuffer, n_bytes)`` over samples in submission order, inflating up to ``depth``
    samples ahead on ``threads`` worker threads.  Buffers are recycled: hand one back with :meth:`release`."""

    def __init__(self, items, path_of=lambda it: it, threads=None, depth=None, pinned=True, errors="raise"):
        """``path_of(item)`` may return None: the item has no file (its bytes are already somewhere) and is passed
        through with ``buf = None``.  ``errors="yield"``: a sample that cannot be read is yielded as ``(index, item,
        exception, -1)`` instead of ending the iteration, so that a batch survives a damaged file the way the reference's
        per-sample loop does (image.py:1020-1027)."""
        self.items = list(items)
        self.path_of = path_of
        self.errors = errors
        self.threads = threads or max(1, min(len(os.sched_getaffinity(0)), 16))
        self.depth = depth or self.threads + 1
        # fewer samples than threads: the spare threads split single gzip members (pigz-written files) into pieces
        self.piece_threads = max(1, self.threads // max(1, min(len(self.items), self.threads)))
        self.pinned = pinned
        self._free = deque()
        self._lock = threading.Lock()
        self.pool = ThreadPoolExecutor(max_workers=self.threads, thread_name_prefix="vk-inflate")

    def _job(self, it):
        path = self.path_of(it)
        if path is None:
            return None, 0
        with self._lock:
            buf = self._free.popleft() if self._free else None
        if buf is None:
            buf = _cached_buffer(self.pinned)
        try:
            n = inflate_into(path, buf, self.piece_threads)
        except BaseException:
            self.release(buf)
            raise
        return buf, n

    def release(self, buf):
        if buf is None or isinstance(buf, BaseException):
            return
        with self._lock:
            self._free.append(buf)

    def __iter__(self):
        pending = deque()
        nxt = 0
        try:
            while nxt < len(self.items) or pending:
                while nxt < len(self.items) and len(pending) < self.depth:
                    pending.append((nxt, self.items[nxt], self.pool.submit(self._job, self.items[nxt])))
                    nxt += 1
                i, it, fut = pending.popleft()
                try:
                    buf, n = fut.result()
                except Exception as exc:
                    if self.errors != "yield":
                        raise
                    buf, n = exc, -1
                yield i, it, buf, n
        finally:
            for _, _, fut in pending:
                fut.cancel()

    def close(self):
        self.pool.shutdown(wait=True, cancel_futures=True)
        with self._lock:
            free, self._free = list(self._free), deque()
        _return_buffers(free)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
