"""Host feed: threaded inflate into (pinned) buffers, ordering, gzip corner cases.  CPU only."""
import gzip
import os

import numpy as np
import pytest

from tests.helpers import fastq, rand_reads
from varkoder_b200 import feed


def _write(tmp_path, name, data, mode):
    p = tmp_path / name
    if mode == "gz":
        with gzip.open(p, "wb", compresslevel=6) as f:
            f.write(data)
    elif mode == "multi":                     # cat a.gz b.gz, as `cat | pigz` of several inputs may look
        half = len(data) // 2
        with open(p, "wb") as f:
            f.write(gzip.compress(data[:half]))
            f.write(gzip.compress(data[half:]))
    else:
        p.write_bytes(data)
    return p


@pytest.mark.parametrize("mode", ["gz", "multi", "plain"])
def test_inflate_into_roundtrip(tmp_path, mode):
    rng = np.random.default_rng(3)
    data = fastq(rand_reads(rng, 4000, 0, 200))
    p = _write(tmp_path, "s.fq.gz" if mode != "plain" else "s.fq", data, mode)
    buf = feed.PinnedBuffer(0, pinned=False)
    n = feed.inflate_into(p, buf)
    assert n == len(data) and buf.array[:n].tobytes() == data
    # a buffer that is too small grows; a reused larger buffer keeps working
    small = feed.PinnedBuffer(16, pinned=False)
    assert feed.inflate_into(p, small) == len(data) and small.array[:n].tobytes() == data
    assert feed.inflate_into(p, small) == len(data)


def test_inflate_empty_and_isize(tmp_path):
    p = _write(tmp_path, "e.fq.gz", b"", "gz")
    buf = feed.PinnedBuffer(0, pinned=False)
    assert feed.inflate_into(p, buf) == 0
    data = b"@r\nACGT\n+\nIIII\n" * 1000
    p = _write(tmp_path, "x.fq.gz", data, "gz")
    assert feed.gzip_isize(p) == len(data)


def test_sample_feeder_order_and_recycling(tmp_path):
    rng = np.random.default_rng(8)
    datas, paths = [], []
    for i in range(9):
        d = fastq(rand_reads(rng, 200 + 300 * (i % 3), 20, 120))
        datas.append(d)
        paths.append(_write(tmp_path, f"s{i}.fq.gz", d, "gz"))
    seen = []
    bufs = set()
    with feed.SampleFeeder(paths, threads=3, depth=2, pinned=False) as fd:
        for i, it, buf, n in fd:
            assert it == paths[i] and buf.array[:n].tobytes() == datas[i]
            seen.append(i)
            bufs.add(id(buf))
            fd.release(buf)
    assert seen == list(range(9))
    assert len(bufs) <= 4                      # buffers are recycled, not one per sample


def test_sample_feeder_propagates_errors(tmp_path):
    bad = tmp_path / "bad.fq.gz"
    bad.write_bytes(b"\x1f\x8b" + os.urandom(64))
    with feed.SampleFeeder([bad], threads=1, pinned=False) as fd:
        with pytest.raises(Exception):
            list(fd)


# ---------------------------------------------------------------------------------- libvk_feed.so (csrc/vk_inflate.c)
def _deflate(raw, level, strategy):
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strategy)
    return c.compress(raw) + c.flush()


def _corpus():
    rng = np.random.default_rng(11)
    from varkoder_b200 import synth
    return {
        "fastq_fixed": synth.fixed(400_000, 150, seed=3).tobytes(),
        "fastq_ragged": synth.variable(3000, seed=4).tobytes(),
        "random": rng.integers(0, 256, 120_000, dtype=np.uint8).tobytes(),          # stored blocks / long codes
        "zeros": bytes(300_000),                                                    # distance 1, maximal matches
        "text": b"the quick brown fox jumps over the lazy dog. " * 8000,
        "four_symbols": rng.integers(0, 4, 200_000, dtype=np.uint8).tobytes(),
        "skewed": (rng.geometric(0.02, 150_000) % 256).astype(np.uint8).tobytes(),  # code lengths up to 15: sub-tables
        "empty": b"",
        "one_byte": b"A",
        "tiny": b"ACGT\n" * 3,
    }


def test_feed_library_is_built_and_used():
    assert feed.feed_lib() is not None, "libvk_feed.so missing: run make -C varkoder_b200/csrc"


def test_gunzip_equals_zlib_on_every_block_type():
    """stored / fixed / dynamic blocks, every zlib strategy and level, exact-size and roomy output buffers"""
    import zlib
    for name, raw in _corpus().items():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                comp = _deflate(raw, level, strategy)
                for slack in (0, 1000):
                    buf = feed.PinnedBuffer(0, pinned=False)
                    n = feed.gunzip_into(comp, buf, size_hint=len(raw) + slack)
                    assert n == len(raw) and buf.array[:n].tobytes() == raw, (name, level, strategy, slack)


def test_gunzip_multi_member_padding_and_growth():
    raw = _corpus()["fastq_fixed"]
    comp = gzip.compress(raw[:100_000], 6) + gzip.compress(b"", 6) + gzip.compress(raw[100_000:250_000], 1) + bytes(37)
    buf = feed.PinnedBuffer(0, pinned=False)
    n = feed.gunzip_into(comp, buf, size_hint=0)                 # no hint: the buffer grows until the text fits
    assert n == 250_000 and buf.array[:n].tobytes() == raw[:250_000]
    # header fields: FEXTRA, FNAME, FCOMMENT, FHCRC
    body = gzip.compress(raw[:5000], 6)
    hdr = bytes([0x1f, 0x8b, 8, 2 | 4 | 8 | 16]) + body[4:10] + b"\x03\x00abc" + b"name\x00" + b"comment\x00" + b"\x12\x34"
    n = feed.gunzip_into(hdr + body[10:], buf, size_hint=5000)
    assert n == 5000 and buf.array[:n].tobytes() == raw[:5000]


def test_gunzip_rejects_damage():
    """truncation anywhere and single bit flips: never accepted, never a crash (CRC-32 + ISIZE are checked)"""
    rng = np.random.default_rng(5)
    raw = _corpus()["fastq_ragged"]
    comp = gzip.compress(raw, 6)
    buf = feed.PinnedBuffer(len(raw) + 64, pinned=False)
    for cut in (0, 5, 17, 100, len(comp) // 2, len(comp) - 9, len(comp) - 1):
        assert feed.gunzip_into(comp[:cut], buf, size_hint=len(raw)) is None, cut
    for _ in range(400):
        b = bytearray(comp)
        b[int(rng.integers(10, len(comp)))] ^= 1 << int(rng.integers(0, 8))
        got = feed.gunzip_into(bytes(b), buf, size_hint=len(raw))
        assert got is None or buf.array[:got].tobytes() == raw           # a flip in a don't-care header bit may pass
    assert feed.gunzip_into(b"not gzip at all, just text", buf) is None


def test_crc32_equals_zlib():
    import ctypes
    import zlib
    L = feed.feed_lib()
    rng = np.random.default_rng(0)
    for n in list(range(0, 200)) + [1000, 4096, 65537, (1 << 20) + 13]:
        a = rng.integers(0, 256, n + 3, dtype=np.uint8)
        for off in (0, 1, 3):
            b = a[off:off + n]
            assert L.vkf_crc32(0, ctypes.c_void_p(b.ctypes.data), n) == zlib.crc32(b.tobytes()), (n, off)
        h = n // 2
        run = L.vkf_crc32(L.vkf_crc32(0, ctypes.c_void_p(a.ctypes.data), h), ctypes.c_void_p(a[h:].ctypes.data), n - h)
        assert run == zlib.crc32(a[:n].tobytes())


def test_inflate_falls_back_to_zlib(tmp_path, monkeypatch):
    """without the library (or when it declines a file) the zlib path gives the same bytes"""
    data = fastq(rand_reads(np.random.default_rng(8), 500, 0, 120))
    p = tmp_path / "x.fq.gz"
    p.write_bytes(gzip.compress(data))
    monkeypatch.setattr(feed, "_feed_lib", False)
    assert feed.feed_lib() is None
    buf = feed.PinnedBuffer(0, pinned=False)
    n = feed.inflate_into(p, buf)
    assert buf.array[:n].tobytes() == data


# ------------------------------------------------------------------------ one gzip member on several threads
def pigz_like(raw, level=6, block=128 * 1024, independent=False):
    """the stream pigz writes (the reference's clean_reads output, image.py:534-540): ONE member; every block is
    compressed on its own, primed with the previous 32 KiB unless -i, and ends with a sync flush (empty stored block)"""
    import zlib
    out = [b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"]
    for i in range(0, max(len(raw), 1), block):
        zd = raw[max(0, i - 32768):i] if (i and not independent) else b""
        c = (zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd
             else zlib.compressobj(level, zlib.DEFLATED, -15, 8))
        last = i + block >= len(raw)
        out.append(c.compress(raw[i:i + block]) + c.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH))
    out.append(zlib.crc32(raw).to_bytes(4, "little") + (len(raw) & 0xFFFFFFFF).to_bytes(4, "little"))
    return b"".join(out)


@pytest.mark.parametrize("independent", [False, True])
@pytest.mark.parametrize("block", [128 * 1024, 20_000])
def test_parallel_member_equals_serial(independent, block):
    """pieces cut behind pigz's sync points, decoded concurrently with placeholders for the unseen 32 KiB, resolved front
    to back: byte-identical to the serial decode; blocks shorter than the window make placeholders chain through
    several pieces"""
    import zlib
    from varkoder_b200 import synth
    raw = synth.variable(9000, seed=21).tobytes() + b"ACGT" * 50_000 + synth.fixed(600_000, 150, seed=2).tobytes()
    comp = pigz_like(raw, 6, block, independent)
    assert zlib.decompress(comp, 31) == raw
    buf = feed.PinnedBuffer(0, pinned=False)
    for threads, min_piece in ((2, 1 << 16), (5, 1 << 14), (16, 4096)):
        n = feed.gunzip_parallel(comp, buf, len(raw), threads, min_piece=min_piece)
        assert n == len(raw) and buf.array[:n].tobytes() == raw, (threads, min_piece)


def test_parallel_member_declines_what_it_cannot_prove(tmp_path):
    import zlib
    from varkoder_b200 import synth
    raw = synth.fixed(900_000, 150, seed=8).tobytes()
    buf = feed.PinnedBuffer(0, pinned=False)
    # an ordinary gzip stream has no sync points
    assert feed.gunzip_parallel(zlib.compress(raw, 6, 31), buf, len(raw), 8, min_piece=4096) is None
    # two members: the last piece does not end in front of the file's trailer
    two = pigz_like(raw[:500_000]) + pigz_like(raw[500_000:])
    assert feed.gunzip_parallel(two, buf, len(raw), 8, min_piece=4096) is None
    # stored data that CONTAINS the marker bytes: cuts at places that are no block starts -> pieces fail -> declined
    fake = (b"\x00\x00\xff\xff" + bytes(range(256))) * 8000
    c = zlib.compressobj(0, zlib.DEFLATED, 31)
    comp = c.compress(fake) + c.flush()
    assert feed.gunzip_parallel(comp, buf, len(fake), 8, min_piece=4096) is None
    # damaged CRC
    good = pigz_like(raw)
    bad = good[:-8] + bytes([good[-8] ^ 1]) + good[-7:]
    assert feed.gunzip_parallel(good, buf, len(raw), 4, min_piece=1 << 15) == len(raw)
    assert feed.gunzip_parallel(bad, buf, len(raw), 4, min_piece=1 << 15) is None
    # end to end through inflate_into: every one of them still comes out right (serial decoder / zlib take over)
    for name, blob, want in (("a", two, raw), ("b", comp, fake), ("c", good, raw)):
        p = tmp_path / f"{name}.fq.gz"
        p.write_bytes(blob)
        n = feed.inflate_into(p, buf, threads=8)
        assert buf.array[:n].tobytes() == want


def test_crc32_combine():
    import zlib
    L = feed.feed_lib()
    rng = np.random.default_rng(4)
    for la, lb in ((0, 0), (1, 0), (0, 5), (7, 9), (1000, 1), (65536, 100_003), (3, 1 << 20)):
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert L.vkf_crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)


def test_feeder_splits_members_when_samples_are_few(tmp_path):
    from varkoder_b200 import synth
    raw = synth.fixed(3_000_000, 150, seed=12).tobytes()
    p = tmp_path / "big.fq.gz"
    p.write_bytes(pigz_like(raw, 1))
    with feed.SampleFeeder([str(p)], threads=8, pinned=False) as fd:
        assert fd.piece_threads == 8
        for i, it, buf, n in fd:
            assert buf.array[:n].tobytes() == raw
            fd.release(buf)
