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
