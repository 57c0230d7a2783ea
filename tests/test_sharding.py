"""Multi-GPU host logic on CPU: by-sample assignment, record-boundary sharding, and the read-sharded path with a
world_size-2 gloo group.  The CUDA engine cannot run here, so the sharded driver is exercised with a stand-in engine
built on the oracle (tests only); the same driver runs with the real Engine + NCCL in test_gpu_parity / bench."""
import ctypes
import os
import socket

import numpy as np
import pytest

from oracle import dsk, image as oimg
from tests.helpers import fastq, oracle_images, oracle_levels, rand_reads
from varkoder_b200 import _lib, sharding
from varkoder_b200.engine import Params, Result
from varkoder_b200.ladder import LessThanMinimumData, ladder, parse_seed
from varkoder_b200.mapping import get_kmer_mapping


from tests.helpers import OracleEngine  # noqa: E402  (CPU stand-in for the CUDA engine)


def test_assign_samples_lpt():
    sizes = [50, 10, 40, 30, 20, 45, 5]
    owner, loads = sharding.assign_samples(sizes, 3)
    assert sum(loads) == sum(sizes) and max(loads) - min(loads) <= 10
    assert sharding.assign_samples(sizes, 3) == (owner, loads)                  # deterministic
    assert sharding.assign_samples([], 4) == ([], [0, 0, 0, 0])
    owner1, loads1 = sharding.assign_samples(sizes, 1)
    assert set(owner1) == {0} and loads1 == [200]
    with pytest.raises(ValueError):
        sharding.assign_samples(sizes, 0)


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_split_records_boundaries(n_shards):
    rng = np.random.default_rng(5)
    reads = rand_reads(rng, 200, 0, 180)
    for final_newline in (True, False):
        buf = fastq(reads, final_newline=final_newline)
        parts = sharding.split_records(buf, n_shards)
        assert parts[0][0] == 0 and parts[-1][1] == len(buf)
        p = dsk.parse_fastq(buf)
        first = 0
        for (b, e, fr), nxt in zip(parts, parts[1:] + [None]):
            assert b <= e and (nxt is None or nxt[0] == e)
            assert fr == first
            if b < e:
                assert buf[b:b + 1] == b"@"
            first += dsk.parse_fastq(buf[b:e])["n_reads"]
        assert first == p["n_reads"]
    # fewer records than shards, and the empty buffer
    tiny = fastq(["ACGTACGT"])
    parts = sharding.split_records(tiny, 4)
    assert sum(e - b for b, e, _ in parts) == len(tiny)
    assert sharding.split_records(b"", 3) == [(0, 0, 0)] * 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, buf, parts, k, seed, min_bp, max_bp, mapping, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e, _ = parts[rank]
        table = get_kmer_mapping(k, mapping)
        res = sharding.sharded_reads_to_images(OracleEngine(), buf[b:e], Params(k=k, min_bp=min_bp, max_bp=max_bp, seed=seed),
                                               table, want_canon=True)
        q.put((rank, res.levels, res.level_reads, res.level_bases, res.nsites, res.n_reads, res.canon, res.pixels))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_read_sharded_path_gloo(world):
    """N ranks, one sample cut at record boundaries: counts and pixels of every level equal the unsharded oracle."""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(77)
    reads = rand_reads(rng, 1500, 0, 220, p_n=0.01) + ["ACGT" * 200]            # one read longer than breaklength
    buf = fastq(reads)
    k, seed, min_bp, max_bp, mapping = 7, 12345, 20_000, 100_000, "cgr"
    parts = sharding.split_records(buf, world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, buf, parts, k, seed, min_bp, max_bp, mapping, q))
             for r in range(world)]
    for p_ in procs:
        p_.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    whole = dsk.parse_fastq(buf)
    levels = ladder(whole["nsites_ref"], min_bp, max_bp)
    assert len(levels) >= 3
    expect_canon = oracle_levels(buf, k, seed, levels, whole["nsites_ref"])
    expect_pix = oracle_images(expect_canon, get_kmer_mapping(k, mapping).lut)
    for rank, lv, lreads, lbases, nsites, n_reads, canon, pixels in got:
        assert lv == levels and nsites == whole["nsites_ref"] and n_reads == whole["n_reads"]
        assert (canon == expect_canon).all()
        assert (pixels == expect_pix).all()
        # thresholds fitted to the base targets (reformat.sh samplebasestarget): the reads drawn for a level hold its
        # target to within one read (fewer than 65536 reads: one read per priority bucket)
        for lvl, bp in enumerate(levels):
            sel = dsk.select_reads(whole["n_reads"], seed, bp, whole["nsites_ref"], lens=whole["lens"]).astype(bool)
            keep = sel & (whole["lens"] >= k)
            assert lbases[lvl] == int(whole["lens"][keep].sum()) and lreads[lvl] == int(keep.sum())
            assert bp >= whole["nsites_ref"] or abs(int(whole["lens"][sel].sum()) - bp) <= int(whole["lens"].max())
        assert all(a >= b for a, b in zip(lreads, lreads[1:]))                  # nested levels
