"""Multi-GPU host logic of the image hot path (one process per GPU, ``torch.distributed``).

Two ways the path shards:

* **by sample** -- samples are independent in the reference (``multiprocessing.Pool`` over samples,
  varKoder/commands/image.py:1281-1294).  :func:`assign_samples` is a longest-processing-time-first greedy
  over the samples' byte sizes; every rank then runs the single-GPU path on its own samples.  No collective.
* **by read shard inside one sample** -- the k-mer histogram is a commutative integer sum over reads, so a
  very large sample is cut into contiguous record ranges (:func:`split_records`), every rank counts its range
  into per-*segment* forward histograms, and ONE exchange step sums them:
  ``all_reduce(SUM)`` over ``VK_MAX_LEVELS * 4^k`` 64-bit integers (NCCL over NVLink on GPUs, gloo in the CPU
  tests).  Level membership of a read depends only on ``prio64(seed, global read index)`` and on the
  sample-wide base count, so the ranks first exchange two scalars each (records and bases of their shard,
  an ``all_gather``) and pass ``read_index_base`` / ``nsites_override`` to ``vk_count``.

The engine is duck-typed (``upload``, ``parse``, ``count``, ``render``): the product passes
:class:`varkoder_b200.engine.Engine`; the gloo tests pass a CPU stand-in built on the oracle.
"""
from dataclasses import replace

import numpy as np

from . import _lib
from .engine import Params, Result


# ------------------------------------------------------------------------------------------ by sample
def assign_samples(sizes, n_ranks):
    """Longest-processing-time-first greedy.  ``sizes``: bytes (or bases) per sample.
    Returns ``(owner, loads)``: ``owner[i]`` = rank of sample i, ``loads[r]`` = total size on rank r.
    Deterministic (ties: lower sample index first, lower rank first) so every rank computes the same plan."""
    n_ranks = int(n_ranks)
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * n_ranks
    owner = [0] * len(sizes)
    for i in order:
        r = min(range(n_ranks), key=lambda q: (loads[q], q))
        owner[i] = r
        loads[r] += int(sizes[i])
    return owner, loads


# --------------------------------------------------------------------------------------- by read shard
def split_records(buf, n_shards):
    """Cut FASTQ bytes into ``n_shards`` contiguous byte ranges that start at record boundaries (a record = 4
    lines, as split_fastq frames them, image.py:662-667) and hold about the same number of bytes.
    Returns ``[(begin, end, first_record)]``; ranges may be empty when there are fewer records than shards."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    n = int(a.size)
    nl = np.flatnonzero(a == 10)                                   # offsets of '\n'
    rec_start = np.concatenate((np.zeros(1, dtype=np.int64), nl[3::4].astype(np.int64) + 1))
    rec_start = rec_start[rec_start < n]                           # records that hold at least one byte
    n_rec = int(rec_start.size)
    idx = [0]
    for s in range(1, int(n_shards)):
        j = int(np.searchsorted(rec_start, n * s // int(n_shards), side="left"))    # first record at or after the target
        idx.append(min(max(j, idx[-1]), n_rec))
    idx.append(n_rec)
    out = []
    for s in range(int(n_shards)):
        first, last = idx[s], idx[s + 1]
        begin = int(rec_start[first]) if first < n_rec else n
        end = int(rec_start[last]) if last < n_rec else n
        out.append((begin, end, first))
    return out


def _exchange_shard_stats(n_reads, nsites, group=None):
    """all_gather of (records, bases) of every rank's shard -> (read_index_base, nsites_total, per_rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([int(n_reads), int(nsites)], dtype=torch.int64, device=dev)
    allv = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    flat = allv.cpu().tolist()                                   # one read-back for all ranks
    per_rank = [(flat[2 * r], flat[2 * r + 1]) for r in range(world)]
    base = sum(r for r, _ in per_rank[:rank])
    total = sum(s for _, s in per_rank)
    return base, total, per_rank


def sharded_count(engine, shard_bytes, params: Params, seg_hist, group=None):
    """Count this rank's shard and sum the per-segment histograms over all ranks, in place in ``seg_hist``
    (a contiguous int64/uint64 torch tensor of ``VK_MAX_LEVELS * 4^k`` elements on the engine's device).

    Returns the :class:`Result` of the WHOLE sample (levels, realised reads / bases summed over ranks)."""
    import torch
    import torch.distributed as dist
    nk = 4 ** params.k
    if seg_hist.numel() != _lib.VK_MAX_LEVELS * nk or seg_hist.element_size() != 8 or not seg_hist.is_contiguous():
        raise ValueError("seg_hist must be a contiguous 64-bit tensor of VK_MAX_LEVELS * 4^k elements")
    if shard_bytes is not None:          # None: the shard is already resident (Engine.upload / Engine.attach)
        engine.upload(shard_bytes)
    st = engine.parse()
    base, total, per_rank = _exchange_shard_stats(st["n_reads"], st["nsites"], group)
    p = replace(params, read_index_base=base, nsites_override=total)      # 0 bases in total: every shard is empty too
    if p.sampling == _lib.VK_SAMPLING_CALIBRATED and not p.prio_hist:
        # thresholds fitted to the base targets need the base histogram of the WHOLE sample over the priority buckets:
        # every rank builds its shard's, one more all_reduce (512 KiB) sums them
        hist = torch.zeros(_lib.VK_PRIO_BUCKETS, dtype=torch.int64, device=seg_hist.device)
        engine.prio_hist(p, hist.data_ptr())
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        if hist.is_cuda:
            torch.cuda.current_stream(hist.device).synchronize()          # the engine counts on its own stream
        p = replace(p, prio_hist=hist.data_ptr())
    res = engine.count(p, seg_hist.data_ptr())
    # the one exchange step of the path; every rank derives the same ladder, so only its levels are exchanged.  The
    # realised reads / bases per level ride along in the first unused row of the table (same reduction, one collective)
    nl = len(res.levels)
    M = _lib.VK_MAX_LEVELS
    words = seg_hist.view(torch.int64)
    mine = torch.zeros(2 * M, dtype=torch.int64)
    if nl:
        mine[:nl] = torch.tensor(res.level_reads, dtype=torch.int64)
        mine[M:M + nl] = torch.tensor(res.level_bases, dtype=torch.int64)
    if nl < M and nk >= 2 * M:
        tail = words[nl * nk:nl * nk + 2 * M]
        tail.copy_(mine)
        dist.all_reduce(words[:nl * nk + 2 * M], op=dist.ReduceOp.SUM, group=group)
        tot = tail.to("cpu", copy=True)            # copy: on a CPU tensor .cpu() would alias the row zeroed next
        tail.zero_()
    else:
        dist.all_reduce(words[:max(nl, 1) * nk], op=dist.ReduceOp.SUM, group=group)
        tot = mine.to(seg_hist.device)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        tot = tot.cpu()
    n_reads = sum(r for r, _ in per_rank)
    return Result(n_bytes=st["n_bytes"], n_lines=st["n_lines"], n_reads=n_reads, nsites=total,
                  nsites_true=st["nsites_true"], status=res.status, levels=res.levels,
                  level_reads=[int(x) for x in tot[:nl]],
                  level_bases=[int(x) for x in tot[M:M + nl]])


def fused_sharded_reads_to_images(engine, shard, params: Params, table, group=None, on_device=False, n_bytes=None,
                                  max_levels=16, want_canon=False):
    """The read-sharded path as ONE enqueue per rank: the CUDA library issues both exchange steps itself, through NCCL,
    on its own stream between its kernels (vk_sharded_reads_to_images) -- no host synchronisation until the images are
    back.  ``shard``: this rank's records (host bytes, or a device pointer with ``on_device`` / ``n_bytes``).
    The engine's communicator is created on first use (collective: every rank of ``group`` must call this)."""
    import torch.distributed as dist
    if getattr(engine, "comm_world", None) != dist.get_world_size(group):
        engine.comm_init(group)
    return engine.sharded_reads_to_images(shard, params, table, on_device=on_device, n_bytes=n_bytes,
                                          max_levels=max_levels, want_canon=want_canon)


def sharded_reads_to_images(engine, shard_bytes, params: Params, table, seg_hist=None, group=None,
                            render_on_all_ranks=True, want_canon=False):
    """Read-sharded form of ``Engine.reads_to_images`` for ONE sample spread over the ranks of ``group``, built from the
    staged calls (parse / count / render) with ``torch.distributed`` collectives between them: works with any engine and
    any backend (the CPU tests run it under gloo), at the price of a host synchronisation around every exchange.
    On GPUs prefer :func:`fused_sharded_reads_to_images`.
    Every rank passes its own shard (see :func:`split_records`); returns the whole-sample :class:`Result`
    (pixels on every rank, or only on rank 0 when ``render_on_all_ranks`` is false)."""
    import torch
    import torch.distributed as dist
    nk = 4 ** params.k
    if seg_hist is None:
        dev = torch.device("cuda", engine.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        seg_hist = torch.zeros(_lib.VK_MAX_LEVELS * nk, dtype=torch.int64, device=dev)
    res = sharded_count(engine, shard_bytes, params, seg_hist, group)
    if seg_hist.is_cuda:
        torch.cuda.current_stream(seg_hist.device).synchronize()      # the engine renders on its own stream
    if render_on_all_ranks or dist.get_rank(group) == 0:
        canon, pixels = engine.render(table, params.k, len(res.levels), seg_hist.data_ptr(), want_canon=want_canon)
        res.canon, res.pixels = canon, pixels
    return res
