"""End to end FROM GZIP FILES: N synthetic cleaned .fq.gz samples -> threaded inflate into pinned buffers -> GPU -> PNGs on
disk, through stages.images_for_samples (the batch entry point).  Side measurement, not the bench line: it is bound by
gzip inflate on the host cores (one gzip member cannot be split; parallelism is across samples).
usage: python tools/bench_feed.py [n_samples] [bases_per_sample]"""
import gzip
import os
import shutil
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from varkoder_b200 import stages, synth
from varkoder_b200.engine import Engine
from varkoder_b200.mapping import get_kmer_mapping

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_bases = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
tmp = tempfile.mkdtemp(prefix="vkfeed_")
try:
    eng = Engine(0)
    total = synth.fixed_total_bytes(n_bases, 150)
    dev = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    samples = []
    t0 = time.perf_counter()
    gz_bytes = 0
    for i in range(n_samples):
        eng.synth_fastq(dev.data_ptr(), dev.numel(), n_bases, 150, seed=77, first_read=i * 10_000_000)
        raw = dev[:total].cpu().numpy().tobytes()
        p = os.path.join(tmp, f"S{i:03d}.fq.gz")
        with open(p, "wb") as f:
            c = zlib.compressobj(1, zlib.DEFLATED, 31)
            f.write(c.compress(raw) + c.flush())
        gz_bytes += os.path.getsize(p)
        samples.append(dict(sample=f"S{i:03d}", path=p, labels=["x"], base_sd=0.0))
    eng.close()
    print(f"wrote {n_samples} x {n_bases} bases as gzip level 1: {gz_bytes / 1e6:.0f} MB compressed, "
          f"{n_samples * total / 1e6:.0f} MB text, {time.perf_counter() - t0:.1f} s", flush=True)
    table = get_kmer_mapping(7, "varKode")
    threads = len(os.sched_getaffinity(0))
    from varkoder_b200 import feed
    lib = feed.feed_lib()
    for decoder, workers in (("zlib", 1), ("libvk_feed", 1), ("libvk_feed", 3)):
        feed._feed_lib = False if decoder == "zlib" else lib
        out = os.path.join(tmp, f"images_{decoder}_{workers}")
        times = []
        for _ in range(3):                     # a shared box: the best of three says what the code can do
            shutil.rmtree(out, ignore_errors=True)
            t0 = time.perf_counter()
            st = stages.images_for_samples(samples, out, table, k=7, mapping_code="varKode", min_bp=500_000,
                                           max_bp=200_000_000, threads=threads, gpu_workers=workers)
            times.append(time.perf_counter() - t0)
        dt = min(times)
        n_png = sum(len(files) for _, _, files in os.walk(out))
        assert len(st) == n_samples and all("failed_step" not in v for v in st.values())
        print(f"decoder={decoder} gpu_workers={workers} inflate_threads={threads}: {dt:.2f} s  {n_samples * n_bases / dt / 1e9:.2f} Gbases/s  "
              f"{n_samples * total / dt / 1e9:.2f} GB/s of text  {n_png} PNGs  (three runs: {", ".join(f"{x:.2f}" for x in times)} s)", flush=True)
    # ---- ONE large sample written the way pigz writes it (one member, a sync point after every 128 KiB chunk, chunks
    # primed with the previous 32 KiB): the spare threads split the member
    big_bases = 200_000_000
    big_total = synth.fixed_total_bytes(big_bases, 150)
    eng = Engine(0)
    dev = torch.empty(big_total + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fastq(dev.data_ptr(), dev.numel(), big_bases, 150, seed=78)
    raw = dev[:big_total].cpu().numpy().tobytes()
    eng.close()
    p = os.path.join(tmp, "BIG.fq.gz")
    with open(p, "wb") as f:
        f.write(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03")
        blk = 128 * 1024
        for i in range(0, len(raw), blk):
            zd = raw[max(0, i - 32768):i]
            c = zlib.compressobj(1, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd else zlib.compressobj(1, zlib.DEFLATED, -15, 8)
            f.write(c.compress(raw[i:i + blk]) + c.flush(zlib.Z_FINISH if i + blk >= len(raw) else zlib.Z_SYNC_FLUSH))
        f.write(zlib.crc32(raw).to_bytes(4, "little") + (len(raw) & 0xFFFFFFFF).to_bytes(4, "little"))
    print(f"one pigz-like sample: {big_bases} bases, {os.path.getsize(p) / 1e6:.0f} MB compressed, {big_total / 1e6:.0f} MB text", flush=True)
    buf = feed.PinnedBuffer(big_total + 64)
    buf.array[:] = 0
    for label, thr, use_lib in (("zlib", 1, False), ("libvk_feed serial", 1, True), ("libvk_feed pieces", threads, True)):
        feed._feed_lib = lib if use_lib else False
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            n = feed.inflate_into(p, buf, threads=thr)
            best = min(best, time.perf_counter() - t0)
        assert n == big_total and buf.array[:n].tobytes() == raw
        print(f"  inflate, {label} ({thr} threads): {best:.3f} s  {big_total / best / 1e9:.2f} GB/s of text", flush=True)
    feed._feed_lib = lib
    sample = [dict(sample="BIG", path=p, labels=["x"], base_sd=0.0)]
    for _ in range(2):
        out = os.path.join(tmp, "images_big")
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.perf_counter()
        st = stages.images_for_samples(sample, out, get_kmer_mapping(7, "cgr"), k=7, mapping_code="cgr", min_bp=500_000,
                                       max_bp=200_000_000, threads=threads, gpu_workers=1)
        dt = time.perf_counter() - t0
    print(f"  file -> 9 PNGs through images_for_samples: {dt:.3f} s  {big_bases / dt / 1e9:.2f} Gbases/s", flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
