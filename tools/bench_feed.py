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
        t0 = time.perf_counter()
        st = stages.images_for_samples(samples, out, table, k=7, mapping_code="varKode", min_bp=500_000,
                                       max_bp=200_000_000, threads=threads, gpu_workers=workers)
        dt = time.perf_counter() - t0
        n_png = sum(len(files) for _, _, files in os.walk(out))
        assert len(st) == n_samples and all("failed_step" not in v for v in st.values())
        print(f"decoder={decoder} gpu_workers={workers} inflate_threads={threads}: {dt:.2f} s  {n_samples * n_bases / dt / 1e9:.2f} Gbases/s  "
              f"{n_samples * total / dt / 1e9:.2f} GB/s of text  {n_png} PNGs", flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
