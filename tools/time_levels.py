"""count-kernel time with the full ladder against a single level (all reads in one segment: perfectly balanced)"""
import sys
sys.path.insert(0, '.')
import torch
from varkoder_b200 import synth
from varkoder_b200.engine import Engine, Params
from varkoder_b200.mapping import get_kmer_mapping
eng = Engine(0)
n = 200_000_000
total = synth.fixed_total_bytes(n, 150)
dev = torch.empty(total + 64, dtype=torch.uint8, device='cuda')
eng.synth_fastq(dev.data_ptr(), dev.numel(), n, 150, seed=1)
t = get_kmer_mapping(7, 'cgr')
for name, p in (("ladder9", Params(k=7, min_bp=500_000, max_bp=200_000_000, seed=1)),
                ("single", Params(k=7, min_bp=0, max_bp=200_000_000, is_query=True, seed=1)),
                ("ladder3", Params(k=7, min_bp=50_000_000, max_bp=200_000_000, seed=1))):
    for _ in range(3):
        r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
    acc = {}
    for _ in range(20):
        r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
        for kk, v in eng.timings().items():
            acc[kk] = acc.get(kk, 0) + v / 20
    print(name, len(r.levels), {a: round(b, 4) for a, b in acc.items()})
