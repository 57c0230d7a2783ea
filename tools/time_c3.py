import sys, os
sys.path.insert(0, '.')
import torch
from varkoder_b200 import synth
from varkoder_b200.engine import Engine, Params
from varkoder_b200.mapping import get_kmer_mapping
eng = Engine(0)
n = int(os.environ.get("VK_N", "1000000000"))
total = synth.fixed_total_bytes(n, 150)
dev = torch.empty(total + 64, dtype=torch.uint8, device='cuda')
eng.synth_fastq(dev.data_ptr(), dev.numel(), n, 150, seed=1)
t = get_kmer_mapping(9, 'varKode')
p = Params(k=9, min_bp=500_000, max_bp=None, seed=1)
for _ in range(2):
    r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=11)
acc = {}
for _ in range(5):
    r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=11)
    for kk, v in eng.timings().items():
        acc[kk] = acc.get(kk, 0) + v / 5
print('k 9', {a: round(b, 4) for a, b in acc.items()})
