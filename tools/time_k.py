import sys, time, os
sys.path.insert(0, '.')
import torch
from varkoder_b200 import synth
from varkoder_b200.engine import Engine, Params
from varkoder_b200.mapping import get_kmer_mapping
eng = Engine(0)
eng.set_fine_timing(True)
n = int(os.environ.get("VK_N", "200000000"))
total = synth.fixed_total_bytes(n, 150)
dev = torch.empty(total + 64, dtype=torch.uint8, device='cuda')
eng.synth_fastq(dev.data_ptr(), dev.numel(), n, 150, seed=1)
for k in [int(x) for x in sys.argv[1:]] or [8]:
    t = get_kmer_mapping(k, 'cgr')
    p = Params(k=k, min_bp=500_000, max_bp=200_000_000, seed=1)
    for _ in range(3):
        r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
    acc = {}
    for _ in range(10):
        r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
        for kk, v in eng.timings().items():
            acc[kk] = acc.get(kk, 0) + v / 10
    print('k', k, 'COUNT16', os.environ.get('VK_COUNT16', '1'), {a: round(b, 4) for a, b in acc.items()})
