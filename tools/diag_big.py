"""one big sample (VK_N bases, read length 150), two steps: count fallbacks, lane flips, last step's kernel split"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from varkoder_b200 import synth
from varkoder_b200.engine import Engine, Params
from varkoder_b200.mapping import get_kmer_mapping
eng = Engine(0)
n_bases = int(os.environ.get('VK_N', '15000000000'))
total = synth.fixed_total_bytes(n_bases, 150)
d = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
eng.synth_fastq(d.data_ptr(), d.numel(), n_bases, 150, seed=20260118 + 5000, first_read=0)
table = get_kmer_mapping(int(os.environ.get('VK_K', '7')), "cgr")
p = Params(k=int(os.environ.get('VK_K', '7')), min_bp=500_000, max_bp=None, seed=11)
eng.set_fine_timing(True)
for i in range(3):
    r = eng.reads_to_images(d.data_ptr(), p, table, on_device=True, n_bytes=total, max_levels=18, want_canon=True)
    print("step", i, "fallbacks", eng.count_fallbacks(), {k: round(v, 3) for k, v in eng.timings().items()}, "max canon", int(r.canon.max()), flush=True)
