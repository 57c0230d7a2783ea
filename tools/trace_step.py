"""VK_TRACE_EACH=1 python tools/trace_step.py : one 200 Mbp step with a CUDA event behind every launch (stderr)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from varkoder_b200 import synth
from varkoder_b200.engine import Engine, Params
from varkoder_b200.mapping import get_kmer_mapping
eng = Engine(0)
n_bases = int(os.environ.get('VK_N', '200000000'))
total = synth.fixed_total_bytes(n_bases, 150)
devs = []
for j in range(2):
    d = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fastq(d.data_ptr(), d.numel(), n_bases, 150, seed=5, first_read=j * 2_000_000)
    devs.append(d)
table = get_kmer_mapping(7, "cgr")
p = Params(k=7, min_bp=500_000, max_bp=None if os.environ.get('VK_NOMAX') else 200_000_000, seed=1)
table = get_kmer_mapping(7, os.environ.get('VK_MAP', 'cgr'))
for i in range(6):
    print("step", i, file=sys.stderr, flush=True)
    r = eng.reads_to_images(devs[i & 1].data_ptr(), p, table, on_device=True, n_bytes=total, max_levels=16)
print(eng.timings(), file=sys.stderr)
