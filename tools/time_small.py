"""step time of the fused path for small samples (BASELINE configs[3] regime) and with several contexts in flight"""
import sys, time, threading
sys.path.insert(0, '.')
import torch
from varkoder_b200 import synth
from varkoder_b200.engine import Engine, Params
from varkoder_b200.mapping import get_kmer_mapping
t = get_kmer_mapping(7, 'varKode')
for n in (1_000_000, 10_000_000, 30_000_000, 50_000_000, 200_000_000):
    eng = Engine(0)
    eng.set_fine_timing(False)
    total = synth.fixed_total_bytes(n, 150)
    dev = torch.empty(total + 64, dtype=torch.uint8, device='cuda')
    eng.synth_fastq(dev.data_ptr(), dev.numel(), n, 150, seed=1)
    p = Params(k=7, min_bp=500_000, max_bp=200_000_000, seed=1)
    for _ in range(5):
        r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); R = 200
    ms = 0.0
    for _ in range(R):
        r = eng.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
        ms += eng.timings()['total']
    wall = (time.perf_counter() - t0) / R * 1e3
    print(f"n={n:>11d} levels={len(r.levels)} device {ms/R*1e3:8.1f} us  wall {wall*1e3:8.1f} us  {n/(wall*1e-3)/1e9:7.1f} Gbases/s (wall)")
    # T contexts in flight, each its own thread
    for T in (2, 4):
        engs = [Engine(0) for _ in range(T)]
        for e in engs:
            e.set_fine_timing(False)
            e.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
        def work(e):
            for _ in range(R):
                e.reads_to_images(dev.data_ptr(), p, t, on_device=True, n_bytes=total, max_levels=9)
        th = [threading.Thread(target=work, args=(e,)) for e in engs]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for x in th: x.start()
        for x in th: x.join()
        wall = (time.perf_counter() - t0) / (R * T) * 1e3
        print(f"   {T} contexts: {wall*1e3:8.1f} us per sample  {n/(wall*1e-3)/1e9:7.1f} Gbases/s")
        for e in engs: e.close()
    eng.close()
