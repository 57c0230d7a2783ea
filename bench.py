#!/usr/bin/env python
"""bench.py -- Gbases/s, cleaned reads -> varKode/CGR images (k=7), on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference-equivalent CPU path (oracle port)

Workload (BASELINE.json configs[1]): synthetic 200 Mbp cleaned FASTQ samples (read length 150, SURVEY.md
section 8d shape), k=7, CGR mapping, full sub-sample ladder 200M..500K (9 levels) from ONE pass.
A step = the whole hot path over one sample: FASTQ framing -> ladder -> seeded sub-sampling -> k-mer
counting of all levels -> canonical fold -> 9 images (uint8) read back to the host.

value   device-resident: the texts are already in HBM.  EXACTLY K steps are dealt to `--in-flight` contexts
        (default 4, one host thread each: the path of one sample is a chain of short dependent kernels, several
        samples in flight fill its gaps -- how stages.images_for_samples runs a batch).  The region is bracketed
        by barrier + device synchronize on both sides and timed by two CUDA events recorded at those idle points,
        so every gap between steps is inside it.  Every context alternates two samples of its own (423 MB each,
        larger than the 126 MB L2; no two contexts read the same bytes).  `one_context` reports a single
        context's per-sample device interval and wall time beside it.
e2e     same call with the text in pinned HOST memory: H2D copy + kernels + read-back, wall clock, two contexts.
N > 1   one process per GPU (torchrun), each with its own samples (sharded by sample, no data-path
        collective): value = N * K * bases / max-over-ranks time; scaling "weak".
Side measurements: --workload c3 (configs[2]: 1 Gbp, k = 9) and --workload c5 (configs[4]: ONE 30 Gbp sample
read-sharded over the ranks, one NCCL all-reduce of the histograms; scaling "strong").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BASES = 200_000_000
READ_LEN = 150
K = 7
MAPPING = "cgr"
MIN_BP, MAX_BP = 500_000, 200_000_000
BYTES_PER_BASE = (2 * READ_LEN + 17) / READ_LEN          # 2.1133: FASTQ text that must be streamed once
LEVELS = [200_000_000, 100_000_000, 50_000_000, 20_000_000, 10_000_000, 5_000_000, 2_000_000, 1_000_000, 500_000]
METRIC = "Gbases/s reads->varKode images (k=7)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def hbm_probe(torch, nbytes=1 << 30, reps=6):
    """what THIS box's HBM does on a plain device-to-device copy (read + write bytes, best of `reps`, CUDA events): the
    boxes of the pool differ by several per cent, and the framing kernels of this path are HBM-bound, so the number is
    recorded next to the official denominator (MEASURED_PEAKS.json, measured by the driver the same way)."""
    try:
        a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        a.zero_()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.copy_(a)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        del a, b
        torch.cuda.empty_cache()
        return 2 * nbytes / (best * 1e-3) / 1e9
    except Exception:
        return None


class ClockSampler:
    """SM clock and clock-event reasons sampled DURING the timed region: NVML polled from a thread every ~2 ms
    (a step lasts ~0.5 ms, so nvidia-smi's 100 ms loop would miss short runs); nvidia-smi -lms is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max, self.nv, self.h, self.stop_flag, self.thread = None, None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self.h = h
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                 (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                 (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                 (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names:
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 2 ms poll",
                    "power_w_max": max(self.power) if self.power else None}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_path_once(buf, threads):
    """reference-equivalent CPU path on host cores: framing + ladder + seeded sub-sampling + dsk-restated counts for
    every level (C, OpenMP) + the exact make_image arithmetic (numpy/Python) -> pixels.  Returns seconds."""
    from oracle import dsk, image as oimg
    from varkoder_b200.mapping import get_kmer_mapping
    lut = get_kmer_mapping(K, MAPPING).lut
    t0 = time.perf_counter()
    p = dsk.parse_fastq(buf)
    nsites = p["nsites_ref"]
    levels = oimg.ladder(nsites, MIN_BP_CPU, MAX_BP)
    # level thresholds fitted to the base targets (what reformat.sh samplebasestarget amounts to), as the GPU path does
    thr, take_all = dsk.level_thresholds(levels, nsites, 1, lens=p["lens"])
    _, canon = dsk.count_levels(buf, K, 1, thr, take_all, threads=threads)
    imgs = [oimg.image_exact(c, lut) for c in canon]
    dt = time.perf_counter() - t0
    assert len(imgs) == len(levels)
    return dt, nsites, len(levels)


MIN_BP_CPU = MIN_BP


def cpu_sample(n_bases):
    from varkoder_b200 import synth
    return synth.fixed(n_bases, READ_LEN, seed=20260118 + 2000)


def best_thread_count(buf):
    """the host may expose more hardware threads than it really schedules (shared boxes): time the counting
    leg with 1 and with all threads once and keep whichever is faster -- the baseline gets its best case."""
    allt = len(os.sched_getaffinity(0))
    best_t, best_dt = 1, None
    for th in sorted({1, allt}):
        dt, _, _ = cpu_path_once(buf, th)
        if best_dt is None or dt < best_dt:
            best_t, best_dt = th, dt
    return best_t, best_dt


SAMPLE_SIZES = (200_000_000, 100_000_000, 50_000_000, 20_000_000, 10_000_000, 5_000_000)


def cpu_sample_desc(sample_bases, nl, secs):
    return (f"{sample_bases} bases of the workload's synthetic shape (L=150), k=7 cgr, {nl}-level ladder "
            f"{sample_bases}..{MIN_BP}; oracle port of the reference CPU path (framing + seeded sub-sampling + dsk restated "
            f"in C/OpenMP + exact make_image arithmetic), uncompressed FASTQ bytes in host memory, {secs:.2f} s per pass; "
            "the reference's own dsk / dsk2ascii / reformat.sh binaries cannot be installed offline")


def run_cpu_baseline(host_bytes, budget_s=20.0):
    """GPU arm, rank 0, N=1: the CPU path on the SAME 200 Mbp bytes the GPU just processed, repeated for about
    budget_s seconds (at least one pass, at most five)."""
    probe = host_bytes[:synth_total(5_000_000)]
    threads, _ = best_thread_count(probe)
    times = []
    nl = 0
    while not times or (sum(times) < budget_s and len(times) < 5):
        dt, nsites, nl = cpu_path_once(host_bytes, threads)
        times.append(dt)
    best = min(times)
    return {"value": N_BASES / best / 1e9, "unit": "Gbases/s", "cores": threads, "kind": "port",
            "sample": cpu_sample_desc(N_BASES, nl, best) + f"; best of {len(times)} passes"}


def synth_total(n_bases):
    from varkoder_b200 import synth
    return synth.fixed_total_bytes(n_bases, READ_LEN)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size each step's sample so that the whole run ends within a few minutes whatever K is
    probe = cpu_sample(5_000_000)
    threads, probe_dt = best_thread_count(probe)
    speed = 5_000_000 / probe_dt
    steps = max(1, args.steps)
    warm = min(max(args.warmup, 0), 1)
    sample_bases = SAMPLE_SIZES[-1]
    for sz in SAMPLE_SIZES:
        if (steps + warm) * sz / speed <= 100.0 and sz * 0.11e-6 <= 60.0:       # second term: host generator time
            sample_bases = sz
            break
    buf = probe if sample_bases == 5_000_000 else cpu_sample(sample_bases)
    for _ in range(warm):
        cpu_path_once(buf, threads)
    times = []
    nl = 0
    for _ in range(steps):
        dt, nsites, nl = cpu_path_once(buf, threads)
        times.append(dt)
    total = sum(times)
    val = sample_bases * len(times) / total / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gbases/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": threads, "kind": "port",
                         "sample": "each step = " + cpu_sample_desc(sample_bases, nl, total / len(times))},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config():
    if K == 9:
        return {"workload": "configs[2]: 1 Gbp synthetic samples, one per step (the GPU arm keeps `in_flight` of them in flight per GPU), read length 150, k=9, varKode mapping, "
                            "-M 0, ladder 1G..500K (11 levels) from one pass; canonical classes counted in shared memory by CTA pairs",
                "bases_per_step_per_gpu": N_BASES, "bytes_per_base": round(BYTES_PER_BASE, 4), "k": K, "mapping": MAPPING,
                "levels": len(LEVELS), "l2_policy": "input (2.1 GB) larger than L2 (126 MB); no flush needed",
                "parallelism": "by-sample, one process per GPU, no collective"}
    return {"workload": "configs[1]: 200 Mbp synthetic samples, one per step (the GPU arm keeps `in_flight` of them in flight per GPU), read length 150, k=7, cgr mapping, "
                        "full subsample ladder 200M..500K (9 levels) from one pass",
            "bases_per_step_per_gpu": N_BASES, "bytes_per_base": round(BYTES_PER_BASE, 4), "k": K, "mapping": MAPPING,
            "levels": len(LEVELS), "l2_policy": "input (423 MB per sample) larger than L2 (126 MB); every context alternates two samples of its own and no two contexts read the same bytes",
            "parallelism": "by-sample, one process per GPU, no collective"}


def c4_leg(args, world, rank, local, torch, dist, reps=None, in_flight=None):
    """BASELINE configs[3]: a batch of 96 Bembidion-shaped samples (10-50 Mbp each, read lengths 60..280, 1 % of the reads
    shorter than k incl. empty ones; k = 7, varKode, -m 500K, every ladder from the sample's own nsites), generated on
    the devices, dealt to the ranks by size (sharding.assign_samples, longest first); every rank pushes its share through
    `in_flight` contexts that take samples off a common queue.  Device-resident texts, wall clock between barriers, max
    over ranks.  Returns the result dict on rank 0 (None elsewhere)."""
    import threading as _th
    import time as _t
    import numpy as np
    from varkoder_b200 import sharding
    from varkoder_b200.engine import Engine, Params
    from varkoder_b200.mapping import get_kmer_mapping
    rng = np.random.default_rng(20260118 + 4000)
    targets = [int(x) for x in rng.integers(10_000_000, 50_000_001, 96)]
    n_reads = [t * 100 // 16833 for t in targets]              # mean read length 168.33
    owner, loads = sharding.assign_samples(targets, world)
    mine = sorted((i for i in range(len(targets)) if owner[i] == rank), key=lambda i: -targets[i])
    T = in_flight or (1 if args.no_concurrent else max(1, args.in_flight))
    engs = [Engine(local) for _ in range(T)]
    for e in engs:
        e.set_batch_mode(T > 1)                # as stages.images_for_samples does for its worker contexts
    table = get_kmer_mapping(7, "varKode")
    texts, sizes = {}, {}
    first = [0]
    for n in n_reads:
        first.append(first[-1] + n)
    for i in mine:
        nb, bases = engs[0].synth_fastq_variable(None, 0, n_reads[i], seed=20260118 + 4000 + i, first_read=first[i])
        d = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
        engs[0].synth_fastq_variable(d.data_ptr(), d.numel(), n_reads[i], seed=20260118 + 4000 + i, first_read=first[i])
        texts[i] = (d, nb)
        sizes[i] = bases

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lock = _th.Lock()
    errors, levels_seen = [], []

    def run_all():
        queue = list(mine)

        def work(t):
            try:
                while True:
                    with lock:
                        if not queue:
                            return
                        i = queue.pop(0)
                    d, nb = texts[i]
                    r = engs[t].reads_to_images(d.data_ptr(), Params(k=7, min_bp=MIN_BP, max_bp=None, seed=100 + i), table,
                                                on_device=True, n_bytes=nb, max_levels=16)
                    assert r.nsites == sizes[i] and r.status == 0
                    levels_seen.append(len(r.levels))
            except Exception as exc:
                errors.append(exc)
        th = [_th.Thread(target=work, args=(t,)) for t in range(T)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize()
        if errors:
            raise errors[0]

    run_all()                                            # warm-up: every buffer of every context at its final size
    run_all()
    reps = reps or max(1, min(args.steps, 5))
    barrier()
    t0 = _t.perf_counter()
    for _ in range(reps):
        run_all()
    barrier()
    ms = 1e3 * (_t.perf_counter() - t0) / reps
    stats = torch.tensor([ms, float(sum(sizes.values())), float(min(levels_seen)), -float(max(levels_seen))],
                         dtype=torch.float64, device="cuda")
    tot = stats.clone()
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(stats.cpu()[0])
    total = int(tot.cpu()[1])
    per_rank = [0.0] * world
    if world > 1:
        g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([float(sum(sizes.values()))], dtype=torch.float64, device="cuda"))
        per_rank = [float(x.cpu()[0]) for x in g]
    else:
        per_rank = [float(sum(sizes.values()))]
    for e in engs:
        e.close()
    del texts
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"value": total / (ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_batch": ms, "reps": reps, "samples": len(targets),
            "bases_per_batch": total, "in_flight": T, "n_gpus": world,
            "workload": f"configs[3]: one batch of 96 samples of 10-50 Mbp ({total} bases in all; read lengths 60..280, 1 % of "
                        f"the reads shorter than k), k=7, varKode, -m 500K, every ladder from the sample's own nsites, dealt to "
                        f"{world} GPU(s) by size, {T} samples in flight per GPU; device-resident texts",
            "per_rank_bases": per_rank, "balance_max_over_mean": max(per_rank) / (sum(per_rank) / len(per_rank))}


def run_c4(args, world, rank, local, torch, dist):
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = c4_leg(args, world, rank, local, torch, dist)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": r["value"], "unit": "Gbases/s", "n_gpus": world, "steps": r["reps"],
            "warmup": 2, "ms_per_step": r["ms_per_batch"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": r["workload"] + "; a step = the whole batch", "bases_per_step": r["bases_per_batch"],
                       "samples": r["samples"], "in_flight": r["in_flight"],
                       "l2_policy": "every sample is read once per step; the batch of a rank (0.8-6 GB) is larger than L2"},
            "per_rank_bases": r["per_rank_bases"], "balance_max_over_mean": r["balance_max_over_mean"], "clocks": clocks}), flush=True)


def run_c5(args, world, rank, local, torch, dist):
    parity = check_sharded_parity_fresh(local, world, rank, torch, dist) if world > 1 else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = c5_leg(args, world, rank, local, torch, dist, steps=max(1, min(args.steps, 20)))
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": r["value"], "unit": "Gbases/s", "n_gpus": world, "steps": r["steps"], "warmup": 3,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": r["workload"], "bases_per_step": r["bases"], "levels": r["levels"],
                       "l2_policy": "shards are far larger than L2"},
            "level_bases": r["level_bases"], "sharded_parity": parity, "clocks": clocks}), flush=True)


def check_sharded_parity_fresh(local, world, rank, torch, dist):
    from varkoder_b200.engine import Engine
    from varkoder_b200.mapping import get_kmer_mapping
    eng = Engine(local)
    try:
        return check_sharded_parity(eng, get_kmer_mapping(K, MAPPING), world, rank, torch, dist)
    finally:
        if getattr(eng, "comm_world", None):
            eng.comm_destroy()
        eng.close()


def check_sharded_parity(eng, table, world, rank, torch, dist, n_bases=5_000_000):
    """Before anything read-sharded is timed: one ~5 Mbp sample cut at record boundaries into `world` shards, counted
    by all ranks with the real engine and ONE NCCL all-reduce, must give -- on every rank -- the counts and pixels of
    (a) the same bytes pushed through the unsharded path on that rank's GPU and (b), on rank 0, the CPU oracle (used
    here as the checker only).  Any difference aborts the bench."""
    import numpy as np
    from varkoder_b200 import sharding, synth
    from varkoder_b200.engine import Params
    from varkoder_b200.ladder import parse_seed
    buf = synth.fixed(n_bases, READ_LEN, seed=20260118 + 7000)          # the same bytes on every rank
    parts = sharding.split_records(buf, world)
    b, e, _ = parts[rank]
    sp = Params(k=K, min_bp=100_000, max_bp=None, seed=1234)
    rs = sharding.fused_sharded_reads_to_images(eng, buf[b:e], sp, table, want_canon=True)       # the form that is timed
    r_staged = sharding.sharded_reads_to_images(eng, buf[b:e], sp, table, want_canon=True)        # torch.distributed form
    whole = eng.reads_to_images(buf, sp, table, want_canon=True)
    ok = len(rs.levels) >= 5
    for r in (rs, r_staged):
        ok = ok and (r.levels == whole.levels and r.level_bases == whole.level_bases and r.level_reads == whole.level_reads
                     and r.n_reads == whole.n_reads and bool((r.canon == whole.canon).all())
                     and bool((r.pixels == whole.pixels).all()))
    what = "sharded (library-issued NCCL on the engine's stream; and the staged torch.distributed form) == unsharded on every rank"
    if rank == 0:
        from oracle import dsk, image as oimg                            # checker only
        thr, take_all = dsk.level_thresholds(rs.levels, n_bases, parse_seed(1234), lens=dsk.parse_fastq(buf)["lens"])
        _, expect = dsk.count_levels(buf, K, parse_seed(1234), thr, take_all, threads=0)
        ok = ok and bool((rs.canon == expect).all()) and all(
            bool((rs.pixels[l] == oimg.image_exact(expect[l], table.lut)).all()) for l in range(len(rs.levels)))
        what += " == CPU oracle (rank 0)"
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.cpu()[0]) != 1:
        raise SystemExit(f"rank {rank}: read-sharded result differs from the unsharded path / the oracle -- bench aborted")
    return {"status": "ok", "checked": what, "bases": n_bases, "levels": len(rs.levels), "ranks": world,
            "compared": "canonical counts of every level (uint64, bit-exact) and every pixel"}


def e2e_gz_leg(eng, local, torch, n_big=200_000_000, n_small=25_000_000, n_files=16):
    """The reference's real input format, end to end on this box: pigz-written ``.fq.gz`` files on local disk (what
    clean_reads leaves, image.py:529-540) -> stages.images_for_samples -> PNG files on disk.  Two cases: ONE 200 Mbp
    sample (configs[1]; the spare inflate threads split the single gzip member at pigz's sync points) and a batch of
    `n_files` x 25 Mbp.  All host threads the process may use inflate; wall clock, best of three; rank 0's GPU only --
    from files the path is bound by DEFLATE decoding on the host cores, not by the GPU."""
    import shutil
    import tempfile
    from varkoder_b200 import feed, stages
    from varkoder_b200.mapping import get_kmer_mapping
    threads = len(os.sched_getaffinity(0))
    tmp = tempfile.mkdtemp(prefix="vk_e2e_gz_")
    out = {"inflate_threads": threads, "gzip": "pigz-style single member, 128 KiB blocks, zlib level 6",
           "decoder": "libvk_feed.so (own DEFLATE)" if feed.feed_lib() is not None else "zlib"}
    try:
        table = get_kmer_mapping(K, MAPPING)

        def make(name, n_bases, first_read):
            total = synth_total(n_bases)
            d = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
            assert eng.synth_fastq(d.data_ptr(), d.numel(), n_bases, READ_LEN, seed=20260118 + 6000, first_read=first_read) == total
            raw = d[:total].cpu().numpy()
            del d
            comp = feed.pigz_compress(raw, 6, threads)
            p = os.path.join(tmp, name + ".fq.gz")
            with open(p, "wb") as f:
                f.write(comp)
            return p, total, len(comp)

        def timed(samples, n_bases_total, text_bytes, gz_bytes, workers, min_bp, max_bp):
            times = []
            for rep in range(3):
                o = os.path.join(tmp, "images")
                shutil.rmtree(o, ignore_errors=True)
                t0 = time.perf_counter()
                st = stages.images_for_samples(samples, o, table, k=K, mapping_code=MAPPING, min_bp=min_bp, max_bp=max_bp,
                                               threads=threads, gpu_workers=workers, device=local)
                times.append(time.perf_counter() - t0)
                assert len(st) == len(samples) and all("failed_step" not in v for v in st.values())
            n_png = sum(len(fs) for _, _, fs in os.walk(os.path.join(tmp, "images")))
            best = min(times)
            return {"seconds": best, "all_runs_s": [round(x, 3) for x in times], "value": n_bases_total / best / 1e9,
                    "unit": "Gbases/s", "text_gb_per_s": text_bytes / best / 1e9, "gz_mb": round(gz_bytes / 1e6, 1),
                    "text_mb": round(text_bytes / 1e6, 1), "png_files": n_png, "gpu_workers": workers}
        p, total, gz = make("BIG", n_big, 0)
        out["one_sample"] = dict(timed([dict(sample="BIG", path=p, labels=["x"], base_sd=0.0)], n_big, total, gz, 1,
                                       MIN_BP, 200_000_000), bases=n_big, levels=len(LEVELS))
        os.remove(p)
        samples, tot_text, tot_gz = [], 0, 0
        for i in range(n_files):
            p, total, gz = make(f"S{i:02d}", n_small, (i + 1) * 2_000_000)
            samples.append(dict(sample=f"S{i:02d}", path=p, labels=["x"], base_sd=0.0))
            tot_text += total
            tot_gz += gz
        out["batch"] = dict(timed(samples, n_small * n_files, tot_text, tot_gz, 2, MIN_BP, 200_000_000),
                            files=n_files, bases_per_file=n_small)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def c5_leg(args, world, rank, local, torch, dist, total_bases=None, steps=8):
    """BASELINE configs[4]: ONE sample of 30 Gbp (k = 7, CGR, -M 0: 16 levels), read-sharded over the ranks: per step and
    rank one enqueue on the library's stream (framing, ncclAllGather of the shard sizes, ladder, count, ONE ncclAllReduce
    of the histograms, images), one host synchronisation.  Shards generated on the devices.  Wall clock between
    barriers, max over ranks.  Result dict on rank 0."""
    import time as _t
    from varkoder_b200 import sharding, synth
    from varkoder_b200.engine import Engine, Params
    from varkoder_b200.mapping import get_kmer_mapping
    total_bases = total_bases or args.total_bases
    total_reads = (total_bases + READ_LEN - 1) // READ_LEN
    first = total_reads * rank // world
    last = total_reads * (rank + 1) // world
    shard_bases = min((last - first) * READ_LEN, total_bases - first * READ_LEN)
    eng = Engine(local)
    table = get_kmer_mapping(7, "cgr")
    nbytes = synth.fixed_total_bytes(shard_bases, READ_LEN)
    dev = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    assert eng.synth_fastq(dev.data_ptr(), dev.numel(), shard_bases, READ_LEN, seed=20260118 + 5000, first_read=first) == nbytes
    sp = Params(k=7, min_bp=MIN_BP, max_bp=None, seed=11)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if world > 1:
            return sharding.fused_sharded_reads_to_images(eng, dev.data_ptr(), sp, table, on_device=True, n_bytes=nbytes,
                                                          max_levels=18)
        return eng.reads_to_images(dev.data_ptr(), sp, table, on_device=True, n_bytes=nbytes, max_levels=18)

    for _ in range(3):
        rs = step()
    assert rs.levels[0] == total_bases and rs.level_bases[0] == total_bases and rs.n_reads == total_reads
    barrier()
    t0 = _t.perf_counter()
    for _ in range(steps):
        rs = step()
    barrier()
    ms = 1e3 * (_t.perf_counter() - t0) / steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.cpu()[0])
    fallbacks = eng.count_fallbacks()
    split = eng.timings()
    if world > 1:
        eng.comm_destroy()
    eng.close()
    del dev
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"value": total_bases / (ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": ms, "steps": steps, "n_gpus": world,
            "bases": total_bases, "levels": len(rs.levels), "level_bases": rs.level_bases,
            "count_fallbacks": fallbacks, "last_step_timings": split,
            "workload": f"configs[4]: ONE sample of {total_bases} bases (read length 150), k=7, cgr, -M 0 ({len(rs.levels)} levels), "
                        f"read-sharded over {world} GPU(s): {shard_bases} bases ({nbytes / 1e9:.1f} GB of text) resident per GPU"
                        + ("; exchange = ncclAllGather(2 x u64) + ONE ncclAllReduce(u64 x %d), issued by the library on its own "
                           "stream between its kernels" % (18 * 4 ** 7 + 130) if world > 1 else "")}


# ----------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the c4 / c5 / e2e_gz side measurements")
    ap.add_argument("--in-flight", type=int, default=4,
                    help="samples in flight per GPU in the timed region: one context + host thread each (1 = one stream)")
    ap.add_argument("--no-concurrent", action="store_true", help="same as --in-flight 1")
    ap.add_argument("--bases", type=int, default=None, help="debug: smaller sample (invalidates the bench line)")
    ap.add_argument("--total-bases", type=int, default=30_000_000_000, help="c5: bases of the one read-sharded sample")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = BASELINE configs[1] (the bench line); c3 = configs[2]: 1 Gbp, k=9 varKode, -M 0; c5 = configs[4]: ONE "
                         "30 Gbp sample, k=7, read-sharded over the ranks with one NCCL all-reduce; c4 = configs[3]: 96 samples of "
                         "10-50 Mbp dealt to the ranks by size (side measurements)")
    args = ap.parse_args()
    global N_BASES, K, MAPPING, MAX_BP, LEVELS
    if args.workload == "c3":
        N_BASES, K, MAPPING, MAX_BP = 1_000_000_000, 9, "varKode", None
        LEVELS = [1_000_000_000, 500_000_000, 200_000_000, 100_000_000, 50_000_000, 20_000_000, 10_000_000, 5_000_000,
                  2_000_000, 1_000_000, 500_000]
    if args.bases is None:
        args.bases = N_BASES
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from varkoder_b200 import synth
    from varkoder_b200.engine import Engine, Params
    from varkoder_b200.mapping import get_kmer_mapping

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    W = max(3, args.warmup)
    n_bases = args.bases
    if args.workload == "c4":
        run_c4(args, world, rank, local, torch, dist)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "c5":
        run_c5(args, world, rank, local, torch, dist)
        if world > 1:
            dist.destroy_process_group()
        return

    import threading
    T = 1 if args.no_concurrent else max(1, args.in_flight)
    engs = [Engine(local) for _ in range(T)]
    eng = engs[0]
    table = get_kmer_mapping(K, MAPPING)
    params = Params(k=K, min_bp=MIN_BP, max_bp=MAX_BP, seed=1 + rank)
    total = synth.fixed_total_bytes(n_bases, READ_LEN)
    # Every context owns TWO different samples and alternates them: whatever one step leaves in the 126 MB L2 (the tail
    # of a 423 MB text) is of no use to the next step of that context, and no two contexts ever read the same bytes.
    n_reads_sample = (n_bases + READ_LEN - 1) // READ_LEN
    samples = []
    for t in range(T):
        pair = []
        for j in range(2):
            d = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
            first_read = ((rank * T + t) * 2 + j) * n_reads_sample            # distinct reads everywhere
            assert eng.synth_fastq(d.data_ptr(), d.numel(), n_bases, READ_LEN, seed=20260118 + 2000, first_read=first_read) == total
            pair.append(d)
        samples.append(pair)
    devs = samples[0]
    dev = devs[0]
    step_no = [0]

    def run_step(t, i):
        return engs[t].reads_to_images(samples[t][i & 1].data_ptr(), params, table, on_device=True, n_bytes=total,
                                       max_levels=len(LEVELS))

    def step_device():
        r = run_step(0, step_no[0])
        step_no[0] += 1
        return r, eng.timings()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(T):
        for i in range(W):
            res = run_step(t, i)
        assert res.nsites == n_bases and (n_bases != N_BASES or res.levels == LEVELS)
        assert res.pixels.shape == (len(res.levels), table.side, table.side) and int(res.pixels.max()) == 255

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- timed region: EXACTLY K steps (one step = one sample through the whole path, device-resident text), dealt to
    # the T contexts; a barrier + device synchronize on both sides; elapsed time from two CUDA events recorded at those
    # two idle points (so it includes every gap between the steps), and the wall clock next to it.  Inside a step only
    # its first and last event are recorded (events between kernels would serialise the stream and defeat the
    # dependent-launch overlap the library uses).
    for e in engs:
        e.set_fine_timing(False)
    for t in range(T):
        for i in range(2):
            run_step(t, i)
    share = [args.steps // T + (1 if t < args.steps % T else 0) for t in range(T)]
    errors = []

    def work(t):
        try:
            for i in range(share[t]):
                run_step(t, i)
        except Exception as exc:          # surfaced after the join
            errors.append(exc)

    launches0 = sum(e.launch_count() for e in engs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    threads_ = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    barrier()
    ev0.record()
    t_wall0 = time.perf_counter()
    if T == 1:
        work(0)
    else:
        for th in threads_:
            th.start()
        for th in threads_:
            th.join()
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    dev_ms = ev0.elapsed_time(ev1)
    barrier()
    if errors:
        raise errors[0]
    launches = sum(e.launch_count() for e in engs) - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- one context alone (sample latency): sum of the per-step device intervals (first to last CUDA event of a step
    # on the library's stream) and the wall clock of the same loop, which adds the host gaps between dependent steps
    one_steps = max(3, min(args.steps, 200))
    one_dev = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(one_steps):
        _, tm = step_device()
        one_dev += tm["total"]
    torch.cuda.synchronize()
    one_wall = 1e3 * (time.perf_counter() - t0)
    one_stream = {"steps": one_steps, "ms_per_step_device": one_dev / one_steps, "ms_per_step_wall": one_wall / one_steps,
                  "value_device": n_bases * one_steps / (one_dev * 1e-3) / 1e9, "unit": "Gbases/s",
                  "note": "one context, one sample at a time: device interval of a step (its first to last CUDA event) "
                          "and wall clock per step including the host gap between two synchronous calls; rank-local"}

    # ---- per-kernel split of a step (CUDA events between the kernel groups, on the library's stream): gives the
    # count kernel's own duration for the roofline.  Not part of `value`.
    eng.set_fine_timing(True)
    split_steps = max(3, min(args.steps, 200))
    for _ in range(2):
        step_device()
    per_kernel = {}
    for _ in range(split_steps):
        _, tm = step_device()
        for kname, v in tm.items():
            per_kernel[kname] = per_kernel.get(kname, 0.0) + v
    per_kernel = {k2: v / split_steps for k2, v in per_kernel.items()}
    torch.cuda.synchronize()

    # ---- end to end: pinned host text -> vk_reads_to_images (H2D copy, whole path, D2H of the images) -> pixels on the
    # host, wall clock.  Two contexts on two host threads so that one sample's upload overlaps the other's kernels; the
    # PCIe link is the limit either way.
    Te = min(T, 2)
    hosts = []
    for t in range(Te):
        h = torch.empty(total, dtype=torch.uint8).pin_memory()
        h.copy_(samples[t][0][:total])
        hosts.append(h)
    host = hosts[0]
    host_keep = None
    torch.cuda.synchronize()
    r2s = [None] * Te
    for t in range(Te):
        for _ in range(2):
            r2s[t] = engs[t].reads_to_images(hosts[t], params, table, max_levels=len(LEVELS))
    e_share = [args.e2e_steps // Te + (1 if t < args.e2e_steps % Te else 0) for t in range(Te)]

    def e2e_work(t):
        try:
            for _ in range(e_share[t]):
                r2s[t] = engs[t].reads_to_images(hosts[t], params, table, max_levels=len(LEVELS))
        except Exception as exc:
            errors.append(exc)

    e_threads = [threading.Thread(target=e2e_work, args=(t,)) for t in range(Te)]
    barrier()
    t0 = time.perf_counter()
    if Te == 1:
        e2e_work(0)
    else:
        for th in e_threads:
            th.start()
        for th in e_threads:
            th.join()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if errors:
        raise errors[0]
    r2 = r2s[0]
    res0 = eng.reads_to_images(devs[0].data_ptr(), params, table, on_device=True, n_bytes=total, max_levels=len(LEVELS))
    assert (r2.pixels == res0.pixels).all()                # host-buffer path == device-resident path, same sample
    d2h = int(res.pixels.size) + 4096

    # ---- N > 1 only: ONE sample of N x 200 Mbp read-sharded over the ranks (BASELINE configs[4] shape): every rank
    # frames and counts its shard, one NCCL all-reduce sums the per-segment histograms, every rank renders.
    sharded = None
    sharded_parity = None
    if world > 1:
        from varkoder_b200 import sharding
        sharded_parity = check_sharded_parity(eng, table, world, rank, torch, dist)     # aborts the bench on a mismatch
        sp = Params(k=K, min_bp=MIN_BP, max_bp=None, seed=7)
        ML = 16

        def sharded_step():
            return sharding.fused_sharded_reads_to_images(eng, dev.data_ptr(), sp, table, on_device=True, n_bytes=total,
                                                          max_levels=ML)
        for _ in range(3):
            rs = sharded_step()
        assert rs.n_reads == world * n_reads_sample and rs.levels[0] == world * n_bases and rs.level_bases[0] == world * n_bases
        s_steps = max(1, min(args.steps, 100))
        barrier()
        t0 = time.perf_counter()
        for _ in range(s_steps):
            rs = sharded_step()
        barrier()
        sh_ms = 1e3 * (time.perf_counter() - t0) / s_steps
        sharded = {"workload": f"one sample of {world}x{n_bases} bases read-sharded over {world} GPUs, k={K} {MAPPING}, "
                               f"{len(rs.levels)} levels; per step and rank ONE enqueue: framing, ncclAllGather(2 x u64), ladder, "
                               f"count, ONE ncclAllReduce(u64 x {ML * 4 ** K + 130}) on the library's stream, images; one host sync",
                   "steps": s_steps, "ms_per_step_wall": sh_ms, "value": world * n_bases / (sh_ms * 1e-3) / 1e9,
                   "unit": "Gbases/s", "level_bases": rs.level_bases}

    t = torch.tensor([dev_ms, wall_ms, e2e_ms, per_kernel["count"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, count_ms = [float(x) for x in t.cpu()]

    probe_gbs = hbm_probe(torch) if rank == 0 else None
    # ---- side legs (recorded in the same line; none of them feeds `value`): the other BASELINE configurations and the
    # reference's real input format.  Each frees what the main leg holds first; a leg that fails is reported, not fatal.
    side = {}
    if not args.no_side_legs and args.bases == N_BASES and args.workload == "c2":
        if world > 1 and getattr(eng, "comm_world", None):
            eng.comm_destroy()
        for e in engs[1:]:
            e.close()
        del samples, devs, dev, hosts, host_keep
        torch.cuda.empty_cache()
        for name, fn in (("c4", lambda: c4_leg(args, world, rank, local, torch, dist, reps=3,
                                             # small samples: eight contexts per GPU while the host has a core for each worker thread
                                             in_flight=max(T, min(8, max(4, len(os.sched_getaffinity(0)) // max(world, 1)))))),
                         ("c5", lambda: c5_leg(args, world, rank, local, torch, dist, steps=6)),
                         ("e2e_gz", lambda: e2e_gz_leg(eng, local, torch) if rank == 0 else None)):
            try:
                r_leg = fn()
            except Exception as exc:                       # noqa: BLE001 -- recorded in the line
                r_leg = {"error": f"{type(exc).__name__}: {exc}"}
            if world > 1:
                dist.barrier()
            if rank == 0 and r_leg is not None:
                side[name] = r_leg
    if rank == 0:
        peak, peak_src = measured_peaks()
        steps = args.steps
        value = world * n_bases * steps / (dev_ms * 1e-3) / 1e9
        count_s = count_ms * 1e-3
        achieved = n_bases * BYTES_PER_BASE / count_s / 1e9
        # k = 7, reads of one length (this workload): countt_kernel counts, count_kernel<7> is launched behind it and
        # returns at once (the choice is made on the device, vk_countt.cuh); VK_COUNT_LANES=0 keeps the flat-lane kernel
        lanes_env = os.environ.get("VK_COUNT_LANES", "-1")
        k7_name = "count_kernel<7,smem>" if lanes_env == "0" else ("countt_kernel<16>" if lanes_env in ("-1", "2") else f"count kernel VK_COUNT_LANES={lanes_env}")
        # k = 9, reads of one length: countt9_kernel (VK_COUNT_LANES9=0 keeps count9h_kernel)
        k9_name = "count9h_kernel" if os.environ.get("VK_COUNT_LANES9", "-1") == "0" else "countt9_kernel"
        kernel_name = (k7_name if K == 7 else f"count_kernel<{K},smem>") if K <= 7 else (k9_name if K == 9 else f"count16_kernel<{K}>")
        traffic = traffic_src = None
        try:
            with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
                ent = json.load(f).get(kernel_name)
            if ent and n_bases == N_BASES:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": steps, "warmup": W,
            "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": workload_config(),
            "e2e": {"value": world * n_bases * args.e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
                    "h2d_bytes_per_step": total, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "ms_per_step": e2e_ms / args.e2e_steps,
                    "contexts": Te,
                    "note": "uncompressed FASTQ in pinned host memory -> vk_reads_to_images -> pixels on host; wall clock; "
                            "two contexts so that uploads overlap kernels"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n_bases * BYTES_PER_BASE,
                         "hbm_copy_probe_this_box_gbs": probe_gbs,
                         "kernel_ms": count_s * 1e3,
                         "whole_step_frac": (n_bases * BYTES_PER_BASE / (dev_ms * 1e-3 / steps) / 1e9) / peak},
            "ms_per_step_wall": wall_ms / steps,
            "in_flight": T,
            "kernel_ms_per_step": per_kernel,
            "kernel_split_note": f"mean of {split_steps} separate steps with CUDA events between the kernel groups; "
                                 "those events serialise the stream, so the groups sum to more than ms_per_step",
            "level_bases": res.level_bases,
        }
        if sharded is not None:
            out["read_sharded"] = sharded
            out["sharded_parity"] = sharded_parity
        out["one_context"] = one_stream
        out.update(side)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = run_cpu_baseline(host.numpy())
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
