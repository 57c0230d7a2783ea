"""Host-side mirror of the reference's stage functions for the image hot path.

Same names, argument meaning, file-name contract, stats keys and error behaviour as
``split_fastq`` / ``count_kmers`` / ``make_image`` in the reference (varKoder/commands/image.py:629-936), so
``run_clean2img`` (image.py:938-1128) can call these instead; plus :func:`reads_to_images`, the fused form of
steps C-E that does one upload, one pass over the reads and one synchronisation per sample.

What differs, by design (DESIGN.md "Boundary"):
  * no sub-sampled FASTQ files and no dsk HDF5 are produced.  ``split_fastq`` writes one small JSON
    *level descriptor* per ladder level under the same name stem (``<sample>@NNNNNNNNK.fq.vk``) and
    ``count_kmers`` writes the canonical counts of that level as ``<sample>@NNNNNNNNK+k7.fq.npy``; the
    globs of run_clean2img (image.py:1060, 1092) find them exactly as they find the reference's files.
  * all levels of a sample are counted in ONE pass on the GPU the first time ``count_kmers`` sees the
    sample; the other levels are served from that result.
  * the sub-sample of each level is the seeded nested rule of this project, not BBTools' RNG.
All compute is in the CUDA library; nothing here falls back to a CPU implementation of it.
"""
import hashlib
import json
import os
import sys
import time
from collections import OrderedDict
from pathlib import Path

import numpy as np

from .engine import Engine, Params
from .ladder import (BP_KMER_SEP, LABELS_SEP, QUAL_THRESH, SAMPLE_BP_SEP, LessThanMinimumData, image_name,
                     ladder, level_tag)
from . import quality
from .mapping import as_pixel_table

_ENGINES = {}


def eprint(*args, **kwargs):
    print(*args, file=sys.stderr, **kwargs)


def default_engine(device=None):
    """one Engine per (process, device): CUDA contexts do not survive fork (image.py:1281 uses a fork pool),
    so the engine is created lazily inside whichever process first needs it."""
    if device is None:
        device = int(os.environ.get("VARKODER_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    key = (os.getpid(), device)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device)
    return _ENGINES[key]


_PINNED = {}


def read_clean_fastq(path):
    """bytes of a cleaned FASTQ (``.fq.gz`` as written by clean_reads, image.py:529-540, or plain text) as a uint8
    view of a page-locked buffer that this process reuses from call to call (the view is valid until the next call):
    the H2D copy of a pinned buffer runs at PCIe speed, a pageable one at a fraction of it."""
    from .feed import PinnedBuffer, inflate_into
    buf = _PINNED.get(os.getpid())
    if buf is None:
        buf = _PINNED[os.getpid()] = PinnedBuffer(0, pinned=True)
    # one sample at a time: every host thread may work on this file (a pigz-written member is decoded in pieces)
    n = inflate_into(path, buf, threads=len(os.sched_getaffinity(0)))
    return buf.array[:n]


def _strip_suffixes(p):
    p = Path(p)
    return str(p.name.removesuffix("".join(p.suffixes)))           # image.py:753, 840


def write_png(pixels, path, labels=(), base_sd=0, base_sd_thresh=QUAL_THRESH, mapping_code="varKode"):
    """image.py:920-930: mode L, tEXt keys in the reference's order, optimize=True."""
    from PIL import Image
    from PIL.PngImagePlugin import PngInfo
    img = Image.fromarray(np.ascontiguousarray(pixels, dtype=np.uint8), mode="L")
    meta = PngInfo()
    meta.add_text("varkoderKeywords", LABELS_SEP.join(labels))
    meta.add_text("varkoderBaseFreqSd", str(base_sd))
    meta.add_text("varkoderLowQualityFlag", str(base_sd > base_sd_thresh))
    meta.add_text("varkoderMapping", mapping_code)
    img.save(Path(path), optimize=True, pnginfo=meta)


def _stage_times(tm):
    """(splitting_time, kmer_counting_time) in seconds from Engine.timings().  With the per-kernel events on
    (Engine.set_fine_timing(True)) the stages are the kernel groups; by default a sample is ONE graph launch and only its
    upload and its total are known: splitting_time is the upload, the counting time carries all the GPU work of the step
    (framing, sub-sampling, counting, rendering)."""
    if tm["count"] > 0:
        return (tm["upload"] + tm["parse"] + tm["plan_bucket"]) / 1e3, (tm["count"] + tm["reduce_fold"]) / 1e3
    return tm["upload"] / 1e3, max(tm["total"] - tm["upload"], 0.0) / 1e3


def _image_folder(outfolder, outfile, subfolder_levels):
    outfolder = Path(outfolder)
    if subfolder_levels:
        hsh = list(hashlib.md5(outfile.encode("UTF-8")).hexdigest())      # image.py:850-853
        for _ in range(subfolder_levels):
            outfolder = outfolder / hsh.pop()
    outfolder.mkdir(exist_ok=True, parents=True)
    return outfolder


# --------------------------------------------------------------------------------------------- fused
def reads_to_images(infile, sample, outfolder, kmer_mapping, k=7, mapping_code="varKode", min_bp=50000,
                    max_bp=None, is_query=False, seed=None, labels=(), base_sd=0, base_sd_thresh=QUAL_THRESH,
                    subfolder_levels=0, overwrite=False, verbose=False, engine=None, fastq_bytes=None):
    """Steps C-E of run_clean2img for one sample (image.py:1005-1125) in one GPU pass.

    Returns the stats dict of the three stages merged (same keys as the reference) and writes
    ``<outfolder>[/h/...]/<sample>@NNNNNNNNK+<mapping>+k<k>.png`` for every ladder level.
    Raises ``Exception("Input file has less than minimum data.")`` exactly when split_fastq does.
    """
    eng = engine or default_engine()
    table = as_pixel_table(kmer_mapping, mapping_code)
    if table.k != int(k):
        raise ValueError(f"pixel table is for k={table.k}, asked for k={k}")
    t0 = time.perf_counter()
    data = fastq_bytes if fastq_bytes is not None else read_clean_fastq(infile)
    params = Params(k=int(k), min_bp=int(min_bp), max_bp=None if max_bp is None else int(max_bp),
                    is_query=bool(is_query), seed=seed)
    res = eng.reads_to_images(data, params, table)
    if res.status != 0:
        eprint("Post-cleaning input file " + str(infile) + " has less than " + str(min_bp)
               + "bp, decrease --min_bp if you want to produce an image.")
        raise LessThanMinimumData()
    measured_sd = None
    if base_sd is None:                 # no fastp report: the quality flag from the reads themselves (quality.py)
        measured_sd = base_sd = quality.base_frequency_sd(eng.base_content())
    t1 = time.perf_counter()
    written = []
    for lvl, bp in enumerate(res.levels):
        outfile = image_name(sample, bp, mapping_code, k)
        folder = _image_folder(outfolder, outfile, subfolder_levels)
        if not overwrite and (folder / outfile).is_file():
            eprint("File exists. Skipping image for file:", outfile)
            continue
        write_png(res.pixels[lvl], folder / outfile, labels, base_sd, base_sd_thresh, mapping_code)
        written.append(folder / outfile)
    t2 = time.perf_counter()
    tm = eng.timings()
    stats = OrderedDict()
    split_s, count_s = _stage_times(tm)
    stats["splitting_time"] = split_s
    stats["splitting_bp_per_file"] = ",".join(str(x) for x in res.levels)
    stats[str(k) + "mer_counting_time"] = count_s
    stats["k" + str(k) + "_img_time"] = tm["render"] / 1e3 + (t2 - t1)
    if measured_sd is not None:
        stats["base_frequencies_sd"] = measured_sd          # the key run_clean2img sets (image.py:1096)
    if verbose:
        eprint(f"varkoder_b200: {sample}: {res.nsites} bp, levels {res.levels}, "
               f"read+gpu {t1 - t0:.3f}s, png {t2 - t1:.3f}s")
    return stats


# --------------------------------------------------------------------------------- stage-by-stage mirror
_SAMPLE = {}      # the sample currently resident on the GPU of this process: source path -> state


def _resident(source, eng):
    """the framed text of ``source`` on the GPU.  The engine keeps ONE text; anything else that ran on it since
    (another sample, a fused call, a sharded count) has replaced ours -- its text generation tells -- and the sample is
    uploaded again rather than counted from someone else's reads."""
    st = _SAMPLE.get("state")
    key = (str(source), os.path.getmtime(source), id(eng))
    if st is None or st["key"] != key or st["gen"] != getattr(eng, "text_generation", None):
        data = read_clean_fastq(source)
        eng.upload(data)
        stats = eng.parse()
        st = dict(key=key, stats=stats, counts=st["counts"] if st is not None and st["key"] == key else {},
                  gen=getattr(eng, "text_generation", None))
        _SAMPLE["state"] = st
    return st


def split_fastq(infile, outprefix, outfolder, min_bp=50000, max_bp=None, is_query=False, seed=None,
                overwrite=False, verbose=False, n_threads=1, engine=None):
    """image.py:629-725.  Counts the bases on the GPU, builds the ladder, writes one level descriptor per level."""
    start = time.perf_counter()
    eng = engine or default_engine()
    st = _resident(infile, eng)
    nsites = st["stats"]["nsites"]
    try:
        sites_per_file = ladder(nsites, min_bp, max_bp, is_query)
    except LessThanMinimumData:
        eprint("Post-cleaning input file " + str(infile) + " has less than " + str(min_bp)
               + "bp, decrease --min_bp if you want to produce an image.")
        raise
    outfs = [Path(outfolder) / (outprefix + SAMPLE_BP_SEP + level_tag(bp) + ".fq.vk") for bp in sites_per_file]
    if all(f.is_file() for f in outfs):
        if not overwrite:
            eprint("Files exist. Skipping subsampling for file:", str(infile))
            return OrderedDict()
    Path(outfolder).mkdir(exist_ok=True, parents=True)
    for lvl, (bp, f) in enumerate(zip(sites_per_file, outfs)):
        desc = dict(source=str(infile), level=lvl, level_bp=int(bp), sites_per_file=[int(x) for x in sites_per_file],
                    nsites=int(nsites), min_bp=int(min_bp), max_bp=None if max_bp is None else int(max_bp),
                    is_query=bool(is_query), seed=str(seed) if seed is not None else None)
        with open(f, "w") as fh:
            json.dump(desc, fh)
    stats = OrderedDict()
    stats["splitting_time"] = time.perf_counter() - start
    stats["splitting_bp_per_file"] = ",".join(str(x) for x in sites_per_file)
    return stats


def count_kmers(infile, outfolder, threads=1, k=7, overwrite=False, verbose=False, engine=None):
    """image.py:727-806.  ``infile`` is a level descriptor written by :func:`split_fastq`."""
    start = time.perf_counter()
    outfolder = Path(outfolder)
    outfolder.mkdir(exist_ok=True)
    outpath = outfolder / (_strip_suffixes(infile) + BP_KMER_SEP + "k" + str(k) + ".fq.npy")
    if not overwrite and outpath.is_file():
        eprint("File exists. Skipping kmer counting for file:", str(infile))
        return OrderedDict()
    with open(infile) as fh:
        desc = json.load(fh)
    eng = engine or default_engine()
    st = _resident(desc["source"], eng)
    ckey = (int(k), desc["min_bp"], desc["max_bp"], desc["is_query"], desc["seed"])
    if ckey not in st["counts"]:
        params = Params(k=int(k), min_bp=desc["min_bp"], max_bp=desc["max_bp"], is_query=desc["is_query"],
                        seed=desc["seed"])
        res = eng.count(params)
        res.raise_if_less_than_min()
        canon, _ = eng.render(None, int(k), len(res.levels))
        st["counts"][ckey] = (res.levels, canon)
    levels, canon = st["counts"][ckey]
    if levels != desc["sites_per_file"]:
        raise RuntimeError("level descriptor does not match the ladder of its source file")
    np.save(outpath, canon[desc["level"]])
    stats = OrderedDict()
    stats[str(k) + "mer_counting_time"] = time.perf_counter() - start
    return stats


def make_image(infile, outfolder, kmer_mapping, threads=1, overwrite=False, verbose=False, labels=[], base_sd=0,
               base_sd_thresh=QUAL_THRESH, subfolder_levels=0, mapping_code="varKode", engine=None):
    """image.py:808-936.  ``infile`` holds the canonical counts written by :func:`count_kmers`."""
    in_basename = _strip_suffixes(infile)
    in_base1, in_k = in_basename.split(BP_KMER_SEP)                     # image.py:841 (ValueError if malformed)
    outfile = in_base1 + BP_KMER_SEP + mapping_code + BP_KMER_SEP + in_k + ".png"
    outfolder = _image_folder(outfolder, outfile, subfolder_levels)
    if not overwrite and (outfolder / outfile).is_file():
        eprint("File exists. Skipping image for file:", str(infile))
        return OrderedDict()
    start = time.perf_counter()
    table = as_pixel_table(kmer_mapping, mapping_code)
    canon = np.load(infile)
    if canon.shape != (4 ** table.k,):
        raise IndexError("k-mer counts do not match the k-mer size of the mapping")      # caller catches IndexError
    eng = engine or default_engine()
    pixels = eng.render_counts(table, canon)[0]
    write_png(pixels, outfolder / outfile, labels, base_sd, base_sd_thresh, mapping_code)
    stats = OrderedDict()
    stats["k" + str(table.k) + "_img_time"] = time.perf_counter() - start
    return stats


# ------------------------------------------------------------------------------------------ batch
def images_for_samples(samples, outfolder, kmer_mapping, k=7, mapping_code="varKode", min_bp=50000, max_bp=None,
                       is_query=False, seeds=None, subfolder_levels=0, overwrite=False, threads=None, engine=None,
                       on_error=None, gpu_workers=1, device=None, on_result=None, write_png_of=None):
    """Steps C-E of run_clean2img for MANY samples (the loop of ImageCommand.process_samples, image.py:1265-1294) on
    one GPU: samples are inflated ahead by worker threads into pinned memory (varkoder_b200.feed), pushed through the
    GPU by ``gpu_workers`` threads that each own a context, and their PNGs are written by the inflate pool off the
    critical path.  From gzip files the batch is bound by inflate on the host cores, so one GPU worker is the default;
    with inputs that are already in memory the path of one sample is a chain of short dependent kernels and three or
    four samples in flight fill the gaps.

    ``samples``: iterable of dicts ``{"sample": name, "path": clean .fq(.gz), "labels": [...], "base_sd": float}``;
    instead of ``"path"`` a sample may carry ``"data"`` (its uncompressed bytes in host memory: bytes / numpy uint8) or
    ``"device": (pointer, n_bytes)`` (the text already resident on this GPU, 16-byte aligned).
    ``"base_sd": None`` measures the quality flag from the reads on the GPU (quality.py) instead of taking it from a
    fastp report.
    ``seeds``: per-sample seeds (default: the sample's position).  ``engine``: use this one context only.
    ``on_result(sample, Result)``: called with every finished sample's counts (``Result.canon``) and pixels.
    ``write_png_of(sample) -> bool``: which samples get their PNG files (default: all).

    Returns ``{sample: stats}`` (submission order) with the reference's stats keys.  Failures stay with their sample,
    as in run_clean2img: too little data gives ``{"failed_step": "split"}`` (image.py:1020-1027), a file that cannot be
    read ``{"failed_step": "split", "error": ...}``, an error on the GPU side (a read beyond the supported length, a
    CUDA error) ``{"failed_step": "image", "error": ...}``; ``on_error(sample, exc)`` is called when given and the batch
    goes on with the next sample."""
    import threading
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    from .feed import SampleFeeder
    table = as_pixel_table(kmer_mapping, mapping_code)
    samples = list(samples)
    if engine is not None:
        gpu_workers = 1
    if device is None:
        device = int(os.environ.get("VARKODER_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    tls = threading.local()
    made = []
    want_canon = on_result is not None

    def gpu_job(i, s, buf, n, feeder):
        eng = engine
        if eng is None:
            eng = getattr(tls, "eng", None)
            if eng is None:
                eng = tls.eng = Engine(device)
                eng.set_batch_mode(gpu_workers > 1)
                made.append(eng)
        seed = seeds[i] if seeds is not None else i
        params = Params(k=int(k), min_bp=int(min_bp), max_bp=None if max_bp is None else int(max_bp),
                        is_query=bool(is_query), seed=seed)
        t0 = time.perf_counter()
        try:
            if "device" in s:
                ptr, nb = s["device"]
                res = eng.reads_to_images(int(ptr), params, table, on_device=True, n_bytes=int(nb), want_canon=want_canon)
            elif buf is None:
                res = eng.reads_to_images(s["data"], params, table, want_canon=want_canon)
            else:
                res = eng.reads_to_images(buf.array[:n], params, table, want_canon=want_canon)
            tm = eng.timings()
            sd = s.get("base_sd", 0)
            if sd is None and res.status == 0:              # no fastp report for this sample: measure it (quality.py)
                sd = quality.base_frequency_sd(eng.base_content())
        finally:
            feeder.release(buf)
        return res, tm, time.perf_counter() - t0, sd

    all_stats = OrderedDict()
    png_jobs = []

    def fail(s, step, exc):
        if on_error is not None:
            on_error(s, exc)
        st = OrderedDict(failed_step=step)
        if not isinstance(exc, LessThanMinimumData):
            st["error"] = f"{type(exc).__name__}: {exc}"
        all_stats[str(s["sample"])] = st

    def finish(s, fut, feeder):
        name = str(s["sample"])
        try:
            res, tm, dt, sd = fut.result()
        except Exception as exc:                     # this sample only; the batch goes on (image.py:1020-1027, 1111-1118)
            fail(s, "image", exc)
            return
        if res.status != 0:
            fail(s, "split", LessThanMinimumData())
            return
        stats = OrderedDict()
        split_s, count_s = _stage_times(tm)
        stats["splitting_time"] = split_s
        stats["splitting_bp_per_file"] = ",".join(str(x) for x in res.levels)
        stats[str(k) + "mer_counting_time"] = count_s
        stats["k" + str(k) + "_img_time"] = dt
        if s.get("base_sd", 0) is None:
            stats["base_frequencies_sd"] = sd
        all_stats[name] = stats
        if on_result is not None:
            on_result(s, res)
        if write_png_of is not None and not write_png_of(s):
            return
        for lvl, bp in enumerate(res.levels):
            outfile = image_name(name, bp, mapping_code, k)
            folder = _image_folder(outfolder, outfile, subfolder_levels)
            if not overwrite and (folder / outfile).is_file():
                continue
            png_jobs.append(feeder.pool.submit(write_png, res.pixels[lvl].copy(), folder / outfile,
                                               s.get("labels", ()), sd, QUAL_THRESH, mapping_code))

    gpu_pool = ThreadPoolExecutor(max_workers=max(1, int(gpu_workers)), thread_name_prefix="vk-gpu")
    try:
        with SampleFeeder(samples, path_of=lambda s: s.get("path"), threads=threads, errors="yield") as feeder:
            inflight = deque()
            for i, s, buf, n in feeder:
                if n < 0:                            # the file could not be read / inflated: split_fastq would have raised
                    fail(s, "split", buf)
                    continue
                inflight.append((s, gpu_pool.submit(gpu_job, i, s, buf, n, feeder)))
                while len(inflight) > gpu_workers:
                    s0, fut = inflight.popleft()
                    finish(s0, fut, feeder)
            while inflight:
                s0, fut = inflight.popleft()
                finish(s0, fut, feeder)
            for j in png_jobs:
                j.result()
    finally:
        gpu_pool.shutdown(wait=True)
        for e in made:
            e.close()
    # submission order, whatever order the samples finished or failed in
    return OrderedDict((str(s["sample"]), all_stats[str(s["sample"])]) for s in samples if str(s["sample"]) in all_stats)
