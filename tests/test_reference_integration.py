"""The reference's OWN caller drives the replacement: the unmodified ``run_clean2img``
(/root/reference/varKoder/commands/image.py:938-1141) is imported, its three stage names are rebound to
``varkoder_b200.stages`` exactly as INTEGRATION.md section 2a tells a maintainer to do, ``clean_reads`` (upstream of
the path, needs fastp) is stubbed, and steps C-E run: globs at :1060 / :1092, stats accumulation, ``failed_step``
handling at :1020-1027.  No GPU here: the stage functions talk to a CPU stand-in engine built on the oracle
(tests/helpers.OracleEngine) -- what is under test is the host-side contract (names, files, stats keys, errors), the
CUDA engine behind the same calls is covered by tests/test_gpu_parity.py.

Runs only where /root/reference is mounted (the build container); skipped on the GPU box."""
import gzip
import json
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import dsk, image as oimg, ref_shim
from tests.helpers import OracleEngine, oracle_images, oracle_levels
from varkoder_b200 import stages, synth
from varkoder_b200.ladder import LessThanMinimumData, ladder, parse_seed

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted (GPU box)")


def _fastp_report(path):
    """a fastp-shaped report (what clean_reads leaves next to the clean file, image.py:1093-1096)"""
    curves = {b: [0.25 + 0.001 * ((i * (j + 1)) % 7) for i in range(60)] for j, b in enumerate("ATCG")}
    curves["N"] = [0.0] * 60
    json.dump({"read1_after_filtering": {"content_curves": curves}}, open(path, "w"))


def _args(command="image", k=7, mapping="cgr", min_bp="20K", max_bp=None):      # cli.py:496-501 maps "--max-bp 0" to None
    return SimpleNamespace(cpus_per_thread=2, kmer_size=k, max_bp=max_bp, min_bp=min_bp, no_adapter=False, no_merge=False,
                           no_deduplicate=False, trim_bp="10,10", overwrite=False, verbose=False, command=command,
                           no_image=False, kmer_mapping=mapping)


@pytest.fixture()
def rebound(monkeypatch):
    image, utils, _ = ref_shim.load()
    eng = OracleEngine()
    monkeypatch.setattr(stages, "default_engine", lambda device=None: eng)
    stages._SAMPLE.clear()
    # rebind the three names; nothing else of the reference is touched
    monkeypatch.setattr(image, "split_fastq", stages.split_fastq)
    monkeypatch.setattr(image, "count_kmers", stages.count_kmers)
    monkeypatch.setattr(image, "make_image", stages.make_image)
    return image, utils


def _run(image, utils, tmp_path, sample, buf, args, row_index=3, labels=("Bembidion", "sp1")):
    inter = tmp_path / "inter"
    (inter / "clean_reads").mkdir(parents=True, exist_ok=True)
    clean = inter / "clean_reads" / (sample + ".fq.gz")

    def fake_clean_reads(infiles, outpath, **kw):             # step B is upstream of the path (fastp, pigz)
        with gzip.open(outpath, "wb", compresslevel=1) as f:
            f.write(buf)
        _fastp_report(inter / "clean_reads" / (sample + "_fastp_unpaired.json"))
        return OrderedDict(cleaning_time=0.0)

    image.clean_reads, real = fake_clean_reads, image.clean_reads
    try:
        kmer_mapping = utils.get_kmer_mapping(args.kmer_size, args.kmer_mapping)
        row = (row_index, {"sample": sample, "labels": list(labels), "files": ["raw_R1.fq.gz"]})
        rng = np.random.default_rng(5)
        stats = image.run_clean2img(row, kmer_mapping, args, rng, inter, OrderedDict(), tmp_path / "stats.csv",
                                    tmp_path / "images", 0)
    finally:
        image.clean_reads = real
    seed = str(row_index) + str(np.random.default_rng(5).integers(low=0, high=2**32))
    return stats, inter, clean, seed


def test_run_clean2img_with_rebound_stage_functions(rebound, tmp_path):
    from PIL import Image
    image, utils = rebound
    buf = synth.variable(2500, seed=77).tobytes()
    args = _args()
    stats, inter, clean, seed = _run(image, utils, tmp_path, "S1", buf, args)
    st = stats["S1"]
    p = dsk.parse_fastq(buf)
    levels = ladder(p["nsites_ref"], 20_000, None)
    assert len(levels) >= 4
    # stats contract: the keys run_clean2img accumulates, no failure recorded
    assert "failed_step" not in st
    assert st["splitting_bp_per_file"] == ",".join(str(x) for x in levels)
    for key in ("cleaning_time", "splitting_time", "7mer_counting_time", "k7_img_time", "base_frequencies_sd"):
        assert key in st, key
    assert st["7mer_counting_time"] > 0 and st["k7_img_time"] > 0
    # the globs of image.py:1060 and :1092 found one placeholder per level, under the reference's name stems
    assert sorted(f.name for f in (inter / "split_fastqs").glob("S1@*")) == sorted(
        "S1@" + str(int(bp / 1000)).rjust(8, "0") + "K.fq.vk" for bp in set(levels))
    assert len(list((inter / "7mer_counts").glob("S1@*"))) == len(set(levels))
    # PNGs: reference names, pixels of the oracle for the reads the seeded rule selects, reference metadata
    canon = oracle_levels(buf, 7, parse_seed(seed), levels, p["nsites_ref"])
    table = stages.as_pixel_table(utils.get_kmer_mapping(7, "cgr"), "cgr")
    pix = oracle_images(canon, table.lut)
    base_sd = image.get_basefrequency_sd((inter / "clean_reads").glob("S1_fastp_*.json"))
    for lvl, bp in enumerate(levels):
        f = tmp_path / "images" / ("S1@" + str(int(bp / 1000)).rjust(8, "0") + "K+cgr+k7.png")
        img = Image.open(f)
        assert img.mode == "L" and (np.asarray(img) == pix[lvl]).all()
        assert list(img.info)[:4] == ["varkoderKeywords", "varkoderBaseFreqSd", "varkoderLowQualityFlag", "varkoderMapping"]
        assert img.info["varkoderKeywords"] == "Bembidion;sp1" and img.info["varkoderMapping"] == "cgr"
        assert img.info["varkoderBaseFreqSd"] == str(base_sd) and img.info["varkoderLowQualityFlag"] == str(base_sd > 0.01)
        # the consumer-side name contract (core/utils.py:123-149)
        meta = utils.get_metadata_from_img_filename(f)
        assert meta["sample"] == "S1" and meta["bp"] == int(bp / 1000) * 1000 and meta["img_kmer_size"] == 7
    # second run without --overwrite: every stage skips (empty stats), as the reference's stages do
    stats2, *_ = _run(image, utils, tmp_path, "S1", buf, args)
    assert "splitting_time" not in stats2["S1"] and stats2["S1"]["7mer_counting_time"] == 0 and stats2["S1"]["k7_img_time"] == 0


def test_run_clean2img_records_split_failure(rebound, tmp_path):
    """too little data: split_fastq raises, run_clean2img records failed_step: split and skips the sample
    (image.py:1020-1027)"""
    image, utils = rebound
    buf = synth.variable(300, seed=5).tobytes()
    stats, inter, *_ = _run(image, utils, tmp_path, "tiny", buf, _args(min_bp="10M", max_bp="200M"))
    assert stats["tiny"]["failed_step"] == "split"
    assert not (tmp_path / "images").exists() or not list((tmp_path / "images").glob("tiny@*"))


def test_run_clean2img_query_mode(rebound, tmp_path):
    """args.command == 'query' (query.py:154-165 calls the same function): one level of min(nsites, max_bp), no min_bp"""
    image, utils = rebound
    buf = synth.variable(1500, seed=9).tobytes()
    args = _args(command="query", k=6, mapping="varKode", max_bp="100K")
    stats, inter, clean, seed = _run(image, utils, tmp_path, "Q", buf, args, row_index=0)
    assert stats["Q"]["splitting_bp_per_file"] == "100000"
    files = list((tmp_path / "images").glob("Q@*"))
    assert [f.name for f in files] == ["Q@00000100K+varKode+k6.png"]
    from PIL import Image
    p = dsk.parse_fastq(buf)
    canon = oracle_levels(buf, 6, parse_seed(seed), [100_000], p["nsites_ref"])
    table = stages.as_pixel_table(utils.get_kmer_mapping(6, "varKode"), "varKode")
    assert (np.asarray(Image.open(files[0])) == oracle_images(canon, table.lut)[0]).all()
