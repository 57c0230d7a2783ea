"""CPU suite, part 1: the oracle itself -- pinned against the reference's golden vectors.

These tests never touch the product; they establish that ``oracle/`` restates the reference so that the
GPU parity tests have something trustworthy to be compared with.
"""
import json
import os

import numpy as np
import pytest

from oracle import dsk, image as oimg, ref_shim
from tests.helpers import fastq, rand_reads

K_ALL = [5, 6, 7]


def test_oracle_c_matches_bruteforce():
    rng = np.random.default_rng(7)
    reads = rand_reads(rng, 60, 0, 90, p_n=0.03) + ["", "ACG", "N" * 20, "acgtacgtacgtNNacgtacgtacgt", "A" * 40]
    buf = fastq(reads)
    for k in (5, 7, 9):
        c = dsk.canonical_counts(buf, k)
        b = dsk.brute_canonical_counts(buf, k)
        assert (c == b).all()
        # conservation: every window of k valid bases inside a piece is counted once per strand
        assert int(c.sum()) == int(b.sum())


def test_two_independent_restatements_agree():
    """dsk_oracle.c (rolling window, C) and oracle/dsk_numpy.py (flat window enumeration + bincount, numpy) restate the same
    rules with different algorithms: framing, membership, counts and fold must agree on every kind of input -- N, lower
    case, reads longer than the break length, empty and too-short reads, no final newline, empty file, Bembidion-shaped and
    fixed-length synthetic samples.  (The GPU suite repeats this at the BASELINE size, 200 Mbp.)"""
    from oracle import dsk_numpy
    from varkoder_b200 import synth
    rng = np.random.default_rng(5)
    bufs = [fastq(rand_reads(rng, 400, 0, 90, p_n=0.03) + ["acgtn" * 20, "ACGT" * 400, "G" * 700, "", "ACGTAC"], final_newline=False),
            synth.variable(6000, seed=9).tobytes(), synth.fixed(3_000_000, 150, seed=2).tobytes(), b"", b"@x\n"]
    for buf in bufs:
        p = dsk.parse_fastq(buf)
        s, l = dsk_numpy.frame(buf)
        assert (s == p["starts"]).all() and (l == p["lens"]).all()
        nsites = p["nsites_ref"]
        levels = oimg.ladder(nsites, 3000, None) if nsites > 3000 else [max(nsites, 1)]
        for k, calibrated in ((5, True), (7, True), (7, False), (8, True), (9, True)):
            thr, take_all = dsk.level_thresholds(levels, nsites, 77, lens=p["lens"] if calibrated else None)
            _, c = dsk.count_levels(buf, k, 77, thr, take_all, threads=0)
            assert (c == dsk_numpy.count_levels(buf, k, 77, thr, take_all)).all(), (len(buf), k)


def test_oracle_breaklength_and_selection():
    rng = np.random.default_rng(3)
    reads = rand_reads(rng, 8, 900, 1700, p_n=0.002)
    buf = fastq(reads)
    sel = np.array([1, 0, 1, 1, 0, 0, 1, 1], dtype=np.uint8)
    for k in (6, 7):
        assert (dsk.canonical_counts(buf, k, select=sel) == dsk.brute_canonical_counts(buf, k, select=sel)).all()
        # with breaklength off there are strictly more k-mers (those spanning a cut)
        assert dsk.canonical_counts(buf, k, breaklen=0).sum() > dsk.canonical_counts(buf, k).sum()


def test_oracle_framing_matches_python_line_loop():
    # the reference counts bases with `for nlines, l in enumerate(f): if nlines % 4 == 1: len(l) - 1`
    cases = [
        fastq(["ACGT", "GG", ""]),
        fastq(["ACGT", "GGA"], final_newline=False),
        b"@h\nACGT",                      # unterminated sequence line: reference counts 3
        b"@h\n",                          # header only
        b"",
        b"@h\nAC\n+\n@@\n@h2\nGT\n+\n++\n",   # qualities starting with '@' and '+'
        fastq(["ACGT"]) + b"@trunc\nACG\n+",
    ]
    for buf in cases:
        expect = 0
        for i, l in enumerate(buf.splitlines(keepends=True)):
            if i % 4 == 1:
                expect += len(l) - 1
        p = dsk.parse_fastq(buf)
        assert p["nsites_ref"] == expect, buf
        assert p["n_lines"] == len(buf.splitlines())


def test_prio_hash_c_equals_python():
    for seed, r in [(0, 0), (1, 2), (2**64 - 1, 2**40), (12345678901234567890, 77)]:
        assert dsk.prio(seed, r) == dsk.prio_py(seed, r)


def test_ladder_golden(golden_dir):
    with open(os.path.join(golden_dir, "ladder.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 15
    for c in cases:
        if "raises" in c:
            with pytest.raises(Exception, match="less than minimum data"):
                oimg.ladder(c["nsites"], c["min_bp"], c["max_bp"], c["is_query"])
        else:
            sites = oimg.ladder(c["nsites"], c["min_bp"], c["max_bp"], c["is_query"])
            assert sites == c["sites"]
            assert ["x@" + oimg.level_tag(b) + ".fq.gz" for b in sites] == c["names"]


@pytest.mark.parametrize("k", [5, 6, 7, 8, 9])
@pytest.mark.parametrize("mapping", ["varKode", "cgr"])
def test_make_image_golden(golden_dir, k, mapping):
    """pixels written by the UNMODIFIED reference make_image == exact-integer restatement."""
    z = np.load(os.path.join(golden_dir, f"make_image_k{k}_{mapping}.npz"))
    from varkoder_b200.mapping import get_kmer_mapping
    lut = get_kmer_mapping(k, mapping).lut
    if k <= 7:
        assert (np.load(os.path.join(golden_dir, f"lut_k{k}_{mapping}.npy")) == lut).all() or mapping == "cgr"
    names = sorted({n.split("__")[0] for n in z.files})
    assert names
    for name in names:
        canon = z[name + "__counts"].astype(np.uint64)
        px = z[name + "__pixels"]
        assert (oimg.image_exact(canon, lut) == px).all(), name
        if k <= 6:
            assert (oimg.image_float(canon, lut) == px).all(), name


def test_docs_png_properties(golden_dir):
    """the reference's own shipped example images: histogram shape of the rank transform and
    the varKode <-> cgr geometry (remap identity)."""
    from PIL import Image
    from varkoder_b200.mapping import get_kmer_mapping
    d = os.path.join(golden_dir, "docs_png")
    vk = get_kmer_mapping(7, "varKode").lut
    cg = get_kmer_mapping(7, "cgr").lut
    samples = sorted({f.split("+")[0] for f in os.listdir(d)})
    assert len(samples) == 3
    rc = np.array([oimg.revcomp_index(i, 7) for i in range(4 ** 7)])
    canon_of = np.minimum(np.arange(4 ** 7), rc)
    for s in samples:
        a = np.array(Image.open(os.path.join(d, s + "+varKode+k7.png")))
        b = np.array(Image.open(os.path.join(d, s + "+cgr+k7.png")))
        assert a.shape == (91, 91) and b.shape == (128, 128)
        h = np.bincount(a.ravel(), minlength=256)
        assert h[0] == 0 and h[1] == 0 and h[2] >= 89           # 89 unused pixels share the lowest used level
        assert 28 <= np.median(h[3:]) <= 36 and a.max() == 255  # 8281 / 256 = 32.3 pixels per grey level
        # the shipped cgr image is the varKode image re-scattered by convert.remap (convert.py:34-77): pixels of
        # the same canonical class carry the same grey level => pins both tables' geometry incl. the y flip
        val = np.zeros(4 ** 7, dtype=np.int64)
        used = vk >= 0
        val[canon_of[vk[used]]] = a[used]
        assert (val[canon_of[cg]] == b).all()


@pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted (GPU box)")
def test_restatement_against_live_reference(tmp_path):
    """re-run the unmodified reference on fresh random counts (not the committed goldens)."""
    from PIL import Image
    _, utils, _ = ref_shim.load()
    rng = np.random.default_rng(99)
    for k, mapping in [(5, "varKode"), (6, "cgr")]:
        table = utils.get_kmer_mapping(k, mapping)
        lut = oimg.lut_from_table(table)
        n = 4 ** k
        rcs = np.array([oimg.revcomp_index(i, k) for i in range(n)])
        canon = rng.integers(0, 50, n).astype(np.uint64)[np.minimum(np.arange(n), rcs)]
        png, _ = ref_shim.reference_make_image(dsk.dsk2ascii_text(canon, k), tmp_path, table, k=k, mapping_code=mapping)
        assert (np.array(Image.open(png)) == oimg.image_exact(canon, lut)).all()


@pytest.mark.parametrize("k", [5, 6, 7])
def test_remap_golden(golden_dir, k):
    """oracle restatement of convert.remap against the outputs of the imported reference (oracle/make_golden_remap.py)"""
    z = np.load(os.path.join(golden_dir, f"remap_k{k}.npz"))
    luts = {m: np.load(os.path.join(golden_dir, f"lut_k{k}_{m}.npy")) for m in ("varKode", "cgr")}
    keys = sorted(set(n.rsplit("__", 1)[0] for n in z.files))
    assert len(keys) == 10
    for key in keys:
        d, _, mode = key.split("__")
        src, dst = d.split("_to_")
        got = oimg.remap_exact(z[key + "__in"], luts[src], luts[dst], k, src == "cgr", dst == "cgr", mode == "sum")
        assert (got == z[key + "__out"]).all(), key


@pytest.mark.parametrize("k", [8, 9])
def test_remap_golden_large_k(golden_dir, k):
    """the same for the large images (182^2 / 256^2, 363^2 / 512^2); their pixel tables are not stored as fixtures --
    the product's tables are used, which the make_image goldens of k = 8, 9 pin"""
    from varkoder_b200.mapping import get_kmer_mapping
    z = np.load(os.path.join(golden_dir, f"remap_k{k}.npz"))
    luts = {m: get_kmer_mapping(k, m).lut for m in ("varKode", "cgr")}
    keys = sorted(set(n.rsplit("__", 1)[0] for n in z.files))
    assert len(keys) == 4
    for key in keys:
        d, _, mode = key.split("__")
        src, dst = d.split("_to_")
        got = oimg.remap_exact(z[key + "__in"], luts[src], luts[dst], k, src == "cgr", dst == "cgr", mode == "sum")
        assert (got == z[key + "__out"]).all(), key


def test_quality_flag_golden(golden_dir):
    """get_basefrequency_sd (image.py:49-88) run on fastp-shaped reports built from the oracle's counts
    (oracle/make_golden_quality.py): the numpy restatement of the counts and the host arithmetic of
    varkoder_b200.quality reproduce the stored counts and the reference's float64 value bit for bit."""
    from oracle.make_golden_quality import fastq_case
    from varkoder_b200 import quality
    cases = json.load(open(os.path.join(golden_dir, "base_sd.json")))
    assert len(cases) >= 5
    for c in cases:
        buf = fastq_case(c["seed"], c["n_reads"], c["len_lo"], c["len_hi"], c["bias"])
        p = dsk.parse_fastq(buf)
        counts = oimg.base_content(buf, p["starts"], p["lens"], 5, 40)
        assert counts.astype(int).tolist() == c["counts_5_40"], c["name"]
        with np.errstate(all="ignore"):
            sd = quality.base_frequency_sd(counts)
        if c["base_sd"] is None:
            assert np.isnan(sd), c["name"]
        else:
            assert float(sd).hex() == c["base_sd_hex"], c["name"]
        assert quality.low_quality_flag(sd) == (c["base_sd"] is not None and c["base_sd"] > 0.01)
