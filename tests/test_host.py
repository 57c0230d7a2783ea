"""CPU suite, part 2: host logic of the product and the C-ABI library (no compute calls without a GPU)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from varkoder_b200 import _lib, ladder as vl, mapping as vm, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.load()
    assert L.vk_abi_version() == 2
    with open(os.path.join(ROOT, "include", "varkoder_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(vk_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.VkParams) == 4 * 4 + 5 * 8 + 2 * 4 + 8      # + sampling, reserved0, prio_hist (ABI 2)
    assert ctypes.sizeof(_lib.VkStats) == 5 * 8
    assert ctypes.sizeof(_lib.VkResult) == 5 * 8 + 8 + 3 * 64 * 8


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from varkoder_b200.engine import Engine, VkError
    with pytest.raises(VkError):
        Engine(0)


def test_product_ladder_matches_reference_golden(golden_dir):
    with open(os.path.join(golden_dir, "ladder.json")) as f:
        cases = json.load(f)
    for c in cases:
        if "raises" in c:
            with pytest.raises(Exception, match="Input file has less than minimum data."):
                vl.ladder(c["nsites"], c["min_bp"], c["max_bp"], c["is_query"])
        else:
            sites = vl.ladder(c["nsites"], c["min_bp"], c["max_bp"], c["is_query"])
            assert sites == c["sites"]
            assert ["x@" + vl.level_tag(b) + ".fq.gz" for b in sites] == c["names"]
    assert vl.image_name("s", 200_000_000, "cgr", 7) == "s@00200000K+cgr+k7.png"
    assert vl.parse_seed("954294967295") == 954294967295 and vl.parse_seed(None) == 0


def test_product_ladder_matches_oracle_sweep():
    from oracle import image as oimg
    rng = np.random.default_rng(0)
    for _ in range(300):
        ns = int(10 ** rng.uniform(2, 11))
        mn = int(10 ** rng.uniform(2, 7))
        mx = None if rng.random() < 0.3 else int(10 ** rng.uniform(3, 9))
        q = bool(rng.random() < 0.2)
        try:
            a = oimg.ladder(ns, mn, mx, q)
        except Exception:
            a = "raise"
        try:
            b = vl.ladder(ns, mn, mx, q)
        except Exception:
            b = "raise"
        assert a == b, (ns, mn, mx, q)


@pytest.mark.parametrize("k", [5, 6, 7])
def test_pixel_tables_match_reference_golden(golden_dir, k):
    for m in ("varKode", "cgr"):
        t = vm.get_kmer_mapping(k, m)
        g = np.load(os.path.join(golden_dir, f"lut_k{k}_{m}.npy"))
        rc = vm.revcomp_index(np.arange(4 ** k), k)
        canon = np.minimum(np.arange(4 ** k), rc)
        assert t.lut.shape == g.shape
        assert ((t.lut < 0) == (g < 0)).all()
        used = g >= 0
        assert (canon[t.lut[used]] == canon[g[used]]).all()
    sides = {5: (23, 32), 6: (46, 64), 7: (91, 128)}[k]
    assert (vm.get_kmer_mapping(k, "varKode").side, vm.get_kmer_mapping(k, "cgr").side) == sides


def test_pixel_table_errors():
    with pytest.raises(Exception, match='method must be "varKode" or "cgr"'):
        vm.get_kmer_mapping(7, "nope")
    with pytest.raises(ValueError):
        vm.get_kmer_mapping(4, "cgr")


def test_lut_from_dataframe_roundtrip():
    import pandas as pd
    k = 5
    t = vm.get_kmer_mapping(k, "cgr")
    rows = []
    side = t.side
    letters = "ACGT"
    for r in range(side):
        for c in range(side):
            idx = int(t.lut[r, c])
            s = "".join(letters[(idx >> (2 * (k - 1 - i))) & 3] for i in range(k))
            rows.append((s, c, side - 1 - r))
            rcs = "".join({"A": "T", "C": "G", "G": "C", "T": "A"}[ch] for ch in reversed(s))
            rows.append((rcs, c, side - 1 - r))
    df = pd.DataFrame(rows, columns=["kmer", "x", "y"]).set_index("kmer")
    lut = vm.lut_from_dataframe(df)
    rc = vm.revcomp_index(np.arange(4 ** k), k)
    canon = np.minimum(np.arange(4 ** k), rc)
    assert (canon[lut] == canon[t.lut]).all()
    bad = df.copy()
    bad.iloc[0, bad.columns.get_loc("x")] = int(bad.iloc[5]["x"])
    bad.iloc[0, bad.columns.get_loc("y")] = int(bad.iloc[5]["y"])
    with pytest.raises(ValueError):
        vm.lut_from_dataframe(bad)


def test_synthetic_generator_shape_and_determinism():
    from oracle import dsk
    a = synth.fixed(30_000, 150, seed=5)
    b = synth.fixed(30_000, 150, seed=5)
    assert (a == b).all() and len(a) == synth.fixed_total_bytes(30_000, 150)
    p = dsk.parse_fastq(a)
    assert p["nsites_ref"] == 30_000 and p["n_reads"] == 200
    text = a.tobytes()
    assert text.startswith(b"@S0000000000\n") and text.endswith(b"\n")
    bases = np.concatenate([a[s:s + l] for s, l in zip(p["starts"], p["lens"])])
    frac = {c: (bases == ord(c)).mean() for c in "ACGTN"}
    assert 0.25 < frac["A"] < 0.33 and 0.17 < frac["C"] < 0.23 and frac["N"] < 0.004
    big = synth.fixed(4_000_000, 150, seed=5)             # vectorised path == per-record path
    assert (big[:len(a) - 317] == a[:len(a) - 317]).all()
    v = synth.variable(500, seed=1)
    pv = dsk.parse_fastq(v)
    assert pv["n_reads"] == 500 and pv["lens"].min() < 7 and pv["lens"].max() <= 280


@pytest.mark.parametrize("k", [5, 6, 7, 8])
def test_remap_plan_matches_oracle(golden_dir, k):
    """the product's join of the two pixel tables (mapping.remap_plan) against the oracle's restatement, which is
    pinned to the reference's convert.remap by tests/golden/remap_k*.npz (test_oracle.py)"""
    from oracle import image as oimg
    for src, dst in (("varKode", "cgr"), ("cgr", "varKode")):
        s0, s1, mult, shape = vm.remap_plan(k, src, dst)
        e0, e1, em = oimg.remap_plan(vm.get_kmer_mapping(k, src).lut, vm.get_kmer_mapping(k, dst).lut, k, src == "cgr", dst == "cgr")
        assert shape == vm.get_kmer_mapping(k, dst).lut.shape
        assert (s0 == e0).all() and (s1 == e1).all() and (mult == em).all()
    with pytest.raises(Exception, match="Input and output mapping must be one of"):
        vm.remap_plan(k, "varKode", "nope")


def test_roofline_traffic_is_recorded_per_kernel():
    """bench.py's ``roofline.traffic`` is looked up by the NAME of the kernel that dominates the workload (round 1 reported the
    k = 7 kernel's DRAM bytes on the k = 9 line): every kernel a default run can name has its own entry, taken from an ncu
    capture under profiles/, of the right order of magnitude for its workload (2.1133 text bytes per base)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "kernel_traffic.json")) as f:
        traffic = json.load(f)
    bases = {"countt_kernel<16>": 200_000_000, "count_kernel<7,smem>": 200_000_000, "countt9_kernel": 1_000_000_000,
             "count9h_kernel": 1_000_000_000}
    for name, n in bases.items():
        ent = traffic[name]
        assert ent["dram_bytes_per_launch"] == ent["read"] + ent["write"]
        assert 0.8 <= ent["dram_bytes_per_launch"] / (n * 2.1133) <= 1.1, name
        src = ent["source"].split(" ")[0]
        assert os.path.exists(os.path.join(root, src)), src
