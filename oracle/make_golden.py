"""Generate tests/golden/* by running the UNMODIFIED reference (imported from /root/reference).

TEST INFRASTRUCTURE ONLY; run in the build container (``python -m oracle.make_golden``).  The GPU box has
no /root/reference, so the vectors are committed.  What each file pins:

  make_image_k{K}_{mapping}.npz   inputs: canon_full counts (lexicographic index); outputs: the pixels of the
                                  PNG written by the reference ``make_image`` (image.py:808-936) when its
                                  dsk2ascii subprocess is replaced by the oracle's text dump of those counts.
                                  Pins R4-R7 + R8/R9 (join, scatter orientation, rank scaling, tables).
  ladder.json                     ``split_fastq`` ladders and file names (image.py:669-709) incl. the raise.
  docs_png/*.png                  the reference's own shipped example images (docs/*.png), used for the
                                  remap-identity and histogram-shape properties.
  lut_k{K}_{mapping}.npy          pixel->k-mer tables derived from get_kmer_mapping (utils.py:152-217)
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import dsk, image as oimg, ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def count_cases(k, rng):
    """adversarial canonical count vectors (returned as canon_full over all 4^k lexicographic indices)."""
    n = 4 ** k
    rc = np.array([oimg.revcomp_index(i, k) for i in range(n)])
    rep = np.minimum(np.arange(n), rc)

    def sym(v):                      # make canon_full symmetric: value of the class representative
        return v[rep]

    cases = {}
    cases["poisson20"] = sym(rng.poisson(20, n).astype(np.uint64))
    cases["sparse"] = sym((rng.random(n) < 0.03) * rng.integers(1, 5, n).astype(np.uint64))
    cases["heavy_tail"] = sym(np.floor(np.exp(rng.normal(6, 2.5, n))).astype(np.uint64))
    cases["ties"] = sym(rng.integers(0, 3, n).astype(np.uint64))
    cases["all_zero"] = np.zeros(n, dtype=np.uint64)
    cases["all_equal"] = np.full(n, 7, dtype=np.uint64)
    big = rng.integers(0, 2 ** 31 - 1, n).astype(np.uint64)
    cases["big"] = sym(big)
    one = np.zeros(n, dtype=np.uint64)
    one[rep[n // 3]] = 12345
    one[rc[rep[n // 3]]] = 12345
    cases["single"] = one
    return cases


def main():
    if not ref_shim.available():
        raise SystemExit("reference not mounted; goldens can only be made in the build container")
    image, utils, convert = ref_shim.load()
    os.makedirs(GOLD, exist_ok=True)
    from PIL import Image

    rng = np.random.default_rng(20260118)
    plan = {5: None, 6: None, 7: None, 8: ["poisson20", "sparse"], 9: ["poisson20"]}
    for k, only in plan.items():
        for mapping in ("varKode", "cgr"):
            table = utils.get_kmer_mapping(k, mapping)
            lut = oimg.lut_from_table(table)
            np.save(os.path.join(GOLD, f"lut_k{k}_{mapping}.npy"), lut)
            cases = count_cases(k, rng)
            if only is not None:
                cases = {c: cases[c] for c in only}
            out = {}
            for name, canon in cases.items():
                text = dsk.dsk2ascii_text(canon, k)
                with tempfile.TemporaryDirectory() as d:
                    png, _ = ref_shim.reference_make_image(text, d, table, k=k, mapping_code=mapping,
                                                           labels=["a", "b"], base_sd=0.02)
                    img = Image.open(png)
                    assert img.mode == "L"
                    px = np.array(img)
                    meta = dict(img.info)
                store = canon.astype(np.uint32) if canon.max() < 2 ** 32 else canon
                out[name + "__counts"] = store
                out[name + "__pixels"] = px
                # the restatement must already agree, or the golden is useless as a pin
                assert (oimg.image_exact(canon, lut) == px).all(), (k, mapping, name)
            np.savez_compressed(os.path.join(GOLD, f"make_image_k{k}_{mapping}.npz"), **out)
            print("k", k, mapping, "cases", len(cases), "side", lut.shape, "meta", meta)

    # ladders
    lad = []
    for nsites, mn, mx, q in [
        (250_000_000, 500_000, 200_000_000, False), (200_000_000, 500_000, 200_000_000, False),
        (37_123_456, 500_000, 200_000_000, False), (10_000_000, 500_000, 200_000_000, False),
        (10_000_001, 500_000, 200_000_000, False), (10_000_000, 9_000_000, 10_000_000, False),
        (600_000, 500_000, 200_000_000, False), (500_001, 500_000, 200_000_000, False),
        (500_000, 500_000, 200_000_000, False), (400_000, 500_000, 200_000_000, False),
        (400_000, 500_000, 200_000_000, True), (30_000_000, 500_000, 20_000_000, False),
        (30_000_000, 50_000, 20_000_000, True), (1_234_567, 50_000, None, False),
        (999, 50_000, None, False), (5_000_001, 500_000, 200_000_000, False),
        (2_000_000, 500_000, 1_000_000, False), (123_456_789, 1_000, 200_000_000, False),
    ]:
        try:
            sites, names = ref_shim.reference_ladder(nsites, mn, mx, q)
            lad.append(dict(nsites=nsites, min_bp=mn, max_bp=mx, is_query=q, sites=sites, names=names))
        except Exception as e:  # the reference raises a bare Exception (image.py:680)
            lad.append(dict(nsites=nsites, min_bp=mn, max_bp=mx, is_query=q, raises=str(e)))
        # restatement check
        try:
            mine = oimg.ladder(nsites, mn, mx, q)
        except Exception as e:
            mine = str(e)
        assert mine == lad[-1].get("sites", lad[-1].get("raises")), lad[-1]
    with open(os.path.join(GOLD, "ladder.json"), "w") as f:
        json.dump(lad, f, indent=1)
    print("ladders", len(lad))

    # shipped example images
    dst = os.path.join(GOLD, "docs_png")
    os.makedirs(dst, exist_ok=True)
    src = os.path.join(ref_shim.REFERENCE_ROOT, "docs")
    for fn in sorted(os.listdir(src)):
        if fn.endswith(".png") and "@" in fn:
            shutil.copyfile(os.path.join(src, fn), os.path.join(dst, fn))
            os.chmod(os.path.join(dst, fn), 0o644)
    print("done")


if __name__ == "__main__":
    main()
