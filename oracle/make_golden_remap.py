"""Generate tests/golden/remap_k{K}.npz by running the UNMODIFIED reference ``convert.remap`` (convert.py:34-77,
imported from /root/reference).  TEST INFRASTRUCTURE ONLY; run in the build container:

    python -m oracle.make_golden_remap

Inputs are images that are consistent under reverse complement (every pixel that shows k-mer K or rc K holds the same
value -- true for everything ``varKoder image`` writes): for those the reference's result does not depend on the row
order of its pandas merge.  One input per k additionally is an ARBITRARY image in the varKode layout (varKode -> cgr
is order-independent for any image because the varKode table draws K and rc K on one pixel).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import image as oimg, ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def consistent_image(lut, k, rng, kind):
    """uint8 image in the layout of ``lut`` whose value depends only on the canonical class of the pixel's k-mer"""
    n = 4 ** k
    rc = np.array([oimg.revcomp_index(i, k) for i in range(n)])
    rep = np.minimum(np.arange(n), rc)
    if kind == "uniform":
        per_class = rng.integers(0, 256, n).astype(np.uint8)
    else:                                   # many small values: exercises the uint8 wrap of sum_rc less, ties more
        per_class = rng.integers(0, 40, n).astype(np.uint8)
    val = per_class[rep]
    img = np.zeros(lut.shape, dtype=np.uint8)
    used = lut >= 0
    img[used] = val[lut[used]]
    return img


def main():
    if not ref_shim.available():
        raise SystemExit("reference not mounted; goldens can only be made in the build container")
    _, utils, convert = ref_shim.load()
    from PIL import Image
    rng = np.random.default_rng(20260119)
    for k in (5, 6, 7):
        if os.path.exists(os.path.join(GOLD, f"remap_k{k}.npz")) and "--all" not in sys.argv:
            continue                           # committed fixture: only --all makes it again
        luts = {m: oimg.lut_from_table(utils.get_kmer_mapping(k, m)) for m in ("varKode", "cgr")}
        out = {}
        for src, dst in (("varKode", "cgr"), ("cgr", "varKode")):
            for kind in ("uniform", "small"):
                img = consistent_image(luts[src], k, rng, kind)
                for sum_rc in (False, True):
                    with np.errstate(all="ignore"):
                        res = np.array(convert.remap(Image.fromarray(img, mode="L"), k, src, dst, sum_rc=sum_rc))
                    key = f"{src}_to_{dst}__{kind}__{'sum' if sum_rc else 'plain'}"
                    out[key + "__in"] = img
                    out[key + "__out"] = res.astype(np.uint8)
        arb = rng.integers(0, 256, luts["varKode"].shape).astype(np.uint8)
        for sum_rc in (False, True):
            with np.errstate(all="ignore"):
                res = np.array(convert.remap(Image.fromarray(arb, mode="L"), k, "varKode", "cgr", sum_rc=sum_rc))
            key = f"varKode_to_cgr__arbitrary__{'sum' if sum_rc else 'plain'}"
            out[key + "__in"] = arb
            out[key + "__out"] = res.astype(np.uint8)
        np.savez_compressed(os.path.join(GOLD, f"remap_k{k}.npz"), **out)
        print("k", k, "cases", len(out) // 2)
    # k = 8, 9 (images of 182^2 / 256^2 and 363^2 / 512^2 pixels): fewer cases, small values (the fixtures stay small)
    for k in (8, 9):
        rng = np.random.default_rng(20260119 + k)          # own stream: independent of whether k = 5..7 were made in this run
        luts = {m: oimg.lut_from_table(utils.get_kmer_mapping(k, m)) for m in ("varKode", "cgr")}
        out = {}
        for src, dst in (("varKode", "cgr"), ("cgr", "varKode")):
            img = consistent_image(luts[src], k, rng, "small")
            for sum_rc in (False, True):
                with np.errstate(all="ignore"):
                    res = np.array(convert.remap(Image.fromarray(img, mode="L"), k, src, dst, sum_rc=sum_rc))
                key = f"{src}_to_{dst}__small__{'sum' if sum_rc else 'plain'}"
                out[key + "__in"] = img
                out[key + "__out"] = res.astype(np.uint8)
        np.savez_compressed(os.path.join(GOLD, f"remap_k{k}.npz"), **out)
        print("k", k, "cases", len(out) // 2)


if __name__ == "__main__":
    main()
