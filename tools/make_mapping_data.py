"""Derive the varKode pixel tables shipped in ``varkoder_b200/data/varkode_lut.npz``.

The varKode layout is an opaque learned embedding distributed with the reference as
``varKoder/kmer_mapping/{5..9}mer_mapping.parquet`` (columns kmer, x, y; loaded by get_kmer_mapping,
core/utils.py:163-165).  It cannot be computed, so this script (run once, in the build container)
converts each table into the compact form the GPU path uploads: for every pixel of the FINAL image
(row = H-1-y, col = x; image.py:911-913) the lexicographic index of one k-mer of the canonical class shown
there, or -1 for unused pixels.  In a drop-in deployment the caller may instead pass the reference's own
DataFrame; ``varkoder_b200.mapping.lut_from_dataframe`` performs the same conversion at run time.
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from varkoder_b200.mapping import lut_from_dataframe  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/varKoder/kmer_mapping"
out = {}
for k in range(5, 10):
    df = pd.read_parquet(os.path.join(src, f"{k}mer_mapping.parquet")).set_index("kmer")
    lut = lut_from_dataframe(df)
    out[f"k{k}"] = lut.astype(np.int32)
    print(k, lut.shape, int((lut < 0).sum()), "unused")
np.savez_compressed(os.path.join(ROOT, "varkoder_b200", "data", "varkode_lut.npz"), **out)
