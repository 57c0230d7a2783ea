"""Generate tests/golden/base_sd.json by running the UNMODIFIED reference ``get_basefrequency_sd``
(varKoder/commands/image.py:49-88, imported from /root/reference) on fastp-shaped JSON reports whose content curves are
built from the oracle's per-position base counts of small synthetic FASTQ files.  TEST INFRASTRUCTURE ONLY; run in the
build container:

    python -m oracle.make_golden_quality

Each case stores the FASTQ recipe (seed, read lengths, bias), the integer counts [35, 5] and the value the reference
returned, so the tests can pin (1) the oracle's counts -> curves -> sd arithmetic and (2) the GPU kernel's counts.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import dsk, image as oimg, ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def fastq_case(seed, n_reads, len_lo, len_hi, bias):
    """reads of length uniform in [len_lo, len_hi]; ``bias`` tilts the base composition along the first 40 positions
    (what a low-quality library looks like to the flag); 1 % N; some lower-case bases"""
    rng = np.random.default_rng(seed)
    out = bytearray()
    for r in range(n_reads):
        L = int(rng.integers(len_lo, len_hi + 1))
        pos = np.arange(L)
        pa = 0.25 + bias * np.cos(pos / 6.0) * (pos < 40)
        probs = np.stack([pa, (1 - pa) / 3, (1 - pa) / 3, (1 - pa) / 3], axis=1)
        u = rng.random(L)
        idx = (u[:, None] > np.cumsum(probs, axis=1)).sum(axis=1).clip(0, 3)
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[idx].copy()
        seq[rng.random(L) < 0.01] = ord("N")
        low = rng.random(L) < 0.02
        seq[low] |= 0x20
        out += b"@r%d\n" % r + seq.tobytes() + b"\n+\n" + b"I" * L + b"\n"
    return bytes(out)


CASES = [
    dict(name="flat_150", seed=1, n_reads=3000, len_lo=150, len_hi=150, bias=0.0),
    dict(name="biased_150", seed=2, n_reads=3000, len_lo=150, len_hi=150, bias=0.08),
    dict(name="ragged_0_90", seed=3, n_reads=4000, len_lo=0, len_hi=90, bias=0.03),
    dict(name="short_0_30", seed=4, n_reads=2000, len_lo=0, len_hi=30, bias=0.05),     # curves end before position 40
    dict(name="tiny_0_5", seed=5, n_reads=50, len_lo=0, len_hi=5, bias=0.0),            # nothing reaches position 5
]


def fastp_report(curves, section):
    """the slice of a fastp JSON that the reference reads: content_curves of A, T, C, G, N, GC from cycle 0"""
    a, t, c, g = (list(map(float, row)) for row in curves)
    return {section: {"content_curves": {"A": a, "T": t, "C": c, "G": g, "N": [0.0] * len(a), "GC": [x + y for x, y in zip(c, g)]}}}


def main():
    image, _, _ = ref_shim.load()
    out = []
    for case in CASES:
        rec = {k: v for k, v in case.items()}
        buf = fastq_case(case["seed"], case["n_reads"], case["len_lo"], case["len_hi"], case["bias"])
        p = dsk.parse_fastq(buf)
        full = oimg.base_content(buf, p["starts"], p["lens"], 0, 64)         # from cycle 0, as fastp reports it
        reach = full[:, 4]
        n = int(np.count_nonzero(reach))
        curves = full[:n, :4].T.astype(np.float64) / reach[:n].astype(np.float64)[None, :]
        with tempfile.TemporaryDirectory() as d:
            f = os.path.join(d, "s_fastp_unpaired.json")
            json.dump(fastp_report(curves, "read1_after_filtering"), open(f, "w"))
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref = float(image.get_basefrequency_sd([f]))
        rec["counts_5_40"] = full[5:40].astype(int).tolist()
        rec["base_sd"] = None if np.isnan(ref) else ref
        rec["base_sd_hex"] = None if np.isnan(ref) else float(ref).hex()
        out.append(rec)
        print(case["name"], ref)
    json.dump(out, open(os.path.join(GOLD, "base_sd.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
