"""The reference's low-quality flag from the reads themselves.

``get_basefrequency_sd`` (varKoder/commands/image.py:49-88) opens the fastp JSON written by ``clean_reads``, takes the
``content_curves`` of A, T, C, G (per read position: share of that base among all bases at the position), keeps
positions 5..39 and returns ``np.std(content[:, 5:40], axis=1).mean()`` -- averaged over the report's
``merged_and_filtered`` and ``read1_after_filtering`` sections when both exist.  ``run_clean2img`` stores it as
``base_frequencies_sd`` and passes it to ``make_image`` for the PNG keys ``varkoderBaseFreqSd`` and
``varkoderLowQualityFlag`` (``base_sd > 0.01``, image.py:1093-1096, :920-930).

Here the curves come from the cleaned reads the path has framed anyway (``vk_base_content``: integer numerators and
denominators, one extra kernel over the read table), so the flag no longer depends on a fastp report being around.
The float64 tail below is the reference's expression, fed with curves built the way fastp builds them (count / reads
that reach the position, positions beyond the longest read absent).  What differs by construction: one curve over ALL
cleaned reads instead of the mean of fastp's per-section curves; for unpaired input the two are the same numbers.
"""
from __future__ import annotations

import numpy as np

POS_BEGIN, POS_END = 5, 40          # content[:, 5:40]  (image.py:70, :81)
BASES = ("A", "T", "C", "G")        # order of fastp's content_curves keys


def content_curves(counts, pos_begin=POS_BEGIN):
    """counts: [P, 5] (A, T, C, G, reads reaching the position) for positions pos_begin.. -> float64 [4, P'] where P'
    stops at the longest read, like the arrays in a fastp report."""
    counts = np.asarray(counts, dtype=np.uint64)
    reach = counts[:, 4]
    n = int(np.count_nonzero(reach))              # reach is non-increasing along the positions
    if n and not (reach[:n] > 0).all():
        raise ValueError("reads-per-position column must be non-increasing")
    # rows contiguous along the positions, like the array the reference builds from the report's lists: numpy's
    # reductions then add in the same order and the result is bit-identical (tests/golden/base_sd.json)
    num = np.ascontiguousarray(counts[:n, :4].T, dtype=np.float64)
    return num / reach[:n].astype(np.float64)[None, :]


def base_frequency_sd(counts):
    """the reference's ``base_sd`` (image.py:66-71) from vk_base_content counts of positions 5..39"""
    content = content_curves(counts)
    if content.shape[1] == 0:
        return float("nan")                        # np.std of an empty slice, as the reference would return
    return float(np.std(content, axis=1).mean())


def low_quality_flag(base_sd, base_sd_thresh=0.01):
    """``varkoderLowQualityFlag`` (make_image, image.py:927)"""
    return bool(base_sd > base_sd_thresh)
