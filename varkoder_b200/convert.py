"""GPU mirror of ``remap`` in the reference's convert command (varKoder/commands/convert.py:34-77): move the pixels of
varKode images to the CGR layout and back.  Same name, arguments and return type (a PIL image) as the reference, plus
:func:`remap_arrays` for whole batches.  The gather runs in ``vk_remap`` (csrc/vk_image.cuh); the join of the two pixel
tables is precomputed once per (k, direction) on the host (``mapping.remap_plan``).  No CPU fallback.
"""
import numpy as np

from .mapping import remap_plan


def remap_arrays(images, k, in_mapping, out_mapping, sum_rc=False, engine=None):
    """uint8 array [n, H, W] (or [H, W]) in the ``in_mapping`` layout -> [n, H', W'] in the ``out_mapping`` layout"""
    from .stages import default_engine
    eng = engine or default_engine()
    src0, src1, mult, shape = remap_plan(int(k), in_mapping, out_mapping)
    a = np.asarray(images, dtype=np.uint8)
    single = a.ndim == 2
    out = eng.remap(a, src0, src1, mult, shape, sum_rc=sum_rc)
    return out[0] if single else out


def remap(img, k, in_mapping, out_mapping, sum_rc=False, engine=None):
    """convert.py:34-77.  ``img``: PIL image (mode L) as written by ``varKoder image``; returns the remapped PIL image."""
    from PIL import Image
    return Image.fromarray(remap_arrays(np.array(img), k, in_mapping, out_mapping, sum_rc, engine), mode="L")
