"""B200-native implementation of varKoder's image hot path.

cleaned FASTQ reads -> sub-sample ladder -> canonical k-mer counts -> varKode / CGR pixels (PNG),
i.e. steps C-E of ``run_clean2img`` (reference: varKoder/commands/image.py:1005-1125).

The compute lives in hand-written sm_100a CUDA kernels behind a C ABI (``include/varkoder_b200.h``,
``varkoder_b200/libvarkoder_b200.so``); this package is the Python host side, mirroring the reference's
own stage functions (``split_fastq``, ``count_kmers``, ``make_image``) so it can be dropped into
``run_clean2img``.  There is no CPU fallback: importing :mod:`varkoder_b200.engine` without the built
library, or creating an :class:`Engine` without a B200, raises.
"""
from .ladder import level_tag, image_name                  # noqa: F401  (ladder() lives in varkoder_b200.ladder)
from .mapping import PixelTable, get_kmer_mapping           # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # the CUDA-backed parts are imported lazily so that host-only logic (ladder, tables, names)
    # can be used and tested where no GPU exists
    if name in ("Engine", "Params"):
        from . import engine
        return getattr(engine, name)
    if name in ("split_fastq", "count_kmers", "make_image", "reads_to_images", "default_engine"):
        from . import stages
        return getattr(stages, name)
    if name in ("remap", "remap_arrays"):
        from . import convert
        return getattr(convert, name)
    raise AttributeError(name)
