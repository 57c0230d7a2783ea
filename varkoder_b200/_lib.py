"""ctypes binding of ``libvarkoder_b200.so`` (C ABI declared in ``include/varkoder_b200.h``).

Fails loudly: there is no fallback implementation of the hot path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvarkoder_b200.so")

VK_MAX_LEVELS = 64
VK_BREAKLENGTH = 500
VK_OK = 0
VK_LADDER_LESS_THAN_MIN = 1
VK_SAMPLING_EXPECTED = 0
VK_SAMPLING_CALIBRATED = 1
VK_PRIO_BUCKETS = 65536
VK_ABI_VERSION = 2

# every symbol include/varkoder_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vk_abi_version", "vk_last_error", "vk_ctx_create", "vk_ctx_destroy", "vk_set_mapping", "vk_upload",
    "vk_attach", "vk_parse", "vk_count", "vk_prio_hist", "vk_render", "vk_render_counts", "vk_reads_to_images", "vk_device_pixels", "vk_remap",
    "vk_base_content",
    "vk_last_timings", "vk_set_fine_timing", "vk_set_batch_mode", "vk_launch_count", "vk_bucket_retries", "vk_count_fallbacks", "vk_synth_fastq",
    "vk_synth_fastq_variable", "vk_graph_stats",
    "vk_comm_unique_id", "vk_comm_init", "vk_comm_destroy", "vk_sharded_reads_to_images",
]


class VkParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("is_query", C.c_int32), ("has_max_bp", C.c_int32), ("breaklength", C.c_int32),
                ("min_bp", C.c_uint64), ("max_bp", C.c_uint64), ("seed", C.c_uint64),
                ("read_index_base", C.c_uint64), ("nsites_override", C.c_uint64),
                ("sampling", C.c_int32), ("reserved0", C.c_int32), ("prio_hist", C.c_uint64)]


class VkStats(C.Structure):
    _fields_ = [("n_bytes", C.c_uint64), ("n_lines", C.c_uint64), ("n_reads", C.c_uint64),
                ("nsites", C.c_uint64), ("nsites_true", C.c_uint64)]


class VkResult(C.Structure):
    _fields_ = [("stats", VkStats), ("status", C.c_int32), ("n_levels", C.c_int32),
                ("level_bp", C.c_uint64 * VK_MAX_LEVELS), ("level_reads", C.c_uint64 * VK_MAX_LEVELS),
                ("level_bases", C.c_uint64 * VK_MAX_LEVELS)]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library or raise LibraryMissing -- never substitutes anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `make -C varkoder_b200/csrc` (or __graft_entry__.build()). "
            "varkoder_b200 has no CPU fallback for the hot path.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.vk_abi_version.restype = C.c_int
    L.vk_last_error.restype = C.c_char_p
    L.vk_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.vk_ctx_destroy.argtypes = [vp]
    L.vk_set_mapping.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.vk_upload.argtypes = [vp, vp, C.c_uint64]
    L.vk_attach.argtypes = [vp, vp, C.c_uint64]
    L.vk_parse.argtypes = [vp, C.POINTER(VkStats)]
    L.vk_count.argtypes = [vp, C.POINTER(VkParams), vp, C.POINTER(VkResult)]
    L.vk_prio_hist.argtypes = [vp, C.POINTER(VkParams), vp]
    L.vk_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.vk_render_counts.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.vk_reads_to_images.argtypes = [vp, vp, C.c_uint64, C.c_int, C.POINTER(VkParams), C.c_int, C.c_int,
                                     C.POINTER(VkResult), vp, vp]
    L.vk_device_pixels.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.vk_remap.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp, vp, C.c_int, vp]
    L.vk_base_content.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.vk_last_timings.argtypes = [vp, C.POINTER(C.c_float)]
    L.vk_set_fine_timing.argtypes = [vp, C.c_int]
    L.vk_set_batch_mode.argtypes = [vp, C.c_int]
    L.vk_launch_count.argtypes = [vp]
    L.vk_launch_count.restype = C.c_uint64
    L.vk_bucket_retries.argtypes = [vp]
    L.vk_bucket_retries.restype = C.c_uint64
    L.vk_count_fallbacks.argtypes = [vp]
    L.vk_count_fallbacks.restype = C.c_uint64
    L.vk_synth_fastq.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64,
                                 C.POINTER(C.c_uint64)]
    L.vk_synth_fastq_variable.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.vk_comm_unique_id.argtypes = [vp]
    L.vk_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.vk_comm_destroy.argtypes = [vp]
    L.vk_sharded_reads_to_images.argtypes = [vp, vp, C.c_uint64, C.c_int, C.POINTER(VkParams), C.c_int, C.c_int,
                                             C.POINTER(VkResult), vp, vp]
    L.vk_graph_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("vk_last_error", "vk_launch_count", "vk_bucket_retries", "vk_count_fallbacks"):
            fn.restype = C.c_int
    _lib = L
    return L
