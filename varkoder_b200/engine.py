"""Python face of one GPU context of the hot path (thin over the C ABI; all compute is CUDA)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .ladder import LessThanMinimumData, parse_seed
from .mapping import PixelTable

TIMING_KEYS = ("upload", "parse", "plan_bucket", "count", "reduce_fold", "render", "readback", "total")


class VkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"varkoder_b200 error {code}: {msg}")
        self.code = code


@dataclass
class Params:
    """arguments of split_fastq / count_kmers that shape the work (image.py:629-640, 727-734)."""
    k: int = 7
    min_bp: int = 50000
    max_bp: int | None = None
    is_query: bool = False
    seed: int | str | None = None
    breaklength: int = _lib.VK_BREAKLENGTH
    read_index_base: int = 0
    nsites_override: int = 0
    sampling: int = _lib.VK_SAMPLING_CALIBRATED      # thresholds fitted to the base targets (reformat.sh samplebasestarget)
    prio_hist: int = 0                               # read-sharded samples: device pointer of the sample-wide histogram

    def to_c(self):
        p = _lib.VkParams()
        p.k = int(self.k)
        p.is_query = 1 if self.is_query else 0
        p.has_max_bp = 0 if self.max_bp is None else 1
        p.breaklength = int(self.breaklength)
        p.min_bp = int(self.min_bp)
        p.max_bp = 0 if self.max_bp is None else int(self.max_bp)
        p.seed = parse_seed(self.seed)
        p.read_index_base = int(self.read_index_base)
        p.nsites_override = int(self.nsites_override)
        p.sampling = int(self.sampling)
        p.prio_hist = int(self.prio_hist)
        return p


@dataclass
class Result:
    n_bytes: int
    n_lines: int
    n_reads: int
    nsites: int
    nsites_true: int
    status: int
    levels: list        # sites_per_file
    level_reads: list
    level_bases: list
    canon: np.ndarray | None = None      # [L, 4^k] uint64, lex index
    pixels: np.ndarray | None = None     # [L, side, side] uint8

    def raise_if_less_than_min(self):
        if self.status == _lib.VK_LADDER_LESS_THAN_MIN:
            raise LessThanMinimumData()


def _result_from(r, canon=None, pixels=None):
    n = r.n_levels
    return Result(r.stats.n_bytes, r.stats.n_lines, r.stats.n_reads, r.stats.nsites, r.stats.nsites_true,
                  r.status, [int(r.level_bp[i]) for i in range(n)], [int(r.level_reads[i]) for i in range(n)],
                  [int(r.level_bases[i]) for i in range(n)], canon, pixels)


def _host_ptr(buf):
    """(pointer, nbytes, keepalive) of bytes / bytearray / memoryview / numpy uint8 / torch CPU tensor."""
    if hasattr(buf, "data_ptr") and hasattr(buf, "is_cuda"):          # torch tensor
        if buf.is_cuda:
            raise TypeError("expected host memory")
        t = buf.contiguous()
        return t.data_ptr(), t.numel() * t.element_size(), t
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf).view(np.uint8)
    return a.ctypes.data, a.size, a


class Engine:
    """One context = one GPU, one stream.  Not thread-safe."""

    def __init__(self, device=0):
        self._L = _lib.load()
        self._ctx = C.c_void_p()
        self.device = int(device)
        self._check(self._L.vk_ctx_create(self.device, C.byref(self._ctx)))
        self._slots = {}          # (k, method, id(lut)) -> slot
        self._slot_keys = [None] * 4
        self._next_slot = 0
        self._keep = None
        self.text_generation = 0      # bumped whenever the context's resident text is replaced (stages._resident)

    def close(self):
        if self._ctx:
            self._L.vk_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.VK_OK:
            raise VkError(rc, self._L.vk_last_error().decode("utf-8", "replace"))

    # ------------------------------------------------------------------ pixel tables
    def mapping_slot(self, table: PixelTable):
        key = (table.k, table.method, table.lut.ctypes.data)
        if key in self._slots:
            return self._slots[key]
        slot = self._next_slot
        self._next_slot = (self._next_slot + 1) % 4
        old = self._slot_keys[slot]
        if old is not None:
            self._slots.pop(old, None)
        self._check(self._L.vk_set_mapping(self._ctx, slot, table.k, table.side, table.lut.ctypes.data))
        self._slots[key] = slot
        self._slot_keys[slot] = key
        return slot

    # ------------------------------------------------------------------ stages
    def upload(self, host_bytes):
        ptr, n, keep = _host_ptr(host_bytes)
        self.text_generation += 1
        self._check(self._L.vk_upload(self._ctx, ptr, n))
        return n

    def attach(self, dev_ptr, n_bytes, keepalive=None):
        self._keep = keepalive
        self.text_generation += 1
        self._check(self._L.vk_attach(self._ctx, dev_ptr, n_bytes))

    def parse(self):
        s = _lib.VkStats()
        self._check(self._L.vk_parse(self._ctx, C.byref(s)))
        return dict(n_bytes=s.n_bytes, n_lines=s.n_lines, n_reads=s.n_reads, nsites=s.nsites,
                    nsites_true=s.nsites_true)

    def count(self, params: Params, seg_hist_ptr=None):
        r = _lib.VkResult()
        p = params.to_c()
        self._check(self._L.vk_count(self._ctx, C.byref(p), seg_hist_ptr, C.byref(r)))
        return _result_from(r)

    def prio_hist(self, params: Params, hist_ptr):
        """read-sharded samples, calibrated thresholds: ADD the base histogram of this buffer's reads over the 2^16 priority
        buckets to ``hist_ptr`` (device memory, VK_PRIO_BUCKETS uint64, zeroed by the caller before the first shard); after
        ``parse``.  The sum over the shards goes to ``count`` as ``Params.prio_hist``."""
        p = params.to_c()
        self._check(self._L.vk_prio_hist(self._ctx, C.byref(p), hist_ptr))

    def render(self, table: PixelTable | None, k, n_levels, seg_hist_ptr=None, want_canon=True):
        nk = 4 ** k
        canon = np.empty((n_levels, nk), dtype=np.uint64) if want_canon else None
        pixels = None
        slot = 0
        if table is not None:
            slot = self.mapping_slot(table)
            pixels = np.empty((n_levels, table.side, table.side), dtype=np.uint8)
        if n_levels:
            self._check(self._L.vk_render(self._ctx, slot, k, n_levels, seg_hist_ptr,
                                          canon.ctypes.data if canon is not None else None,
                                          pixels.ctypes.data if pixels is not None else None))
        return canon, pixels

    def render_counts(self, table: PixelTable, canon):
        canon = np.ascontiguousarray(canon, dtype=np.uint64)
        if canon.ndim == 1:
            canon = canon[None, :]
        n = canon.shape[0]
        slot = self.mapping_slot(table)
        pixels = np.empty((n, table.side, table.side), dtype=np.uint8)
        self._check(self._L.vk_render_counts(self._ctx, slot, table.k, n, canon.ctypes.data, pixels.ctypes.data))
        return pixels

    def reads_to_images(self, text, params: Params, table: PixelTable, on_device=False, n_bytes=None,
                        max_levels=16, want_canon=False):
        """whole path, one host synchronisation.  text: host buffer, or a device pointer (int) with n_bytes."""
        slot = self.mapping_slot(table)
        if on_device:
            ptr, n, keep = int(text), int(n_bytes), None
        else:
            ptr, n, keep = _host_ptr(text)
        nk = 4 ** params.k
        L = _lib.VK_MAX_LEVELS
        pixels = np.empty((L, table.side, table.side), dtype=np.uint8)
        canon = np.empty((L, nk), dtype=np.uint64) if want_canon else None
        r = _lib.VkResult()
        p = params.to_c()
        self.text_generation += 1
        self._check(self._L.vk_reads_to_images(self._ctx, ptr, n, 1 if on_device else 0, C.byref(p), slot,
                                               int(max_levels), C.byref(r),
                                               canon.ctypes.data if canon is not None else None, pixels.ctypes.data))
        del keep
        nl = r.n_levels
        return _result_from(r, canon[:nl] if canon is not None else None, pixels[:nl])

    # ------------------------------------------------------------------ read-sharded samples
    def comm_init(self, group=None):
        """Give this context an NCCL communicator over the ranks of a torch.distributed group (collective).  The 128-byte
        id is made by rank 0 and handed round with the group's own broadcast; after that the library talks to NCCL itself,
        on its own stream (vk_sharded_reads_to_images)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            self._check(self._L.vk_comm_unique_id(buf))
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = C.create_string_buffer(box[0], 128)
        self._check(self._L.vk_comm_init(self._ctx, uid, rank, world))
        self.comm_world, self.comm_rank = world, rank

    def comm_destroy(self):
        self._check(self._L.vk_comm_destroy(self._ctx))
        self.comm_world = None

    def sharded_reads_to_images(self, text, params: Params, table: PixelTable, on_device=False, n_bytes=None,
                                max_levels=16, want_canon=False):
        """collective form of reads_to_images for ONE sample whose records are spread over the ranks (this rank passes its
        shard): framing, all-gather of the shard sizes, counting, ONE all-reduce of the histograms, images -- a single
        enqueue on the context's stream, one host synchronisation at the end.  Needs comm_init() first."""
        slot = self.mapping_slot(table)
        if on_device:
            ptr, n, keep = int(text), int(n_bytes), None
        else:
            ptr, n, keep = _host_ptr(text)
        nk = 4 ** params.k
        L = _lib.VK_MAX_LEVELS
        pixels = np.empty((L, table.side, table.side), dtype=np.uint8)
        canon = np.empty((L, nk), dtype=np.uint64) if want_canon else None
        r = _lib.VkResult()
        p = params.to_c()
        self.text_generation += 1
        self._check(self._L.vk_sharded_reads_to_images(self._ctx, ptr, n, 1 if on_device else 0, C.byref(p), slot,
                                                       int(max_levels), C.byref(r),
                                                       canon.ctypes.data if canon is not None else None, pixels.ctypes.data))
        del keep
        nl = r.n_levels
        return _result_from(r, canon[:nl] if canon is not None else None, pixels[:nl])

    def remap(self, images, src0, src1, mult, n_out_shape, sum_rc=False):
        """batch of uint8 images [n, H_in, W_in] -> [n, H_out, W_out] through a remap plan (mapping.remap_plan)"""
        imgs = np.ascontiguousarray(images, dtype=np.uint8)
        if imgs.ndim == 2:
            imgs = imgs[None]
        n, n_in = imgs.shape[0], imgs.shape[1] * imgs.shape[2]
        src0 = np.ascontiguousarray(src0, dtype=np.int32)
        src1 = np.ascontiguousarray(src1, dtype=np.int32)
        mult = np.ascontiguousarray(mult, dtype=np.uint8)
        n_out = int(n_out_shape[0]) * int(n_out_shape[1])
        if src0.size != n_out or src1.size != n_out or mult.size != 2 * n_out:
            raise ValueError("remap plan does not match the output shape")
        out = np.empty((n, int(n_out_shape[0]), int(n_out_shape[1])), dtype=np.uint8)
        self._check(self._L.vk_remap(self._ctx, n, n_in, n_out, imgs.ctypes.data, src0.ctypes.data, src1.ctypes.data,
                                     mult.ctypes.data, 1 if sum_rc else 0, out.ctypes.data))
        return out

    def base_content(self, pos_begin=5, pos_end=40):
        """per-position base content of the framed reads: uint64 [pos_end - pos_begin, 5], columns A, T, C, G, reads
        reaching the position (vk_base_content; the input of quality.base_frequency_sd)"""
        out = np.zeros((int(pos_end) - int(pos_begin), 5), dtype=np.uint64)
        self._check(self._L.vk_base_content(self._ctx, int(pos_begin), int(pos_end), out.ctypes.data))
        return out

    def device_pixels(self):
        """the images of the last render as a torch uint8 CUDA tensor [levels, side, side] that ALIASES the context's
        buffer (valid until the next render): the hand-off to a classifier that stays on the GPU (query.py:283-314)."""
        import torch
        ptr, nl, side = C.c_void_p(), C.c_int32(), C.c_int32()
        self._check(self._L.vk_device_pixels(self._ctx, C.byref(ptr), C.byref(nl), C.byref(side)))

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (nl.value, side.value, side.value), "typestr": "|u1",
                                      "data": (ptr.value, False), "version": 2, "strides": None}
        if nl.value == 0:
            return torch.empty((0, side.value, side.value), dtype=torch.uint8, device=f"cuda:{self.device}")
        return torch.as_tensor(v, device=f"cuda:{self.device}")

    # ------------------------------------------------------------------ misc
    def timings(self):
        ms = (C.c_float * 8)()
        self._check(self._L.vk_last_timings(self._ctx, ms))
        return dict(zip(TIMING_KEYS, [float(x) for x in ms]))

    def set_fine_timing(self, on):
        """events between kernel groups (per-kernel split of timings()) on/off; off lets dependent launches overlap"""
        self._check(self._L.vk_set_fine_timing(self._ctx, 1 if on else 0))

    def set_batch_mode(self, on):
        """this engine is one of several that work on one GPU at once: small samples leave the SMs they cannot fill to
        the others (vk_set_batch_mode)"""
        self._check(self._L.vk_set_batch_mode(self._ctx, 1 if on else 0))

    def launch_count(self):
        return int(self._L.vk_launch_count(self._ctx))

    def graph_stats(self):
        """(steps submitted as one CUDA graph, graphs captured, state: 1 on / 0 off / -1 capture failed)"""
        a, b, st = C.c_uint64(), C.c_uint64(), C.c_int32()
        self._check(self._L.vk_graph_stats(self._ctx, C.byref(a), C.byref(b), C.byref(st)))
        return int(a.value), int(b.value), int(st.value)

    def count_fallbacks(self):
        return int(self._L.vk_count_fallbacks(self._ctx))

    def bucket_retries(self):
        return int(self._L.vk_bucket_retries(self._ctx))

    def synth_fastq(self, dev_ptr, capacity, n_bases, read_len=150, seed=0, first_read=0):
        n = C.c_uint64()
        self._check(self._L.vk_synth_fastq(self._ctx, dev_ptr, capacity, n_bases, read_len, seed, first_read,
                                           C.byref(n)))
        return int(n.value)

    def synth_fastq_variable(self, dev_ptr, capacity, n_reads, seed=0, first_read=0, min_len=60, max_len=280,
                             short_frac=0.01, k=7):
        """device twin of synth.variable(); dev_ptr=None only reports the size.  -> (bytes, bases)"""
        n, nb = C.c_uint64(), C.c_uint64()
        self._check(self._L.vk_synth_fastq_variable(self._ctx, dev_ptr, capacity, n_reads, seed, first_read, min_len,
                                                    max_len, int(short_frac * 10000), k, C.byref(n), C.byref(nb)))
        return int(n.value), int(nb.value)
