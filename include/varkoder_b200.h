/*
 * varkoder_b200.h -- C ABI of the B200-native varKoder image hot path.
 *
 * Scope: cleaned FASTQ reads -> (sub-sample ladder) -> canonical k-mer counts -> varKode / CGR pixels,
 * i.e. steps C-E of run_clean2img in the reference (varKoder/commands/image.py:1005-1125).  The reference
 * has no FFI for this path: it shells out to reformat.sh, dsk and dsk2ascii and finishes in pandas/numpy.
 * Each entry point below names the reference code it stands in for.  Plain pointers and sizes only; every
 * function returns 0 on success or a negative VK_E* code, with a thread-local message in vk_last_error().
 *
 * All device work of one context runs on one internal CUDA stream of one GPU.  A context is not
 * thread-safe; use one per (thread, device).  Host buffers may be pageable or pinned.
 *
 * K-mer index convention of every histogram that crosses this boundary ("lex index"): first base most
 * significant, A=0 C=1 G=2 T=3, so index = sum code(b_i) * 4^(k-1-i).
 */
#ifndef VARKODER_B200_H
#define VARKODER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VK_ABI_VERSION 2
#define VK_MAX_LEVELS 64 /* ladder levels: 3 per decade, far more than any real sample needs */
#define VK_MIN_K 5       /* ImageCommand.__init__ accepts k in [5, 9] (image.py:1209) */
#define VK_MAX_K 9
#define VK_BREAKLENGTH 500 /* reformat.sh breaklength=500 (image.py:586) */

enum {
    VK_OK = 0,
    VK_EINVAL = -1,   /* bad argument */
    VK_ECUDA = -2,    /* CUDA runtime error (message holds cudaGetErrorString) */
    VK_ENOMEM = -3,
    VK_ESTATE = -4,   /* call order violated (e.g. count before parse) */
    VK_ERANGE = -5    /* input outside supported range (buffer >= 2^40 B; more ladder levels than the caller allowed for) */
};

/* status of the ladder, vk_result.status */
enum {
    VK_LADDER_OK = 0,
    VK_LADDER_LESS_THAN_MIN = 1 /* split_fastq raises Exception("Input file has less than minimum data.") image.py:674-680 */
};

typedef struct vk_ctx vk_ctx;

/* Options of one sample; mirrors the arguments of split_fastq (image.py:629-640) and count_kmers (:727-734). */
typedef struct vk_params {
    int32_t k;             /* kmer_size, 5..9 */
    int32_t is_query;      /* split_fastq(is_query=...): single level, no min_bp check */
    int32_t has_max_bp;    /* 0 <=> max_bp is None (cli.py:496-501 maps "--max-bp 0" to None) */
    int32_t breaklength;   /* reformat.sh breaklength; VK_BREAKLENGTH for parity, 0 disables */
    uint64_t min_bp;
    uint64_t max_bp;
    uint64_t seed;         /* split_fastq(seed=...) reduced mod 2^64 */
    uint64_t read_index_base; /* global index of this buffer's first record (read-sharded samples) */
    uint64_t nsites_override; /* 0: ladder from this buffer's own base count; else the sample-wide total */
    /* How a level's reads are drawn (reformat.sh samplebasestarget=<bp>, image.py:582-596, writes reads until the base
     * target is met).  VK_SAMPLING_CALIBRATED (default of the Python layer): the level's priority threshold is fitted to
     * the sample -- a histogram of the reads' bases over 2^16 priority buckets gives the bucket in which the cumulative
     * base count reaches the target, the threshold is interpolated inside it -- so the realised bases meet the target to
     * within a couple of reads.  VK_SAMPLING_EXPECTED: threshold = bp * 2^64 / nsites, met in expectation only (round 1). */
    int32_t sampling;
    int32_t reserved0;
    /* read-sharded samples: device pointer to the histogram (VK_PRIO_BUCKETS uint64) of the WHOLE sample (vk_prio_hist of every shard,
     * summed by the caller); 0: the histogram of this buffer's own reads */
    uint64_t prio_hist;
} vk_params;
#define VK_SAMPLING_EXPECTED 0
#define VK_SAMPLING_CALIBRATED 1
#define VK_PRIO_BUCKETS 65536

/* What parsing found; mirrors the first loop of split_fastq (image.py:662-667). */
typedef struct vk_stats {
    uint64_t n_bytes;
    uint64_t n_lines;      /* lines as Python's binary line iterator yields them */
    uint64_t n_reads;      /* sequence lines (line index % 4 == 1) */
    uint64_t nsites;       /* sum(len(line) - 1) over sequence lines, exactly as the reference counts */
    uint64_t nsites_true;  /* bases really present (differs by one only for an unterminated last sequence line) */
} vk_stats;

/* Ladder and realised sub-samples; level 0 is the largest. */
typedef struct vk_result {
    vk_stats stats;
    int32_t status;        /* VK_LADDER_* */
    int32_t n_levels;
    uint64_t level_bp[VK_MAX_LEVELS];    /* sites_per_file (image.py:669-695): the target named in the file */
    uint64_t level_reads[VK_MAX_LEVELS]; /* reads realised in each (nested) level, reads shorter than k excluded */
    uint64_t level_bases[VK_MAX_LEVELS]; /* bases realised in each level */
} vk_result;

int vk_abi_version(void);
const char* vk_last_error(void);

int vk_ctx_create(int device, vk_ctx** out);
int vk_ctx_destroy(vk_ctx* ctx);

/*
 * Pixel table of one (k, mapping): replaces get_kmer_mapping / get_cgr (core/utils.py:152-217) and the
 * join + groupby + scatter of make_image (image.py:900-913).  lut[row * side + col] = lex index of a k-mer
 * of the canonical class shown at that pixel of the FINAL image (row = H-1-y, col = x), or -1 if unused.
 * slot: 0..3, lets a context keep several tables (e.g. varKode and cgr) resident.
 */
int vk_set_mapping(vk_ctx* ctx, int slot, int k, int side, const int32_t* lut_host);

/*
 * Input: the uncompressed bytes of <int>/clean_reads/<sample>.fq.gz (image.py:971).
 * vk_upload copies host bytes into a context-owned device buffer (one stream-ordered cudaMemcpyAsync; at PCIe speed
 * when the host buffer is page-locked);
 * vk_attach uses caller-owned device memory (16-byte aligned; readable up to the next 16-byte boundary).
 */
int vk_upload(vk_ctx* ctx, const void* host_bytes, uint64_t n_bytes);
int vk_attach(vk_ctx* ctx, const void* dev_bytes, uint64_t n_bytes);

/* FASTQ framing + base count: the gzip line loop of split_fastq (image.py:662-667). Synchronises. */
int vk_parse(vk_ctx* ctx, vk_stats* out);

/*
 * Ladder (image.py:669-695), seeded nested sub-sampling (stands in for reformat.sh, image.py:577-627) and
 * k-mer counting of every level in one pass over the reads (stands in for dsk, image.py:771-796).
 * seg_hist_dev: nullable caller-owned DEVICE buffer of VK_MAX_LEVELS * 4^k uint64 that receives (is
 * overwritten with) per-segment forward-strand histograms in the internal index order; segment s holds
 * the reads that are in levels 0..s but not in s+1, so sums over ranks/shards are meaningful (NCCL
 * all-reduce between vk_count and vk_render).  NULL uses a context-owned buffer.  Synchronises.
 */
int vk_count(vk_ctx* ctx, const vk_params* params, uint64_t* seg_hist_dev, vk_result* out);

/*
 * Read-sharded samples with VK_SAMPLING_CALIBRATED: the base histogram over the priority buckets of THIS buffer's reads
 * (after vk_parse; params->seed and params->read_index_base name the reads' global indices), ADDED to
 * hist_dev[VK_PRIO_BUCKETS] (caller-owned device memory, zeroed by the caller).  The caller sums the shards' histograms
 * (all-reduce) and passes the result to vk_count as params->prio_hist.  Stands in for the first half of what reformat.sh
 * samplebasestarget does (image.py:582-596): knowing how many bases the reads drawn so far hold.  Synchronises.
 */
int vk_prio_hist(vk_ctx* ctx, const vk_params* params, uint64_t* hist_dev);

/*
 * Canonical fold + pixel mapping + rank scaling for n_levels levels (stands in for dsk2ascii and
 * make_image's numpy half, image.py:875-919).
 * canon_host: nullable, n_levels * 4^k uint64, canon[l][K] = canon[l][rc K] = abundance of the canonical
 *             class of K in level l (lex index) -- what dsk2ascii would print for min(K, rc K).
 * pixels_host: nullable, n_levels * side * side uint8, the PNG pixels (mode "L") of each level.
 * Synchronises.
 */
int vk_render(vk_ctx* ctx, int slot, int k, int n_levels, const uint64_t* seg_hist_dev,
              uint64_t* canon_host, uint8_t* pixels_host);

/*
 * Pixel mapping + rank scaling from canonical counts that are already on the host (lex index, n * 4^k uint64,
 * as returned in canon_host above): the part of make_image after the dsk2ascii dump (image.py:897-919).
 */
int vk_render_counts(vk_ctx* ctx, int slot, int k, int n, const uint64_t* canon_host, uint8_t* pixels_host);

/*
 * Whole path without intermediate host synchronisation: parse -> ladder -> count -> render.
 * text: host bytes (on_device = 0; copied in the timed path) or device bytes (on_device = 1).
 * Outputs as in vk_count / vk_render; pixels_host must hold VK_MAX_LEVELS * side * side bytes unless
 * max_levels_out is smaller (then at most that many levels are copied back).
 */
int vk_reads_to_images(vk_ctx* ctx, const void* text, uint64_t n_bytes, int on_device, const vk_params* params,
                       int slot, int max_levels_out, vk_result* result, uint64_t* canon_host, uint8_t* pixels_host);

/*
 * The pixels of the last vk_reads_to_images / vk_render call as they lie in DEVICE memory (n_levels x side x side
 * uint8, row-major, level 0 first), for a consumer that stays on the GPU: `varKoder query` feeds the images to a
 * classifier (query.py:283-314) and need not go through PNG files.  The pointer belongs to the context and is valid
 * until the next call that renders or until the context is destroyed.
 */
int vk_device_pixels(vk_ctx* ctx, const uint8_t** dev_pixels, int32_t* n_levels, int32_t* side);

/*
 * convert.remap (varKoder/commands/convert.py:34-77) for a batch of images: move pixels from one pixel table to another
 * (varKode <-> cgr).  The join of the two tables is passed as, per OUTPUT pixel, two source pixels (src0, src1; -1 =
 * none) and their multiplicities mult[2 p], mult[2 p + 1] (varkoder_b200/mapping.py: remap_plan).
 * sum_rc = 0: out[p] = in[src0[p]] (0 where src0 < 0).
 * sum_rc = 1: out[p] = (mult0 * in[src0] + mult1 * in[src1]) mod 256, then rescaled over the image as
 *             uint8((a - min) / max * 255) in float64, exactly as the reference does.
 * in_host: n_images x n_in_pixels uint8; out_host: n_images x n_out_pixels uint8.  Synchronises.
 */
int vk_remap(vk_ctx* ctx, int n_images, uint32_t n_in_pixels, uint32_t n_out_pixels, const uint8_t* in_host,
             const int32_t* src0, const int32_t* src1, const uint8_t* mult, int sum_rc, uint8_t* out_host);

/* Per-position base content of the framed reads of the current text -- the numbers fastp's "content_curves" are made
 * of, which the reference reads back from the fastp report to set varkoderBaseFreqSd / varkoderLowQualityFlag
 * (get_basefrequency_sd, varKoder/commands/image.py:49-88; used at :1093-1096).
 * counts_host[(p - pos_begin) * 5 + j], pos_begin <= p < pos_end (at most 64 positions per call):
 *   j = 0..3  reads whose base at position p is A, T, C, G (fastp's classes: byte & 7 == 1, 4, 3, 7),
 *   j = 4     reads that have a position p at all (the curve's denominator; N and anything else only count here).
 * Needs framing: after vk_parse, vk_count or vk_reads_to_images on the current text.  Synchronises. */
int vk_base_content(vk_ctx* ctx, int32_t pos_begin, int32_t pos_end, uint64_t* counts_host);

/* Device time of the last vk_reads_to_images / stage call, per kernel group, in milliseconds (CUDA events on
 * the context stream): [0] upload (H2D), [1] parse, [2] plan+bucket, [3] count kernel, [4] slab reduce+fold,
 * [5] render, [6] read-back (D2H), [7] total. */
int vk_last_timings(vk_ctx* ctx, float* ms8);

/* on: CUDA events are recorded between the kernel groups so that vk_last_timings can split a step.  An event between two
 * kernels is a full stream dependency: it defeats programmatic dependent launch across it and keeps the step from being
 * submitted as one CUDA graph.  off (default) keeps only the first, the post-upload and the last event ([0] upload and
 * [7] total; the other entries read 0). */
int vk_set_fine_timing(vk_ctx* ctx, int on);

/* on: this context is one of several that push samples through the same GPU at once (ImageCommand.process_samples runs
 * its samples through a pool, image.py:1236-1294).  A small sample then takes at most ceil(n_reads / 3000) count CTAs
 * instead of one per SM: the SMs it could not keep busy anyway are left to the kernels of the other samples in flight
 * (config 4, 96 samples of 10-50 Mbp: +8..12 % per batch).  Samples of 450 k reads and more are not affected.  off: default. */
int vk_set_batch_mode(vk_ctx* ctx, int on);

/* Number of kernels this library launched on the context since creation (bench.py's gpu_launches). */
uint64_t vk_launch_count(vk_ctx* ctx);

/* vk_reads_to_images submits a step as ONE CUDA graph (captured once per k / pixel table / level count, replayed for every
 * sample; what differs between samples travels in a device-resident argument block).  *launches: steps submitted as a
 * graph so far; *captures: graphs captured; *state: 1 graphs in use, 0 switched off (VK_GRAPH=0), -1 a capture failed and
 * the context fell back to plain launches.  Any pointer may be NULL. */
int vk_graph_stats(vk_ctx* ctx, uint64_t* launches, uint64_t* captures, int32_t* state);

/* How often a step had to be repeated because a ladder segment received more reads than the region sized from its
 * expected share (8 sigma + slack): always 0 in practice; tests force it with VK_TEST_TIGHT_BUCKETS=1. */
uint64_t vk_bucket_retries(vk_ctx* ctx);

/* How often a count had to be repeated with the exact kernel because a 16-bit shared-memory bin of the fire-and-forget
 * kernel wrapped (k = 8, k = 7 in pair mode; a flood of one k-mer): 0 on ordinary reads. */
uint64_t vk_count_fallbacks(vk_ctx* ctx);

/* Deterministic synthetic FASTQ (bench / tests): fills a DEVICE buffer.
 * Fixed read length L: record r occupies bytes [r*(2L+17), (r+1)*(2L+17)); returns bytes written in *n_out. */
int vk_synth_fastq(vk_ctx* ctx, void* dev_bytes, uint64_t capacity, uint64_t n_bases, int read_len, uint64_t seed,
                   uint64_t first_read, uint64_t* n_out);

/*
 * Read-sharded samples (one sample spread over the GPUs of a box): every rank frames and counts its
 * contiguous shard of the records, and two exchange steps run ON THE CONTEXT'S STREAM between the kernels, with no host
 * synchronisation in between: an all-gather of two integers per rank (records and bases of the shard -> sample-wide base
 * count for the ladder, global index of the shard's first record for the seeded priorities) and ONE all-reduce of the
 * per-segment histograms with the per-segment totals behind them.  NCCL is opened at run time (dlopen of libnccl.so.2, the
 * same library a torch.distributed process already holds); this library does not link against it.
 *
 * vk_comm_unique_id: 128 bytes made on one rank and handed to all (any transport: torch.distributed, MPI, a file).
 * vk_comm_init:      collective over all ranks; one communicator per context.
 * vk_sharded_reads_to_images: collective; arguments as vk_reads_to_images, `text` = this rank's shard (whole records).
 *   Every rank receives the sample-wide result (ladder, realised reads / bases, counts, pixels); stats.n_reads is the
 *   sample's record count, stats.n_bytes / n_lines / nsites describe the local shard.  max_levels_out rows are exchanged;
 *   a ladder with more levels is an error (VK_ERANGE), so pass the bound of the largest sample (16 covers 30 Gbp).
 */
int vk_comm_unique_id(void* out128);
int vk_comm_init(vk_ctx* ctx, const void* unique_id128, int rank, int world);
int vk_comm_destroy(vk_ctx* ctx);
int vk_sharded_reads_to_images(vk_ctx* ctx, const void* text, uint64_t n_bytes, int on_device, const vk_params* params,
                               int slot, int max_levels_out, vk_result* result, uint64_t* canon_host, uint8_t* pixels_host);

/* Same generator with VARIABLE read lengths (the shape of fastp-cleaned, merged
 * reads): lengths uniform min_len..max_len, short_per_10000 / 10000 of the reads shorter than k (0..k-1 bases, empty
 * reads included).  dev_bytes = NULL only reports the size.  *n_out = bytes, *n_bases_out = bases (nullable). */
int vk_synth_fastq_variable(vk_ctx* ctx, void* dev_bytes, uint64_t capacity, uint64_t n_reads, uint64_t seed,
                            uint64_t first_read, int min_len, int max_len, int short_per_10000, int k, uint64_t* n_out,
                            uint64_t* n_bases_out);

#ifdef __cplusplus
}
#endif
#endif /* VARKODER_B200_H */
