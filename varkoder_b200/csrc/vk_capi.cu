// vk_capi.cu -- C ABI (include/varkoder_b200.h) over the sm_100a kernels of the varKoder image hot path.
// Host side of one context: device buffers, one stream, kernel sequencing, timing events.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <string>
#include <vector>

#include <nccl.h>      // types only: the library is opened at run time (dlopen), nothing links against it

#include "vk_bucket.cuh"
#include "vk_common.cuh"
#include "vk_count.cuh"
#include "vk_countu.cuh"
#include "vk_countt.cuh"
#include "vk_countt9.cuh"
#include "vk_image.cuh"
#include "vk_parse.cuh"
#include "vk_quality.cuh"
#include "vk_sample.cuh"
#include "vk_synth.cuh"

namespace {

thread_local std::string g_err;

struct CudaError {
    cudaError_t e;
    const char* what;
    int line;
};
#define CU(x)                                                 \
    do {                                                      \
        cudaError_t e_ = (x);                                 \
        if (e_ != cudaSuccess) throw CudaError{e_, #x, __LINE__}; \
    } while (0)

struct ApiError {
    int code;
    std::string msg;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;      // elements
    bool ensure(size_t n)      // true: the buffer moved (captured graphs that hold its address are stale)
    {
        if (n <= cap) return false;
        if (p) CU(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        CU(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        return true;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Mapping {
    int k = 0, side = 0;
    DevBuf<int32_t> lut;
};

constexpr size_t kPlanPad = 8192;
static_assert(sizeof(vk::Plan) <= kPlanPad, "Plan must fit its slot of the outbox");

enum { EV_START = 0, EV_UPLOAD, EV_PARSE, EV_BUCKET, EV_COUNT, EV_FOLD, EV_RENDER, EV_DONE, EV_N };

}  // namespace

struct vk_ctx {
    int device = 0;
    int n_sms = vk::kNumSMs;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_N] = {};
    bool ev_valid[EV_N] = {};
    uint64_t launches = 0;
    int count_threads = 1024, count_ctas_per_sm = 1;      // count-kernel launch shape (VK_COUNT_THREADS / VK_COUNT_CTAS)
    int count_extra = 0;            // k <= 8: CTAs launched beyond the resident ones (VK_COUNT_EXTRA); they start when the
                                    // CTAs of the small segments have finished and level the tail of the big ones.
                                    // Measured at k = 7: 0 -> 139.0 us, 1 -> 138.4, 2 -> 139.3, 3 and more -> 155+: off.
    int count_extra9 = 6;           // k = 9: extra PAIRS (VK_COUNT_EXTRA9).  74 pairs over 11 segments leave six pairs with
                                    // 1.5 % of the work; 2.76 ms (0) -> 2.64 (2) -> 2.59 (6) -> 2.66 (10) per Gbp
    int reads_per_cta_env = 0;      // what VK_COUNT_READS_PER_CTA asked for (vk_set_batch_mode(0) goes back to it)
    int reads_per_cta = 0;          // VK_COUNT_READS_PER_CTA: > 0 caps the count CTAs of a small sample (plan_kernel)
    int count_grid() const { return n_sms * count_ctas_per_sm + count_extra; }

    // "outbox": [Plan, padded to kPlanPad bytes][pixels of every level] contiguous on the device and mirrored in pinned
    // host memory, so that the fused path reads everything back with ONE device-to-host copy
    uint8_t* out_d = nullptr;
    uint8_t* out_h = nullptr;       // pinned
    size_t out_cap = 0;             // pixel bytes the outbox holds
    vk::Plan* plan_d = nullptr;     // = out_d
    vk::Plan* plan_h = nullptr;     // = out_h
    uint8_t* pix_d() const { return out_d + kPlanPad; }
    uint8_t* pix_h() const { return out_h + kPlanPad; }

    DevBuf<uint8_t> text_own;
    const uint8_t* text = nullptr;  // device
    uint64_t n_bytes = 0;
    bool have_text = false, parsed = false, counted = false;
    int last_levels = 0, last_side = 0;     // what pix_d() holds (vk_device_pixels)
    bool exact_layout = false;      // segment regions sized for every read (set after a bucket overflow)
    bool use_count16 = true;        // VK_COUNT16=0: k = 8 with global atomics instead of 16-bit shared-memory bins
    bool use_pairs = false;         // VK_COUNT_PAIRS=1: k = 7 through 8-mer pairs in 16-bit bins (exact; halves the
                                    // shared-memory traffic but costs more instructions: 168 vs 143 us, profiles/r01_notes.md)
    bool use_lanes = false;         // VK_COUNT_LANES=1: k = 7 with one read per lane, pairs, uniform fast path (countu_kernel)
    bool k7_lanes = false;          // k = 7, automatic choice: the context's last sample was one for countt_kernel (reads of one length)
    bool in_sharded = false;        // inside vk_sharded_reads_to_images: the ranks repeat steps together, the choice is made on the device
    int lanes9_mode = -1;           // VK_COUNT_LANES9: -1 (default) countt9_kernel or count9h_kernel, by the sample (the context's k7_lanes); 0 count9h_kernel; 1 countt9_kernel always
    // k = 9 through countt9_kernel in this step?  (one CTA per SM: 74 pairs, no late pairs -- the plan's pair count follows)
    bool use_countt9() const
    {
        if (!use_count16 || use_packed || !use_fast || count_is_safe()) return false;
        return lanes9_mode > 0 || (lanes9_mode < 0 && k7_lanes && !in_sharded);
    }
    unsigned countt_knobs = 0;      // VK_COUNTT_KNOBS: experiments of countt_kernel (vk_countt.cuh)
    int lanes_mode = -1;            // VK_COUNT_LANES: -1 (default) countt_kernel or the flat-lane kernel, by the sample; 0 the flat-lane kernel;
                                    // 1 countu_kernel; 2 countt_kernel for every sample; 3 countt_kernel with IMAD.HI shifts in the classification (experiment)
    bool use_fast = true;           // VK_COUNT_FAST=0: 16-bit bins always through returning adds + drains (exact in one go)
    bool count_safe = false;        // set for the repeat of a step whose fire-and-forget count reported a wrapped bin
    // Texts of this size and more go to the exact kernels at once: the smallest text whose 16-bit bins wrapped in this
    // context (a 15 Gbp shard puts 50 M pairs into the 2^15 words of a CTA; a 7-mer at ten times the mean wraps one), less
    // a quarter.  Without it every step of such a sample would run the wrapped count first (36 against 19.5 ms at 15 Gbp).
    // The rule is forgotten after 32 steps it sent to the exact kernels (the samples may have changed).
    uint64_t safe_from_bytes = ~0ull;
    uint32_t safe_by_size_steps = 0;
    bool count_is_safe() const { return count_safe || n_bytes >= safe_from_bytes; }
    uint64_t count_fallbacks = 0;
    uint64_t lanes_flips = 0;       // steps repeated because countt_kernel refused the sample
    bool use_pdl = true;            // VK_PDL=0 disables programmatic dependent launch
    bool test_tight = false;        // VK_TEST_TIGHT_BUCKETS=1: undersized regions, exercises the retry (tests only)
    uint64_t bucket_retries = 0;
    int counted_k = 0;

    // per-step arguments of the kernels (vk::StepArgs): written by the host into pinned memory, copied to the device by
    // the first node of every step
    vk::StepArgs* args_d = nullptr;
    vk::StepArgs* args_h = nullptr;
    bool use_packed = false;        // VK_PACKED=1 (measured, a net loss); 0: the count kernels classify the text themselves (round-1 path) instead of
                                    // reading the 2-bit codes + validity bits the framing pass writes
    bool use_chunks = false;        // VK_CHUNKS=1: k <= 7 counts from a chunk table the scatter kernel writes (one descriptor per
                                    // 32-byte chunk) instead of working the chunk stream out of the segment-sorted READ table in
                                    // the count kernel.  Measured (profiles/r02_notes.md): count 139 -> 135 us, scatter 23 -> 51 us: off
    DevBuf<uint64_t> chunks;        // chunk table: one descriptor per 32-byte chunk of every selected read
    bool chunk_mode(int k) const { return use_chunks && k <= 7 && !use_packed && n_bytes < vk::kChunkMaxText; }
    DevBuf<uint4> codes;            // 2-bit codes of every text byte, 16 B per 64 text bytes (parse_mask_kernel<true>)
    DevBuf<uint2> valid;            // validity bits, 8 B per 64 text bytes
    // one step = one CUDA graph: captured once per (k, pixel table, levels, layout) and replayed for every sample
    bool zero_unused_rows = true;   // the count enqueue writes zeros to the segment rows beyond the ladder (false inside the fused step)
    bool use_graph = true;          // VK_GRAPH=0: launch the kernels of a step one by one
    struct StepGraph {
        int k, slot, side, max_levels, exact, packed;
        uint64_t generation;
        cudaGraphExec_t exec;
        uint32_t kernels;
    };
    std::vector<StepGraph> graphs;
    uint64_t generation = 0;        // bumped whenever a device buffer moves
    uint64_t graph_launches = 0, graph_captures = 0;
    int graph_failed = 0;           // capture or instantiation failed once: stay on plain launches
    bool capturing = false;
    uint32_t captured_kernels = 0;
    struct TraceEv { cudaEvent_t e; const void* fn; unsigned grid, block; };
    std::vector<TraceEv> trace;
    bool trace_each = false;        // VK_TRACE_EACH=1

    // read-sharded samples: an NCCL communicator of this context's own (vk_comm_init), collectives on the context's stream
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    DevBuf<unsigned long long> shard_table;      // [world][2] (records, bases) after the all-gather; [world] = send slot

    DevBuf<uint64_t> tile_status, masks, starts, ends, sorted;      // tile_status = exclusive newline prefix per tile
    DevBuf<uint32_t> tile_count, warp_count;
    DevBuf<uint32_t> slabs;
    DevBuf<unsigned long long> seg_hist, canon, vals, bins;
    DevBuf<uint8_t> remap_in, remap_out, remap_mult;
    DevBuf<int32_t> remap_src;
    DevBuf<unsigned long long> content;
    DevBuf<uint64_t> long_list;                  // indices of the reads of 2^24 bases or more of the current sample (vk_bucket.cuh)
    DevBuf<unsigned long long> prio_hist;        // bases per priority bucket of the current sample (vk_sample.cuh)
    DevBuf<uint64_t> synth_off;     // vk_synth_fastq_variable: record offsets
    Mapping maps[4];

    bool fine_timing = false;       // vk_set_fine_timing: events between the kernel groups (they serialise the stream and keep
                                    // the step from being submitted as one graph); off: first / upload / last event only
    void mark(int e)
    {
        if (!fine_timing && e != EV_START && e != EV_UPLOAD && e != EV_DONE) { ev_valid[e] = false; return; }
        if (capturing) return;
        CU(cudaEventRecord(ev[e], stream));
        ev_valid[e] = true;
    }
};

namespace {

void set_device(vk_ctx* c) { CU(cudaSetDevice(c->device)); }

// Every kernel of the path starts with griddepcontrol.wait (vk::pdl_wait) and is launched with programmatic stream
// serialisation: the next kernel's CTAs are scheduled while the previous kernel drains, so launch latency and kernel
// prologues overlap the tail of their predecessor (a dozen dependent launches per sample).  VK_PDL=0 turns it off.
template <typename... KArgs, typename... Args>
void launch(vk_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = c->use_pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    if (c->capturing) ++c->captured_kernels; else ++c->launches;
    if (c->trace_each && !c->capturing) {                 // VK_TRACE_EACH=1: an event behind every launch (debugging aid)
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        CU(cudaEventRecord(e, c->stream));
        c->trace.push_back({e, (const void*)kernel, grid.x, block.x});
    }
}

void trace_dump(vk_ctx* c)
{
    if (!c->trace_each || c->trace.empty()) return;
    cudaStreamSynchronize(c->stream);
    float prev = 0.f;
    for (size_t i = 0; i < c->trace.size(); ++i) {
        float t = 0.f;
        if (c->ev_valid[EV_START]) cudaEventElapsedTime(&t, c->ev[EV_START], c->trace[i].e);
        cudaFuncAttributes fa;
        const char* nm = "?";
        (void)fa;
        fprintf(stderr, "[vk trace] %2zu  +%8.1f us  (at %8.1f us)  grid %u x %u  fn %p %s\n", i, (t - prev) * 1e3, t * 1e3,
                c->trace[i].grid, c->trace[i].block, c->trace[i].fn, nm);
        prev = t;
        cudaEventDestroy(c->trace[i].e);
    }
    c->trace.clear();
}


// ---- per-step argument block ----------------------------------------------------------------------------
// Fills the pinned StepArgs and enqueues its copy to the device (the first node of a step).  `n_bytes`, the text pointer
// and the sample's options are the ONLY things that vary between steps of one context.
void enqueue_args(vk_ctx* c, const vk_params* params)
{
    using namespace vk;
    StepArgs& a = *c->args_h;
    memset(&a, 0, sizeof(a));
    a.text = c->text;
    a.n_bytes = c->n_bytes;
    a.n_tiles = (uint32_t)((c->n_bytes + kParseTileBytes - 1) / kParseTileBytes);
    if (params) a.pa.p = *params;
    a.pa.n_bytes = c->n_bytes;
    a.pa.cap_reads = c->starts.cap;
    a.pa.cap_sorted = c->sorted.cap;
    a.pa.n_count_ctas = (uint32_t)c->count_grid();
    a.pa.reads_per_cta = (uint32_t)c->reads_per_cta;
    if (params && params->k == 9 && c->use_count16)                             // k = 9 counts in CTA pairs (count9h_kernel)
        a.pa.n_count_ctas = (uint32_t)(c->n_sms * c->count_ctas_per_sm / 2 + (c->use_countt9() ? 0 : c->count_extra9));
    a.pa.exact_layout = c->exact_layout ? 1u : 0u;
    a.pa.test_tight = c->test_tight ? 1u : 0u;
    a.pa.shard_table = nullptr;
    a.pa.cap_chunks = (params && c->chunk_mode(params->k)) ? c->chunks.cap : 0;
}
// first kernel of a step: StepArgs host -> device, and (rescan) the framing accumulators cleared (vk_parse.cuh K0)
void enqueue_begin(vk_ctx* c, bool rescan)
{
    launch(c, vk::step_begin_kernel, dim3(8), dim3(1024), 0, reinterpret_cast<const uint32_t*>(c->args_h),
           reinterpret_cast<uint32_t*>(c->args_d), c->plan_d, c->tile_count.p, (uint32_t)c->tile_count.cap, rescan ? 1 : 0,
           c->prio_hist.p);
    CU(cudaGetLastError());
}

// buffers of the framing pass for a text of n bytes (never allocates inside a graph capture: called before it)
void ensure_parse_buffers(vk_ctx* c, uint64_t n)
{
    using namespace vk;
    const uint32_t n_tiles = (uint32_t)((n + kParseTileBytes - 1) / kParseTileBytes);
    c->generation += c->tile_status.ensure((size_t)n_tiles + 1);
    c->generation += c->masks.ensure((size_t)n_tiles * kParseThreads + 1);
    c->generation += c->tile_count.ensure((size_t)n_tiles + 1);
    c->generation += c->warp_count.ensure((size_t)n_tiles * kParseWarps + 1);
    if (c->use_packed) {
        c->generation += c->codes.ensure((size_t)n_tiles * kParseThreads + 1);
        c->generation += c->valid.ensure((size_t)n_tiles * kParseThreads + 1);
    }
}

// ---- K1: framing -------------------------------------------------------------------------------------
// Grids do not depend on the sample (grid-stride loops over sa->n_tiles), so the same launches serve every step.
void enqueue_parse(vk_ctx* c, bool rescan = true, bool with_plan = true)
{
    using namespace vk;
    enqueue_begin(c, rescan);
    if (rescan) {
        const int grid = c->n_sms * 16;
        if (c->use_packed)
            launch(c, parse_mask_kernel<true>, dim3(grid), dim3(kParseThreads), 0, (const StepArgs*)c->args_d, c->masks.p,
                   c->tile_count.p, c->warp_count.p, c->codes.p, c->valid.p);
        else
            launch(c, parse_mask_kernel<false>, dim3(grid), dim3(kParseThreads), 0, (const StepArgs*)c->args_d, c->masks.p,
                   c->tile_count.p, c->warp_count.p, (uint4*)nullptr, (uint2*)nullptr);
        CU(cudaGetLastError());
        launch(c, parse_scan_kernel, dim3(1), dim3(1024), 0, c->tile_count.p, (const StepArgs*)c->args_d, c->tile_status.p, c->plan_d);
        CU(cudaGetLastError());
        const int egrid = c->n_sms * 32;
        launch(c, parse_emit_kernel, dim3(egrid), dim3(256), 0, c->masks.p, c->tile_status.p, c->warp_count.p,
               (const StepArgs*)c->args_d, 0, c->starts.p, c->ends.p, c->plan_d);
        CU(cudaGetLastError());
    }
    if (!with_plan) return;        // read-sharded: the shards' statistics are exchanged first
    launch(c, plan_kernel, dim3(1), dim3(64), 0, (const StepArgs*)c->args_d, c->starts.p, c->ends.p, c->plan_d);
    CU(cudaGetLastError());
}

// entries the segment-sorted read table needs for n reads in n_levels segments (plan_kernel's layout rule)
uint64_t sorted_need(uint64_t n_reads, uint64_t n_levels, bool exact)
{
    const uint64_t stride = (n_reads + vk::kUnitReads - 1) / vk::kUnitReads * vk::kUnitReads + vk::kUnitReads;
    if (exact) return n_levels * stride;
    uint64_t root = 1;
    while (root * root < n_reads) root += root < 1024 ? 1 : root / 64 + 1;       // >= sqrt(n_reads)
    return n_reads + n_levels * (8 * root + 1024 + 2 * vk::kUnitReads);
}

void ensure_tables_for(vk_ctx* c, uint64_t n_reads_hint)
{
    c->generation += c->starts.ensure(n_reads_hint);
    c->generation += c->ends.ensure(n_reads_hint);
    if (!c->exact_layout) c->generation += c->sorted.ensure(sorted_need(n_reads_hint, vk::kMaxLevels, false));
}

void ensure_count_buffers(vk_ctx* c, int k)
{
    const uint32_t nk = 1u << (2 * k);
    // chunk table: ~ one descriptor per 32 bases + two per read with the 8-sigma slack of every region; from the text
    // size alone (half of it bases, records of ~300 bytes) -- a sample that needs more says so (chunk_table_small)
    if (c->chunk_mode(k)) c->generation += c->chunks.ensure((size_t)(c->n_bytes / 40) + (1u << 18));
    if (k <= 7 || (k == 8 && c->use_count16)) c->generation += c->slabs.ensure((size_t)c->count_grid() * nk);
    if (k == 9 && c->use_count16)
        c->generation += c->slabs.ensure((size_t)2 * (c->n_sms * c->count_ctas_per_sm / 2 + c->count_extra9) * 65536u);
}

// dynamic shared memory of every count kernel variant, set once per context (not inside a capture)
template <int K>
void prepare_count_kernels()
{
    using namespace vk;
    constexpr uint32_t NK = 1u << (2 * K);
    if constexpr (K <= 7) {
        const int smem = (int)(0x10000 + (size_t)(NK + 32) * sizeof(uint32_t));
        CU(cudaFuncSetAttribute(count_kernel<K, kSmem32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(count_kernel<K, kSmem32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(countd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if constexpr (K == 7) {
            CU(cudaFuncSetAttribute(countp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((32768 + 16384) * sizeof(uint32_t))));
            CU(cudaFuncSetAttribute(countu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)countu_smem_bytes()));
            CU(cudaFuncSetAttribute((countt_kernel<16, 0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)countt_smem_bytes<16>()));
            CU(cudaFuncSetAttribute((countt_kernel<16, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)countt_smem_bytes<16>()));
            CU(cudaFuncSetAttribute((countt_kernel<16, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)countt_smem_bytes<16>()));
        }
    }
    if constexpr (K == 7 || K == 8) {
        const int smem = (int)((size_t)(32768 + (K == 7 ? 16384 : 0)) * sizeof(uint32_t));
        CU(cudaFuncSetAttribute((count16_kernel<K, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute((count16_kernel<K, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute((count16_kernel<K, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute((count16_kernel<K, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    if constexpr (K == 9) {
        const int smem = (int)((size_t)32768 * sizeof(uint32_t) + 2048);
        CU(cudaFuncSetAttribute(countt9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)countt9_smem_bytes()));
        CU(cudaFuncSetAttribute(count9h_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(count9h_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
}
void prepare_kernels()
{
    prepare_count_kernels<5>();
    prepare_count_kernels<6>();
    prepare_count_kernels<7>();
    prepare_count_kernels<8>();
    prepare_count_kernels<9>();
    CU(cudaFuncSetAttribute(vk::bucket_scatter_kernel<vk::kBucketItemsChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(vk::kStageChunks * sizeof(uint64_t))));
    // the cluster image kernel: up to 8 slices of 2048 keys
    CU(cudaFuncSetAttribute(vk::image_kernel_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)((size_t)vk::kImgCluster * 2048 * sizeof(unsigned long long))));
}

// ---- K1b + K2 + K3 -----------------------------------------------------------------------------------
// countt_kernel in its two forms.  The CTAs share a text in proportion to the segments' sizes: below kEpochBytes a CTA sees
// at most 29 MB of text = 6.5 M pairs, and a 16-bit word of its table then wraps (2^15 hits) only when one pair of bins holds
// 0.5 % of all pairs -- what the overflow check and the exact recount are for -- so the kernel for such texts carries no epoch
// code (the same loop with the flushes compiled in measured 108.5 against 106.1 us); larger texts flush every kTEpochUnits units.
constexpr uint64_t kEpochBytes = 4ull << 30;
void launch_countt(vk_ctx* c, dim3 grid, const vk::StepArgs* sa, const uint64_t* srt, uint32_t pol)
{
    using namespace vk;
    if (c->n_bytes < kEpochBytes && (pol >> 16) == 0u)
        launch(c, (countt_kernel<16, 0, false>), grid, dim3(512), countt_smem_bytes<16>(), sa, srt, c->plan_d, c->slabs.p, pol);
    else
        launch(c, (countt_kernel<16, 0, true>), grid, dim3(512), countt_smem_bytes<16>(), sa, srt, c->plan_d, c->slabs.p, pol);
}

template <int K, bool PACKED>
void launch_count(vk_ctx* c, unsigned long long* seg_hist)
{
    using namespace vk;
    constexpr uint32_t NK = 1u << (2 * K);
    const dim3 grid(c->count_grid()), block(c->count_threads);
    const uint64_t total = (uint64_t)kMaxLevels * NK;
    const StepArgs* sa = c->args_d;
    const PackedSrc pk = {reinterpret_cast<const uint2*>(c->codes.p), reinterpret_cast<const uint32_t*>(c->valid.p)};
    if constexpr (K == 7) {
        // k = 7, one read per lane, pairs (vk_countu.cuh); a wrapped bin repeats the count with the u32 kernel
        if (c->use_lanes && !PACKED && !c->chunk_mode(7) && c->use_fast && !c->count_is_safe()) {
            const uint32_t pol = 1u | ((c->countt_knobs & ~1u) << 8);
            const uint64_t* srt = (const uint64_t*)c->sorted.p;
            if (c->lanes_mode == 2) launch_countt(c, grid, sa, srt, pol);
            else if (c->lanes_mode == 3) launch(c, (countt_kernel<16, 1, true>), grid, dim3(512), countt_smem_bytes<16>(), sa, srt, c->plan_d, c->slabs.p, pol);
            else
                launch(c, countu_kernel, grid, block, countu_smem_bytes(), sa, (const uint64_t*)c->sorted.p, c->plan_d, c->slabs.p, 1u);
            c->mark(EV_COUNT);
            launch(c, reduce_slabs_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, c->slabs.p, c->plan_d, NK, seg_hist, c->zero_unused_rows ? 1 : 0);
            return;
        }
        // k = 7 in read-aligned pairs from the chunk table (countp_kernel); a wrapped bin repeats the count with the u32 kernel
        if (c->use_pairs && c->chunk_mode(7) && c->use_fast && !c->count_is_safe()) {
            const size_t smem = (size_t)(32768 + 16384) * sizeof(uint32_t);
            launch(c, countp_kernel, grid, block, smem, sa, (const uint64_t*)c->chunks.p, c->plan_d, c->slabs.p);
            c->mark(EV_COUNT);
            launch(c, reduce_slabs_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, c->slabs.p, c->plan_d, NK, seg_hist, c->zero_unused_rows ? 1 : 0);
            return;
        }
    }
    if constexpr (K == 7 || K == 8) {
        // 16-bit bins in shared memory: k = 8 directly, k = 7 through pairs (vk_count.cuh).  The fire-and-forget form by
        // default; after it reported a wrapped bin (count_safe) the exact form -- for k = 7 that is the u32 kernel below.
        const bool fast = c->use_fast && !c->count_is_safe();
        if (K == 8 ? c->use_count16 : (c->use_pairs && !c->chunk_mode(7) && (fast || !c->use_fast))) {
            const size_t smem = (size_t)(32768 + (K == 7 ? 16384 : 0)) * sizeof(uint32_t);
            if (fast) launch(c, (count16_kernel<K, PACKED, true>), grid, block, smem, sa, pk, c->sorted.p, c->plan_d, c->slabs.p);
            else launch(c, (count16_kernel<K, PACKED, false>), grid, block, smem, sa, pk, c->sorted.p, c->plan_d, c->slabs.p);
            c->mark(EV_COUNT);
            launch(c, reduce_slabs_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, c->slabs.p, c->plan_d, NK, seg_hist, c->zero_unused_rows ? 1 : 0);
            return;
        }
    }
    if constexpr (K == 9) {
        if (c->use_count16) {
            // canonical classes in two halves, one per CTA of a pair (vk_count.cuh)
            const bool t9 = c->use_countt9();
            const unsigned pairs = (unsigned)(c->n_sms * c->count_ctas_per_sm / 2 + (t9 ? 0 : c->count_extra9));
            const size_t smem = (size_t)32768 * sizeof(uint32_t) + 2048;      // + padding to a 2 KiB shared address
            // countt9_kernel for samples of one read length, count9h_kernel for the others: the context goes by its last sample,
            // as for k = 7 (count9h_kernel reports a sample that was one for countt9_kernel, countt9_kernel refuses one that is not)
            if (t9)
                launch(c, countt9_kernel, dim3(2 * pairs), dim3(512), countt9_smem_bytes(), sa, (const uint64_t*)c->sorted.p, c->plan_d, c->slabs.p,
                       c->lanes9_mode > 0 ? 1u : 2u);
            else
                launch(c, (count9h_kernel<PACKED>), dim3(2 * pairs), block, smem, sa, pk, c->sorted.p, c->plan_d, c->slabs.p,
                       (c->lanes9_mode < 0 && !c->in_sharded && c->use_fast && !c->count_is_safe()) ? 3u : 0u);
            c->mark(EV_COUNT);
            launch(c, reduce_slabs9h_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, c->slabs.p, c->plan_d, seg_hist);
            return;
        }
    }
    if constexpr (K <= 7) {
        // the histogram sits at a 64 KiB-aligned shared address (vk_count.cuh): up to 64 KiB of padding in front
        const size_t smem = 0x10000 + (size_t)(NK + 32) * sizeof(uint32_t);
        if (c->chunk_mode(K)) launch(c, countd_kernel<K>, grid, block, smem, sa, (const uint64_t*)c->chunks.p, c->plan_d, c->slabs.p);
        else {
            // k = 7, text input: the one-read-per-lane kernel (vk_countt.cuh) for samples whose reads have one length, this
            // kernel for the others.  Which it is shows on the device only (the scatter kernel leaves the longest read in
            // the plan), so the context goes by its LAST sample: this kernel counts and reports when the sample was one
            // for countt_kernel (lanes_verdict bit 0: the next step launches that kernel instead); countt_kernel refuses
            // a sample that is not (bit 1) and with_table_retry repeats the step with this kernel.  Samples of a context
            // come from one sequencing run as a rule: one repeated step per change of kind, no idle launch otherwise.
            // A read-sharded sample (all ranks must agree) launches both and lets the device pick (policy 2).
            // VK_COUNT_LANES=0: always this kernel.
            uint32_t lanes_policy = 0;
            if constexpr (K == 7 && !PACKED) {
                if (c->lanes_mode < 0 && c->use_fast && !c->count_is_safe() && !c->use_pairs) {
                    const uint64_t* srt = (const uint64_t*)c->sorted.p;
                    if (c->in_sharded) {
                        lanes_policy = 2u;
                        launch_countt(c, grid, sa, srt, 2u | ((c->countt_knobs & ~1u) << 8));
                    } else if (c->k7_lanes) {
                        launch_countt(c, grid, sa, srt, 2u | ((c->countt_knobs | 1u) << 8));
                        c->mark(EV_COUNT);
                        launch(c, reduce_slabs_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, c->slabs.p, c->plan_d, NK, seg_hist, c->zero_unused_rows ? 1 : 0);
                        return;
                    } else lanes_policy = 3u;
                }
            }
            launch(c, (count_kernel<K, kSmem32, PACKED>), grid, block, smem, sa, pk, c->sorted.p, c->plan_d, c->slabs.p, seg_hist, lanes_policy);
        }
        c->mark(EV_COUNT);
        launch(c, reduce_slabs_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, c->slabs.p, c->plan_d, NK, seg_hist, c->zero_unused_rows ? 1 : 0);
    } else {
        launch(c, zero_u64_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, seg_hist, total);
        c->mark(EV_BUCKET);
        launch(c, (count_kernel<K, kGlobal, PACKED>), grid, block, 0, sa, pk, c->sorted.p, c->plan_d, c->slabs.p, seg_hist, 0u);
        c->mark(EV_COUNT);
    }
}

// between: what a read-sharded sample does between the histogram and the thresholds (sum the shards' histograms)
void enqueue_count(vk_ctx* c, int k, unsigned long long* seg_hist, const std::function<void()>& between = nullptr)
{
    using namespace vk;
    const int bgrid = c->n_sms * 8;
    // thresholds fitted to the base targets (vk_sample.cuh; both kernels return at once for VK_SAMPLING_EXPECTED): the base
    // histogram over the priority buckets, then the fit.  A read-sharded sample sums the shards' histograms in between;
    // a caller that brings the sample's histogram (params.prio_hist) only needs the fit.
    launch(c, prio_hist_kernel, dim3(c->n_sms * 4), dim3(kPrioHistThreads), 0, c->starts.p, c->ends.p, (const StepArgs*)c->args_d,
           (const Plan*)c->plan_d, c->prio_hist.p, 0);
    CU(cudaGetLastError());
    if (between) between();
    launch(c, thr_calibrate_kernel, dim3(kCalibCtas), dim3(1024), 0, (const StepArgs*)c->args_d, (const unsigned long long*)c->prio_hist.p,
           c->prio_hist.p + VK_PRIO_BUCKETS, c->plan_d);
    CU(cudaGetLastError());
    if (c->chunk_mode(k))
        launch(c, bucket_scatter_kernel<kBucketItemsChunk>, dim3(2 * bgrid), dim3(kBucketThreads), kStageChunks * sizeof(uint64_t),
               c->starts.p, c->ends.p, (const StepArgs*)c->args_d, 0, c->sorted.p, c->chunks.p, c->plan_d, c->long_list.p);
    else
        launch(c, bucket_scatter_kernel<kBucketItems>, dim3(bgrid), dim3(kBucketThreads), 0, c->starts.p, c->ends.p,
               (const StepArgs*)c->args_d, 0, c->sorted.p, (uint64_t*)nullptr, c->plan_d, c->long_list.p);
    CU(cudaGetLastError());
    // reads of 2^24 bases or more, listed by the scatter kernel: several table entries each (returns at once when there are none)
    launch(c, long_reads_kernel, dim3(4), dim3(256), 0, c->starts.p, c->ends.p, (const StepArgs*)c->args_d, 0, c->sorted.p, c->plan_d,
           (const uint64_t*)c->long_list.p);
    CU(cudaGetLastError());
    c->mark(EV_BUCKET);
    const bool pk = c->use_packed;
    switch (k) {
    case 5: pk ? launch_count<5, true>(c, seg_hist) : launch_count<5, false>(c, seg_hist); break;
    case 6: pk ? launch_count<6, true>(c, seg_hist) : launch_count<6, false>(c, seg_hist); break;
    case 7: pk ? launch_count<7, true>(c, seg_hist) : launch_count<7, false>(c, seg_hist); break;
    case 8: pk ? launch_count<8, true>(c, seg_hist) : launch_count<8, false>(c, seg_hist); break;
    case 9: pk ? launch_count<9, true>(c, seg_hist) : launch_count<9, false>(c, seg_hist); break;
    default: throw ApiError{VK_EINVAL, "k must be 5..9"};
    }
    CU(cudaGetLastError());
}

// ---- K4 ----------------------------------------------------------------------------------------------
uint32_t next_pow2(uint32_t v)
{
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

void ensure_outbox(vk_ctx* c, size_t n);

void ensure_render_buffers(vk_ctx* c, const Mapping& m, int k, int levels)
{
    const uint32_t nk = 1u << (2 * k);
    const uint32_t n_pix = (uint32_t)m.side * (uint32_t)m.side;
    c->generation += c->canon.ensure((size_t)levels * nk);
    ensure_outbox(c, (size_t)levels * n_pix);
    if (n_pix > 16384) {
        c->generation += c->vals.ensure((size_t)levels * next_pow2(n_pix));
        c->generation += c->bins.ensure((size_t)levels * 256);
    }
}

// levels: number of levels to render (grid size); canon must hold levels * 4^k
// live_plan: the ladder of the step that is being enqueued (levels is then an upper bound); nullptr: exactly `levels` levels
void enqueue_render(vk_ctx* c, const Mapping& m, int k, int levels, const unsigned long long* seg_hist, const vk::Plan* live_plan = nullptr)
{
    using namespace vk;
    const uint32_t nk = 1u << (2 * k);
    const uint32_t n_pix = (uint32_t)m.side * (uint32_t)m.side;
    const uint32_t n_pad = next_pow2(n_pix);
    if (!c->capturing) ensure_render_buffers(c, m, k, levels);
    c->last_levels = levels;
    c->last_side = m.side;
    if (seg_hist) {
        launch(c, fold_kernel, dim3((nk + 255) / 256), dim3(256), 0, seg_hist, k, levels, c->canon.p, live_plan);
        CU(cudaGetLastError());
    }
    c->mark(EV_FOLD);
    if (n_pix <= 16384) {
        // one cluster of 8 CTAs per level; slice = next power of two of a eighth of the pixels
        uint32_t S = next_pow2((n_pix + vk::kImgCluster - 1) / vk::kImgCluster);
        if (S < 64) S = 64;
        const size_t smem = (size_t)vk::kImgCluster * S * sizeof(unsigned long long);
        const unsigned threads = S / 2 > 1024 ? 1024 : S / 2;
        launch(c, image_kernel_cluster, dim3(vk::kImgCluster, levels), dim3(threads), smem, c->canon.p, m.lut.p, nk, n_pix, S, c->pix_d(), live_plan);
        CU(cudaGetLastError());
    } else {
        dim3 g1((n_pad + 255) / 256, levels);
        launch(c, image_gather_kernel, dim3(g1), dim3(256), 0, c->canon.p, m.lut.p, nk, n_pix, n_pad, c->vals.p);
        CU(cudaGetLastError());
        dim3 gt(n_pad / kSortTile, levels);
        launch(c, bitonic_tile_kernel, dim3(gt), dim3(1024), 0, c->vals.p, n_pad, 2, kSortTile, kSortTile / 2);
        CU(cudaGetLastError());
        for (uint32_t size = 2 * kSortTile; size <= n_pad; size <<= 1) {
            uint32_t stride = size >> 1;
            for (; (stride >> 1) >= kSortTile; stride >>= 2) {           // two strides per pass while both are global
                dim3 gs((n_pad / 4 + 255) / 256, levels);
                launch(c, bitonic_global_step2, dim3(gs), dim3(256), 0, c->vals.p, n_pad, size, stride);
                CU(cudaGetLastError());
            }
            if (stride >= kSortTile) {
                dim3 gs((n_pad / 2 + 255) / 256, levels);
                launch(c, bitonic_global_step, dim3(gs), dim3(256), 0, c->vals.p, n_pad, size, stride);
                CU(cudaGetLastError());
            }
            launch(c, bitonic_tile_kernel, dim3(gt), dim3(1024), 0, c->vals.p, n_pad, size, size, kSortTile / 2);
            CU(cudaGetLastError());
        }
        launch(c, image_bins_kernel, dim3(levels), dim3(256), 0, c->vals.p, n_pix, n_pad, c->bins.p);
        CU(cudaGetLastError());
        dim3 gd((n_pix + 255) / 256, levels);
        launch(c, image_digitize_kernel, dim3(gd), dim3(256), 0, c->canon.p, m.lut.p, nk, n_pix, c->bins.p, c->pix_d());
        CU(cudaGetLastError());
    }
    c->mark(EV_RENDER);
}

void fill_stats(const vk::Plan& p, vk_stats* s)
{
    s->n_bytes = p.n_bytes;
    s->n_lines = p.n_lines;
    s->n_reads = p.n_reads;
    s->nsites = p.nsites_ref;
    s->nsites_true = p.nsites_true;
}

void fill_result(const vk::Plan& p, vk_result* r)
{
    memset(r, 0, sizeof(*r));
    fill_stats(p, &r->stats);
    r->status = p.status;
    r->n_levels = p.n_levels;
    uint64_t reads = 0, bases = 0;
    for (int l = p.n_levels - 1; l >= 0; --l) {
        reads += p.seg_reads[l] - p.seg_extra[l];          // entries - the extra entries of reads cut into several
        bases += p.seg_bases[l];
        r->level_bp[l] = p.level_bp[l];
        r->level_reads[l] = reads;
        r->level_bases[l] = bases;
    }
}

void fetch_plan(vk_ctx* c)
{
    CU(cudaMemcpyAsync(c->plan_h, c->plan_d, sizeof(vk::Plan), cudaMemcpyDeviceToHost, c->stream));
}

// grow the outbox to hold n pixel bytes; the Plan travels with it (stream is idle or the copy is stream-ordered)
void ensure_outbox(vk_ctx* c, size_t n)
{
    if (c->out_d && n <= c->out_cap) return;
    const size_t want = n + n / 4 + (1u << 20);
    uint8_t* nd = nullptr;
    uint8_t* nh = nullptr;
    CU(cudaMalloc(&nd, kPlanPad + want));
    CU(cudaMallocHost(&nh, kPlanPad + want));
    if (c->out_d) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(nd, c->out_d, kPlanPad, cudaMemcpyDeviceToDevice));
        memcpy(nh, c->out_h, kPlanPad);
        CU(cudaFree(c->out_d));
        CU(cudaFreeHost(c->out_h));
    } else {
        CU(cudaMemset(nd, 0, kPlanPad));
        memset(nh, 0, kPlanPad);
    }
    c->out_d = nd;
    c->out_h = nh;
    c->out_cap = want;
    ++c->generation;
    c->plan_d = reinterpret_cast<vk::Plan*>(nd);
    c->plan_h = reinterpret_cast<vk::Plan*>(nh);
}

const Mapping& get_mapping(vk_ctx* c, int slot, int k)
{
    if (slot < 0 || slot >= 4) throw ApiError{VK_EINVAL, "mapping slot must be 0..3"};
    const Mapping& m = c->maps[slot];
    if (m.k != k || m.side <= 0) throw ApiError{VK_ESTATE, "no pixel table of this k in the mapping slot (vk_set_mapping)"};
    return m;
}

void check_params(const vk_params* p)
{
    if (!p) throw ApiError{VK_EINVAL, "params is NULL"};
    if (p->k < VK_MIN_K || p->k > VK_MAX_K) throw ApiError{VK_EINVAL, "k must be 5..9"};
    if (p->breaklength != 0 && p->breaklength < 32) throw ApiError{VK_EINVAL, "breaklength must be 0 or >= 32"};
    if (p->sampling != VK_SAMPLING_EXPECTED && p->sampling != VK_SAMPLING_CALIBRATED) throw ApiError{VK_EINVAL, "sampling must be VK_SAMPLING_EXPECTED or VK_SAMPLING_CALIBRATED"};
}

uint64_t reads_bound(uint64_t n_bytes) { return n_bytes / 24 + 1024; }

// ---- NCCL, opened at run time -----------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // the soname first: a process that already uses NCCL (torch.distributed) gets the very same library
        const char* names[] = {getenv("VK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nme : names) {
            if (!nme || !*nme) continue;
            api.handle = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather)
        throw ApiError{VK_ESTATE, "libnccl.so.2 could not be opened (set VK_NCCL_LIB to its path)"};
    return api;
}
#define NC(x)                                                                                             \
    do {                                                                                                  \
        ncclResult_t r_ = (x);                                                                            \
        if (r_ != ncclSuccess) {                                                                          \
            NcclApi& a_ = nccl_api();                                                                     \
            throw ApiError{VK_ECUDA, std::string("NCCL error in " #x ": ") + (a_.GetErrorString ? a_.GetErrorString(r_) : "?")}; \
        }                                                                                                 \
    } while (0)

template <typename F>
int guarded(F&& f)
{
    try {
        f();
        return VK_OK;
    } catch (const CudaError& e) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at vk_capi.cu:%d", (int)e.e, cudaGetErrorString(e.e), e.what, e.line);
        g_err = buf;
        cudaGetLastError();
        return e.e == cudaErrorMemoryAllocation ? VK_ENOMEM : VK_ECUDA;
    } catch (const ApiError& e) {
        g_err = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        g_err = e.what();
        return VK_EINVAL;
    }
}

// run parse (+ optional later stages via `rest`) and redo everything once if the read table was too small
template <typename F>
void with_table_retry(vk_ctx* c, F&& body)
{
    for (int attempt = 0; attempt < 4; ++attempt) {
        body();
        CU(cudaStreamSynchronize(c->stream));
        const bool t_over = c->plan_h->table_overflow != 0, b_over = c->plan_h->bucket_overflow != 0;
        const bool k_small = c->plan_h->chunk_table_small != 0 && !t_over;
        const bool c_over = c->plan_h->count_overflow != 0 && !t_over && !b_over && !k_small;
        // k = 7: countt_kernel was launched alone and the sample was not one for it (vk_countt.cuh): nothing was counted
        const bool l_ref = (c->plan_h->lanes_verdict & 2u) != 0 && !t_over && !b_over && !k_small && !c_over;
        if (!t_over && !b_over && !c_over && !k_small && !l_ref) {
            if (c->plan_h->lanes_verdict & 1u) c->k7_lanes = true;       // (the flat-lane kernel counted a sample of one read length)
            if (!c->count_safe && c->n_bytes >= c->safe_from_bytes && ++c->safe_by_size_steps >= 32u) {
                c->safe_from_bytes = ~0ull;                              // (try the 16-bit bins again)
                c->safe_by_size_steps = 0;
            }
            c->exact_layout = false;
            c->count_safe = false;
            return;
        }
        if (attempt == 3) { c->count_safe = false; throw ApiError{VK_ERANGE, "read table overflow after resize"}; }
        if (l_ref) {
            c->k7_lanes = false;
            ++c->lanes_flips;
        } else if (c_over) {
            // a 16-bit bin of the fire-and-forget count kernel wrapped (a flood of one k-mer): count again, exactly
            c->count_safe = true;
            if (c->n_bytes - c->n_bytes / 4 < c->safe_from_bytes) c->safe_from_bytes = c->n_bytes - c->n_bytes / 4;
            ++c->count_fallbacks;
        } else if (t_over) ensure_tables_for(c, c->plan_h->n_reads + 16);          // exact size is known now
        else if (k_small) c->generation += c->chunks.ensure((size_t)c->plan_h->chunks_needed + 256);
        else {
            // a segment outgrew its expected-size region (or the table was too small for the layout): give every
            // segment room for every read
            c->exact_layout = true;
            ++c->bucket_retries;
            c->generation += c->sorted.ensure(sorted_need(c->plan_h->n_reads, (uint64_t)std::max(c->plan_h->n_levels, 1), true));
            if (c->plan_h->chunks_needed)       // chunk table in use: every region then holds every chunk (plan_kernel's bound)
                c->generation += c->chunks.ensure((size_t)std::max(c->plan_h->n_levels, 1) *
                                                  (size_t)(c->plan_h->nsites_true / 32 + 2 * c->plan_h->n_reads + 4608));
        }
    }
}

void drop_graphs(vk_ctx* c)
{
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

// The kernels of one whole step (arguments copy, framing, ladder, scatter, count, reduce, fold, images, read-back) in
// stream order; the body of both the plain path and the graph capture.
void enqueue_step(vk_ctx* c, const Mapping& m, int k, int max_levels_out)
{
    const size_t n_pix = (size_t)m.side * m.side;
    enqueue_parse(c);
    c->mark(EV_PARSE);
    c->zero_unused_rows = false;                 // nobody reads the rows beyond the ladder in this form
    enqueue_count(c, k, c->seg_hist.p);
    c->zero_unused_rows = true;
    // the number of levels is only known on the device: render max_levels_out, rows beyond the ladder
    // come from all-zero segments and are ignored by the caller
    enqueue_render(c, m, k, max_levels_out, c->seg_hist.p, c->plan_d);
    // Plan + pixels in one copy
    CU(cudaMemcpyAsync(c->out_h, c->out_d, kPlanPad + (size_t)max_levels_out * n_pix, cudaMemcpyDeviceToHost, c->stream));
}

// Executable graph of a step for this (k, table, levels, layout); captured on first use, replayed afterwards.  Everything
// that differs between samples reaches the kernels through the StepArgs block, so one graph serves every sample until a
// device buffer moves (generation).  Returns nullptr when graphs are off or could not be built (plain launches then).
vk_ctx::StepGraph* step_graph(vk_ctx* c, const Mapping& m, int slot, int k, int max_levels_out)
{
    if (!c->use_graph || c->graph_failed || c->fine_timing) return nullptr;
    for (auto& g : c->graphs)
        if (g.k == k && g.slot == slot && g.side == m.side && g.max_levels == max_levels_out && g.exact == (int)c->exact_layout &&
            g.packed == ((int)c->use_packed | (c->count_is_safe() ? 2 : 0) | (c->k7_lanes ? 4 : 0) | (c->n_bytes >= kEpochBytes ? 8 : 0)) && g.generation == c->generation)
            return &g;
    // stale graphs (a buffer moved) are of no use any more
    for (size_t i = 0; i < c->graphs.size();) {
        if (c->graphs[i].generation != c->generation) {
            cudaGraphExecDestroy(c->graphs[i].exec);
            c->graphs.erase(c->graphs.begin() + i);
        } else ++i;
    }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    c->capturing = true;
    c->captured_kernels = 0;
    bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
        try {
            enqueue_step(c, m, k, max_levels_out);
        } catch (...) {
            ok = false;
        }
        if (cudaStreamEndCapture(c->stream, &graph) != cudaSuccess || !graph) ok = false;
    }
    c->capturing = false;
    if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();
        if (exec) cudaGraphExecDestroy(exec);
        c->graph_failed = 1;
        return nullptr;
    }
    ++c->graph_captures;
    c->graphs.push_back({k, slot, m.side, max_levels_out, (int)c->exact_layout, (int)c->use_packed | (c->count_is_safe() ? 2 : 0) | (c->k7_lanes ? 4 : 0) | (c->n_bytes >= kEpochBytes ? 8 : 0), c->generation, exec,
                         c->captured_kernels});
    return &c->graphs.back();
}

}  // namespace

extern "C" {

int vk_abi_version(void) { return VK_ABI_VERSION; }
const char* vk_last_error(void) { return g_err.c_str(); }

int vk_ctx_create(int device, vk_ctx** out)
{
    return guarded([&] {
        if (!out) throw ApiError{VK_EINVAL, "out is NULL"};
        int n = 0;
        CU(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) throw ApiError{VK_EINVAL, "no such CUDA device"};
        CU(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) throw ApiError{VK_ECUDA, "this library is built for sm_100a (B200) only"};
        vk_ctx* c = new vk_ctx();
        c->device = device;
        c->n_sms = prop.multiProcessorCount;
        if (const char* e = getenv("VK_COUNT_THREADS")) c->count_threads = atoi(e);
        if (const char* e = getenv("VK_COUNT_CTAS")) c->count_ctas_per_sm = atoi(e);
        if (const char* e = getenv("VK_COUNT_EXTRA")) c->count_extra = std::max(-64, std::min(64, atoi(e)));      // (negative: SMs left to the other samples in flight)
        if (const char* e = getenv("VK_COUNT_EXTRA9")) c->count_extra9 = std::max(0, std::min(64, atoi(e)));
        if (const char* e = getenv("VK_COUNT_READS_PER_CTA")) c->reads_per_cta = std::max(0, atoi(e));
        c->reads_per_cta_env = c->reads_per_cta;
        if (const char* e = getenv("VK_TEST_TIGHT_BUCKETS")) c->test_tight = atoi(e) != 0;
        if (const char* e = getenv("VK_PDL")) c->use_pdl = atoi(e) != 0;
        if (const char* e = getenv("VK_COUNT16")) c->use_count16 = atoi(e) != 0;
        if (const char* e = getenv("VK_COUNT_PAIRS")) c->use_pairs = atoi(e) != 0;
        if (const char* e = getenv("VK_COUNT_LANES")) { c->lanes_mode = atoi(e); c->use_lanes = c->lanes_mode > 0; }
        if (const char* e = getenv("VK_COUNT_LANES9")) c->lanes9_mode = atoi(e);
        if (const char* e = getenv("VK_COUNTT_KNOBS")) c->countt_knobs = (unsigned)strtoul(e, nullptr, 0);
        if (const char* e = getenv("VK_COUNT_FAST")) c->use_fast = atoi(e) != 0;
        if (const char* e = getenv("VK_PACKED")) c->use_packed = atoi(e) != 0;
        if (const char* e = getenv("VK_CHUNKS")) c->use_chunks = atoi(e) != 0;
        if (const char* e = getenv("VK_GRAPH")) c->use_graph = atoi(e) != 0;
        if (const char* e = getenv("VK_TRACE_EACH")) { c->trace_each = atoi(e) != 0; if (c->trace_each) c->use_graph = false; }
        if (c->count_threads < 32 || c->count_threads > 1024 || c->count_threads % 32 || c->count_ctas_per_sm < 1 || c->count_ctas_per_sm > 3)
            throw ApiError{VK_EINVAL, "bad VK_COUNT_THREADS / VK_COUNT_CTAS"};
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (int i = 0; i < EV_N; ++i) CU(cudaEventCreate(&c->ev[i]));
        ensure_outbox(c, 0);
        CU(cudaMalloc(&c->args_d, sizeof(vk::StepArgs)));
        c->long_list.ensure(vk::kLongListCap);
        c->prio_hist.ensure(VK_PRIO_BUCKETS + 1024);      // + the block sums of thr_calibrate_kernel
        CU(cudaMallocHost(&c->args_h, sizeof(vk::StepArgs)));
        prepare_kernels();
        *out = c;
    });
}

int vk_ctx_destroy(vk_ctx* c)
{
    if (!c) return VK_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    if (c->comm) {
        try { nccl_api().CommDestroy(c->comm); } catch (...) {}
        c->comm = nullptr;
    }
    c->shard_table.release();
    if (c->args_d) cudaFree(c->args_d);
    if (c->args_h) cudaFreeHost(c->args_h);
    c->codes.release();
    c->valid.release();
    c->chunks.release();
    c->text_own.release();
    c->tile_status.release();
    c->masks.release();
    c->tile_count.release();
    c->warp_count.release();
    c->starts.release();
    c->ends.release();
    c->sorted.release();
    c->slabs.release();
    c->seg_hist.release();
    c->canon.release();
    c->vals.release();
    c->bins.release();
    c->remap_in.release();
    c->remap_out.release();
    c->remap_mult.release();
    c->remap_src.release();
    c->content.release();
    c->prio_hist.release();
    c->long_list.release();
    c->synth_off.release();
    for (auto& m : c->maps) m.lut.release();
    if (c->out_h) cudaFreeHost(c->out_h);
    if (c->out_d) cudaFree(c->out_d);
    for (int i = 0; i < EV_N; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return VK_OK;
}

int vk_set_mapping(vk_ctx* c, int slot, int k, int side, const int32_t* lut_host)
{
    return guarded([&] {
        if (!c || !lut_host) throw ApiError{VK_EINVAL, "NULL argument"};
        if (slot < 0 || slot >= 4) throw ApiError{VK_EINVAL, "mapping slot must be 0..3"};
        if (k < VK_MIN_K || k > VK_MAX_K || side <= 0 || side > 1024) throw ApiError{VK_EINVAL, "bad k or side"};
        const int32_t nk = 1 << (2 * k);
        const size_t n = (size_t)side * side;
        for (size_t i = 0; i < n; ++i)
            if (lut_host[i] < -1 || lut_host[i] >= nk) throw ApiError{VK_EINVAL, "pixel table entry outside [-1, 4^k)"};
        set_device(c);
        Mapping& m = c->maps[slot];
        m.lut.ensure(n);
        CU(cudaMemcpyAsync(m.lut.p, lut_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        m.k = k;
        m.side = side;
    });
}

int vk_upload(vk_ctx* c, const void* host_bytes, uint64_t n_bytes)
{
    return guarded([&] {
        if (!c || (!host_bytes && n_bytes)) throw ApiError{VK_EINVAL, "NULL argument"};
        if (n_bytes >> 40) throw ApiError{VK_ERANGE, "buffers of 2^40 bytes or more are not supported"};
        set_device(c);
        c->mark(EV_START);
        c->ev_valid[EV_DONE] = false;
        c->text_own.ensure(n_bytes + 64);
        if (n_bytes) CU(cudaMemcpyAsync(c->text_own.p, host_bytes, n_bytes, cudaMemcpyHostToDevice, c->stream));
        c->mark(EV_UPLOAD);
        c->text = c->text_own.p;
        c->n_bytes = n_bytes;
        c->have_text = true;
        c->parsed = c->counted = false;
    });
}

int vk_attach(vk_ctx* c, const void* dev_bytes, uint64_t n_bytes)
{
    return guarded([&] {
        if (!c || (!dev_bytes && n_bytes)) throw ApiError{VK_EINVAL, "NULL argument"};
        if (n_bytes >> 40) throw ApiError{VK_ERANGE, "buffers of 2^40 bytes or more are not supported"};
        if ((uintptr_t)dev_bytes & 15) throw ApiError{VK_EINVAL, "device text must be 16-byte aligned"};
        c->text = static_cast<const uint8_t*>(dev_bytes);
        c->n_bytes = n_bytes;
        c->have_text = true;
        c->parsed = c->counted = false;
        c->ev_valid[EV_UPLOAD] = false;
    });
}

int vk_parse(vk_ctx* c, vk_stats* out)
{
    return guarded([&] {
        if (!c || !out) throw ApiError{VK_EINVAL, "NULL argument"};
        if (!c->have_text) throw ApiError{VK_ESTATE, "vk_parse before vk_upload / vk_attach"};
        set_device(c);
        ensure_tables_for(c, reads_bound(c->n_bytes));
        ensure_parse_buffers(c, c->n_bytes);
        with_table_retry(c, [&] {
            if (!c->ev_valid[EV_UPLOAD]) { c->mark(EV_START); c->mark(EV_UPLOAD); }
            enqueue_args(c, nullptr);
            enqueue_parse(c);
            c->mark(EV_PARSE);
            fetch_plan(c);
        });
        c->parsed = true;
        fill_stats(*c->plan_h, out);
    });
}

int vk_count(vk_ctx* c, const vk_params* p, uint64_t* seg_hist_dev, vk_result* out)
{
    return guarded([&] {
        if (!c || !out) throw ApiError{VK_EINVAL, "NULL argument"};
        check_params(p);
        if (!c->have_text) throw ApiError{VK_ESTATE, "vk_count before vk_upload / vk_attach"};
        set_device(c);
        const uint32_t nk = 1u << (2 * p->k);
        unsigned long long* sh = reinterpret_cast<unsigned long long*>(seg_hist_dev);
        if (!sh) {
            c->generation += c->seg_hist.ensure((size_t)vk::kMaxLevels * nk);
            sh = c->seg_hist.p;
        }
        ensure_tables_for(c, reads_bound(c->n_bytes));
        ensure_parse_buffers(c, c->n_bytes);
        ensure_count_buffers(c, p->k);
        with_table_retry(c, [&] {
            if (!c->ev_valid[EV_UPLOAD]) { c->mark(EV_START); c->mark(EV_UPLOAD); }
            enqueue_args(c, p);
            enqueue_parse(c, !c->parsed);         // framing found by vk_parse is kept; the ladder needs the parameters
            c->mark(EV_PARSE);
            enqueue_count(c, p->k, sh);
            fetch_plan(c);
            c->parsed = false;                     // a retry (table overflow) must rescan
        });
        if (c->plan_h->long_reads) throw ApiError{VK_ERANGE, "more than 4096 reads of 2^24 bases or more (or one such read in chunk-table mode)"};
        c->parsed = true;
        c->counted = true;
        c->counted_k = p->k;
        fill_result(*c->plan_h, out);
    });
}

// Base histogram over the priority buckets of this buffer's reads, added to the caller's device buffer (read-sharded
// samples with calibrated thresholds: the caller sums the shards' histograms and hands the sum to vk_count).
int vk_prio_hist(vk_ctx* c, const vk_params* p, uint64_t* hist_dev)
{
    return guarded([&] {
        if (!c || !hist_dev) throw ApiError{VK_EINVAL, "NULL argument"};
        check_params(p);
        if (!c->have_text) throw ApiError{VK_ESTATE, "vk_prio_hist before vk_upload / vk_attach"};
        if (!c->parsed) throw ApiError{VK_ESTATE, "vk_prio_hist before vk_parse"};
        set_device(c);
        enqueue_args(c, p);
        enqueue_begin(c, false);
        launch(c, vk::prio_hist_kernel, dim3(c->n_sms * 4), dim3(vk::kPrioHistThreads), 0, c->starts.p, c->ends.p, (const vk::StepArgs*)c->args_d,
               (const vk::Plan*)c->plan_d, reinterpret_cast<unsigned long long*>(hist_dev), 1);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
    });
}

int vk_render(vk_ctx* c, int slot, int k, int n_levels, const uint64_t* seg_hist_dev, uint64_t* canon_host, uint8_t* pixels_host)
{
    return guarded([&] {
        if (!c) throw ApiError{VK_EINVAL, "NULL argument"};
        if (n_levels < 0 || n_levels > VK_MAX_LEVELS) throw ApiError{VK_EINVAL, "n_levels out of range"};
        if (k < VK_MIN_K || k > VK_MAX_K) throw ApiError{VK_EINVAL, "k must be 5..9"};
        set_device(c);
        const unsigned long long* sh = reinterpret_cast<const unsigned long long*>(seg_hist_dev);
        if (!sh) {
            if (!c->counted || c->counted_k != k) throw ApiError{VK_ESTATE, "vk_render without histograms: call vk_count first"};
            sh = c->seg_hist.p;
        }
        if (n_levels == 0) return;
        const uint32_t nk = 1u << (2 * k);
        const Mapping* m = pixels_host ? &get_mapping(c, slot, k) : nullptr;
        if (m) {
            enqueue_render(c, *m, k, n_levels, sh);
        } else {
            c->generation += c->canon.ensure((size_t)n_levels * nk);
            launch(c, vk::fold_kernel, dim3((nk + 255) / 256), dim3(256), 0, sh, k, n_levels, c->canon.p, (const vk::Plan*)nullptr);
            CU(cudaGetLastError());
        }
        if (canon_host)
            CU(cudaMemcpyAsync(canon_host, c->canon.p, (size_t)n_levels * nk * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        if (pixels_host)
            CU(cudaMemcpyAsync(pixels_host, c->pix_d(), (size_t)n_levels * m->side * m->side, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    });
}

int vk_render_counts(vk_ctx* c, int slot, int k, int n, const uint64_t* canon_host, uint8_t* pixels_host)
{
    return guarded([&] {
        if (!c || !canon_host || !pixels_host) throw ApiError{VK_EINVAL, "NULL argument"};
        if (n < 1 || n > VK_MAX_LEVELS) throw ApiError{VK_EINVAL, "n out of range"};
        if (k < VK_MIN_K || k > VK_MAX_K) throw ApiError{VK_EINVAL, "k must be 5..9"};
        set_device(c);
        const Mapping& m = get_mapping(c, slot, k);
        const uint32_t nk = 1u << (2 * k);
        c->generation += c->canon.ensure((size_t)n * nk);
        CU(cudaMemcpyAsync(c->canon.p, canon_host, (size_t)n * nk * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        enqueue_render(c, m, k, n, nullptr);      // canon is already in place: no fold
        CU(cudaMemcpyAsync(pixels_host, c->pix_d(), (size_t)n * m.side * m.side, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    });
}

int vk_reads_to_images(vk_ctx* c, const void* text, uint64_t n_bytes, int on_device, const vk_params* p, int slot,
                       int max_levels_out, vk_result* result, uint64_t* canon_host, uint8_t* pixels_host)
{
    int rc = on_device ? vk_attach(c, text, n_bytes) : VK_OK;
    if (rc != VK_OK) return rc;
    return guarded([&] {
        if (!c || !result) throw ApiError{VK_EINVAL, "NULL argument"};
        check_params(p);
        if (max_levels_out < 1 || max_levels_out > VK_MAX_LEVELS) throw ApiError{VK_EINVAL, "max_levels_out out of range"};
        set_device(c);
        const int k = p->k;
        const uint32_t nk = 1u << (2 * k);
        const Mapping& m = get_mapping(c, slot, k);
        const size_t n_pix = (size_t)m.side * m.side;
        if (!on_device) {
            if (n_bytes >> 40) throw ApiError{VK_ERANGE, "buffers of 2^40 bytes or more are not supported"};
            c->text_own.ensure(n_bytes + 64);
        }
        with_table_retry(c, [&] {
            // every buffer at its size before anything is enqueued (a capture must not allocate); after an overflow the
            // tables have grown and the graph is captured again
            c->generation += c->seg_hist.ensure((size_t)vk::kMaxLevels * nk);
            ensure_tables_for(c, reads_bound(n_bytes));
            ensure_parse_buffers(c, n_bytes);
            ensure_count_buffers(c, k);
            ensure_render_buffers(c, m, k, max_levels_out);
            if (!on_device) {
                c->text = c->text_own.p;
                c->n_bytes = n_bytes;
                c->have_text = true;
            }
            enqueue_args(c, p);
            vk_ctx::StepGraph* g = step_graph(c, m, slot, k, max_levels_out);
            c->mark(EV_START);
            if (!on_device && n_bytes) CU(cudaMemcpyAsync(c->text_own.p, text, n_bytes, cudaMemcpyHostToDevice, c->stream));
            c->mark(EV_UPLOAD);
            if (g) {
                CU(cudaGraphLaunch(g->exec, c->stream));
                c->launches += g->kernels;
                ++c->graph_launches;
                c->last_levels = max_levels_out;
                c->last_side = m.side;
            } else {
                enqueue_step(c, m, k, max_levels_out);
            }
            c->mark(EV_DONE);
        });
        trace_dump(c);
        if (c->plan_h->long_reads) throw ApiError{VK_ERANGE, "more than 4096 reads of 2^24 bases or more (or one such read in chunk-table mode)"};
        c->parsed = c->counted = true;
        c->counted_k = k;
        fill_result(*c->plan_h, result);
        int nl = result->n_levels;
        if (nl > max_levels_out) {
            // rare: more levels than the caller guessed; render the rest with a second pass
            enqueue_render(c, m, k, nl, c->seg_hist.p);
            CU(cudaMemcpyAsync(c->pix_h(), c->pix_d(), (size_t)nl * n_pix, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
        c->last_levels = nl;
        if (pixels_host && nl > 0) memcpy(pixels_host, c->pix_h(), (size_t)nl * n_pix);
        if (canon_host && nl > 0) {
            CU(cudaMemcpyAsync(canon_host, c->canon.p, (size_t)nl * nk * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
    });
}

int vk_base_content(vk_ctx* c, int32_t pos_begin, int32_t pos_end, uint64_t* counts_host)
{
    return guarded([&] {
        if (!c || !counts_host) throw ApiError{VK_EINVAL, "NULL argument"};
        if (pos_begin < 0 || pos_end <= pos_begin || pos_end - pos_begin > vk::kContentMaxPos)
            throw ApiError{VK_EINVAL, "positions must satisfy 0 <= begin < end <= begin + 64"};
        if (!c->have_text || !c->parsed) throw ApiError{VK_ESTATE, "vk_base_content needs framed reads: call vk_parse, vk_count or vk_reads_to_images first"};
        set_device(c);
        const uint32_t n_pos = (uint32_t)(pos_end - pos_begin);
        const size_t n = (size_t)n_pos * 5;
        c->content.ensure(n);
        CU(cudaMemsetAsync(c->content.p, 0, n * sizeof(unsigned long long), c->stream));
        const uint64_t n_reads = c->plan_h->n_reads;
        if (n_reads) {
            const unsigned grid = (unsigned)std::min<uint64_t>((n_reads + vk::kContentThreads - 1) / vk::kContentThreads,
                                                               (uint64_t)c->n_sms * 8);
            launch(c, vk::base_content_kernel, dim3(grid), dim3(vk::kContentThreads), 0, c->text, (const uint64_t*)c->starts.p,
                   (const uint64_t*)c->ends.p, (const vk::Plan*)c->plan_d, (uint32_t)pos_begin, n_pos, c->content.p);
            CU(cudaGetLastError());
        }
        CU(cudaMemcpyAsync(counts_host, c->content.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    });
}

int vk_last_timings(vk_ctx* c, float* ms8)
{
    return guarded([&] {
        if (!c || !ms8) throw ApiError{VK_EINVAL, "NULL argument"};
        for (int i = 0; i < 8; ++i) ms8[i] = 0.f;
        for (int i = 1; i < EV_N; ++i) {
            if (c->ev_valid[i] && c->ev_valid[i - 1]) {
                float t = 0.f;
                if (cudaEventElapsedTime(&t, c->ev[i - 1], c->ev[i]) == cudaSuccess) ms8[i - 1] = t;
            }
        }
        const int last = c->ev_valid[EV_DONE] ? EV_DONE : EV_RENDER;
        if (!c->ev_valid[last]) return;
        if (c->ev_valid[EV_START] && c->ev_valid[last]) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, c->ev[EV_START], c->ev[last]) == cudaSuccess) ms8[7] = t;
        }
        cudaGetLastError();
    });
}

int vk_device_pixels(vk_ctx* c, const uint8_t** dev_pixels, int32_t* n_levels, int32_t* side)
{
    return guarded([&] {
        if (!c || !dev_pixels || !n_levels || !side) throw ApiError{VK_EINVAL, "NULL argument"};
        if (c->last_side <= 0) throw ApiError{VK_ESTATE, "nothing has been rendered on this context yet"};
        set_device(c);
        CU(cudaStreamSynchronize(c->stream));
        *dev_pixels = c->pix_d();
        *n_levels = c->last_levels;
        *side = c->last_side;
    });
}

int vk_remap(vk_ctx* c, int n_images, uint32_t n_in, uint32_t n_out, const uint8_t* in_host, const int32_t* src0,
             const int32_t* src1, const uint8_t* mult, int sum_rc, uint8_t* out_host)
{
    return guarded([&] {
        if (!c || !in_host || !src0 || !src1 || !mult || !out_host) throw ApiError{VK_EINVAL, "NULL argument"};
        if (n_images < 1 || n_in == 0 || n_out == 0 || n_in > (1u << 20) || n_out > (1u << 20))
            throw ApiError{VK_EINVAL, "bad image count or size"};
        for (uint32_t p = 0; p < n_out; ++p)
            if (src0[p] < -1 || src1[p] < -1 || src0[p] >= (int32_t)n_in || src1[p] >= (int32_t)n_in)
                throw ApiError{VK_EINVAL, "remap source pixel outside the input image"};
        set_device(c);
        c->remap_in.ensure((size_t)n_images * n_in);
        c->remap_out.ensure((size_t)n_images * n_out);
        c->remap_src.ensure((size_t)2 * n_out);
        c->remap_mult.ensure((size_t)2 * n_out);
        CU(cudaMemcpyAsync(c->remap_in.p, in_host, (size_t)n_images * n_in, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->remap_src.p, src0, sizeof(int32_t) * n_out, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->remap_src.p + n_out, src1, sizeof(int32_t) * n_out, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->remap_mult.p, mult, (size_t)2 * n_out, cudaMemcpyHostToDevice, c->stream));
        launch(c, vk::remap_kernel, dim3(n_images), dim3(256), 0, c->remap_in.p, n_in, c->remap_src.p, c->remap_src.p + n_out,
               c->remap_mult.p, n_out, sum_rc, c->remap_out.p);
        CU(cudaMemcpyAsync(out_host, c->remap_out.p, (size_t)n_images * n_out, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    });
}

int vk_set_fine_timing(vk_ctx* c, int on)
{
    if (!c) return VK_EINVAL;
    c->fine_timing = on != 0;
    return VK_OK;
}

int vk_set_batch_mode(vk_ctx* c, int on)
{
    if (!c) return VK_EINVAL;
    c->reads_per_cta = on ? 3000 : c->reads_per_cta_env;
    return VK_OK;
}

uint64_t vk_launch_count(vk_ctx* c) { return c ? c->launches : 0; }

int vk_graph_stats(vk_ctx* c, uint64_t* launches, uint64_t* captures, int32_t* state)
{
    if (!c) return VK_EINVAL;
    if (launches) *launches = c->graph_launches;
    if (captures) *captures = c->graph_captures;
    if (state) *state = !c->use_graph ? 0 : (c->graph_failed ? -1 : 1);
    return VK_OK;
}
uint64_t vk_bucket_retries(vk_ctx* c) { return c ? c->bucket_retries : 0; }
uint64_t vk_count_fallbacks(vk_ctx* c) { return c ? c->count_fallbacks : 0; }

int vk_synth_fastq(vk_ctx* c, void* dev_bytes, uint64_t capacity, uint64_t n_bases, int read_len, uint64_t seed,
                   uint64_t first_read, uint64_t* n_out)
{
    return guarded([&] {
        if (!c || !dev_bytes || !n_out) throw ApiError{VK_EINVAL, "NULL argument"};
        if (read_len < 1 || read_len > 100000 || n_bases == 0) throw ApiError{VK_EINVAL, "bad read_len / n_bases"};
        set_device(c);
        const uint64_t L = (uint64_t)read_len;
        const uint64_t n_reads = (n_bases + L - 1) / L;
        const uint64_t last_len = n_bases - (n_reads - 1) * L;
        const uint64_t total = (n_reads - 1) * (2 * L + 17) + 2 * last_len + 17;
        if (total > capacity) throw ApiError{VK_EINVAL, "synthetic FASTQ does not fit the buffer"};
        const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->n_sms * 16);
        launch(c, vk::synth_fixed_kernel, dim3(grid), dim3(256), 0, static_cast<uint8_t*>(dev_bytes), total, n_reads, (uint32_t)L,
                                                            (uint32_t)last_len, seed, first_read);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
        *n_out = total;
    });
}

int vk_comm_unique_id(void* out128)
{
    return guarded([&] {
        if (!out128) throw ApiError{VK_EINVAL, "NULL argument"};
        ncclUniqueId id;
        NC(nccl_api().GetUniqueId(&id));
        memcpy(out128, &id, sizeof(id));
    });
}

int vk_comm_init(vk_ctx* c, const void* unique_id128, int rank, int world)
{
    return guarded([&] {
        if (!c || !unique_id128) throw ApiError{VK_EINVAL, "NULL argument"};
        if (world < 1 || rank < 0 || rank >= world) throw ApiError{VK_EINVAL, "bad rank / world"};
        set_device(c);
        NcclApi& api = nccl_api();
        if (c->comm) { NC(api.CommDestroy(c->comm)); c->comm = nullptr; }
        ncclUniqueId id;
        memcpy(&id, unique_id128, sizeof(id));
        NC(api.CommInitRank(&c->comm, world, id, rank));
        c->comm_rank = rank;
        c->comm_world = world;
        c->shard_table.ensure((size_t)2 * world + 2);
    });
}

int vk_comm_destroy(vk_ctx* c)
{
    return guarded([&] {
        if (!c) throw ApiError{VK_EINVAL, "NULL argument"};
        if (c->comm) {
            set_device(c);
            CU(cudaStreamSynchronize(c->stream));
            NC(nccl_api().CommDestroy(c->comm));
            c->comm = nullptr;
        }
        c->comm_world = 1;
        c->comm_rank = 0;
    });
}

int vk_sharded_reads_to_images(vk_ctx* c, const void* text, uint64_t n_bytes, int on_device, const vk_params* p, int slot,
                               int max_levels_out, vk_result* result, uint64_t* canon_host, uint8_t* pixels_host)
{
    int rc = on_device ? vk_attach(c, text, n_bytes) : VK_OK;
    if (rc != VK_OK) return rc;
    return guarded([&] {
        if (!c || !result) throw ApiError{VK_EINVAL, "NULL argument"};
        check_params(p);
        if (!c->comm) throw ApiError{VK_ESTATE, "vk_sharded_reads_to_images before vk_comm_init"};
        if (max_levels_out < 1 || max_levels_out > VK_MAX_LEVELS) throw ApiError{VK_EINVAL, "max_levels_out out of range"};
        set_device(c);
        NcclApi& api = nccl_api();
        const int k = p->k;
        const uint32_t nk = 1u << (2 * k);
        const Mapping& m = get_mapping(c, slot, k);
        const size_t n_pix = (size_t)m.side * m.side;
        if (!on_device) {
            if (n_bytes >> 40) throw ApiError{VK_ERANGE, "buffers of 2^40 bytes or more are not supported"};
            c->text_own.ensure(n_bytes + 64);
        }
        const size_t n_hist = (size_t)max_levels_out * nk;                    // rows that are exchanged
        struct InSharded { vk_ctx* c; explicit InSharded(vk_ctx* x) : c(x) { c->in_sharded = true; } ~InSharded() { c->in_sharded = false; } } in_sharded_guard(c);
        with_table_retry(c, [&] {
            c->generation += c->seg_hist.ensure((size_t)vk::kMaxLevels * nk + vk::kShardTail);
            ensure_tables_for(c, reads_bound(n_bytes));
            ensure_parse_buffers(c, n_bytes);
            ensure_count_buffers(c, k);
            ensure_render_buffers(c, m, k, max_levels_out);
            if (!on_device) {
                c->text = c->text_own.p;
                c->n_bytes = n_bytes;
                c->have_text = true;
            }
            enqueue_args(c, p);
            c->args_h->pa.shard_table = c->shard_table.p;
            c->args_h->pa.shard_rank = (uint32_t)c->comm_rank;
            c->args_h->pa.shard_world = (uint32_t)c->comm_world;
            c->mark(EV_START);
            if (!on_device && n_bytes) CU(cudaMemcpyAsync(c->text_own.p, text, n_bytes, cudaMemcpyHostToDevice, c->stream));
            c->mark(EV_UPLOAD);
            // framing of this shard, then the shards' (records, bases) to every rank: exchange step 1, 16 bytes per rank
            enqueue_parse(c, true, /*with_plan=*/false);
            unsigned long long* const send = c->shard_table.p + 2 * (size_t)c->comm_world;
            launch(c, vk::shard_stats_kernel, dim3(1), dim3(32), 0, (const vk::StepArgs*)c->args_d, (const vk::Plan*)c->plan_d, send);
            NC(api.AllGather(send, c->shard_table.p, 2, ncclUint64, c->comm, c->stream));
            launch(c, vk::plan_kernel, dim3(1), dim3(64), 0, (const vk::StepArgs*)c->args_d, c->starts.p, c->ends.p, c->plan_d);
            CU(cudaGetLastError());
            c->mark(EV_PARSE);
            // calibrated thresholds: the shards' base histograms over the priority buckets are summed first (512 KiB)
            const bool calibrated = p->sampling == VK_SAMPLING_CALIBRATED && p->prio_hist == 0;
            enqueue_count(c, k, c->seg_hist.p, [&]() {
                if (calibrated)
                    NC(api.AllReduce(c->prio_hist.p, c->prio_hist.p, VK_PRIO_BUCKETS, ncclUint64, ncclSum, c->comm, c->stream));
            });
            // exchange step 2: the histograms of every ladder segment + the per-segment totals, ONE all-reduce
            unsigned long long* const tail = c->seg_hist.p + n_hist;
            launch(c, vk::shard_tail_kernel, dim3(1), dim3(256), 0, c->plan_d, tail, 0);
            NC(api.AllReduce(c->seg_hist.p, c->seg_hist.p, n_hist + vk::kShardTail, ncclUint64, ncclSum, c->comm, c->stream));
            launch(c, vk::shard_tail_kernel, dim3(1), dim3(256), 0, c->plan_d, tail, 1);
            launch(c, vk::zero_u64_kernel, dim3(1), dim3(256), 0, tail, (uint64_t)vk::kShardTail);      // the row belongs to a segment again
            enqueue_render(c, m, k, max_levels_out, c->seg_hist.p, c->plan_d);
            CU(cudaMemcpyAsync(c->out_h, c->out_d, kPlanPad + (size_t)max_levels_out * n_pix, cudaMemcpyDeviceToHost, c->stream));
            c->mark(EV_DONE);
        });
        if (c->plan_h->long_reads) throw ApiError{VK_ERANGE, "more than 4096 reads of 2^24 bases or more (or one such read in chunk-table mode)"};
        c->parsed = c->counted = true;
        c->counted_k = k;
        fill_result(*c->plan_h, result);
        result->stats.n_reads = c->plan_h->total_reads;
        int nl = result->n_levels;
        if (nl > max_levels_out) throw ApiError{VK_ERANGE, "more ladder levels than max_levels_out (read-sharded samples exchange max_levels_out rows)"};
        c->last_levels = nl;
        if (pixels_host && nl > 0) memcpy(pixels_host, c->pix_h(), (size_t)nl * n_pix);
        if (canon_host && nl > 0) {
            CU(cudaMemcpyAsync(canon_host, c->canon.p, (size_t)nl * nk * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
    });
}

int vk_synth_fastq_variable(vk_ctx* c, void* dev_bytes, uint64_t capacity, uint64_t n_reads, uint64_t seed,
                            uint64_t first_read, int min_len, int max_len, int short_per_10000, int k, uint64_t* n_out,
                            uint64_t* n_bases_out)
{
    return guarded([&] {
        if (!c || !n_out) throw ApiError{VK_EINVAL, "NULL argument"};
        if (min_len < 1 || max_len < min_len || max_len > 100000 || short_per_10000 < 0 || short_per_10000 > 10000 || k < 1 ||
            n_reads == 0 || n_reads > (1ull << 32))
            throw ApiError{VK_EINVAL, "bad generator arguments"};
        set_device(c);
        c->synth_off.ensure(n_reads + 1);
        const int grid = (int)std::min<uint64_t>((n_reads + 255) / 256, (uint64_t)c->n_sms * 16);
        launch(c, vk::synth_var_sizes_kernel, dim3(grid), dim3(256), 0, c->synth_off.p, n_reads, seed, first_read,
               (uint32_t)min_len, (uint32_t)max_len, (uint32_t)short_per_10000, (uint32_t)k);
        launch(c, vk::synth_scan_kernel, dim3(1), dim3(1024), 0, c->synth_off.p, n_reads);
        CU(cudaGetLastError());
        uint64_t total = 0;
        CU(cudaMemcpyAsync(&total, c->synth_off.p + n_reads, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        *n_out = total;
        if (n_bases_out) *n_bases_out = (total - 17ull * n_reads) / 2;
        if (!dev_bytes) return;                         // size query
        if (total > capacity) throw ApiError{VK_EINVAL, "synthetic FASTQ does not fit the buffer"};
        const int g2 = (int)std::min<uint64_t>((n_reads + 7) / 8, (uint64_t)c->n_sms * 32);
        launch(c, vk::synth_var_kernel, dim3(g2), dim3(256), 0, static_cast<uint8_t*>(dev_bytes), (const uint64_t*)c->synth_off.p,
               n_reads, seed, first_read);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
    });
}

}  // extern "C"
