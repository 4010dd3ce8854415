// C ABI of mapdamage_b200 (include/mapdamage_b200.h): context, staging slots,
// launches.  No CPU path: without a CUDA device mdg_create fails.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>  // types only; the library is loaded with dlopen

#include "../../include/mapdamage_b200.h"
#include "mdg_count.cuh"
#include "mdg_swar.cuh"
#include "mdg_stage.cuh"
#include "mdg_planes.cuh"
#include "mdg_planes_ws.cuh"
#include "mdg_rescale.cuh"
#include "mdg_synth.cuh"
#include "mdg_inflate_dev.cuh"

namespace {

thread_local std::string g_create_error;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *name : names) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot load libnccl: ") + dlerror();
            return;
        }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy)
            api.error = "libnccl lacks a required symbol";
    });
    return api;
}

// Device copy of one batch; arrays live in one allocation.
struct DeviceArrays {
    void *block = nullptr;
    size_t bytes = 0;
    int64_t cap_reads = 0, cap_cigar = 0, cap_bases = 0;
    int64_t n_cigar = 0, n_bases = 0;
    unsigned long long *scan_tmp = nullptr;  // inside `block`: per-256-read totals for the layout scan
    bool has_qual = false;
    mdg::DevBatch view{};
};

struct Slot {
    cudaStream_t stream = nullptr;
    DeviceArrays arrays;
    // rescale outputs
    float *mr_out = nullptr;
    uint8_t *status_out = nullptr;
    // sparse mode (mdg_rescale_submit_sparse): the quality bytes that changed, instead of the whole array
    uint32_t *change_at = nullptr;
    uint8_t *change_q = nullptr;
    unsigned long long *n_changes = nullptr;       // device counter
    unsigned long long *n_changes_host = nullptr;  // page-locked copy, filled behind the kernels
    int64_t change_cap = 0;
};

}  // namespace

// Work list of one launch stream: reads the bit-sliced kernel hands to the general kernel.
struct WorkList {
    cudaStream_t stream = nullptr;
    uint32_t *reads = nullptr;
    unsigned long long *count = nullptr;
    int64_t cap = 0;
    // several libraries: reads grouped by library
    uint32_t *by_library = nullptr;         // [cap]
    unsigned long long *lib_scratch = nullptr;  // counts [n_lib] | offsets [n_lib + 1] | cursors [n_lib]
    // reads with one short indel, left by the bit-plane kernel to count_staged_kernel's indel variant
    uint32_t *indel_reads = nullptr;            // [cap], library l from offsets[l] on
    unsigned long long *indel_count = nullptr;  // [n_lib] | bounds [2 n_lib] = {first, last} of every library's stretch
};

struct mdg_dev_batch {
    DeviceArrays arrays;
    // results of mdg_rescale_resident (one entry per read), kept with the batch for the BAM writer
    float *res_mr = nullptr;
    uint8_t *res_status = nullptr;
    int64_t res_cap = 0;
};

struct mdg_ctx {
    mdg_config cfg{};
    int sm_count = 0;
    size_t smem_optin = 0;
    std::string error;
    cudaStream_t compute = nullptr;
    std::vector<Slot> slots;
    int next_slot = 0;
    // reference genome
    mdg::DevRef ref{};
    void *ref_block = nullptr;
    uint64_t ref_total_bases = 0;
    int64_t ref_words = 0;  // 32-bit words of the genome image (without the padding around it)
    uint32_t ref_min_contig = 0;
    // tables: one allocation [misincorp | dnacomp | lghist]
    unsigned long long *tables = nullptr;
    size_t n_mis = 0, n_comp = 0, n_lg = 0;
    mdg::CountTables count_tables{};
    void *aux_block = nullptr;  // overflow rows, counters, error flag, rescale stats, fast-path worklist counter
    unsigned long long *rescale_stats = nullptr;
    // rescale model
    void *model_block = nullptr;
    mdg::RescaleModel model{};
    unsigned long long *rescale_hist = nullptr;  // [2][n_slots][94] | [2][94] | [4], see mdg_fetch_rescale_hist
    size_t n_hist = 0;
    // launch geometry
    bool shared_slab = false;
    size_t slab_bytes = 0;
    int general_grid = 0;
    // bit-sliced kernel for gap-free reads; complex reads go through a per-stream work list
    bool swar_enabled = false, force_general = false;
    int swar_max_threads = 256, swar_reads = 1, swar_blocks_per_sm = 2;
    // staged bit-sliced kernel (mdg_stage.cuh): the default for gap-free reads
    bool staged_enabled = false;
    mdg::StagedGeom staged{};
    size_t staged_smem_plain = 0, staged_smem_qual = 0;
    int staged_tile_plain = 0, staged_tile_qual = 0, staged_threads = 512, staged_blocks_per_sm = 1;
    // reads with one indel: staged too (three planes, smaller tiles) or left to the general kernel.  Chosen per
    // launch from what the previous launches met (MDG_STAGE_INDELS=0/1 pins it): both give the same tables.
    bool staged_indels = false, staged_indels_auto = true;
    unsigned long long *indel_seen_dev = nullptr;  // reads with one indel met by the parse phase so far
    unsigned long long *indel_seen_host = nullptr; // pinned copy, refreshed after every launch
    int64_t reads_launched = 0;
    mdg::SwarGeom swar{};
    size_t swar_smem = 0;
    // bit-plane kernel (mdg_planes.cuh): the default for gap-free reads when no quality mask is asked for
    bool planes_enabled = false;
    mdg::PlaneGeom planes{};
    size_t planes_smem = 0;
    // its warp-specialised form (mdg_planes_ws.cuh): producer teams stage, consumer warps count
    int ws_variant = -1;  // index into WS_VARIANTS, -1: off
    mdg::PlaneGeom ws{};
    size_t ws_smem = 0;
    // its two-library form (a read's library picks counters and tables; smaller tiles: two sets of event tables)
    bool ws_qual = true;  // -Q batches go through it too (MDG_PLANES_QUAL=0: through count_staged_kernel)
    int ws_variant_libraries = -1;
    mdg::PlaneGeom ws_libraries_geom{};
    size_t ws_smem_libraries = 0;
    void *planes_block = nullptr;  // genome as bit planes
    size_t ref_words_bytes = 0;    // size of one genome image
    std::vector<WorkList> worklists;
    // measurement
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> kernel_events;  // pairs
    size_t kernel_events_used = 0;
    int64_t launches = 0;
    // multi-GPU: the all-reduce writes the sums over ranks into `reduced`; the accumulators stay local
    ncclComm_t comm = nullptr;
    unsigned long long *reduced = nullptr;
    bool reduced_valid = false;  // `reduced` holds the sum of what every rank has counted so far
};

namespace {

int fail(mdg_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    else g_create_error = buf;
    return code;
}

#define MDG_CUDA(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, MDG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Carves the arrays of a batch with the given capacities out of one allocation.
int alloc_arrays(mdg_ctx *ctx, DeviceArrays &a, int64_t reads, int64_t cigar, int64_t bases, bool with_qual)
{
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t at = off;
        off += align_up(bytes + 16);
        return at;
    };
    size_t o_flag = take(reads * 2), o_tid = take(reads * 4), o_pos = take(reads * 4), o_lib = take(reads * 2);
    size_t o_lseq = take(reads * 4), o_boff = take(reads * 4), o_coff = take((reads + 1) * 4);
    size_t o_cig = take(cigar * 4), o_seq = take(bases / 2 + 8), o_qual = with_qual ? take(bases + 8) : 0;
    size_t o_tlen = take(reads * 4), o_mtid = take(reads * 4), o_mpos = take(reads * 4);
    size_t o_scan = take(((size_t)(reads + 255) / 256 + 2) * 16);
    MDG_CUDA(ctx, cudaMalloc(&a.block, off));
    a.bytes = off;
    a.cap_reads = reads; a.cap_cigar = cigar; a.cap_bases = bases; a.has_qual = with_qual;
    char *p = (char *)a.block;
    a.view.flag = (const uint16_t *)(p + o_flag);
    a.view.tid = (const int32_t *)(p + o_tid);
    a.view.pos = (const int32_t *)(p + o_pos);
    a.view.lib = (const uint16_t *)(p + o_lib);
    a.view.l_seq = (const uint32_t *)(p + o_lseq);
    a.view.base_off = (const uint32_t *)(p + o_boff);
    a.view.cigar_off = (const uint32_t *)(p + o_coff);
    a.view.cigar = (const uint32_t *)(p + o_cig);
    a.view.seq4 = (const uint8_t *)(p + o_seq);
    a.view.qual = with_qual ? (const uint8_t *)(p + o_qual) : nullptr;
    a.view.tlen = (const int32_t *)(p + o_tlen);
    a.view.mtid = (const int32_t *)(p + o_mtid);
    a.view.mpos = (const int32_t *)(p + o_mpos);
    a.scan_tmp = (unsigned long long *)(p + o_scan);
    return MDG_OK;
}

int check_batch(mdg_ctx *ctx, const mdg_batch *h)
{
    if (!h) return fail(ctx, MDG_ERR_ARGUMENT, "batch is NULL");
    if (h->n_reads < 0 || h->n_cigar < 0 || h->n_bases < 0 || (h->n_bases & 1))
        return fail(ctx, MDG_ERR_ARGUMENT, "batch sizes must be non-negative and n_bases even");
    if (h->n_reads >= (1ll << 31) || h->n_cigar >= (1ll << 32) || h->n_bases >= (1ll << 32))
        return fail(ctx, MDG_ERR_ARGUMENT, "batch too large: split it (offsets are 32-bit)");
    if (h->n_reads && (!h->flag || !h->pos || (h->n_cigar && !h->cigar) || (h->n_bases && !h->seq4)))
        return fail(ctx, MDG_ERR_ARGUMENT, "batch lacks a required array (flag, pos, cigar, seq4)");
    if (h->n_reads && !h->cigar_off && h->n_cigar != h->n_reads && h->n_cigar != 1)
        return fail(ctx, MDG_ERR_ARGUMENT, "cigar_off may only be NULL when every read has exactly one CIGAR op "
                                           "(n_cigar = n_reads words, or n_cigar = 1: the one word all reads share)");
    return MDG_OK;
}

// Queues the host->device copies of one batch on `stream`.
// The counting pass reads neither the mate fields nor (with min_qual = 0) the qualities: they stay on the host.
int copy_batch(mdg_ctx *ctx, DeviceArrays &a, const mdg_batch *h, cudaStream_t stream, bool mates, bool quals)
{
    if (h->n_reads > a.cap_reads || h->n_cigar > a.cap_cigar || h->n_bases > a.cap_bases)
        return fail(ctx, MDG_ERR_CAPACITY,
                    "batch (%lld reads, %lld CIGAR ops, %lld bases) exceeds the slot capacity (%lld, %lld, %lld)",
                    (long long)h->n_reads, (long long)h->n_cigar, (long long)h->n_bases, (long long)a.cap_reads,
                    (long long)a.cap_cigar, (long long)a.cap_bases);
    const int64_t n = h->n_reads;
    a.view.n_reads = n;
    a.n_cigar = h->n_cigar;
    a.n_bases = h->n_bases;
    if (!n) return MDG_OK;
#define MDG_H2D(field, bytes) \
    MDG_CUDA(ctx, cudaMemcpyAsync((void *)a.view.field, h->field, (size_t)(bytes), cudaMemcpyHostToDevice, stream))
#define MDG_FILL(field, byte, bytes) \
    MDG_CUDA(ctx, cudaMemsetAsync((void *)a.view.field, byte, (size_t)(bytes), stream))
    MDG_H2D(flag, n * 2);
    MDG_H2D(pos, n * 4);
    const bool shared_cigar = !h->cigar_off && h->n_cigar == 1 && n > 1;  // one CIGAR word for every read
    if (shared_cigar) {
        if (n > a.cap_cigar) return fail(ctx, MDG_ERR_CAPACITY, "batch of %lld reads exceeds the slot's CIGAR capacity (%lld)", (long long)n, (long long)a.cap_cigar);
        mdg::fill_words<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, stream>>>((uint32_t *)a.view.cigar, n, h->cigar[0]);
        MDG_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        a.n_cigar = n;
    } else {
        MDG_H2D(cigar, h->n_cigar * 4);
    }
    MDG_H2D(seq4, h->n_bases / 2);
    if (h->tid) MDG_H2D(tid, n * 4);
    else MDG_FILL(tid, 0, n * 4);
    // optional arrays: what a NULL pointer stands for is made on the device instead of crossing PCIe
    if (h->lib) MDG_H2D(lib, n * 2);
    else MDG_FILL(lib, 0, n * 2);
    if (h->tlen) MDG_H2D(tlen, n * 4);
    else MDG_FILL(tlen, 0, n * 4);
    if (mates) {
        if (h->mtid) MDG_H2D(mtid, n * 4);
        else MDG_FILL(mtid, 0xFF, n * 4);
        if (h->mpos) MDG_H2D(mpos, n * 4);
        else MDG_FILL(mpos, 0xFF, n * 4);
    }
    if (h->cigar_off) MDG_H2D(cigar_off, (n + 1) * 4);
    if (h->base_off) MDG_H2D(base_off, n * 4);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (h->l_seq) {
        MDG_H2D(l_seq, n * 4);
    } else {
        mdg::lseq_from_cigar<<<blocks, 256, 0, stream>>>(a.view.cigar, h->cigar_off ? a.view.cigar_off : nullptr, n,
                                                         (uint32_t *)a.view.l_seq);
        MDG_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    if (!h->cigar_off || !h->base_off) {
        mdg::layout_block_totals<<<blocks, 256, 0, stream>>>(a.view.l_seq, n, a.scan_tmp);
        mdg::synth_scan_totals<<<1, 1024, 0, stream>>>(a.scan_tmp, (int64_t)blocks);
        mdg::layout_fill<<<blocks, 256, 0, stream>>>(a.view.l_seq, n, a.scan_tmp, h->base_off ? nullptr : (uint32_t *)a.view.base_off,
                                                     h->cigar_off ? nullptr : (uint32_t *)a.view.cigar_off,
                                                     (uint64_t)h->n_bases, ctx->count_tables.error_flag);
        MDG_CUDA(ctx, cudaGetLastError());
        ctx->launches += 3;
    }
#undef MDG_H2D
#undef MDG_FILL
    if (quals && a.has_qual && h->qual)
        MDG_CUDA(ctx, cudaMemcpyAsync((void *)a.view.qual, h->qual, (size_t)h->n_bases, cudaMemcpyHostToDevice, stream));
    return MDG_OK;
}

int next_kernel_events(mdg_ctx *ctx, cudaEvent_t *start, cudaEvent_t *stop)
{
    if (ctx->kernel_events_used + 2 > ctx->kernel_events.size()) {
        for (int i = 0; i < 2; ++i) {
            cudaEvent_t e;
            MDG_CUDA(ctx, cudaEventCreate(&e));
            ctx->kernel_events.push_back(e);
        }
    }
    *start = ctx->kernel_events[ctx->kernel_events_used++];
    *stop = ctx->kernel_events[ctx->kernel_events_used++];
    return MDG_OK;
}

typedef void (*SwarKernelStaged)(mdg::DevBatch, mdg::DevRef, mdg::CountParams, mdg::CountTables, mdg::StagedGeom, uint32_t *,
                                 unsigned long long *, mdg::SwarSubset);
typedef void (*SwarKernel)(mdg::DevBatch, mdg::DevRef, mdg::CountParams, mdg::CountTables, mdg::SwarGeom, uint32_t *,
                           unsigned long long *, mdg::SwarSubset);

// variants: with / without the quality mask; blocks of up to `max_threads` threads, each counting `reads`
// reads per loop iteration (fewer threads leave more registers for more reads in flight)
SwarKernel swar_kernel(bool qual, int max_threads, int reads)
{
#define MDG_VARIANT(T, R) \
    if (max_threads == T && reads == R) return qual ? mdg::count_swar_kernel<true, T, R> : mdg::count_swar_kernel<false, T, R>;
    MDG_VARIANT(512, 1)
#undef MDG_VARIANT
    // the default: two co-resident 256-thread blocks per SM, one stages its tile while the other counts
    if (max_threads == 256 && reads == 1) return qual ? mdg::count_swar_kernel<true, 256, 1, 2> : mdg::count_swar_kernel<false, 256, 1, 2>;
    return nullptr;
}

// staged kernel variants: quality mask, indel reads staged too (third plane), block size / co-resident blocks
SwarKernelStaged staged_kernel(bool qual, bool indel, int threads)
{
    if (threads == 256) {
        if (qual) return mdg::count_staged_kernel<true, true, 256, 2>;
        return indel ? mdg::count_staged_kernel<false, true, 256, 2> : mdg::count_staged_kernel<false, false, 256, 2>;
    }
    if (threads == 512) {
        if (qual) return mdg::count_staged_kernel<true, true, 512, 1>;
        return indel ? mdg::count_staged_kernel<false, true, 512, 1> : mdg::count_staged_kernel<false, false, 512, 1>;
    }
    return nullptr;
}

int worklist_for(mdg_ctx *ctx, cudaStream_t stream, int64_t n_reads, WorkList **out)
{
    WorkList *wl = nullptr;
    for (auto &w : ctx->worklists)
        if (w.stream == stream) wl = &w;
    if (!wl) {
        ctx->worklists.emplace_back();
        wl = &ctx->worklists.back();
        wl->stream = stream;
        MDG_CUDA(ctx, cudaMalloc(&wl->count, (size_t)ctx->cfg.n_libraries * 8));
        MDG_CUDA(ctx, cudaMalloc(&wl->indel_count, (size_t)ctx->cfg.n_libraries * 24));
        if (ctx->cfg.n_libraries > 1)
            MDG_CUDA(ctx, cudaMalloc(&wl->lib_scratch, ((size_t)3 * ctx->cfg.n_libraries + 1) * 8));
    }
    if (wl->cap < n_reads) {
        MDG_CUDA(ctx, cudaStreamSynchronize(stream));
        cudaFree(wl->reads);
        cudaFree(wl->indel_reads);
        wl->reads = nullptr;
        wl->indel_reads = nullptr;
        wl->cap = 0;
        MDG_CUDA(ctx, cudaMalloc(&wl->reads, (size_t)n_reads * 4));
        MDG_CUDA(ctx, cudaMalloc(&wl->indel_reads, (size_t)n_reads * 4));
        if (ctx->cfg.n_libraries > 1) {
            cudaFree(wl->by_library);
            wl->by_library = nullptr;
            MDG_CUDA(ctx, cudaMalloc(&wl->by_library, (size_t)n_reads * 4));
        }
        wl->cap = n_reads;
    }
    *out = wl;
    return MDG_OK;
}

// The count tables as library `lib` sees them: its own slabs at index 0 of the library axis.
mdg::CountTables lib_tables(const mdg_ctx *ctx, int lib)
{
    mdg::CountTables t = ctx->count_tables;
    const size_t L = ctx->cfg.length, A = ctx->cfg.around;
    t.misincorp += (size_t)lib * 4 * MDG_N_CLASSES * L;
    t.dnacomp += (size_t)lib * 16 * (L + A);
    t.lghist += (size_t)lib * 4 * ctx->cfg.lg_bins;
    return t;
}

// compiled shapes of the warp-specialised bit-plane kernel: producer teams x warps per team + consumer warps
using WsKernel = void (*)(mdg::DevBatch, mdg::DevRef, mdg::CountParams, mdg::CountTables, mdg::PlaneGeom, uint32_t *, unsigned long long *,
                          uint32_t *, unsigned long long *, mdg::SwarSubset);
struct WsVariant {
    const char *name;
    int teams, team_warps, cons_warps, nw_anchor;  // nw_anchor = ceil((L + A) / 32) is compiled in
    // one library per launch; every read's own of two libraries.  First index: 0 genome in L2 (one rolled stage loop), 1 a
    // genome that does not fit (all of a window's gathers in flight before its first word); second index: 1 stages reads
    // with one insertion / deletion itself
    WsKernel kernel[2][2], kernel_two_libraries[2][2];
    WsKernel kernel_qual[2], kernel_two_libraries_qual[2];  // with the -Q mask (first index; never stage one-indel reads)
};
const WsVariant WS_VARIANTS[] = {
#define MDG_WS(teams, warps, cons, nwa, libs)                                                                                         \
    {                                                                                                                                \
        {mdg::count_planes_ws_kernel<teams, warps, cons, nwa, libs, false, false, false>, mdg::count_planes_ws_kernel<teams, warps, cons, nwa, libs, false, true, false>}, \
        {mdg::count_planes_ws_kernel<teams, warps, cons, nwa, libs, true, false, false>, mdg::count_planes_ws_kernel<teams, warps, cons, nwa, libs, true, true, false>}    \
    }
#define MDG_WSQ(teams, warps, cons, nwa, libs) \
    {mdg::count_planes_ws_kernel<teams, warps, cons, nwa, libs, false, false, true>, mdg::count_planes_ws_kernel<teams, warps, cons, nwa, libs, true, false, true>}
#define MDG_WS_NONE {{nullptr, nullptr}, {nullptr, nullptr}}
#define MDG_WSQ_NONE {nullptr, nullptr}
    {"2x9+8", 2, 9, 8, 3, MDG_WS(2, 9, 8, 3, 1), MDG_WS_NONE, MDG_WSQ(2, 9, 8, 3, 1), MDG_WSQ_NONE},  // the default: 0.39 ms per 4 M 100 bp reads
    {"2x9+8", 2, 9, 8, 2, MDG_WS(2, 9, 8, 2, 1), MDG_WS_NONE, MDG_WSQ(2, 9, 8, 2, 1), MDG_WSQ_NONE},
    {"2x8+8", 2, 8, 8, 3, MDG_WS(2, 8, 8, 3, 1), MDG_WS(2, 8, 8, 3, 2), MDG_WSQ(2, 8, 8, 3, 1), MDG_WSQ(2, 8, 8, 3, 2)},
    {"2x8+8", 2, 8, 8, 2, MDG_WS(2, 8, 8, 2, 1), MDG_WS(2, 8, 8, 2, 2), MDG_WSQ(2, 8, 8, 2, 1), MDG_WSQ(2, 8, 8, 2, 2)},
#undef MDG_WS
#undef MDG_WS_NONE
#undef MDG_WSQ
#undef MDG_WSQ_NONE
};

// The counting kernels over one device batch.
int launch_count(mdg_ctx *ctx, const mdg::DevBatch &view, bool has_qual, cudaStream_t stream)
{
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "mdg_set_reference must be called before counting");
    if (view.n_reads == 0) return MDG_OK;
    ctx->reduced_valid = false;  // the local tables move on: sums over ranks need a new mdg_allreduce_tables
    mdg::DevBatch b = view;
    if (!has_qual) b.qual = nullptr;
    mdg::CountParams p{ctx->cfg.length, ctx->cfg.around, ctx->cfg.min_qual, ctx->cfg.n_libraries, ctx->cfg.lg_bins};
    cudaEvent_t e0, e1;
    int rc = next_kernel_events(ctx, &e0, &e1);
    if (rc) return rc;
    MDG_CUDA(ctx, cudaEventRecord(e0, stream));
    if (ctx->swar_enabled && !ctx->force_general) {
        WorkList *wl = nullptr;
        rc = worklist_for(ctx, stream, b.n_reads, &wl);
        if (rc) return rc;
        MDG_CUDA(ctx, cudaMemsetAsync(wl->count, 0, (size_t)ctx->cfg.n_libraries * 8, stream));
        const int64_t n_tiles = (b.n_reads + ctx->swar.tile - 1) / ctx->swar.tile;
        const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * ctx->swar_blocks_per_sm, n_tiles);
        const bool q = b.qual && p.min_qual > 0;
        const int nl = ctx->cfg.n_libraries;
        if (ctx->staged_enabled && ctx->staged_indels_auto && !ctx->staged_indels && ctx->indel_seen_host) {
            // what earlier launches reported (the copy may lag a launch or two; it only steers speed)
            const unsigned long long seen = *(volatile unsigned long long *)ctx->indel_seen_host;
            if (seen * 50 > (unsigned long long)ctx->reads_launched && seen > 1000) ctx->staged_indels = true;
        }
        // the bit-plane kernel counts the gap-free reads (no quality mask); reads with one short indel come back in
        // a list for the staged kernel's indel variant
        // (under a quality mask: its warp-specialised form only, MDG_PLANES_QUAL=0: the staged kernel as in round 1)
        const bool use_planes = ctx->planes_enabled && (!q || (ctx->ws_variant >= 0 && ctx->ws_qual));
        if (use_planes) MDG_CUDA(ctx, cudaMemsetAsync(wl->indel_count, 0, (size_t)nl * 24, stream));
        // one launch of the bit-sliced kernel over a library's reads (or all reads) into the tables `tl`
        auto launch_bitsliced = [&](const mdg::CountTables &tl, const mdg::SwarSubset &subset) {
            if (use_planes && ctx->ws_variant >= 0) {
                const bool together = !subset.list && subset.offsets;  // every library in this launch
                mdg::PlaneGeom pg = together ? ctx->ws_libraries_geom : ctx->ws;
                pg.indel_seen = ctx->indel_seen_dev;
                // genome image larger than what stays in L2: prefetch each read's genome entries while it is parsed, and the
                // stage that has all of a window's gathers in flight at once
                const bool big = ctx->ref_words_bytes > ((size_t)48 << 20);
                if (!getenv("MDG_PLANES_PREFETCH") && big) pg.prefetch_bases |= 2;
                // reads with one insertion / deletion staged by this kernel once the data have shown such reads
                // (MDG_PLANES_INDELS=0 / 1 pins it: A/B, tests)
                const char *indel_env = getenv("MDG_PLANES_INDELS");
                const int indels = indel_env ? atoi(indel_env) != 0 : ctx->staged_indels;
                const char *gather_env = getenv("MDG_PLANES_GATHER");
                const int gather = gather_env ? atoi(gather_env) != 0 : big;
                const WsVariant &v = WS_VARIANTS[together ? ctx->ws_variant_libraries : ctx->ws_variant];
                const int64_t tiles = (b.n_reads + pg.tile - 1) / pg.tile;
                const int pgrid = (int)std::min<int64_t>((int64_t)ctx->sm_count, (tiles + v.teams - 1) / v.teams);
                (q ? (together ? v.kernel_two_libraries_qual : v.kernel_qual)[gather]
                   : (together ? v.kernel_two_libraries : v.kernel)[gather][indels])<<<pgrid, pg.threads, together ? ctx->ws_smem_libraries : ctx->ws_smem, stream>>>(
                    b, ctx->ref, p, tl, pg, wl->reads, wl->count, wl->indel_reads, wl->indel_count, subset);
            } else if (use_planes) {
                mdg::PlaneGeom pg = ctx->planes;
                pg.indel_seen = ctx->indel_seen_dev;
                const int64_t tiles = (b.n_reads + pg.tile - 1) / pg.tile;
                const int pgrid = (int)std::min<int64_t>((int64_t)ctx->sm_count * (pg.threads == 256 ? 2 : 1), tiles);
                if (pg.threads == 256)
                    mdg::count_planes_kernel<256><<<pgrid, pg.threads, ctx->planes_smem, stream>>>(b, ctx->ref, p, tl, pg, wl->reads, wl->count,
                                                                                                   wl->indel_reads, wl->indel_count, subset);
                else
                    mdg::count_planes_kernel<512><<<pgrid, pg.threads, ctx->planes_smem, stream>>>(b, ctx->ref, p, tl, pg, wl->reads, wl->count,
                                                                                                   wl->indel_reads, wl->indel_count, subset);
            } else if (ctx->staged_enabled) {
                // three planes (quality mask or indel reads staged) leave room for fewer reads per tile
                const bool three = q || ctx->staged_indels;
                mdg::StagedGeom sg = ctx->staged;
                sg.indel_seen = ctx->indel_seen_dev;
                sg.tile = three ? ctx->staged_tile_qual : ctx->staged_tile_plain;
                const int64_t tiles = (b.n_reads + sg.tile - 1) / sg.tile;
                const int sgrid = (int)std::min<int64_t>((int64_t)ctx->sm_count * ctx->staged_blocks_per_sm, tiles);
                staged_kernel(q, ctx->staged_indels, ctx->staged_threads)<<<sgrid, sg.threads, three ? ctx->staged_smem_qual : ctx->staged_smem_plain, stream>>>(
                    b, ctx->ref, p, tl, sg, wl->reads, wl->count, subset);
            } else {
                swar_kernel(q, ctx->swar_max_threads, ctx->swar_reads)<<<grid, ctx->swar.threads, ctx->swar_smem, stream>>>(
                    b, ctx->ref, p, tl, ctx->swar, wl->reads, wl->count, subset);
            }
        };
        if (nl == 1) {
            launch_bitsliced(ctx->count_tables, mdg::SwarSubset{nullptr, nullptr, 0});
            ctx->launches += 1;
        } else {
            // group the reads by library, then one pass per library into that library's tables
            unsigned long long *counts = wl->lib_scratch, *offsets = counts + nl, *cursors = offsets + nl + 1;
            MDG_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)nl * 8, stream));
            const int pgrid = (int)std::min<int64_t>((b.n_reads + 1023) / 1024, (int64_t)ctx->sm_count * 8);
            mdg::library_count_kernel<<<pgrid, 256, (size_t)nl * 4, stream>>>(b, nl, counts, ctx->count_tables.error_flag);
            mdg::library_offsets_kernel<<<1, 256, 0, stream>>>(counts, nl, offsets, cursors);
            if (use_planes && ctx->ws_variant >= 0 && ctx->ws_variant_libraries >= 0) {
                // one launch: a read's library picks its counters and tables; the offsets place the per-library lists of
                // the reads left to the other kernels
                MDG_CUDA(ctx, cudaGetLastError());
                launch_bitsliced(ctx->count_tables, mdg::SwarSubset{nullptr, offsets, 0});
                ctx->launches += 3;
            } else {
                mdg::library_scatter_kernel<<<(unsigned)((b.n_reads + 1023) / 1024), 256, (size_t)nl * 8, stream>>>(b, nl, cursors,
                                                                                                                    wl->by_library);
                MDG_CUDA(ctx, cudaGetLastError());
                for (int lib = 0; lib < nl; ++lib) launch_bitsliced(lib_tables(ctx, lib), mdg::SwarSubset{wl->by_library, offsets, lib});
                ctx->launches += 3 + nl;
            }
        }
        MDG_CUDA(ctx, cudaGetLastError());
        if (use_planes) {
            // the one-indel reads: count_staged_kernel's three-plane variant over each library's stretch of the list
            unsigned long long *const bounds = wl->indel_count + nl;
            mdg::indel_bounds_kernel<<<1, 64, 0, stream>>>(wl->indel_count, nl > 1 ? wl->lib_scratch + nl : nullptr, nl, bounds);
            mdg::StagedGeom sg = ctx->staged;
            sg.indel_seen = nullptr;
            sg.tile = ctx->staged_tile_qual;
            const int64_t tiles = (b.n_reads + sg.tile - 1) / sg.tile;
            const int sgrid = (int)std::min<int64_t>((int64_t)ctx->sm_count * ctx->staged_blocks_per_sm, tiles);
            for (int lib = 0; lib < nl; ++lib)
                staged_kernel(q, true, ctx->staged_threads)<<<sgrid, sg.threads, ctx->staged_smem_qual, stream>>>(
                    b, ctx->ref, p, nl == 1 ? ctx->count_tables : lib_tables(ctx, lib), sg, wl->reads, wl->count,
                    mdg::SwarSubset{wl->indel_reads, bounds + lib, lib});
            MDG_CUDA(ctx, cudaGetLastError());
            ctx->launches += 1 + nl;
        }
        ctx->reads_launched += b.n_reads;
        if (ctx->staged_enabled && ctx->staged_indels_auto && !ctx->staged_indels && ctx->indel_seen_host)
            MDG_CUDA(ctx, cudaMemcpyAsync(ctx->indel_seen_host, ctx->indel_seen_dev, 8, cudaMemcpyDeviceToHost, stream));
        // reads with indels / skips: the general kernel over the work list(s) (returns at once when empty)
        const int ggrid = (int)std::min<int64_t>(ctx->general_grid, (b.n_reads + 7) / 8);
        if (nl == 1 || !ctx->shared_slab) {
            const unsigned long long *offsets = nl == 1 ? nullptr : wl->lib_scratch + nl;
            if (nl > 1) {
                // no room for a shared-memory slab: one launch per library all the same, into the global tables
                for (int lib = 0; lib < nl; ++lib)
                    mdg::count_general_kernel<false><<<ggrid, 256, 0, stream>>>(b, ctx->ref, p, lib_tables(ctx, lib), wl->reads,
                                                                                wl->count, offsets, lib);
                ctx->launches += nl;
            } else if (ctx->shared_slab) {
                mdg::count_general_kernel<true><<<ggrid, 256, ctx->slab_bytes, stream>>>(b, ctx->ref, p, ctx->count_tables,
                                                                                         wl->reads, wl->count, nullptr, 0);
                ctx->launches += 1;
            } else {
                mdg::count_general_kernel<false><<<ggrid, 256, 0, stream>>>(b, ctx->ref, p, ctx->count_tables, wl->reads,
                                                                            wl->count, nullptr, 0);
                ctx->launches += 1;
            }
        } else {
            for (int lib = 0; lib < nl; ++lib)
                mdg::count_general_kernel<true><<<ggrid, 256, ctx->slab_bytes, stream>>>(b, ctx->ref, p, lib_tables(ctx, lib),
                                                                                         wl->reads, wl->count,
                                                                                         wl->lib_scratch + nl, lib);
            ctx->launches += nl;
        }
        MDG_CUDA(ctx, cudaGetLastError());
    } else {
        int grid = (int)std::min<int64_t>(ctx->general_grid, (b.n_reads + 7) / 8);
        if (grid < 1) grid = 1;
        if (ctx->shared_slab && ctx->cfg.n_libraries == 1)
            mdg::count_general_kernel<true><<<grid, 256, ctx->slab_bytes, stream>>>(b, ctx->ref, p, ctx->count_tables, nullptr, nullptr,
                                                                                    nullptr, 0);
        else
            mdg::count_general_kernel<false><<<grid, 256, 0, stream>>>(b, ctx->ref, p, ctx->count_tables, nullptr, nullptr,
                                                                       nullptr, 0);
        ctx->launches += 1;
        MDG_CUDA(ctx, cudaGetLastError());
    }
    MDG_CUDA(ctx, cudaEventRecord(e1, stream));
    return MDG_OK;
}

int check_device_errors(mdg_ctx *ctx)
{
    int32_t flag = 0;
    MDG_CUDA(ctx, cudaMemcpy(&flag, ctx->count_tables.error_flag, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) {
        int32_t zero = 0;
        cudaMemcpy(ctx->count_tables.error_flag, &zero, sizeof zero, cudaMemcpyHostToDevice);
        switch (flag) {
        case mdg::DATA_ERR_LIB: return fail(ctx, MDG_ERR_DATA, "a read's library index is >= n_libraries");
        case mdg::DATA_ERR_TID: return fail(ctx, MDG_ERR_DATA, "a mapped read has no CIGAR or a reference id outside the genome");
        case mdg::DATA_ERR_QUAL: return fail(ctx, MDG_ERR_DATA, "a base quality above 93 cannot be rescaled");
        case mdg::DATA_ERR_LAYOUT: return fail(ctx, MDG_ERR_DATA, "n_bases does not match the read lengths of a batch passed without base_off");
        case mdg::DATA_ERR_CLIP: return fail(ctx, MDG_ERR_DATA, "quality and sequence mismatch: soft clip behind a hard clip (reference rescale.py:266-273 fails the same way)");
        default: return fail(ctx, MDG_ERR_DATA, "device reported data error %d", flag);
        }
    }
    return MDG_OK;
}

}  // namespace

extern "C" {

int mdg_abi_version(void) { return MDG_ABI_VERSION; }

const char *mdg_last_error(const mdg_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int mdg_device_pci_bus_id(int32_t device, char *buf, int32_t cap)
{
    if (!buf || cap < 16) return MDG_ERR_ARGUMENT;
    if (cudaDeviceGetPCIBusId(buf, cap, device) != cudaSuccess) {
        cudaGetLastError();
        return MDG_ERR_NO_DEVICE;
    }
    return MDG_OK;
}

void *mdg_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void mdg_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

int mdg_create(mdg_ctx **out, const mdg_config *cfg)
{
    if (!out || !cfg) return fail(nullptr, MDG_ERR_ARGUMENT, "mdg_create: NULL argument");
    *out = nullptr;
    if (cfg->length < 1 || cfg->around < 0 || cfg->n_libraries < 1 || cfg->lg_bins < 1 || cfg->min_qual < 0 ||
        cfg->n_libraries > 65535)
        return fail(nullptr, MDG_ERR_ARGUMENT, "mdg_create: invalid length/around/min_qual/n_libraries/lg_bins");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, MDG_ERR_NO_DEVICE, "no CUDA device available (%s); mapdamage_b200 has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (cfg->device < 0 || cfg->device >= n_dev)
        return fail(nullptr, MDG_ERR_ARGUMENT, "mdg_create: device %d out of range (0..%d)", cfg->device, n_dev - 1);
    mdg_ctx *ctx = new (std::nothrow) mdg_ctx();
    if (!ctx) return fail(nullptr, MDG_ERR_ARGUMENT, "out of host memory");
    ctx->cfg = *cfg;
    if (ctx->cfg.n_slots < 1) ctx->cfg.n_slots = 2;
    auto bail = [&](int code) {
        g_create_error = ctx->error;
        mdg_destroy(ctx);
        return code;
    };
#define MDG_CREATE_CUDA(call)                                                                       \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fail(ctx, MDG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));               \
            return bail(MDG_ERR_CUDA);                                                              \
        }                                                                                           \
    } while (0)
    MDG_CREATE_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    MDG_CREATE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    MDG_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking));
    MDG_CREATE_CUDA(cudaEventCreate(&ctx->ev[0]));
    MDG_CREATE_CUDA(cudaEventCreate(&ctx->ev[1]));

    const size_t L = cfg->length, A = cfg->around, nl = cfg->n_libraries;
    ctx->n_mis = nl * 4 * MDG_N_CLASSES * L;
    ctx->n_comp = nl * 16 * (L + A);
    ctx->n_lg = nl * 4 * (size_t)cfg->lg_bins;
    const size_t n_tables = ctx->n_mis + ctx->n_comp + ctx->n_lg;
    MDG_CREATE_CUDA(cudaMalloc(&ctx->tables, n_tables * 8));
    MDG_CREATE_CUDA(cudaMemset(ctx->tables, 0, n_tables * 8));
    const int64_t overflow_cap = 1 << 20;
    const size_t aux_bytes = 256 + 64 + overflow_cap * 16;
    MDG_CREATE_CUDA(cudaMalloc(&ctx->aux_block, aux_bytes));
    MDG_CREATE_CUDA(cudaMemset(ctx->aux_block, 0, aux_bytes));
    char *aux = (char *)ctx->aux_block;
    ctx->count_tables.misincorp = ctx->tables;
    ctx->count_tables.dnacomp = ctx->tables + ctx->n_mis;
    ctx->count_tables.lghist = ctx->tables + ctx->n_mis + ctx->n_comp;
    ctx->count_tables.lg_overflow_count = (unsigned long long *)(aux + 0);
    ctx->count_tables.error_flag = (int32_t *)(aux + 8);
    ctx->rescale_stats = (unsigned long long *)(aux + 64);
    ctx->count_tables.lg_overflow_rows = (int32_t *)(aux + 256 + 64);
    ctx->count_tables.lg_overflow_cap = overflow_cap;

    // general kernel: shared-memory slab when there is one library and it fits
    ctx->slab_bytes = (4 * MDG_N_CLASSES * L + 16 * (L + A) + 4 * MDG_LG_SMEM_BINS) * 4;
    ctx->shared_slab = ctx->slab_bytes + 1024 <= ctx->smem_optin;  // one library's slab fits a block
    int per_sm = 0;
    if (ctx->shared_slab) {
        MDG_CREATE_CUDA(cudaFuncSetAttribute(mdg::count_general_kernel<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->slab_bytes));
        MDG_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mdg::count_general_kernel<true>, 256,
                                                                      ctx->slab_bytes));
    } else {
        MDG_CREATE_CUDA(
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mdg::count_general_kernel<false>, 256, 0));
    }
    if (per_sm < 1) per_sm = 1;
    ctx->general_grid = ctx->sm_count * per_sm;
    // bit-sliced kernel: one library, a shared-memory general slab for its work list, geometry that fits a block
    {
        mdg::SwarGeom &g = ctx->swar;
        g.words = (cfg->around + cfg->length + 7) / 8;
        // block size / reads per iteration; MDG_SWAR_VARIANT="threads,reads" picks another compiled variant
        ctx->swar_max_threads = 256;  // two co-resident blocks per SM
        ctx->swar_reads = 1;
        if (const char *venv = getenv("MDG_SWAR_VARIANT")) {
            int vt = 0, vr = 0;
            if (sscanf(venv, "%d,%d", &vt, &vr) == 2 && swar_kernel(false, vt, vr)) {
                ctx->swar_max_threads = vt;
                ctx->swar_reads = vr;
            }
        }
        g.slots = (ctx->swar_max_threads / (2 * g.words)) & ~1;
        if (nl <= mdg::PARTITION_MAX_LIBS && g.slots >= 2 && cfg->around <= 64 && cfg->length < 32768) {
            g.work_threads = 2 * g.words * g.slots;
            g.threads = (g.work_threads + 31) / 32 * 32;
            const char *tile_env = getenv("MDG_SWAR_TILE");
            const int tile_max = tile_env ? atoi(tile_env) : 2048;
            const int blocks_per_sm = ctx->swar_max_threads == 256 ? 2 : 1;
            ctx->swar_blocks_per_sm = blocks_per_sm;
            for (int tile : {2048, 1024, 512}) {
                if (tile > tile_max) continue;
                const size_t bytes = ((size_t)mdg::SWAR_L2_WORDS * g.threads + (size_t)tile * 5 + 4 * MDG_LG_SMEM_BINS + 4 * L + 8) * 4;
                if (bytes * blocks_per_sm + 1024 * blocks_per_sm <= ctx->smem_optin + (blocks_per_sm > 1 ? 1024 : 0)) {
                    g.tile = tile;
                    ctx->swar_smem = bytes;
                    break;
                }
            }
            const char *uenv = getenv("MDG_SWAR_UNIFORM");
            g.uniform = !(uenv && uenv[0] == '0');
            const char *fenv = getenv("MDG_SWAR_FLUSH_TILES");
            g.flush_tiles = fenv ? atoi(fenv) : 0;
            if (g.tile) {
                for (bool q : {false, true})
                    MDG_CREATE_CUDA(cudaFuncSetAttribute(swar_kernel(q, ctx->swar_max_threads, ctx->swar_reads),
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->swar_smem));
                ctx->swar_enabled = true;
            }
        }
        // staged kernel geometry
        {
            mdg::StagedGeom &sg = ctx->staged;
            sg.words = g.words;
            sg.uniform = g.uniform;
            sg.flush_tiles = g.flush_tiles;
            const char *kenv = getenv("MDG_KERNEL");
            const char *tenv = getenv("MDG_STAGE_THREADS");
            ctx->staged_threads = tenv && staged_kernel(false, false, atoi(tenv)) ? atoi(tenv) : 512;
            const char *ienv = getenv("MDG_STAGE_INDELS");
            ctx->staged_indels = ienv ? ienv[0] == '1' : false;
            ctx->staged_indels_auto = !ienv;
            ctx->staged_blocks_per_sm = ctx->staged_threads == 256 ? 2 : 1;
            const int wpr = 2 * sg.words;
            sg.threads = ctx->staged_threads / 32 * 32;
            const bool fits_block = wpr * 2 <= sg.threads && cfg->length + cfg->around <= 2048 && cfg->around <= 64;
            if (ctx->swar_enabled && fits_block && !(kenv && !strcmp(kenv, "swar"))) {
                const size_t budget = (ctx->smem_optin + 1024) / ctx->staged_blocks_per_sm - 1024;
                for (int with_qual = 0; with_qual < 2; ++with_qual) {
                    const size_t nw = with_qual ? 3 : 2;
                    const size_t fixed = ((size_t)32 * sg.threads + (size_t)wpr * 2 * 96 + 2 * wpr + 4 * MDG_LG_SMEM_BINS + 4 * L + 16) * 4;
                    const size_t per_read = (4 + (size_t)(wpr | 1) * nw + 1 + (with_qual ? 2 : 0)) * 4;
                    int tile = 0;
                    if (fixed + 64 * per_read <= budget) tile = (int)std::min<size_t>(2048, (budget - fixed) / per_read / 32 * 32);
                    if (const char *tile_env2 = getenv("MDG_STAGE_TILE")) tile = std::min(tile, std::max(32, atoi(tile_env2)));
                    const size_t bytes = fixed + (size_t)tile * per_read;
                    (with_qual ? ctx->staged_tile_qual : ctx->staged_tile_plain) = tile;
                    (with_qual ? ctx->staged_smem_qual : ctx->staged_smem_plain) = bytes;
                    if (tile) {
                        // with_qual sizes the three-plane layout: the quality-mask kernel and the indel-staging one
                        MDG_CREATE_CUDA(cudaFuncSetAttribute(staged_kernel(with_qual != 0, with_qual != 0, ctx->staged_threads),
                                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                        if (with_qual)
                            MDG_CREATE_CUDA(cudaFuncSetAttribute(staged_kernel(false, true, ctx->staged_threads),
                                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                    }
                }
                ctx->staged_enabled = ctx->staged_tile_plain >= 64 && ctx->staged_tile_qual >= 64;
                if (ctx->staged_enabled) {
                    MDG_CREATE_CUDA(cudaMalloc(&ctx->indel_seen_dev, 8));
                    MDG_CREATE_CUDA(cudaMemset(ctx->indel_seen_dev, 0, 8));
                    MDG_CREATE_CUDA(cudaHostAlloc((void **)&ctx->indel_seen_host, 8, cudaHostAllocDefault));
                    *ctx->indel_seen_host = 0;
                }
            }
        }
        // bit-plane kernel geometry
        {
            mdg::PlaneGeom &pg = ctx->planes;
            const char *kenv = getenv("MDG_KERNEL");
            const char *pt_env = getenv("MDG_PLANES_THREADS");  // 256: two co-resident blocks per SM (A/B)
            pg.threads = pt_env && atoi(pt_env) == 256 ? 256 : 512;
            const size_t smem_budget = pg.threads == 256 ? (ctx->smem_optin + 1024) / 2 - 1024 : ctx->smem_optin;
            const char *pf_env = getenv("MDG_PLANES_PREFETCH");
            pg.prefetch_bases = pf_env ? atoi(pf_env) : 1;  // bit 0: bases of the tile after next (one-role kernel); bits 1, 2: A/B switches of the warp-specialised kernel
            pg.uniform = g.uniform;
            pg.flush_tiles = g.flush_tiles;
            pg.nw_anchor = (cfg->length + cfg->around + 31) / 32;
            pg.row_words = 16 * pg.nw_anchor + 4;
            const int pairs = (pg.threads >> 7) * 32;
            const size_t fixed = ((size_t)mdg::PL_WIDE * mdg::PL_CLASSES * pg.threads + 4 * pg.nw_anchor + 4 * MDG_LG_SMEM_BINS + 4 * L + 64 +
                                  (size_t)2 * 12 * 64 * pg.nw_anchor) * 4 + 256;
            // per read: the staged row, the record, two list slots, and 56 bytes of the tile's seq4 stretch (reads of up
            // to about 110 bases on average; a tile of longer reads takes its bases from global memory)
            const char *slab_env = getenv("MDG_PLANES_SLAB");
            const size_t seq_per_read = slab_env && slab_env[0] == '0' ? 0 : 56;
            const size_t per_read = ((size_t)pg.row_words + 4 + 2) * 4 + seq_per_read;
            int tile = 0;
            if (fixed + 192 * per_read <= smem_budget) tile = (int)std::min<size_t>(1024, (smem_budget - fixed) / per_read / 32 * 32);
            // whole rounds of the block: a tile of 576 reads would leave 448 threads idle in the second round of every phase
            if (tile > pg.threads) tile = tile / pg.threads * pg.threads;
            if (const char *tile_env3 = getenv("MDG_PLANES_TILE")) tile = std::min(tile, std::max(32, atoi(tile_env3)));
            pg.tile = tile;
            pg.seq_words = (int)((size_t)tile * seq_per_read / 4 / 4 * 4);
            ctx->planes_smem = fixed + (size_t)tile * per_read;
            const bool fits = ctx->staged_enabled && 2 * pg.nw_anchor * 2 <= pairs && cfg->around <= 64 && tile >= 192 &&
                              (size_t)tile * pg.row_words >= (size_t)2 * 20 * 64 * pg.nw_anchor;
            if (fits && !(kenv && (!strcmp(kenv, "swar") || !strcmp(kenv, "staged")))) {
                if (pg.threads == 256)
                    MDG_CREATE_CUDA(cudaFuncSetAttribute(mdg::count_planes_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)ctx->planes_smem));
                else
                    MDG_CREATE_CUDA(cudaFuncSetAttribute(mdg::count_planes_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)ctx->planes_smem));
                ctx->planes_enabled = true;
            }
        }
        // warp-specialised bit-plane kernel: MDG_PLANES_WS=<variant name> (or 0: off)
        if (ctx->planes_enabled) {
            const char *ws_env = getenv("MDG_PLANES_WS");
            const char *want = ws_env ? ws_env : "2x9+8";  // the default; MDG_PLANES_WS=0: the one-role kernel (count_planes_kernel)
            const char *lib_env = getenv("MDG_PLANES_WS_LIBS");  // 0: two libraries as two launches over index lists (A/B, tests)
            const char *slab_env = getenv("MDG_PLANES_SLAB");
            for (int i = 0; i < (int)(sizeof(WS_VARIANTS) / sizeof(WS_VARIANTS[0])); ++i) {
                const WsVariant &v = WS_VARIANTS[i];
                if (v.nw_anchor != ctx->planes.nw_anchor) continue;
                mdg::PlaneGeom wg = ctx->planes;
                wg.threads = (v.teams * v.team_warps + v.cons_warps) * 32;
                wg.tile = v.team_warps * 32;
                wg.seq_words = slab_env && slab_env[0] == '0' ? 0 : wg.tile * 56 / 4;
                const int pairs = (v.cons_warps * 32 >> 7) * 32;
                const size_t bytes = mdg::planes_ws_smem(v.teams, v.team_warps, v.cons_warps, (int)L, wg.nw_anchor, 1);
                if (ctx->ws_variant < 0 && !strcmp(want, v.name) && bytes <= ctx->smem_optin && 2 * wg.nw_anchor * 2 <= pairs) {
                    for (int which = 0; which < 4; ++which)
                        MDG_CREATE_CUDA(cudaFuncSetAttribute(v.kernel[which >> 1][which & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                    for (int gather = 0; gather < 2; ++gather)
                        MDG_CREATE_CUDA(cudaFuncSetAttribute(v.kernel_qual[gather], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                    ctx->ws = wg;
                    ctx->ws_smem = bytes;
                    ctx->ws_variant = i;
                }
                // two libraries in one launch: the first shape whose two sets of event tables fit next to the stage buffers
                // and that gives every group (library, strand) a read slot in the two-window layout
                const size_t bytes_two = mdg::planes_ws_smem(v.teams, v.team_warps, v.cons_warps, (int)L, wg.nw_anchor, 2);
                if (ctx->ws_variant_libraries < 0 && nl == 2 && v.kernel_two_libraries[0][0] && strcmp(want, "0") && !(lib_env && lib_env[0] == '0') &&
                    bytes_two <= ctx->smem_optin && 2 * wg.nw_anchor * 2 * 2 <= pairs) {
                    for (int which = 0; which < 4; ++which)
                        MDG_CREATE_CUDA(cudaFuncSetAttribute(v.kernel_two_libraries[which >> 1][which & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_two));
                    for (int gather = 0; gather < 2; ++gather)
                        MDG_CREATE_CUDA(cudaFuncSetAttribute(v.kernel_two_libraries_qual[gather], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_two));
                    ctx->ws_libraries_geom = wg;
                    ctx->ws_smem_libraries = bytes_two;
                    ctx->ws_variant_libraries = i;
                }
            }
            if (ctx->ws_variant < 0) ctx->ws_variant_libraries = -1;
            if (const char *qual_env = getenv("MDG_PLANES_QUAL")) ctx->ws_qual = atoi(qual_env) != 0;
        }
        const char *env = getenv("MDG_FORCE_GENERAL");
        ctx->force_general = env && env[0] == '1';
    }

    // staging slots
    ctx->slots.resize(ctx->cfg.n_slots);
    if (cfg->max_reads > 0) {
        for (auto &slot : ctx->slots) {
            MDG_CREATE_CUDA(cudaStreamCreateWithFlags(&slot.stream, cudaStreamNonBlocking));
            int64_t bases = (cfg->max_bases + 1) & ~1ll;
            if (alloc_arrays(ctx, slot.arrays, cfg->max_reads, cfg->max_cigar_ops, bases, true)) return bail(MDG_ERR_CUDA);
            MDG_CREATE_CUDA(cudaMalloc(&slot.mr_out, cfg->max_reads * 4 + 8));
            MDG_CREATE_CUDA(cudaMalloc(&slot.status_out, cfg->max_reads + 8));
        }
    }
#undef MDG_CREATE_CUDA
    *out = ctx;
    return MDG_OK;
}

void mdg_destroy(mdg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    if (ctx->comm && nccl_api().CommDestroy) nccl_api().CommDestroy(ctx->comm);
    for (auto &slot : ctx->slots) {
        if (slot.stream) cudaStreamDestroy(slot.stream);
        cudaFree(slot.arrays.block);
        cudaFree(slot.mr_out);
        cudaFree(slot.status_out);
        cudaFree(slot.change_at);
        cudaFree(slot.change_q);
        cudaFree(slot.n_changes);
        if (slot.n_changes_host) cudaFreeHost(slot.n_changes_host);
    }
    cudaFree(ctx->indel_seen_dev);
    if (ctx->indel_seen_host) cudaFreeHost(ctx->indel_seen_host);
    for (auto &w : ctx->worklists) {
        cudaFree(w.reads);
        cudaFree(w.indel_reads);
        cudaFree(w.indel_count);
        cudaFree(w.count);
        cudaFree(w.by_library);
        cudaFree(w.lib_scratch);
    }
    for (cudaEvent_t e : ctx->kernel_events) cudaEventDestroy(e);
    if (ctx->ev[0]) cudaEventDestroy(ctx->ev[0]);
    if (ctx->ev[1]) cudaEventDestroy(ctx->ev[1]);
    cudaFree(ctx->ref_block);
    cudaFree(ctx->planes_block);
    cudaFree(ctx->tables);
    cudaFree(ctx->reduced);
    cudaFree(ctx->aux_block);
    cudaFree(ctx->model_block);
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    cudaGetLastError();
    delete ctx;
}

int mdg_set_reference(mdg_ctx *ctx, const uint8_t *packed, int64_t n_bytes, const uint64_t *contig_off,
                      const uint32_t *contig_len, int32_t n_contigs)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    if (!packed || !contig_off || !contig_len || n_contigs < 1 || n_bytes < 0)
        return fail(ctx, MDG_ERR_ARGUMENT, "mdg_set_reference: NULL or empty argument");
    for (int c = 0; c < n_contigs; ++c) {
        if (contig_off[c] & 7) return fail(ctx, MDG_ERR_ARGUMENT, "contig %d does not start on a multiple of 8 bases", c);
        uint64_t end = contig_off[c] + ((uint64_t)contig_len[c] + 7) / 8 * 8;
        if (end > (uint64_t)n_bytes * 2)
            return fail(ctx, MDG_ERR_ARGUMENT, "contig %d extends past the packed genome (%lld bytes)", c, (long long)n_bytes);
    }
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(ctx->ref_block);
    ctx->ref_block = nullptr;
    ctx->ref = mdg::DevRef{};
    // 4 KB of "not a base" on both sides: the kernels read whole words around an alignment (up to L + A bases away)
    const size_t pad = 4096;
    size_t words_bytes = align_up((size_t)n_bytes + pad);
    ctx->ref_words_bytes = words_bytes;
    size_t off_bytes = align_up((size_t)n_contigs * 8);
    size_t len_bytes = align_up((size_t)n_contigs * 4);
    MDG_CUDA(ctx, cudaMalloc(&ctx->ref_block, pad + words_bytes + off_bytes + len_bytes));
    char *p = (char *)ctx->ref_block + pad;
    MDG_CUDA(ctx, cudaMemset(ctx->ref_block, 0x77, pad + words_bytes));
    MDG_CUDA(ctx, cudaMemcpy(p, packed, (size_t)n_bytes, cudaMemcpyHostToDevice));
    // device image: one-hot nibbles (A,C,G,T = 1,2,4,8 as in BAM; 0 = anything else)
    mdg::ref_to_one_hot_kernel<<<ctx->sm_count * 8, 256, 0, ctx->compute>>>((uint32_t *)ctx->ref_block,
                                                                             (int64_t)((pad + words_bytes) / 4));
    MDG_CUDA(ctx, cudaGetLastError());
    // the same image as bit planes (32 bases per uint4), padding included
    cudaFree(ctx->planes_block);
    ctx->planes_block = nullptr;
    MDG_CUDA(ctx, cudaMalloc(&ctx->planes_block, pad + words_bytes));
    mdg::ref_planes_kernel<<<ctx->sm_count * 8, 256, 0, ctx->compute>>>((const uint32_t *)ctx->ref_block, (int64_t)((pad + words_bytes) / 16),
                                                                         (uint4 *)ctx->planes_block);
    MDG_CUDA(ctx, cudaGetLastError());
    MDG_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    MDG_CUDA(ctx, cudaMemcpy(p + words_bytes, contig_off, (size_t)n_contigs * 8, cudaMemcpyHostToDevice));
    MDG_CUDA(ctx, cudaMemcpy(p + words_bytes + off_bytes, contig_len, (size_t)n_contigs * 4, cudaMemcpyHostToDevice));
    ctx->ref.words = (const uint32_t *)p;
    ctx->ref.planes = (const uint4 *)((char *)ctx->planes_block + pad);
    ctx->ref.contig_off = (const uint64_t *)(p + words_bytes);
    ctx->ref.contig_len = (const uint32_t *)(p + words_bytes + off_bytes);
    ctx->ref.n_contigs = n_contigs;
    ctx->ref_words = (int64_t)((n_bytes + 3) / 4);
    ctx->ref_total_bases = 0;
    ctx->ref_min_contig = 0xffffffffu;
    for (int c = 0; c < n_contigs; ++c) {
        ctx->ref_total_bases += contig_len[c];
        ctx->ref_min_contig = std::min(ctx->ref_min_contig, contig_len[c]);
    }
    return MDG_OK;
}

// A random genome made on the device (benchmarks with a genome larger than L2: 3.1 Gbp is 1.55 GB here): contig c has
// contig_len[c] uniform A/C/G/T bases, a pure function of (seed, c, position).
int mdg_synth_reference(mdg_ctx *ctx, const uint32_t *contig_len, int32_t n_contigs, uint64_t seed)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    if (!contig_len || n_contigs < 1) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_synth_reference: NULL or empty argument");
    std::vector<uint64_t> off((size_t)n_contigs);
    uint64_t total = 0;
    for (int c = 0; c < n_contigs; ++c) {
        off[(size_t)c] = total;
        total += ((uint64_t)contig_len[c] + 7) / 8 * 8;
    }
    const int64_t n_bytes = (int64_t)(total / 2);
    // an all-"not a base" image through the ordinary path, then the contigs are overwritten in place
    std::vector<uint8_t> blank(1 << 20, 0x77);
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(ctx->ref_block);
    cudaFree(ctx->planes_block);
    ctx->ref_block = ctx->planes_block = nullptr;
    ctx->ref = mdg::DevRef{};
    const size_t pad = 4096;
    const size_t words_bytes = align_up((size_t)n_bytes + pad), off_bytes = align_up((size_t)n_contigs * 8), len_bytes = align_up((size_t)n_contigs * 4);
    ctx->ref_words_bytes = words_bytes;
    MDG_CUDA(ctx, cudaMalloc(&ctx->ref_block, pad + words_bytes + off_bytes + len_bytes));
    MDG_CUDA(ctx, cudaMalloc(&ctx->planes_block, pad + words_bytes));
    char *p = (char *)ctx->ref_block + pad;
    MDG_CUDA(ctx, cudaMemset(ctx->ref_block, 0, pad + words_bytes));  // one-hot image: 0 = not a base
    for (int c = 0; c < n_contigs; ++c) {
        mdg::synth_reference_kernel<<<ctx->sm_count * 8, 256, 0, ctx->compute>>>((uint32_t *)p + off[(size_t)c] / 8, contig_len[c],
                                                                                  mdg::mix64(seed + 0x1000003ull * (uint64_t)c));
        ctx->launches += 1;
    }
    mdg::ref_planes_kernel<<<ctx->sm_count * 8, 256, 0, ctx->compute>>>((const uint32_t *)ctx->ref_block, (int64_t)((pad + words_bytes) / 16),
                                                                         (uint4 *)ctx->planes_block);
    MDG_CUDA(ctx, cudaGetLastError());
    MDG_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    MDG_CUDA(ctx, cudaMemcpy(p + words_bytes, off.data(), (size_t)n_contigs * 8, cudaMemcpyHostToDevice));
    MDG_CUDA(ctx, cudaMemcpy(p + words_bytes + off_bytes, contig_len, (size_t)n_contigs * 4, cudaMemcpyHostToDevice));
    ctx->ref.words = (const uint32_t *)p;
    ctx->ref.planes = (const uint4 *)((char *)ctx->planes_block + pad);
    ctx->ref.contig_off = (const uint64_t *)(p + words_bytes);
    ctx->ref.contig_len = (const uint32_t *)(p + words_bytes + off_bytes);
    ctx->ref.n_contigs = n_contigs;
    ctx->ref_words = (int64_t)((n_bytes + 3) / 4);
    ctx->ref_total_bases = 0;
    ctx->ref_min_contig = 0xffffffffu;
    for (int c = 0; c < n_contigs; ++c) {
        ctx->ref_total_bases += contig_len[c];
        ctx->ref_min_contig = std::min(ctx->ref_min_contig, contig_len[c]);
    }
    return MDG_OK;
}

// The genome image as the device holds it (one-hot nibbles, low nibble = even base; 0 = not a base), n_bytes of it.
int mdg_reference_download(mdg_ctx *ctx, uint8_t *one_hot, int64_t n_bytes)
{
    if (!ctx || !one_hot || n_bytes < 0) return MDG_ERR_ARGUMENT;
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "no reference on the device");
    if (n_bytes > ctx->ref_words * 4) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_reference_download: the image has %lld bytes", (long long)(ctx->ref_words * 4));
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaMemcpy(one_hot, ctx->ref.words, (size_t)n_bytes, cudaMemcpyDeviceToHost));
    return MDG_OK;
}

int mdg_genome_composition(mdg_ctx *ctx, uint64_t *counts4)
{
    if (!ctx || !counts4) return MDG_ERR_ARGUMENT;
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "mdg_set_reference must be called before mdg_genome_composition");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    unsigned long long *dev = nullptr;
    MDG_CUDA(ctx, cudaMalloc(&dev, 32));
    cudaError_t e = cudaMemsetAsync(dev, 0, 32, ctx->compute);
    if (e == cudaSuccess) {
        mdg::genome_composition_kernel<<<ctx->sm_count * 8, 256, 0, ctx->compute>>>(ctx->ref.words, ctx->ref_words, dev);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts4, dev, 32, cudaMemcpyDeviceToHost, ctx->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->compute);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(ctx, MDG_ERR_CUDA, "mdg_genome_composition failed: %s", cudaGetErrorString(e));
    return MDG_OK;
}

int mdg_count_submit(mdg_ctx *ctx, const mdg_batch *host)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    int rc = check_batch(ctx, host);
    if (rc) return rc;
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "mdg_set_reference must be called before counting");
    if (ctx->slots.empty() || !ctx->slots[0].stream)
        return fail(ctx, MDG_ERR_STATE, "context was created without staging slots (max_reads = 0)");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    Slot &slot = ctx->slots[ctx->next_slot];
    ctx->next_slot = (ctx->next_slot + 1) % (int)ctx->slots.size();
    MDG_CUDA(ctx, cudaStreamSynchronize(slot.stream));  // the slot's previous batch is done
    const bool quals = host->qual != nullptr && ctx->cfg.min_qual > 0;
    rc = copy_batch(ctx, slot.arrays, host, slot.stream, false, quals);
    if (rc) return rc;
    return launch_count(ctx, slot.arrays.view, quals, slot.stream);
}

int mdg_batch_upload(mdg_ctx *ctx, const mdg_batch *host, mdg_dev_batch **out)
{
    if (!ctx || !out) return MDG_ERR_ARGUMENT;
    *out = nullptr;
    int rc = check_batch(ctx, host);
    if (rc) return rc;
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    mdg_dev_batch *d = new (std::nothrow) mdg_dev_batch();
    if (!d) return fail(ctx, MDG_ERR_ARGUMENT, "out of host memory");
    const int64_t cigar_words = !host->cigar_off && host->n_cigar == 1 ? std::max<int64_t>(host->n_reads, 1) : host->n_cigar;
    rc = alloc_arrays(ctx, d->arrays, host->n_reads, cigar_words, host->n_bases, host->qual != nullptr);
    if (!rc) rc = copy_batch(ctx, d->arrays, host, ctx->compute, true, true);
    if (!rc && cudaStreamSynchronize(ctx->compute) != cudaSuccess)
        rc = fail(ctx, MDG_ERR_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc) {
        cudaFree(d->arrays.block);
        delete d;
        return rc;
    }
    *out = d;
    return MDG_OK;
}

int mdg_batch_free(mdg_ctx *ctx, mdg_dev_batch *batch)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    if (!batch) return MDG_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(batch->arrays.block);
    cudaFree(batch->res_mr);
    cudaFree(batch->res_status);
    delete batch;
    return MDG_OK;
}

int mdg_count_resident(mdg_ctx *ctx, const mdg_dev_batch *batch)
{
    if (!ctx || !batch) return MDG_ERR_ARGUMENT;
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return launch_count(ctx, batch->arrays.view, batch->arrays.has_qual, ctx->compute);
}

int mdg_sync(mdg_ctx *ctx)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    for (auto &slot : ctx->slots)
        if (slot.stream) MDG_CUDA(ctx, cudaStreamSynchronize(slot.stream));
    MDG_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
#ifdef MDG_PHASE_CLOCKS
    {
        unsigned int pp[16];
        if (cudaMemcpyFromSymbol(pp, mdg::mdg_plane_phase_dump, sizeof(pp)) == cudaSuccess && pp[0])
            fprintf(stderr, "plane kernel phase clocks (block 3, thread 0): parse %u sync %u stage %u sync %u count %u sync+flush %u\n", pp[0], pp[1],
                    pp[2], pp[3], pp[4], pp[5]);
    }
    {
        unsigned int pc[24];
        if (cudaMemcpyFromSymbol(pc, mdg::mdg_phase_dump, sizeof(pc)) == cudaSuccess) {
            static const char *names[12] = {"parse", "sync", "prefetch", "stage", "sync", "count", "sync", "rest",
                                            "worklist", "mode", "-", "-"};
            for (int w = 0; w < 2; ++w) {
                fprintf(stderr, "phase clocks %s warp:", w ? "last " : "first");
                for (int i = 0; i < 10; ++i) fprintf(stderr, " %s %u", names[i], pc[12 * w + i]);
                fprintf(stderr, "\n");
            }
        }
    }
#endif
    return check_device_errors(ctx);
}

int mdg_reset_tables(mdg_ctx *ctx)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    int rc = mdg_sync(ctx);
    if (rc) return rc;
    MDG_CUDA(ctx, cudaMemset(ctx->tables, 0, (ctx->n_mis + ctx->n_comp + ctx->n_lg) * 8));
    MDG_CUDA(ctx, cudaMemset(ctx->aux_block, 0, 256 + 64));
    ctx->reduced_valid = false;
    return MDG_OK;
}

int mdg_fetch_tables(mdg_ctx *ctx, uint64_t *misincorp, uint64_t *dnacomp, uint64_t *lghist)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    int rc = mdg_sync(ctx);
    if (rc) return rc;
    // after mdg_allreduce_tables (and until this rank counts again): the sums over all ranks
    const unsigned long long *from = ctx->reduced_valid ? ctx->reduced : ctx->tables;
    if (misincorp) MDG_CUDA(ctx, cudaMemcpy(misincorp, from, ctx->n_mis * 8, cudaMemcpyDeviceToHost));
    if (dnacomp) MDG_CUDA(ctx, cudaMemcpy(dnacomp, from + ctx->n_mis, ctx->n_comp * 8, cudaMemcpyDeviceToHost));
    if (lghist) MDG_CUDA(ctx, cudaMemcpy(lghist, from + ctx->n_mis + ctx->n_comp, ctx->n_lg * 8, cudaMemcpyDeviceToHost));
    return MDG_OK;
}

int64_t mdg_fetch_lg_overflow(mdg_ctx *ctx, int32_t *rows, int64_t max_rows)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    int rc = mdg_sync(ctx);
    if (rc) return rc;
    unsigned long long count = 0;
    if (cudaMemcpy(&count, ctx->count_tables.lg_overflow_count, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(ctx, MDG_ERR_CUDA, "overflow count copy failed");
    if ((int64_t)count > ctx->count_tables.lg_overflow_cap)
        return fail(ctx, MDG_ERR_CAPACITY, "more than %lld fragment lengths >= lg_bins; raise lg_bins",
                    (long long)ctx->count_tables.lg_overflow_cap);
    int64_t n = std::min<int64_t>((int64_t)count, max_rows);
    if (rows && n > 0 &&
        cudaMemcpy(rows, ctx->count_tables.lg_overflow_rows, (size_t)n * 16, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(ctx, MDG_ERR_CUDA, "overflow rows copy failed");
    return (int64_t)count;
}

int mdg_set_rescale_model(mdg_ctx *ctx, const uint8_t *lut, const double *inc, int32_t len5p, int32_t len3p)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    if (!lut || !inc || len5p < 0 || len3p < 0) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_set_rescale_model: bad argument");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(ctx->model_block);
    ctx->model_block = nullptr;
    const int n_slots = 1 + len5p + len3p;
    const size_t inc_bytes = align_up((size_t)2 * n_slots * 8), lut_bytes = align_up((size_t)2 * n_slots * 94);
    ctx->n_hist = (size_t)2 * n_slots * 94 + 2 * 94 + 4;
    MDG_CUDA(ctx, cudaMalloc(&ctx->model_block, inc_bytes + lut_bytes + ctx->n_hist * 8));
    char *p = (char *)ctx->model_block;
    ctx->rescale_hist = (unsigned long long *)(p + inc_bytes + lut_bytes);
    MDG_CUDA(ctx, cudaMemset(ctx->rescale_hist, 0, ctx->n_hist * 8));
    MDG_CUDA(ctx, cudaMemcpy(p, inc, (size_t)2 * n_slots * 8, cudaMemcpyHostToDevice));
    // the host table holds exactly 2 * n_slots * 94 bytes; the rounded-up tail of the device copy is zeroed
    MDG_CUDA(ctx, cudaMemset(p + inc_bytes, 0, lut_bytes));
    MDG_CUDA(ctx, cudaMemcpy(p + inc_bytes, lut, (size_t)2 * n_slots * 94, cudaMemcpyHostToDevice));
    ctx->model.inc = (const double *)p;
    ctx->model.lut = (const uint8_t *)(p + inc_bytes);
    ctx->model.len5p = len5p;
    ctx->model.len3p = len3p;
    ctx->model.n_slots = n_slots;
    return MDG_OK;
}

// The rescale kernels over one device batch whose qualities are rewritten in place; mr / status (device, one entry per
// read) and the optional change list are filled on `stream`.
static int launch_rescale(mdg_ctx *ctx, const mdg::DevBatch &view, float *d_mr, uint8_t *d_status, uint32_t *change_at,
                          uint8_t *change_q, unsigned long long *n_changes, int64_t change_cap, cudaStream_t stream)
{
    const int64_t n = view.n_reads;
    if (n == 0) return MDG_OK;
    const size_t n_sub = (size_t)2 * ctx->model.n_slots * 94;
    uint8_t *const dev_qual = const_cast<uint8_t *>(view.qual);
    mdg::RescaleOut out{dev_qual, d_mr, d_status, ctx->rescale_stats, ctx->count_tables.error_flag,
                        ctx->rescale_hist, ctx->rescale_hist + n_sub, ctx->rescale_hist + n_sub + 2 * 94,
                        change_at, change_q, n_changes, (unsigned long long)change_cap};
    WorkList *wl = nullptr;
    int rc = worklist_for(ctx, stream, n, &wl);
    if (rc) return rc;
    MDG_CUDA(ctx, cudaMemsetAsync(wl->count, 0, 8, stream));
    if (n_changes) MDG_CUDA(ctx, cudaMemsetAsync(n_changes, 0, 8, stream));
    cudaEvent_t e0, e1;
    rc = next_kernel_events(ctx, &e0, &e1);
    if (rc) return rc;
    MDG_CUDA(ctx, cudaEventRecord(e0, stream));
    const bool general_only = getenv("MDG_RESCALE_GENERAL") != nullptr;  // A/B and tests: every record by the warp kernel
    const int warp_grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 8, (n + 7) / 8);
    if (!general_only) {
        // block histograms in shared memory when the model's table fits (it does for any sensible --rescale-length)
        const size_t hist_words = n_sub + 2 * 94;
        const bool shared_hist = hist_words * 4 <= 48 * 1024;
        const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 8, (n + 255) / 256);
        mdg::rescale_gapfree_kernel<<<grid, 256, shared_hist ? hist_words * 4 : 0, stream>>>(
            view, ctx->ref, ctx->model, out, wl->reads, wl->count, shared_hist ? (int)hist_words : 0);
        MDG_CUDA(ctx, cudaGetLastError());
        mdg::rescale_kernel<<<warp_grid, 256, 0, stream>>>(view, ctx->ref, ctx->model, out, wl->reads, wl->count);
        ctx->launches += 2;
    } else {
        mdg::rescale_kernel<<<warp_grid, 256, 0, stream>>>(view, ctx->ref, ctx->model, out, nullptr, nullptr);
        ctx->launches += 1;
    }
    MDG_CUDA(ctx, cudaGetLastError());
    MDG_CUDA(ctx, cudaEventRecord(e1, stream));
    return MDG_OK;
}

static int rescale_submit(mdg_ctx *ctx, const mdg_batch *host, uint8_t *qual_out, float *mr_out, uint8_t *status_out, int32_t *ticket)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    int rc = check_batch(ctx, host);
    if (rc) return rc;
    const bool sparse = ticket != nullptr;
    if ((!sparse && !qual_out) || !mr_out || !status_out) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_rescale_submit: NULL output");
    if (!host->qual) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_rescale_submit: batch has no quality array");
    if (!ctx->model.lut) return fail(ctx, MDG_ERR_STATE, "mdg_set_rescale_model must be called before rescaling");
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "mdg_set_reference must be called before rescaling");
    if (ctx->slots.empty() || !ctx->slots[0].stream)
        return fail(ctx, MDG_ERR_STATE, "context was created without staging slots (max_reads = 0)");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int slot_index = ctx->next_slot;
    Slot &slot = ctx->slots[slot_index];
    ctx->next_slot = (ctx->next_slot + 1) % (int)ctx->slots.size();
    MDG_CUDA(ctx, cudaStreamSynchronize(slot.stream));
    if (sparse) {
        *ticket = slot_index;
        // a batch in which more than a quarter of all bases change is not sequencing data
        const int64_t want = host->n_bases / 4 + 4096;
        if (slot.change_cap < want) {
            cudaFree(slot.change_at);
            cudaFree(slot.change_q);
            slot.change_at = nullptr; slot.change_q = nullptr; slot.change_cap = 0;
            const int64_t cap = std::max<int64_t>(want, ctx->cfg.max_bases / 4 + 4096);
            MDG_CUDA(ctx, cudaMalloc(&slot.change_at, (size_t)cap * 4));
            MDG_CUDA(ctx, cudaMalloc(&slot.change_q, (size_t)cap));
            slot.change_cap = cap;
        }
        if (!slot.n_changes) {
            MDG_CUDA(ctx, cudaMalloc(&slot.n_changes, 8));
            MDG_CUDA(ctx, cudaHostAlloc((void **)&slot.n_changes_host, 8, cudaHostAllocDefault));
        }
        *slot.n_changes_host = 0;
    }
    rc = copy_batch(ctx, slot.arrays, host, slot.stream, true, true);
    if (rc) return rc;
    const int64_t n = host->n_reads;
    if (n == 0) return MDG_OK;
    // qualities are rewritten in place in the slot's copy of the batch
    rc = launch_rescale(ctx, slot.arrays.view, slot.mr_out, slot.status_out, sparse ? slot.change_at : nullptr,
                        sparse ? slot.change_q : nullptr, sparse ? slot.n_changes : nullptr, sparse ? slot.change_cap : 0, slot.stream);
    if (rc) return rc;
    if (sparse)
        MDG_CUDA(ctx, cudaMemcpyAsync(slot.n_changes_host, slot.n_changes, 8, cudaMemcpyDeviceToHost, slot.stream));
    else
        MDG_CUDA(ctx, cudaMemcpyAsync(qual_out, slot.arrays.view.qual, (size_t)host->n_bases, cudaMemcpyDeviceToHost, slot.stream));
    MDG_CUDA(ctx, cudaMemcpyAsync(mr_out, slot.mr_out, (size_t)n * 4, cudaMemcpyDeviceToHost, slot.stream));
    MDG_CUDA(ctx, cudaMemcpyAsync(status_out, slot.status_out, (size_t)n, cudaMemcpyDeviceToHost, slot.stream));
    return MDG_OK;
}

int mdg_rescale_submit(mdg_ctx *ctx, const mdg_batch *host, uint8_t *qual_out, float *mr_out, uint8_t *status_out)
{
    return rescale_submit(ctx, host, qual_out, mr_out, status_out, nullptr);
}

int mdg_rescale_submit_sparse(mdg_ctx *ctx, const mdg_batch *host, float *mr_out, uint8_t *status_out, int32_t *ticket)
{
    if (!ticket) return ctx ? fail(ctx, MDG_ERR_ARGUMENT, "mdg_rescale_submit_sparse: NULL ticket") : MDG_ERR_ARGUMENT;
    return rescale_submit(ctx, host, nullptr, mr_out, status_out, ticket);
}

int64_t mdg_rescale_collect(mdg_ctx *ctx, int32_t ticket, uint32_t *change_at, uint8_t *change_q, int64_t cap, uint8_t *patch_qual)
{
    if (!ctx || ticket < 0 || ticket >= (int32_t)ctx->slots.size()) return MDG_ERR_ARGUMENT;
    Slot &slot = ctx->slots[(size_t)ticket];
    if (!slot.n_changes_host) return fail(ctx, MDG_ERR_STATE, "mdg_rescale_collect: no sparse submit on this ticket");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaStreamSynchronize(slot.stream));
    const int64_t n = (int64_t)*slot.n_changes_host;
    if (n > slot.change_cap)
        return fail(ctx, MDG_ERR_CAPACITY, "%lld quality bytes changed, more than the change list holds (%lld): use mdg_rescale_submit",
                    (long long)n, (long long)slot.change_cap);
    if (n > cap || (n && (!change_at || !change_q)))
        return fail(ctx, MDG_ERR_CAPACITY, "mdg_rescale_collect: %lld changes do not fit the caller's arrays (%lld)", (long long)n, (long long)cap);
    if (n) {
        MDG_CUDA(ctx, cudaMemcpyAsync(change_at, slot.change_at, (size_t)n * 4, cudaMemcpyDeviceToHost, slot.stream));
        MDG_CUDA(ctx, cudaMemcpyAsync(change_q, slot.change_q, (size_t)n, cudaMemcpyDeviceToHost, slot.stream));
        MDG_CUDA(ctx, cudaStreamSynchronize(slot.stream));
        if (patch_qual) {
            // scattered single-byte writes over a few hundred MB: a few threads, each its share of the list
            const int n_parts = n >= (1 << 18) ? 8 : 1;
            const int64_t piece = (n + n_parts - 1) / n_parts;
            std::vector<std::thread> pool;
            for (int t = 1; t < n_parts; ++t)
                pool.emplace_back([=] {
                    for (int64_t k = t * piece; k < std::min(n, (t + 1) * piece); ++k) patch_qual[change_at[k]] = change_q[k];
                });
            for (int64_t k = 0; k < std::min(n, piece); ++k) patch_qual[change_at[k]] = change_q[k];
            for (auto &th : pool) th.join();
        }
    }
    return n;
}

int mdg_rescale_resident(mdg_ctx *ctx, mdg_dev_batch *batch, float *mr_out, uint8_t *status_out)
{
    if (!ctx || !batch) return MDG_ERR_ARGUMENT;
    if (!batch->arrays.has_qual || !batch->arrays.view.qual) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_rescale_resident: batch has no quality array");
    if (!ctx->model.lut) return fail(ctx, MDG_ERR_STATE, "mdg_set_rescale_model must be called before rescaling");
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "mdg_set_reference must be called before rescaling");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int64_t n = batch->arrays.view.n_reads;
    if (n > batch->res_cap) {
        MDG_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
        cudaFree(batch->res_mr);
        cudaFree(batch->res_status);
        batch->res_mr = nullptr; batch->res_status = nullptr; batch->res_cap = 0;
        MDG_CUDA(ctx, cudaMalloc(&batch->res_mr, (size_t)batch->arrays.cap_reads * 4 + 8));
        MDG_CUDA(ctx, cudaMalloc(&batch->res_status, (size_t)batch->arrays.cap_reads + 8));
        batch->res_cap = batch->arrays.cap_reads;
    }
    int rc = launch_rescale(ctx, batch->arrays.view, batch->res_mr, batch->res_status, nullptr, nullptr, nullptr, 0, ctx->compute);
    if (rc) return rc;
    if (mr_out) MDG_CUDA(ctx, cudaMemcpyAsync(mr_out, batch->res_mr, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->compute));
    if (status_out) MDG_CUDA(ctx, cudaMemcpyAsync(status_out, batch->res_status, (size_t)n, cudaMemcpyDeviceToHost, ctx->compute));
    return MDG_OK;
}

int mdg_fetch_rescale_stats(mdg_ctx *ctx, uint64_t *stats8)
{
    if (!ctx || !stats8) return MDG_ERR_ARGUMENT;
    int rc = mdg_sync(ctx);
    if (rc) return rc;
    MDG_CUDA(ctx, cudaMemcpy(stats8, ctx->rescale_stats, 64, cudaMemcpyDeviceToHost));
    return MDG_OK;
}

int mdg_fetch_rescale_hist(mdg_ctx *ctx, uint64_t *sub, uint64_t *rev, uint64_t *ref_count)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    if (!ctx->rescale_hist) return fail(ctx, MDG_ERR_STATE, "mdg_set_rescale_model must be called first");
    int rc = mdg_sync(ctx);
    if (rc) return rc;
    const size_t n_sub = (size_t)2 * ctx->model.n_slots * 94;
    if (sub) MDG_CUDA(ctx, cudaMemcpy(sub, ctx->rescale_hist, n_sub * 8, cudaMemcpyDeviceToHost));
    if (rev) MDG_CUDA(ctx, cudaMemcpy(rev, ctx->rescale_hist + n_sub, 2 * 94 * 8, cudaMemcpyDeviceToHost));
    if (ref_count) MDG_CUDA(ctx, cudaMemcpy(ref_count, ctx->rescale_hist + n_sub + 2 * 94, 4 * 8, cudaMemcpyDeviceToHost));
    return MDG_OK;
}

int mdg_synth_batch(mdg_ctx *ctx, const mdg_synth_params *sp, mdg_dev_batch **out)
{
    if (!ctx || !sp || !out) return MDG_ERR_ARGUMENT;
    *out = nullptr;
    if (!ctx->ref.words) return fail(ctx, MDG_ERR_STATE, "mdg_set_reference must be called before mdg_synth_batch");
    const int64_t mix_total = (int64_t)sp->mix[0] + sp->mix[1] + sp->mix[2] + sp->mix[3];
    if (sp->n_reads < 1 || sp->n_reads >= (1ll << 31) || sp->len_lo < 1 || sp->len_hi < sp->len_lo ||
        sp->len_hi > 60000 || mix_total < 1 || mix_total > 65535 || sp->mix[0] < 0 || sp->mix[1] < 0 ||
        sp->mix[2] < 0 || sp->mix[3] < 0 || sp->n_libraries < 1 || (sp->paired && (sp->n_reads & 1)))
        return fail(ctx, MDG_ERR_ARGUMENT, "mdg_synth_batch: invalid parameters");
    if ((int64_t)ctx->ref_min_contig < (int64_t)sp->len_hi + 16)
        return fail(ctx, MDG_ERR_ARGUMENT, "mdg_synth_batch: a contig is shorter than a read (%u < %d + 16)",
                    ctx->ref_min_contig, sp->len_hi);
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    mdg::SynthDev p{};
    p.seed = sp->seed;
    p.n_reads = sp->n_reads;
    p.len_lo = sp->len_lo;
    p.len_hi = sp->len_hi;
    uint32_t cum = 0;
    for (int k = 0; k < 4; ++k) p.mix_cum[k] = cum += (uint32_t)sp->mix[k];
    p.paired = sp->paired;
    p.with_qual = sp->with_qual;
    p.n_lib = sp->n_libraries;
    auto u24 = [](float x) { return (uint32_t)(std::min(std::max(x, 0.f), 1.f) * 16777216.f); };
    p.error_u24 = u24(sp->error_rate);
    p.read_n_u24 = u24(sp->read_n_rate);
    p.filtered_u24 = u24(sp->filtered_rate);
    p.damage0 = sp->damage0;
    p.decay = sp->damage_decay;
    p.genome_bases = ctx->ref_total_bases;
    p.sorted = sp->reserved != 0;

    const int64_t n_blocks = (sp->n_reads + 255) / 256;
    unsigned long long *totals = nullptr;
    MDG_CUDA(ctx, cudaMalloc(&totals, (size_t)(n_blocks + 1) * 16));
    mdg::synth_block_totals<<<(unsigned)n_blocks, 256, 0, ctx->compute>>>(p, totals);
    mdg::synth_scan_totals<<<1, 1024, 0, ctx->compute>>>(totals, n_blocks);
    unsigned long long grand[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(grand, totals + 2 * n_blocks, 16, cudaMemcpyDeviceToHost, ctx->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->compute);
    if (e != cudaSuccess) {
        cudaFree(totals);
        return fail(ctx, MDG_ERR_CUDA, "mdg_synth_batch: sizing pass failed: %s", cudaGetErrorString(e));
    }
    if (grand[0] >= (1ull << 32) || grand[1] >= (1ull << 32)) {
        cudaFree(totals);
        return fail(ctx, MDG_ERR_ARGUMENT, "mdg_synth_batch: %llu bases exceed the 32-bit offsets of one batch; "
                    "generate fewer reads per batch", grand[0]);
    }
    mdg_dev_batch *d = new (std::nothrow) mdg_dev_batch();
    if (!d) {
        cudaFree(totals);
        return fail(ctx, MDG_ERR_ARGUMENT, "out of host memory");
    }
    int rc = alloc_arrays(ctx, d->arrays, sp->n_reads, (int64_t)grand[1], (int64_t)grand[0], sp->with_qual != 0);
    if (rc) {
        cudaFree(totals);
        delete d;
        return rc;
    }
    d->arrays.view.n_reads = sp->n_reads;
    d->arrays.n_cigar = (int64_t)grand[1];
    d->arrays.n_bases = (int64_t)grand[0];
    const mdg::DevBatch &v = d->arrays.view;
    mdg::SynthOut o{(uint16_t *)v.flag, (int32_t *)v.tid, (int32_t *)v.pos, (uint16_t *)v.lib, (uint32_t *)v.l_seq,
                    (uint32_t *)v.base_off, (uint32_t *)v.cigar_off, (uint32_t *)v.cigar, (uint8_t *)v.seq4,
                    (uint8_t *)v.qual, (int32_t *)v.tlen, (int32_t *)v.mtid, (int32_t *)v.mpos};
    mdg::synth_fill<<<(unsigned)n_blocks, 256, 0, ctx->compute>>>(p, ctx->ref, totals, o);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->compute);
    cudaFree(totals);
    if (e != cudaSuccess) {
        cudaFree(d->arrays.block);
        delete d;
        return fail(ctx, MDG_ERR_CUDA, "mdg_synth_batch: fill failed: %s", cudaGetErrorString(e));
    }
    *out = d;
    return MDG_OK;
}

int mdg_batch_sizes(mdg_ctx *ctx, const mdg_dev_batch *batch, int64_t *n_reads, int64_t *n_cigar, int64_t *n_bases)
{
    if (!ctx || !batch) return MDG_ERR_ARGUMENT;
    if (n_reads) *n_reads = batch->arrays.view.n_reads;
    if (n_cigar) *n_cigar = batch->arrays.n_cigar;
    if (n_bases) *n_bases = batch->arrays.n_bases;
    return MDG_OK;
}

int mdg_batch_download(mdg_ctx *ctx, const mdg_dev_batch *batch, const mdg_batch *h)
{
    if (!ctx || !batch) return MDG_ERR_ARGUMENT;
    int rc = check_batch(ctx, h);
    if (rc) return rc;
    const DeviceArrays &a = batch->arrays;
    const int64_t n = a.view.n_reads;
    if (h->n_reads != n || h->n_cigar != a.n_cigar || h->n_bases != a.n_bases)
        return fail(ctx, MDG_ERR_ARGUMENT, "mdg_batch_download: host arrays must be sized by mdg_batch_sizes");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
#define MDG_D2H(field, bytes) \
    if (h->field) MDG_CUDA(ctx, cudaMemcpy((void *)h->field, a.view.field, (size_t)(bytes), cudaMemcpyDeviceToHost))
    MDG_D2H(flag, n * 2);
    MDG_D2H(tid, n * 4);
    MDG_D2H(pos, n * 4);
    MDG_D2H(lib, n * 2);
    MDG_D2H(l_seq, n * 4);
    MDG_D2H(base_off, n * 4);
    MDG_D2H(cigar_off, (n + 1) * 4);
    MDG_D2H(cigar, a.n_cigar * 4);
    MDG_D2H(seq4, a.n_bases / 2);
    MDG_D2H(tlen, n * 4);
    MDG_D2H(mtid, n * 4);
    MDG_D2H(mpos, n * 4);
    if (h->qual) {
        if (!a.has_qual) return fail(ctx, MDG_ERR_ARGUMENT, "mdg_batch_download: the resident batch has no qualities");
        MDG_D2H(qual, a.n_bases);
    }
#undef MDG_D2H
    return MDG_OK;
}

int mdg_nccl_unique_id(void *id128)
{
    if (!id128) return MDG_ERR_ARGUMENT;
    NcclApi &api = nccl_api();
    if (!api.error.empty()) return fail(nullptr, MDG_ERR_NCCL, "%s", api.error.c_str());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, MDG_ERR_NCCL, "ncclGetUniqueId: %s", api.GetErrorString ? api.GetErrorString(r) : "?");
    memcpy(id128, &id, 128);
    return MDG_OK;
}

int mdg_nccl_init(mdg_ctx *ctx, const void *id128, int32_t rank, int32_t n_ranks)
{
    if (!ctx || !id128) return MDG_ERR_ARGUMENT;
    NcclApi &api = nccl_api();
    if (!api.error.empty()) return fail(ctx, MDG_ERR_NCCL, "%s", api.error.c_str());
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = api.CommInitRank(&ctx->comm, n_ranks, id, rank);
    if (r != ncclSuccess) return fail(ctx, MDG_ERR_NCCL, "ncclCommInitRank: %s", api.GetErrorString ? api.GetErrorString(r) : "?");
    return MDG_OK;
}

int mdg_allreduce_tables(mdg_ctx *ctx)
{
    if (!ctx) return MDG_ERR_ARGUMENT;
    if (!ctx->comm) return fail(ctx, MDG_ERR_STATE, "mdg_nccl_init must be called before mdg_allreduce_tables");
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    for (auto &slot : ctx->slots)
        if (slot.stream) MDG_CUDA(ctx, cudaStreamSynchronize(slot.stream));
    NcclApi &api = nccl_api();
    const size_t n = ctx->n_mis + ctx->n_comp + ctx->n_lg;
    // out of place: the accumulators keep this rank's own counts, so counting on and reducing again (or reducing
    // twice) gives the sum over ranks each time instead of re-adding what an earlier call had already summed
    if (!ctx->reduced) MDG_CUDA(ctx, cudaMalloc(&ctx->reduced, n * 8));
    ncclResult_t r = api.AllReduce(ctx->tables, ctx->reduced, n, ncclUint64, ncclSum, ctx->comm, ctx->compute);
    if (r != ncclSuccess) return fail(ctx, MDG_ERR_NCCL, "ncclAllReduce: %s", api.GetErrorString ? api.GetErrorString(r) : "?");
    ctx->reduced_valid = true;
    ctx->launches += 1;
    return MDG_OK;
}

int mdg_event_record(mdg_ctx *ctx, int32_t which)
{
    if (!ctx || which < 0 || which > 1) return MDG_ERR_ARGUMENT;
    MDG_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MDG_CUDA(ctx, cudaEventRecord(ctx->ev[which], ctx->compute));
    return MDG_OK;
}

int mdg_event_elapsed_ms(mdg_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return MDG_ERR_ARGUMENT;
    MDG_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    MDG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
    return MDG_OK;
}

int64_t mdg_launch_count(const mdg_ctx *ctx) { return ctx ? ctx->launches : 0; }

int mdg_last_kernel_ms(mdg_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return MDG_ERR_ARGUMENT;
    int rc = mdg_sync(ctx);
    if (rc) return rc;
    float total = 0.f;
    for (size_t i = 0; i + 1 < ctx->kernel_events_used; i += 2) {
        float t = 0.f;
        MDG_CUDA(ctx, cudaEventElapsedTime(&t, ctx->kernel_events[i], ctx->kernel_events[i + 1]));
        total += t;
    }
    ctx->kernel_events_used = 0;
    *ms = total;
    return MDG_OK;
}

}  // extern "C"

#include <chrono>
#include "mdg_bamdev.cuh"
