// BGZF blocks inflated on the GPU: one thread per block, thousands of blocks per launch.
//
// A DEFLATE stream is a chain of dependent table lookups, hopeless for one thread -- but a BAM file is tens of
// thousands of independent 64 KB blocks, and a launch that decodes them all side by side takes about as long as one
// block takes one thread.  The decoder is the host's (mdg_inflate_core.h, compiled for the device); its tables
// live in a slice of global memory per thread.  The caller checks every block's CRC32 on the host and re-does a
// block that failed here on the CPU.
#pragma once
#include <cuda_runtime.h>

#include "mdg_inflate_core.h"

namespace mdg {

// The host's decode loop lets every stream run its own branches; 32 streams in a warp would then take turns (measured:
// 1.2 lanes active per instruction).  Here every lane makes one small step per iteration of one common loop -- read a
// block header, copy up to four bytes of a pending match, or decode one symbol -- so the lanes stay together.
__device__ inline int64_t inflate_lockstep(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap,
                                           mdg_inflate::InflateScratch &s, bool live)
{
    using namespace mdg_inflate;
    BitReader br;
    br.in = in;
    br.in_end = in + (live ? in_len : 0);
    uint8_t *at = out, *const out_end = out + out_cap;
    bool done = !live, failed = false, need_header = true, final_block = false;
    uint32_t pending = 0;
    const uint8_t *src = nullptr;
    while (__any_sync(0xffffffffu, !done)) {
        // block headers are read by all lanes at once: building the decode tables is thousands of instructions, which
        // 32 lanes arriving one by one would take turns at.  Streams of one file change tables after about as many
        // symbols, so a lane seldom waits long for the others.
        const bool headers_now = __all_sync(0xffffffffu, done || need_header);
        if (done || (need_header && !headers_now)) continue;
        if (need_header) {
            br.refill();
            final_block = br.take(1) != 0;
            const uint32_t type = br.take(2);
            if (br.overrun || type == 3) {
                failed = done = true;
            } else if (type == 0) {
                // stored: back to a byte boundary, bytes the bit buffer holds but has not used go back to the input
                br.consume(br.cnt & 7);
                br.in -= br.cnt >> 3;
                br.bits = 0;
                br.cnt = 0;
                bool ok = br.in_end - br.in >= 4;
                uint32_t len = 0;
                if (ok) {
                    len = br.in[0] | (uint32_t)br.in[1] << 8;
                    const uint32_t nlen = br.in[2] | (uint32_t)br.in[3] << 8;
                    br.in += 4;
                    ok = (len ^ nlen) == 0xFFFFu && (int64_t)len <= br.in_end - br.in && (int64_t)len <= out_end - at;
                }
                if (!ok) {
                    failed = done = true;
                } else {
                    for (uint32_t i = 0; i < len; ++i) at[i] = br.in[i];
                    at += len;
                    br.in += len;
                    done = final_block;
                }
            } else {
                if (!(type == 1 ? fixed_tables(s.t, s) : read_dynamic_header(br, s.t, s))) failed = done = true;
                need_header = false;
            }
        } else if (pending) {
            // bytes a match copies may be ones it wrote itself: at most `distance` of them per step
            const uint32_t distance = (uint32_t)(at - src);
            const uint32_t k = min(min(pending, 4u), distance);
            at[0] = src[0];
            if (k > 1) at[1] = src[1];
            if (k > 2) at[2] = src[2];
            if (k > 3) at[3] = src[3];
            at += k;
            src += k;
            pending -= k;
        } else {
            br.refill();
            const uint32_t e = lookup(s.t.litlen, LITLEN_BITS, br.bits);
            const uint32_t kind = (e >> 5) & 7;
            br.consume((int)(e & 31));
            if (kind == LITERAL) {
                if (br.overrun || at >= out_end) failed = done = true;
                else *at++ = (uint8_t)(e >> 16);
            } else if (kind == END) {
                if (br.overrun) failed = true;
                need_header = true;
                done = final_block || failed;
            } else if (kind == BASE) {
                // one refill leaves 56 bits: a length code with its extra bits and a distance code with its own take 48
                const uint32_t length = (e >> 16) + br.take((int)((e >> 8) & 31));
                const uint32_t d = lookup(s.t.offset, OFFSET_BITS, br.bits);
                br.consume((int)(d & 31));
                const uint32_t offset = (d >> 16) + br.take((int)((d >> 8) & 31));
                if (((d >> 5) & 7) != BASE || br.overrun || offset > (uint32_t)(at - out) || length > (uint32_t)(out_end - at)) {
                    failed = done = true;
                } else {
                    pending = length;
                    src = at - offset;
                }
            } else {
                failed = done = true;
            }
        }
    }
    return failed ? -1 : at - out;
}

__global__ void __launch_bounds__(32) inflate_blocks_kernel(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                                                            const uint32_t *__restrict__ in_len, uint8_t *__restrict__ out,
                                                            const uint64_t *__restrict__ out_off,
                                                            const uint32_t *__restrict__ isize, int32_t *__restrict__ status,
                                                            mdg_inflate::InflateScratch *__restrict__ scratch, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const int at = live ? i : 0;
    const int64_t got = inflate_lockstep(in + in_off[at], (int64_t)in_len[at], out + out_off[at], (int64_t)isize[at],
                                         scratch[at], live);
    if (live) status[i] = got == (int64_t)isize[i] ? 0 : 1;
}

}  // namespace mdg

// Device buffers and a stream of its own: used from the BAM reader's producer thread, next to (not through) an
// mdg_ctx.
struct mdg_inflater {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    uint64_t *d_meta = nullptr;  // in_off[n] | out_off[n] | in_len[n], isize[n], status[n] (32-bit)
    mdg_inflate::InflateScratch *d_scratch = nullptr;
    int n_cap = 0;
    uint8_t *h_meta = nullptr;  // pinned mirror of d_meta
    std::string error;
};

extern "C" {

int mdg_inflater_create(int32_t device, mdg_inflater **out)
{
    if (!out) return MDG_ERR_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return MDG_ERR_NO_DEVICE;
    mdg_inflater *f = new (std::nothrow) mdg_inflater;
    if (!f) return MDG_ERR_ARGUMENT;
    f->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete f;
        return MDG_ERR_CUDA;
    }
    *out = f;
    return MDG_OK;
}

void mdg_inflater_free(mdg_inflater *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    cudaFree(f->d_in);
    cudaFree(f->d_out);
    cudaFree(f->d_meta);
    cudaFree(f->d_scratch);
    cudaFreeHost(f->h_meta);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

const char *mdg_inflater_error(const mdg_inflater *f) { return f ? f->error.c_str() : "no inflater"; }

int mdg_inflate_blocks(mdg_inflater *f, const uint8_t *in, int64_t in_bytes, const uint64_t *in_off, const uint32_t *in_len,
                       uint8_t *out, int64_t out_bytes, const uint64_t *out_off, const uint32_t *isize, int32_t n,
                       int32_t *status)
{
    if (!f || n < 0 || in_bytes < 0 || out_bytes < 0) return MDG_ERR_ARGUMENT;
    if (n == 0) return MDG_OK;
    if (!in || !in_off || !in_len || !out || !out_off || !isize || !status) return MDG_ERR_ARGUMENT;
    for (int32_t i = 0; i < n; ++i)
        if (in_off[i] + in_len[i] > (uint64_t)in_bytes || out_off[i] + isize[i] > (uint64_t)out_bytes) return MDG_ERR_ARGUMENT;
#define MDG_INF_CUDA(call)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            f->error = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            return MDG_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)
    MDG_INF_CUDA(cudaSetDevice(f->device));
    if ((size_t)in_bytes + 16 > f->in_cap) {
        cudaFree(f->d_in);
        f->d_in = nullptr;
        f->in_cap = 0;
        MDG_INF_CUDA(cudaMalloc(&f->d_in, (size_t)in_bytes * 5 / 4 + 16));
        f->in_cap = (size_t)in_bytes * 5 / 4 + 16;
    }
    if ((size_t)out_bytes + 16 > f->out_cap) {
        cudaFree(f->d_out);
        f->d_out = nullptr;
        f->out_cap = 0;
        MDG_INF_CUDA(cudaMalloc(&f->d_out, (size_t)out_bytes * 5 / 4 + 16));
        f->out_cap = (size_t)out_bytes * 5 / 4 + 16;
    }
    if (n > f->n_cap) {
        cudaFree(f->d_meta);
        cudaFree(f->d_scratch);
        cudaFreeHost(f->h_meta);
        f->d_meta = nullptr;
        f->d_scratch = nullptr;
        f->h_meta = nullptr;
        f->n_cap = 0;
        const int cap = n * 5 / 4 + 64;
        MDG_INF_CUDA(cudaMalloc(&f->d_meta, (size_t)cap * 28));
        MDG_INF_CUDA(cudaMalloc(&f->d_scratch, (size_t)cap * sizeof(mdg_inflate::InflateScratch)));
        MDG_INF_CUDA(cudaHostAlloc(&f->h_meta, (size_t)cap * 28, cudaHostAllocDefault));
        f->n_cap = cap;
    }
    // metadata in one pinned block: in_off | out_off | in_len | isize | status
    const size_t cap = (size_t)f->n_cap;
    uint64_t *h_in_off = (uint64_t *)f->h_meta, *h_out_off = h_in_off + cap;
    uint32_t *h_in_len = (uint32_t *)(h_out_off + cap), *h_isize = h_in_len + cap;
    int32_t *h_status = (int32_t *)(h_isize + cap);
    memcpy(h_in_off, in_off, (size_t)n * 8);
    memcpy(h_out_off, out_off, (size_t)n * 8);
    memcpy(h_in_len, in_len, (size_t)n * 4);
    memcpy(h_isize, isize, (size_t)n * 4);
    uint64_t *d_in_off = f->d_meta, *d_out_off = d_in_off + cap;
    uint32_t *d_in_len = (uint32_t *)(d_out_off + cap), *d_isize = d_in_len + cap;
    int32_t *d_status = (int32_t *)(d_isize + cap);
    MDG_INF_CUDA(cudaMemcpyAsync(f->d_meta, f->h_meta, cap * 24, cudaMemcpyHostToDevice, f->stream));
    MDG_INF_CUDA(cudaMemcpyAsync(f->d_in, in, (size_t)in_bytes, cudaMemcpyHostToDevice, f->stream));
    mdg::inflate_blocks_kernel<<<(n + 31) / 32, 32, 0, f->stream>>>(f->d_in, d_in_off, d_in_len, f->d_out, d_out_off, d_isize,
                                                                   d_status, f->d_scratch, n);
    MDG_INF_CUDA(cudaGetLastError());
    MDG_INF_CUDA(cudaMemcpyAsync(out, f->d_out, (size_t)out_bytes, cudaMemcpyDeviceToHost, f->stream));
    MDG_INF_CUDA(cudaMemcpyAsync(h_status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, f->stream));
    MDG_INF_CUDA(cudaStreamSynchronize(f->stream));
    memcpy(status, h_status, (size_t)n * 4);
#undef MDG_INF_CUDA
    return MDG_OK;
}

}  // extern "C"
