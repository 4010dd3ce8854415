// BAM files decoded on the GPU: the device twin of mdg_bamio.cpp's reader.
//
// The reference iterates a pysam.AlignmentFile (reader.py:38,121-132; rescale.py:298-300): htslib inflates the BGZF
// blocks and hands out one record at a time.  Here a whole slab of the file (hundreds of MB, thousands of BGZF blocks)
// is handled per step, and the host only reads the file:
//
//   host   pread the slab into page-locked memory, find the BGZF block boundaries (headers only)
//   GPU    inflate every block (inflate_blocks_kernel, one thread per block), CRC32 of every block's data
//   GPU    find the record boundaries: the records are a chain of length prefixes, so each 32 KB segment of the inflated
//          stream guesses where its first record starts (a run of plausible record headers), walks its records from
//          there, and one warp then checks the guesses against the true chain (segment g must start where segment
//          g - 1 ended) and re-walks any segment that guessed wrong -- the result is exact, the guess only buys speed
//   GPU    scatter the records into the struct-of-arrays batch the counting / rescale kernels read (BAM's CIGAR words
//          and 4-bit SEQ are the batch layout already), read group -> library, "has an MR tag"
//
// The incomplete record at the end of a slab's stream is carried to the front of the next slab's.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace mdg {

constexpr uint32_t BAM_SEGMENT = 32768;         // bytes of inflated stream one walking thread covers
constexpr uint64_t BAM_NO_START = ~0ull;         // "no record starts in this segment"
constexpr int BAM_GUESS_CHAIN = 3;               // plausible record headers in a row that make a guess

__device__ __forceinline__ uint32_t bam_ld16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
__device__ __forceinline__ uint32_t bam_ld32(const uint8_t *p)
{
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

// Does a BAM record plausibly start at stream[o]?  (SAM specification 4.2.)  Only ever a hint: see bam_verify_kernel.
__device__ inline bool bam_plausible(const uint8_t *s, uint64_t len, uint64_t o, int32_t n_ref, uint32_t *size)
{
    if (o + 36 > len) return false;
    const uint8_t *p = s + o;
    const uint32_t bs = bam_ld32(p);
    if (bs < 32 || bs > (1u << 28)) return false;
    const int32_t tid = (int32_t)bam_ld32(p + 4), pos = (int32_t)bam_ld32(p + 8);
    const uint32_t l_name = p[12], n_cig = bam_ld16(p + 16), l_seq = bam_ld32(p + 20);
    const int32_t mtid = (int32_t)bam_ld32(p + 24), mpos = (int32_t)bam_ld32(p + 28);
    if (tid < -1 || tid >= n_ref || mtid < -1 || mtid >= n_ref || pos < -1 || mpos < -1 || l_name < 1) return false;
    if (l_seq > (1u << 28) || 32ull + l_name + 4ull * n_cig + (l_seq + 1) / 2 + (uint64_t)l_seq > bs) return false;
    if (o + 36 + l_name <= len) {
        // the name is printable and NUL-terminated
        if (p[36 + l_name - 1] != 0) return false;
        if (l_name > 1 && (p[36] < 33 || p[36] > 126)) return false;
    }
    *size = bs;
    return true;
}

struct BamWalk {      // what a walk over one segment yields
    uint64_t next;    // where the walk stopped: the first record start at or beyond the segment's end, or the start of
                      // the record the stream ends inside
    uint32_t seen, kept;
    uint32_t stopped; // 1: the stream ends inside the record at `next`; 2: a malformed record at `next`
};

__device__ inline BamWalk bam_walk(const uint8_t *s, uint64_t len, uint64_t from, uint64_t seg_hi, uint32_t drop_flags)
{
    BamWalk w{from, 0, 0, 0};
    uint64_t o = from;
    while (o < seg_hi) {
        if (o + 4 > len) { w.stopped = 1; break; }
        const uint32_t bs = bam_ld32(s + o);
        if (bs < 32) { w.stopped = 2; break; }
        if (o + 4 + (uint64_t)bs > len) { w.stopped = 1; break; }
        const uint32_t flag = bam_ld16(s + o + 18);
        ++w.seen;
        if (!(flag & drop_flags)) ++w.kept;
        o += 4 + (uint64_t)bs;
    }
    w.next = o;
    return w;
}

struct BamSegments {
    uint64_t *start;   // [n_seg] first record start in the segment (guess, then the verified value), or BAM_NO_START
    uint64_t *next;    // [n_seg] where the segment's walk stopped
    uint32_t *seen;    // [n_seg]
    uint32_t *kept;    // [n_seg] records kept (flag & drop_flags == 0); after the verify pass: exclusive prefix sum
    uint32_t *stopped; // [n_seg]
};

struct BamTotals {     // written by bam_verify_kernel
    unsigned long long n_seen, n_kept, tail;  // tail: offset of the incomplete record at the end (= len when there is none)
    unsigned int guesses_wrong, malformed;
    unsigned long long bases, cigars;         // filled by the layout scan: padded base slots and CIGAR words of the kept records
    unsigned int first_bad_library;           // index of the first kept record without a usable read group, or 0xFFFFFFFF
    unsigned int crc_failures;
};

// pass 1: one thread per segment guesses its first record start and walks
__global__ void __launch_bounds__(128) bam_guess_walk_kernel(const uint8_t *__restrict__ s, uint64_t len, uint64_t start0, int32_t n_ref,
                                                             uint32_t drop_flags, BamSegments seg, int64_t n_seg)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_seg) return;
    const uint64_t lo = (uint64_t)g * BAM_SEGMENT, hi = min(len, lo + BAM_SEGMENT);
    uint64_t first = BAM_NO_START;
    if (start0 >= hi) {
        first = BAM_NO_START;  // still inside the header
    } else if (start0 >= lo) {
        first = start0;
    } else {
        for (uint64_t o = lo; o < hi; ++o) {
            uint64_t at = o;
            bool ok = true;
            for (int k = 0; k < BAM_GUESS_CHAIN && ok; ++k) {
                uint32_t size = 0;
                if (at + 36 > len) break;  // nothing left to look at: plausible as far as it goes
                ok = bam_plausible(s, len, at, n_ref, &size);
                at += 4 + (uint64_t)size;
            }
            if (ok && o + 36 <= len) {
                first = o;
                break;
            }
        }
    }
    seg.start[g] = first;
    BamWalk w{0, 0, 0, 0};
    if (first != BAM_NO_START) w = bam_walk(s, len, first, hi, drop_flags);
    seg.next[g] = w.next;
    seg.seen[g] = w.seen;
    seg.kept[g] = w.kept;
    seg.stopped[g] = w.stopped;
}

// pass 2: one warp follows the true chain from start0 through the segments.  A segment whose guess differs from where
// the chain arrives is walked again from there (every lane does the same walk: no divergence, and it is rare).  Leaves
// start[] = verified starts, kept[] = exclusive prefix sums, and the totals.
__global__ void __launch_bounds__(32) bam_verify_kernel(const uint8_t *__restrict__ s, uint64_t len, uint64_t start0, uint32_t drop_flags,
                                                        BamSegments seg, int64_t n_seg, BamTotals *totals)
{
    const int lane = threadIdx.x;
    uint64_t expected = start0;
    unsigned long long n_seen = 0, n_kept = 0;
    unsigned int wrong = 0, malformed = 0;
    bool done = false;  // the chain has reached the incomplete record at the end of the stream
    for (int64_t base = 0; base < n_seg; base += 32) {
        const int64_t mine = base + lane;
        uint64_t g_start = BAM_NO_START, g_next = 0;
        uint32_t g_seen = 0, g_kept = 0, g_stopped = 0;
        if (mine < n_seg) {
            g_start = seg.start[mine];
            g_next = seg.next[mine];
            g_seen = seg.seen[mine];
            g_kept = seg.kept[mine];
            g_stopped = seg.stopped[mine];
        }
        uint64_t out_start = BAM_NO_START;
        uint32_t out_off = 0;
        const int n_here = n_seg - base < 32 ? (int)(n_seg - base) : 32;
        for (int i = 0; i < n_here; ++i) {
            const uint64_t hi = min(len, (uint64_t)(base + i + 1) * BAM_SEGMENT);
            uint64_t st = __shfl_sync(0xffffffffu, g_start, i), nx = __shfl_sync(0xffffffffu, g_next, i);
            uint32_t sn = __shfl_sync(0xffffffffu, g_seen, i), kp = __shfl_sync(0xffffffffu, g_kept, i);
            uint32_t sp = __shfl_sync(0xffffffffu, g_stopped, i);
            uint64_t final_start = BAM_NO_START;
            uint32_t final_kept = 0;
            if (!done && expected < hi) {
                if (st != expected) {
                    const BamWalk w = bam_walk(s, len, expected, hi, drop_flags);
                    nx = w.next; sn = w.seen; kp = w.kept; sp = w.stopped;
                    if (base + i > 0 || st != BAM_NO_START) ++wrong;
                }
                final_start = expected;
                final_kept = kp;
                expected = nx;
                if (sp) {
                    done = true;
                    if (sp == 2) ++malformed;
                }
            }
            if (lane == i) {
                out_start = final_start;
                out_off = (uint32_t)n_kept;
            }
            n_seen += final_start != BAM_NO_START ? sn : 0;
            n_kept += final_kept;
        }
        if (mine < n_seg) {
            seg.start[mine] = out_start;
            seg.kept[mine] = out_off;
        }
    }
    if (lane == 0) {
        totals->n_seen = n_seen;
        totals->n_kept = n_kept;
        totals->tail = done ? expected : min(expected, len);
        if (!done && expected > len) totals->tail = len;  // cannot happen: a record that overruns the stream stops the walk
        totals->guesses_wrong = wrong;
        totals->malformed = malformed;
        totals->first_bad_library = 0xFFFFFFFFu;
    }
}

struct BamFields {     // destination arrays of the kept records (a DeviceArrays view, written here)
    uint64_t *rec_off;
    uint16_t *flag;
    int32_t *tid, *pos, *tlen, *mtid, *mpos;
    uint32_t *l_seq, *n_cigar;  // n_cigar lands in cigar_off[] and is scanned in place
};

// pass 3: every segment walks again from its verified start and writes the fixed fields of its kept records
__global__ void __launch_bounds__(128) bam_fields_kernel(const uint8_t *__restrict__ s, uint64_t len, uint32_t drop_flags, BamSegments seg,
                                                         int64_t n_seg, BamFields f)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_seg) return;
    uint64_t o = seg.start[g];
    if (o == BAM_NO_START) return;
    const uint64_t hi = min(len, (uint64_t)(g + 1) * BAM_SEGMENT);
    uint64_t at = seg.kept[g];
    while (o < hi) {
        if (o + 4 > len) break;
        const uint32_t bs = bam_ld32(s + o);
        if (bs < 32 || o + 4 + (uint64_t)bs > len) break;
        const uint8_t *p = s + o + 4;
        const uint32_t flag = bam_ld16(p + 14);
        if (!(flag & drop_flags)) {
            f.rec_off[at] = o;
            f.flag[at] = (uint16_t)flag;
            f.tid[at] = (int32_t)bam_ld32(p);
            f.pos[at] = (int32_t)bam_ld32(p + 4);
            f.n_cigar[at] = bam_ld16(p + 12);
            f.l_seq[at] = bam_ld32(p + 16);
            f.mtid[at] = (int32_t)bam_ld32(p + 20);
            f.mpos[at] = (int32_t)bam_ld32(p + 24);
            f.tlen[at] = (int32_t)bam_ld32(p + 28);
            ++at;
        }
        o += 4 + (uint64_t)bs;
    }
}

// layout of the batch: base_off[i] = sum of l_seq rounded up to even, cigar_off[i] = sum of n_cigar (in place)
__global__ void __launch_bounds__(256) bam_layout_totals(const uint32_t *__restrict__ l_seq, const uint32_t *__restrict__ n_cigar, int64_t n,
                                                         unsigned long long *totals)
{
    __shared__ uint2 total;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t bases = i < n ? (l_seq[i] + 1) & ~1u : 0, ops = i < n ? n_cigar[i] : 0;
    block_exclusive_scan2(bases, ops, &total);
    if (threadIdx.x == 0) {
        totals[2 * (size_t)blockIdx.x] = total.x;
        totals[2 * (size_t)blockIdx.x + 1] = total.y;
    }
}

__global__ void __launch_bounds__(256) bam_layout_fill(const uint32_t *__restrict__ l_seq, uint32_t *cigar_off, int64_t n,
                                                       const unsigned long long *__restrict__ block_off, int64_t n_blocks, uint32_t *base_off,
                                                       BamTotals *totals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t len = i < n ? l_seq[i] : 0, ops = i < n ? cigar_off[i] : 0;
    const uint2 off = block_exclusive_scan2((len + 1) & ~1u, ops, nullptr);
    if (i < n) {
        base_off[i] = (uint32_t)(block_off[2 * (size_t)blockIdx.x] + off.x);
        cigar_off[i] = (uint32_t)(block_off[2 * (size_t)blockIdx.x + 1] + off.y);
    }
    if (i == 0) {
        totals->bases = block_off[2 * n_blocks];
        totals->cigars = block_off[2 * n_blocks + 1];
        cigar_off[n] = (uint32_t)block_off[2 * n_blocks + 1];
    }
}

struct BamLibraries {   // read group -> library (reader.py:63-81); n == 0: every read is in library 0 (--merge-libraries)
    const char *ids;    // NUL-terminated ids, back to back
    const uint16_t *library;
    int32_t n;
};

// the value of tag `a``b` among the auxiliary fields [aux, end): offset of its type byte, or -1
__device__ inline int64_t bam_find_tag(const uint8_t *aux, const uint8_t *end, uint8_t a, uint8_t b)
{
    const uint8_t *p = aux;
    while (p + 3 <= end) {
        if (p[0] == a && p[1] == b) return p + 2 - aux;
        const uint8_t type = p[2];
        p += 3;
        uint64_t skip = 0;
        switch (type) {
        case 'A': case 'c': case 'C': skip = 1; break;
        case 's': case 'S': skip = 2; break;
        case 'i': case 'I': case 'f': skip = 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = p;
            while (z < end && *z) ++z;
            if (z >= end) return -1;
            skip = (uint64_t)(z - p) + 1;
            break;
        }
        case 'B': {
            if (p + 5 > end) return -1;
            const uint8_t sub = p[0];
            const uint64_t width = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            skip = 5 + (uint64_t)bam_ld32(p + 1) * width;
            break;
        }
        default: return -1;
        }
        if ((uint64_t)(end - p) < skip) return -1;
        p += skip;
    }
    return -1;
}

struct BamScatter {
    const uint64_t *rec_off;
    const uint32_t *l_seq, *base_off, *cigar_off;
    uint32_t *cigar;
    uint8_t *seq4, *qual;   // qual may be null
    uint16_t *lib;
    uint8_t *has_mr;        // may be null
};

// pass 4: one warp per kept record copies its CIGAR, SEQ and QUAL into the batch; lane 0 looks through the tags
__global__ void __launch_bounds__(256) bam_scatter_kernel(const uint8_t *__restrict__ s, int64_t n, BamScatter d, BamLibraries libs,
                                                          BamTotals *totals)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const uint8_t *p = s + d.rec_off[i] + 4;
        const uint32_t bs = bam_ld32(p - 4), l_name = p[8], n_cig = bam_ld16(p + 12), l_seq = d.l_seq[i];
        const uint8_t *cig = p + 32 + l_name, *seq = cig + 4ull * n_cig, *qual = seq + (l_seq + 1) / 2;
        const uint32_t c0 = d.cigar_off[i], b0 = d.base_off[i];
        for (uint32_t k = lane; k < n_cig; k += 32) d.cigar[c0 + k] = bam_ld32(cig + 4ull * k);
        uint8_t *seq_to = d.seq4 + (b0 >> 1);
        for (uint32_t k = lane; k < (l_seq + 1) / 2; k += 32) seq_to[k] = seq[k];
        if (d.qual) {
            uint8_t *q_to = d.qual + b0;
            for (uint32_t k = lane; k < l_seq; k += 32) q_to[k] = qual[k];
            if ((l_seq & 1) && lane == 0) q_to[l_seq] = 0xFF;
        }
        if (lane == 0) {
            const uint8_t *aux = qual + l_seq, *end = p + bs;
            uint16_t lib = 0;
            if (libs.n > 0) {
                lib = 0xFFFF;
                const int64_t at = bam_find_tag(aux, end, 'R', 'G');
                if (at >= 0 && aux[at] == 'Z') {
                    const uint8_t *value = aux + at + 1;
                    const char *id = libs.ids;
                    for (int k = 0; k < libs.n && lib == 0xFFFF; ++k) {
                        const uint8_t *v = value;
                        const char *c = id;
                        while (v < end && *c && *v == (uint8_t)*c) { ++v; ++c; }
                        if (v < end && !*c && !*v) lib = libs.library[k];
                        while (*id) ++id;
                        ++id;
                    }
                }
                if (lib == 0xFFFF) atomicMin(&totals->first_bad_library, (unsigned int)i);
            }
            d.lib[i] = lib;
            if (d.has_mr) d.has_mr[i] = bam_find_tag(aux, end, 'M', 'R') >= 0;
        }
    }
}

// CRC32 (the gzip polynomial, reflected) of every BGZF block's data: one warp per block, each lane a contiguous
// piece, pieces combined as zlib's crc32_combine does: crc(A B) = crc(A) * x^(8 |B|) + crc(B)  (mod P).
__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if (!(a & (m - 1))) break;
        }
        m >>= 1;
        b = b & 1 ? (b >> 1) ^ 0xEDB88320u : b >> 1;
        if (!m) break;
    }
    return p;
}

__device__ inline uint32_t crc_x_pow_bytes(uint64_t n_bytes, const uint32_t *__restrict__ x2n)
{
    // x^(8 n) mod P: x2n[k] = x^(2^k) mod P
    uint32_t p = 1u << 31;  // x^0
    uint64_t n = n_bytes;
    int k = 3;
    while (n) {
        if (n & 1) p = crc_mulmod(x2n[k & 31], p);
        n >>= 1;
        ++k;
    }
    return p;
}

__global__ void __launch_bounds__(256) bam_crc_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ out_off,
                                                      const uint32_t *__restrict__ isize, const uint32_t *__restrict__ want, int32_t n,
                                                      const uint32_t *__restrict__ table, const uint32_t *__restrict__ x2n,
                                                      int32_t *__restrict__ status)
{
    __shared__ uint32_t t[256];
    t[threadIdx.x] = table[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= n) return;
    const uint32_t size = isize[b];
    const uint32_t piece = (size + 31) / 32;
    const uint32_t lo = min(size, lane * piece), hi = min(size, lo + piece);
    const uint8_t *p = data + out_off[b];
    uint32_t crc = 0xFFFFFFFFu;
    for (uint32_t k = lo; k < hi; ++k) crc = t[(crc ^ p[k]) & 0xFF] ^ (crc >> 8);
    crc ^= 0xFFFFFFFFu;  // crc32 of this lane's piece (0 for an empty piece)
    uint32_t len = hi - lo;
    // tree: lane l absorbs lane l + step
    for (int step = 1; step < 32; step <<= 1) {
        const uint32_t other_crc = __shfl_down_sync(0xffffffffu, crc, step);
        const uint32_t other_len = __shfl_down_sync(0xffffffffu, len, step);
        if (!(lane & (2 * step - 1)) && lane + step < 32) {
            if (other_len) crc = crc_mulmod(crc_x_pow_bytes(other_len, x2n), crc) ^ other_crc;
            len += other_len;
        }
    }
    if (lane == 0 && crc != want[b]) status[b] = 2;
}

}  // namespace mdg

// ---------------------------------------------------------------------------------------------------------------------
// host side

struct BamDevSlab {     // one slab of the file in page-locked memory, with its BGZF blocks
    uint8_t *data = nullptr;
    size_t cap = 0, len = 0;
    std::vector<uint64_t> in_off, out_off;
    std::vector<uint32_t> in_len, isize, crc;
    uint64_t inflated = 0;
    bool last = false;
    int error = 0;
    std::string message;
};

struct BamDevDecoded {  // one decoded slab on the device
    uint8_t *stream = nullptr;   // carried-over bytes, then the slab's inflated blocks
    size_t stream_cap = 0;
    uint64_t stream_len = 0;
    DeviceArrays arrays;         // the batch
    uint64_t *rec_off = nullptr; // [cap_reads]
    uint8_t *has_mr = nullptr;   // [cap_reads]
    int64_t rec_cap = 0;
    mdg_dev_batch batch;         // what mdg_bam_stream_next hands out (a view of `arrays`)
    int64_t n = 0;
    int64_t n_seen = 0;
    bool last = false;
    int error = 0;
    std::string message;
};

struct mdg_bam_stream {
    mdg_ctx *ctx = nullptr;
    int fd = -1;
    std::string error;
    size_t slab_bytes = 0;
    uint64_t data_start = 0;  // uncompressed offset of the first record (behind the header)
    int32_t n_ref = 0;
    uint32_t drop_flags = 0;
    bool with_qual = true, want_mr = false;
    // read group -> library
    std::vector<char> lib_ids;
    std::vector<uint16_t> lib_index;
    char *d_lib_ids = nullptr;
    uint16_t *d_lib_index = nullptr;
    // file slabs (page-locked), filled by the reader thread
    std::vector<BamDevSlab> slabs;
    std::vector<int> slab_state;      // 0 free, 1 filled
    int slab_fill = 0, slab_take = 0;
    // decoded slabs, filled by the decoder thread
    BamDevDecoded decoded[2];
    int decoded_state[2] = {0, 0};    // 0 free, 1 ready, 2 handed to the caller
    int decode_fill = 0, decode_take = 0;
    int handed = -1;
    bool finished = false;            // the end of the file has been reported to the caller
    std::mutex mutex;
    std::condition_variable cond;
    bool stop = false;
    std::thread reader, decoder;
    // device scratch of the decoder
    cudaStream_t stream = nullptr;
    uint8_t *d_comp = nullptr;
    size_t d_comp_cap = 0;
    uint64_t *d_meta = nullptr;       // in_off | out_off | in_len | isize | crc | status
    uint8_t *h_meta = nullptr;
    int meta_cap = 0;
    mdg_inflate::InflateScratch *d_scratch = nullptr;
    mdg::BamSegments seg{};
    int64_t seg_cap = 0;
    unsigned long long *d_scan = nullptr;
    int64_t scan_cap = 0;
    mdg::BamTotals *d_totals = nullptr, *h_totals = nullptr;
    uint32_t *d_crc_tables = nullptr; // crc table [256] | x2n [32]
    // carry between slabs: the incomplete record at the end of the previous stream
    std::vector<uint8_t> carry_host;  // only for error messages / tests
    uint64_t carry_len = 0;
    int carry_from = -1;              // decoded[] index whose stream holds the carry at carry_at
    uint64_t carry_at = 0;
    int64_t records_seen = 0, blocks_done = 0, blocks_host = 0, guesses_wrong = 0;
    double t_read = 0, t_decode = 0, t_wait = 0;
    // encoder (mdg_bam_encode_batch): device buffers, two page-locked output buffers, a thread that writes the file
    uint64_t *e_out_off = nullptr;
    size_t e_out_off_cap = 0;
    uint8_t *e_records = nullptr, *e_blocks = nullptr, *e_packed = nullptr;
    size_t e_records_cap = 0, e_blocks_cap = 0, e_packed_cap = 0;
    uint32_t *e_sizes = nullptr;
    size_t e_sizes_cap = 0;
    unsigned long long *e_prefix = nullptr, *e_scan = nullptr, *e_total_host = nullptr;
    size_t e_prefix_cap = 0, e_scan_cap = 0;
    uint8_t *e_host[2] = {nullptr, nullptr};
    size_t e_host_cap[2] = {0, 0};
    int e_turn = 0;
    std::thread e_writer;
    int e_write_rc = 0;
    bool e_attr_set = false;
    int64_t e_bytes_in = 0, e_bytes_out = 0;
    double t_encode = 0, t_write_wait = 0;
};

extern "C" int mdg_bam_write_raw(mdg_bam_writer *w, const uint8_t *blocks, int64_t n_bytes);

namespace {

thread_local std::string g_stream_error;

// inflated bytes one slab may hold: the batch arrays index bases (and this stream) with 32 bits
constexpr uint64_t BAMDEV_STREAM_LIMIT = 3400ull << 20;

// Page-locked buffers are expensive to make (the kernel locks every page: ~0.4 s per GB) and cheap to keep: a stream
// hands its slab and output buffers back to this pool when it closes, and the next stream of the process takes them.
struct PinnedPool {
    std::mutex mutex;
    std::vector<std::pair<uint8_t *, size_t>> free_list;
    uint8_t *take(size_t want, size_t *cap)
    {
        {
            std::lock_guard<std::mutex> lock(mutex);
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].second >= want && (best == free_list.size() || free_list[i].second < free_list[best].second)) best = i;
            if (best < free_list.size() && free_list[best].second <= 2 * want + (64u << 20)) {
                uint8_t *p = free_list[best].first;
                *cap = free_list[best].second;
                free_list.erase(free_list.begin() + (ptrdiff_t)best);
                return p;
            }
        }
        uint8_t *p = nullptr;
        if (cudaHostAlloc((void **)&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        *cap = want;
        return p;
    }
    void give(uint8_t *p, size_t cap)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lock(mutex);
        free_list.emplace_back(p, cap);
    }
};
PinnedPool &pinned_pool()
{
    static PinnedPool *pool = new PinnedPool();  // never destroyed: the CUDA context may be gone by then
    return *pool;
}

int sfail(mdg_bam_stream *s, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) s->error = buf;
    else g_stream_error = buf;
    return code;
}

double stream_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// reader thread: slabs of the file into page-locked memory, cut into BGZF blocks
void bamdev_reader_loop(mdg_bam_stream *s)
{
    std::vector<uint8_t> carry;  // what the previous slab left over: the piece of a BGZF block cut by its end, or
                                 // the blocks that would have inflated past BAMDEV_STREAM_LIMIT
    uint64_t file_at = 0;
    bool done = false, file_ended = false;
    while (!done) {
        int k;
        {
            std::unique_lock<std::mutex> lock(s->mutex);
            s->cond.wait(lock, [&] { return s->stop || s->slab_state[s->slab_fill] == 0; });
            if (s->stop) return;
            k = s->slab_fill;
        }
        BamDevSlab &slab = s->slabs[k];
        const double t0 = stream_now();
        slab.in_off.clear(); slab.out_off.clear(); slab.in_len.clear(); slab.isize.clear(); slab.crc.clear();
        slab.error = 0;
        slab.last = false;
        memcpy(slab.data, carry.data(), carry.size());
        size_t len = carry.size();
        // several preads in flight: one thread copies out of the page cache at a few GB/s only.  What was carried over
        // (the blocks of the slab before that would have inflated past the limit) counts towards this slab.
        const size_t want = file_ended ? 0 : s->slab_bytes > carry.size() ? s->slab_bytes - carry.size() : 0;
        {
            const int n_io = 4;
            const size_t piece = (want + n_io - 1) / n_io;
            size_t got[n_io] = {0};
            std::thread io[n_io];
            for (int t = 0; t < n_io; ++t)
                io[t] = std::thread([&, t] {
                    size_t at = (size_t)t * piece, end = std::min(want, at + piece);
                    while (at < end) {
                        const ssize_t r = pread(s->fd, slab.data + len + at, end - at, (off_t)(file_at + at));
                        if (r <= 0) break;
                        at += (size_t)r;
                    }
                    got[t] = at - std::min(want, (size_t)t * piece);
                });
            for (auto &th : io) th.join();
            size_t total = 0;
            for (int t = 0; t < n_io; ++t) {
                const size_t full = std::min(want, (size_t)(t + 1) * piece) - std::min(want, (size_t)t * piece);
                total += got[t];
                if (got[t] < full) break;  // the file ends in this piece
            }
            file_at += total;
            len += total;
            if (total < want) file_ended = true;
        }
        carry.clear();
        bool cut = false;
        slab.len = len;
        size_t at = 0;
        uint64_t out_off = 0;
        const uint8_t *in = slab.data;
        while (at < len) {
            const size_t left = len - at;
            int64_t total = -1;
            if (left >= 18) {
                if (in[at] != 31 || in[at + 1] != 139 || in[at + 2] != 8 || !(in[at + 3] & 4)) {
                    slab.error = MDG_ERR_DATA;
                    slab.message = "not a BGZF block (bad gzip member header)";
                    break;
                }
                const uint32_t xlen = (uint32_t)in[at + 10] | (uint32_t)in[at + 11] << 8;
                if (left >= 12 + (size_t)xlen) {
                    int64_t bsize = -1;
                    for (size_t x = 0; x + 4 <= xlen;) {
                        const uint32_t slen = (uint32_t)in[at + 12 + x + 2] | (uint32_t)in[at + 12 + x + 3] << 8;
                        if (in[at + 12 + x] == 'B' && in[at + 12 + x + 1] == 'C' && slen == 2 && x + 6 <= xlen)
                            bsize = (int64_t)((uint32_t)in[at + 12 + x + 4] | (uint32_t)in[at + 12 + x + 5] << 8);
                        x += 4 + slen;
                    }
                    if (bsize < 0 || bsize + 1 < 12 + (int64_t)xlen + 8) {
                        slab.error = MDG_ERR_DATA;
                        slab.message = bsize < 0 ? "BGZF block without a BC subfield" : "BGZF block with an impossible size";
                        break;
                    }
                    total = bsize + 1;
                    if (left >= (size_t)total) {
                        const uint8_t *tail = in + at + total - 8;
                        const uint32_t isize = (uint32_t)tail[4] | (uint32_t)tail[5] << 8 | (uint32_t)tail[6] << 16 | (uint32_t)tail[7] << 24;
                        if (isize > 65536) {
                            slab.error = MDG_ERR_DATA;
                            slab.message = "BGZF block claims more than 65536 bytes of data";
                            break;
                        }
                        if (out_off + isize > BAMDEV_STREAM_LIMIT && !slab.in_off.empty()) {
                            // the batch arrays index bases with 32 bits: the rest of the slab waits for the next one
                            carry.assign(in + at, in + len);
                            cut = true;
                            break;
                        }
                        if (isize) {
                            slab.in_off.push_back(at + 12 + xlen);
                            slab.in_len.push_back((uint32_t)((size_t)total - 12 - xlen - 8));
                            slab.isize.push_back(isize);
                            slab.out_off.push_back(out_off);
                            slab.crc.push_back((uint32_t)tail[0] | (uint32_t)tail[1] << 8 | (uint32_t)tail[2] << 16 | (uint32_t)tail[3] << 24);
                            out_off += isize;
                        }
                        at += (size_t)total;
                        continue;
                    }
                }
            }
            if (file_ended) {
                slab.error = MDG_ERR_DATA;
                slab.message = "truncated BGZF block";
            } else {
                carry.assign(in + at, in + len);
            }
            break;
        }
        if (!carry.empty()) slab.len = at;  // what is carried over is not this slab's to copy
        slab.last = file_ended && carry.empty() && !cut;
        slab.inflated = out_off;
        done = slab.last || slab.error;
        {
            std::lock_guard<std::mutex> lock(s->mutex);
            s->slab_state[k] = 1;
            s->slab_fill = (k + 1) % (int)s->slabs.size();
            s->t_read += stream_now() - t0;
        }
        s->cond.notify_all();
    }
}

#define MDG_S_CUDA(s, out, call)                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            (out).error = MDG_ERR_CUDA;                                                        \
            (out).message = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return;                                                                            \
        }                                                                                      \
    } while (0)

template <typename T>
bool bamdev_grow(T *&ptr, size_t &cap, size_t want, size_t slack_num = 5, size_t slack_den = 4)
{
    if (want <= cap) return true;
    cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    const size_t grown = want * slack_num / slack_den + 256;
    if (cudaMalloc((void **)&ptr, grown * sizeof(T)) != cudaSuccess) return false;
    cap = grown;
    return true;
}

// decodes one slab into `out` on the decoder's CUDA stream; synchronous (returns when the batch is complete)
void bamdev_decode(mdg_bam_stream *s, BamDevSlab &slab, BamDevDecoded &out)
{
    out.error = 0;
    out.n = 0;
    out.n_seen = 0;
    out.last = slab.last;
    if (slab.error) {
        out.error = slab.error;
        out.message = slab.message;
        return;
    }
    MDG_S_CUDA(s, out, cudaSetDevice(s->ctx->cfg.device));
    static const bool timing = getenv("MDG_BAM_TIMING") != nullptr;  // per-slab stage times on stderr
    const double tt0 = stream_now();
    double tt1 = tt0, tt2 = tt0, tt3 = tt0, tt4 = tt0;
    const int nb = (int)slab.in_off.size();
    const uint64_t carry = s->carry_len;
    const uint64_t stream_len = carry + slab.inflated;
    if (stream_len >= (1ull << 32) - (1u << 20)) {
        out.error = MDG_ERR_CAPACITY;
        out.message = "a slab inflates to more than 4 GB: use a smaller slab";
        return;
    }
    // device buffers
    if (!bamdev_grow(s->d_comp, s->d_comp_cap, slab.len + 64) || !bamdev_grow(out.stream, out.stream_cap, (size_t)stream_len + 4096)) {
        out.error = MDG_ERR_CUDA;
        out.message = "out of device memory (slab buffers)";
        return;
    }
    if (nb > s->meta_cap) {
        cudaFree(s->d_meta);
        cudaFree(s->d_scratch);
        cudaFreeHost(s->h_meta);
        s->d_meta = nullptr; s->d_scratch = nullptr; s->h_meta = nullptr; s->meta_cap = 0;
        const int cap = nb * 5 / 4 + 64;
        MDG_S_CUDA(s, out, cudaMalloc(&s->d_meta, (size_t)cap * 32));
        MDG_S_CUDA(s, out, cudaMalloc(&s->d_scratch, (size_t)cap * sizeof(mdg_inflate::InflateScratch)));
        MDG_S_CUDA(s, out, cudaHostAlloc(&s->h_meta, (size_t)cap * 32, cudaHostAllocDefault));
        s->meta_cap = cap;
    }
    const size_t cap = (size_t)s->meta_cap;
    uint64_t *h_in_off = (uint64_t *)s->h_meta, *h_out_off = h_in_off + cap;
    uint32_t *h_in_len = (uint32_t *)(h_out_off + cap), *h_isize = h_in_len + cap, *h_crc = h_isize + cap;
    int32_t *h_status = (int32_t *)(h_crc + cap);
    for (int i = 0; i < nb; ++i) {
        h_in_off[i] = slab.in_off[(size_t)i];
        h_out_off[i] = carry + slab.out_off[(size_t)i];
        h_in_len[i] = slab.in_len[(size_t)i];
        h_isize[i] = slab.isize[(size_t)i];
        h_crc[i] = slab.crc[(size_t)i];
    }
    uint64_t *d_in_off = s->d_meta, *d_out_off = d_in_off + cap;
    uint32_t *d_in_len = (uint32_t *)(d_out_off + cap), *d_isize = d_in_len + cap, *d_crc = d_isize + cap;
    int32_t *d_status = (int32_t *)(d_crc + cap);
    cudaStream_t st = s->stream;
    tt1 = stream_now();
    // the carried-over tail of the previous stream goes in front
    if (carry) {
        MDG_S_CUDA(s, out, cudaMemcpyAsync(out.stream, s->decoded[s->carry_from].stream + s->carry_at, (size_t)carry,
                                           cudaMemcpyDeviceToDevice, st));
    }
    if (nb) {
        MDG_S_CUDA(s, out, cudaMemcpyAsync(s->d_meta, s->h_meta, cap * 28, cudaMemcpyHostToDevice, st));
        MDG_S_CUDA(s, out, cudaMemcpyAsync(s->d_comp, slab.data, slab.len, cudaMemcpyHostToDevice, st));
        mdg::inflate_blocks_kernel<<<(nb + 31) / 32, 32, 0, st>>>(s->d_comp, d_in_off, d_in_len, out.stream, d_out_off, d_isize, d_status,
                                                                 s->d_scratch, nb);
        mdg::bam_crc_kernel<<<(nb + 7) / 8, 256, 0, st>>>(out.stream, d_out_off, d_isize, d_crc, nb, s->d_crc_tables, s->d_crc_tables + 256,
                                                          d_status);
        MDG_S_CUDA(s, out, cudaGetLastError());
        MDG_S_CUDA(s, out, cudaMemcpyAsync(h_status, d_status, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        MDG_S_CUDA(s, out, cudaStreamSynchronize(st));
        s->ctx->launches += 2;
        // blocks the device turned down or got wrong: the host decoder, zlib behind it (as in mdg_bamio.cpp)
        std::vector<uint8_t> tmp;
        for (int i = 0; i < nb; ++i) {
            if (h_status[i] == 0) continue;
            tmp.resize(h_isize[i]);
            const uint8_t *in = slab.data + h_in_off[i];
            bool ok = mdg_inflate_raw(in, (int64_t)h_in_len[i], tmp.data(), (int64_t)h_isize[i]) == (int64_t)h_isize[i];
            if (!ok) {
                z_stream z;
                memset(&z, 0, sizeof z);
                if (inflateInit2(&z, -15) == Z_OK) {
                    z.next_in = (Bytef *)in;
                    z.avail_in = (uInt)h_in_len[i];
                    z.next_out = tmp.data();
                    z.avail_out = h_isize[i];
                    ok = inflate(&z, Z_FINISH) == Z_STREAM_END && z.avail_out == 0;
                    inflateEnd(&z);
                }
            }
            if (!ok || (uint32_t)crc32(crc32(0L, Z_NULL, 0), tmp.data(), h_isize[i]) != h_crc[i]) {
                out.error = MDG_ERR_DATA;
                out.message = ok ? "BGZF block fails its CRC32" : "BGZF block does not inflate";
                return;
            }
            MDG_S_CUDA(s, out, cudaMemcpy(out.stream + h_out_off[i], tmp.data(), h_isize[i], cudaMemcpyHostToDevice));
            ++s->blocks_host;
        }
        s->blocks_done += nb;
    } else if (carry) {
        MDG_S_CUDA(s, out, cudaStreamSynchronize(st));
    }
    out.stream_len = stream_len;
    tt2 = stream_now();
    // ---- record boundaries ----
    const uint64_t start0 = s->records_seen == 0 && s->carry_from < 0 ? s->data_start : 0;
    const int64_t n_seg = (int64_t)((stream_len + mdg::BAM_SEGMENT - 1) / mdg::BAM_SEGMENT);
    if (n_seg > s->seg_cap) {
        cudaFree(s->seg.start);
        s->seg = mdg::BamSegments{};
        s->seg_cap = 0;
        const int64_t scap = n_seg * 5 / 4 + 64;
        void *block = nullptr;
        MDG_S_CUDA(s, out, cudaMalloc(&block, (size_t)scap * 28));
        s->seg.start = (uint64_t *)block;
        s->seg.next = s->seg.start + scap;
        s->seg.seen = (uint32_t *)(s->seg.next + scap);
        s->seg.kept = s->seg.seen + scap;
        s->seg.stopped = s->seg.kept + scap;
        s->seg_cap = scap;
    }
    if (start0 > stream_len) {
        if (slab.last) {
            out.error = MDG_ERR_DATA;
            out.message = "truncated BAM header";
            return;
        }
        // the header is longer than this slab's stream: carry everything (cannot happen with slabs of many MB)
        out.error = MDG_ERR_CAPACITY;
        out.message = "the BAM header does not fit one slab";
        return;
    }
    if (n_seg) {
        mdg::bam_guess_walk_kernel<<<(unsigned)((n_seg + 127) / 128), 128, 0, st>>>(out.stream, stream_len, start0, s->n_ref, s->drop_flags,
                                                                                    s->seg, n_seg);
        mdg::bam_verify_kernel<<<1, 32, 0, st>>>(out.stream, stream_len, start0, s->drop_flags, s->seg, n_seg, s->d_totals);
        MDG_S_CUDA(s, out, cudaGetLastError());
        MDG_S_CUDA(s, out, cudaMemcpyAsync(s->h_totals, s->d_totals, sizeof(mdg::BamTotals), cudaMemcpyDeviceToHost, st));
        MDG_S_CUDA(s, out, cudaStreamSynchronize(st));
        s->ctx->launches += 2;
    } else {
        memset(s->h_totals, 0, sizeof(mdg::BamTotals));
        s->h_totals->tail = stream_len;
    }
    const mdg::BamTotals walked = *s->h_totals;
    tt3 = stream_now();
    if (walked.malformed) {
        out.error = MDG_ERR_DATA;
        char buf[128];
        snprintf(buf, sizeof buf, "BAM record %lld is shorter than its fixed part", (long long)(s->records_seen + (int64_t)walked.n_seen));
        out.message = buf;
        return;
    }
    if (slab.last && walked.tail < stream_len) {
        out.error = MDG_ERR_DATA;
        out.message = "BAM stream ends inside a record";
        return;
    }
    s->guesses_wrong += walked.guesses_wrong;
    const int64_t n = (int64_t)walked.n_kept;
    out.n = n;
    out.n_seen = (int64_t)walked.n_seen;
    // what is left over goes to the front of the next slab's stream
    s->carry_len = stream_len - walked.tail;
    s->carry_at = walked.tail;
    s->carry_from = (int)(&out - s->decoded);
    s->records_seen += (int64_t)walked.n_seen;
    if (n == 0) {
        out.arrays.view.n_reads = 0;
        out.arrays.n_cigar = out.arrays.n_bases = 0;
        return;
    }
    if (n >= (1ll << 31) - 1) {
        out.error = MDG_ERR_CAPACITY;
        out.message = "more than 2^31 records in one slab: use a smaller slab";
        return;
    }
    // ---- fixed fields, then the layout ----
    // the variable-length arrays are bounded by the stream itself: every CIGAR word, packed base pair and quality byte
    // of a record is a byte range of the stream
    const int64_t cap_reads = n, cap_cigar = (int64_t)(stream_len / 4) + 1, cap_bases = (int64_t)((stream_len + 1) & ~1ull) + 2 * n;
    if (out.arrays.cap_reads < cap_reads || out.arrays.cap_cigar < cap_cigar || out.arrays.cap_bases < cap_bases ||
        out.arrays.has_qual != s->with_qual) {
        cudaFree(out.arrays.block);
        out.arrays = DeviceArrays{};
        // qualities are at most as many as bases; sized generously once so that later slabs fit
        if (alloc_arrays(s->ctx, out.arrays, cap_reads * 9 / 8 + 1024, cap_cigar * 9 / 8, (cap_bases * 9 / 8) & ~1ll, s->with_qual)) {
            out.error = MDG_ERR_CUDA;
            out.message = "out of device memory (batch arrays): " + s->ctx->error;
            return;
        }
    }
    if (out.rec_cap < n) {
        cudaFree(out.rec_off);
        cudaFree(out.has_mr);
        out.rec_off = nullptr; out.has_mr = nullptr; out.rec_cap = 0;
        const int64_t rcap = n * 9 / 8 + 1024;
        MDG_S_CUDA(s, out, cudaMalloc(&out.rec_off, (size_t)rcap * 8));
        MDG_S_CUDA(s, out, cudaMalloc(&out.has_mr, (size_t)rcap));
        out.rec_cap = rcap;
    }
    const mdg::DevBatch &v = out.arrays.view;
    mdg::BamFields fields{out.rec_off, (uint16_t *)v.flag, (int32_t *)v.tid, (int32_t *)v.pos, (int32_t *)v.tlen, (int32_t *)v.mtid,
                          (int32_t *)v.mpos, (uint32_t *)v.l_seq, (uint32_t *)v.cigar_off};
    mdg::bam_fields_kernel<<<(unsigned)((n_seg + 127) / 128), 128, 0, st>>>(out.stream, stream_len, s->drop_flags, s->seg, n_seg, fields);
    const int64_t n_blocks = (n + 255) / 256;
    if (n_blocks + 2 > s->scan_cap) {
        cudaFree(s->d_scan);
        s->d_scan = nullptr;
        s->scan_cap = 0;
        MDG_S_CUDA(s, out, cudaMalloc(&s->d_scan, (size_t)(n_blocks * 5 / 4 + 16) * 16));
        s->scan_cap = n_blocks * 5 / 4 + 16;
    }
    mdg::bam_layout_totals<<<(unsigned)n_blocks, 256, 0, st>>>(v.l_seq, v.cigar_off, n, s->d_scan);
    mdg::synth_scan_totals<<<1, 1024, 0, st>>>(s->d_scan, n_blocks);
    mdg::bam_layout_fill<<<(unsigned)n_blocks, 256, 0, st>>>(v.l_seq, (uint32_t *)v.cigar_off, n, s->d_scan, n_blocks, (uint32_t *)v.base_off,
                                                             s->d_totals);
    mdg::BamScatter sc{out.rec_off, v.l_seq, v.base_off, v.cigar_off, (uint32_t *)v.cigar, (uint8_t *)v.seq4,
                       s->with_qual ? (uint8_t *)v.qual : nullptr, (uint16_t *)v.lib, s->want_mr ? out.has_mr : nullptr};
    mdg::BamLibraries libs{s->d_lib_ids, s->d_lib_index, (int32_t)s->lib_index.size()};
    const int sgrid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)s->ctx->sm_count * 16);
    mdg::bam_scatter_kernel<<<sgrid, 256, 0, st>>>(out.stream, n, sc, libs, s->d_totals);
    MDG_S_CUDA(s, out, cudaGetLastError());
    MDG_S_CUDA(s, out, cudaMemcpyAsync(s->h_totals, s->d_totals, sizeof(mdg::BamTotals), cudaMemcpyDeviceToHost, st));
    MDG_S_CUDA(s, out, cudaStreamSynchronize(st));
    s->ctx->launches += 5;
    const mdg::BamTotals laid = *s->h_totals;
    tt4 = stream_now();
    if (timing)
        fprintf(stderr, "device slab: %d blocks, %lld records, buffers %.3f s, copy + inflate + crc %.3f s, walk %.3f s, fields + layout + scatter %.3f s (with allocations)\n",
                nb, (long long)n, tt1 - tt0, tt2 - tt1, tt3 - tt2, tt4 - tt3);
    if (laid.bases >= (1ull << 32) || laid.cigars >= (1ull << 32)) {
        out.error = MDG_ERR_CAPACITY;
        out.message = "a slab holds more than 2^32 bases: use a smaller slab";
        return;
    }
    out.arrays.view.n_reads = n;
    out.arrays.n_cigar = (int64_t)laid.cigars;
    out.arrays.n_bases = (int64_t)laid.bases;
    if (laid.first_bad_library != 0xFFFFFFFFu) {
        // reader.py:67-81: the reference's message carries the read's name (and its read group)
        uint64_t off = 0;
        std::vector<uint8_t> rec(4 + 36 + 256);
        std::string name = "?", group;
        if (cudaMemcpy(&off, out.rec_off + laid.first_bad_library, 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
            uint32_t size = 0;
            cudaMemcpy(&size, out.stream + off, 4, cudaMemcpyDeviceToHost);
            rec.resize(4 + (size_t)size);
            if (cudaMemcpy(rec.data(), out.stream + off, rec.size(), cudaMemcpyDeviceToHost) == cudaSuccess && size >= 32) {
                const uint8_t *p = rec.data() + 4;
                const uint32_t l_name = p[8], n_cig = (uint32_t)p[12] | (uint32_t)p[13] << 8;
                const uint32_t l_seq = (uint32_t)p[16] | (uint32_t)p[17] << 8 | (uint32_t)p[18] << 16 | (uint32_t)p[19] << 24;
                name.assign((const char *)p + 32, l_name ? l_name - 1 : 0);
                const uint8_t *aux = p + 32 + l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq, *end = p + size;
                // RG:Z among the tags
                while (aux + 3 <= end) {
                    if (aux[0] == 'R' && aux[1] == 'G' && aux[2] == 'Z') {
                        group.assign((const char *)aux + 3);
                        break;
                    }
                    const uint8_t type = aux[2];
                    aux += 3;
                    size_t skip = 0;
                    if (type == 'A' || type == 'c' || type == 'C') skip = 1;
                    else if (type == 's' || type == 'S') skip = 2;
                    else if (type == 'i' || type == 'I' || type == 'f') skip = 4;
                    else if (type == 'Z' || type == 'H') skip = strnlen((const char *)aux, (size_t)(end - aux)) + 1;
                    else if (type == 'B' && aux + 5 <= end) {
                        const size_t width = (aux[0] == 'c' || aux[0] == 'C') ? 1 : (aux[0] == 's' || aux[0] == 'S') ? 2 : 4;
                        skip = 5 + ((size_t)aux[1] | (size_t)aux[2] << 8 | (size_t)aux[3] << 16 | (size_t)aux[4] << 24) * width;
                    } else break;
                    if ((size_t)(end - aux) < skip) break;
                    aux += skip;
                }
            }
        }
        out.error = MDG_ERR_DATA;
        if (group.empty()) out.message = "Read '" + name + "' has no read-group. Either fix BAM or use --merge-libraries";
        else out.message = "Read '" + name + "' has read-group not listed in BAM header ('" + group + "'); either fix BAM or use --merge-libraries";
    }
}

void bamdev_decoder_loop(mdg_bam_stream *s)
{
    while (true) {
        int k, d;
        {
            std::unique_lock<std::mutex> lock(s->mutex);
            s->cond.wait(lock, [&] { return s->stop || (s->slab_state[s->slab_take] == 1 && s->decoded_state[s->decode_fill] == 0); });
            if (s->stop) return;
            k = s->slab_take;
            d = s->decode_fill;
        }
        const double t0 = stream_now();
        bamdev_decode(s, s->slabs[k], s->decoded[d]);
        const bool done = s->decoded[d].last || s->decoded[d].error;
        {
            std::lock_guard<std::mutex> lock(s->mutex);
            s->slab_state[k] = 0;
            s->slab_take = (k + 1) % (int)s->slabs.size();
            s->decoded_state[d] = 1;
            s->decode_fill = d ^ 1;
            s->t_decode += stream_now() - t0;
        }
        s->cond.notify_all();
        if (done) return;
    }
}

}  // namespace

extern "C" {

int mdg_bam_stream_open(mdg_ctx *ctx, const char *path, uint64_t data_start, int32_t n_references, int64_t slab_bytes,
                        mdg_bam_stream **out)
{
    if (!ctx || !path || !out || n_references < 0) return sfail(nullptr, MDG_ERR_ARGUMENT, "mdg_bam_stream_open: bad argument");
    *out = nullptr;
    mdg_bam_stream *s = new (std::nothrow) mdg_bam_stream();
    if (!s) return sfail(nullptr, MDG_ERR_ARGUMENT, "out of host memory");
    s->ctx = ctx;
    s->data_start = data_start;
    s->n_ref = n_references;
    s->fd = open(path, O_RDONLY);
    if (s->fd < 0) {
        delete s;
        return sfail(nullptr, MDG_ERR_ARGUMENT, "cannot open %s", path);
    }
    struct stat st;
    size_t file_size = 0;
    if (fstat(s->fd, &st) == 0 && S_ISREG(st.st_mode)) file_size = (size_t)st.st_size;
    if (const char *env = getenv("MDG_BAM_DEVICE_SLAB")) slab_bytes = atoll(env);
    if (slab_bytes <= 0) slab_bytes = 2048ll << 20;  // the inflate kernel is latency-bound: the more blocks per launch the better
    if (slab_bytes < (1 << 17)) slab_bytes = 1 << 17;
    // no more page-locked memory than the file needs
    int n_slabs = slab_bytes >= (512ll << 20) ? 2 : 3;
    if (file_size && (size_t)slab_bytes >= file_size) {
        slab_bytes = (int64_t)file_size + 1;
        n_slabs = 2;  // a slab cut short by BAMDEV_STREAM_LIMIT leaves its rest to a second one
    }
    s->slab_bytes = (size_t)slab_bytes;
    auto bail = [&](int code, const char *what) {
        sfail(nullptr, code, "mdg_bam_stream_open: %s", what);
        mdg_bam_stream_close(s);
        return code;
    };
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return bail(MDG_ERR_CUDA, "cudaSetDevice failed");
    s->slabs.resize((size_t)n_slabs);
    s->slab_state.assign((size_t)n_slabs, 0);
    for (auto &slab : s->slabs) {
        // room for the piece of a block carried over from the slab before
        slab.data = pinned_pool().take(s->slab_bytes + (1u << 17), &slab.cap);
        if (!slab.data) return bail(MDG_ERR_CUDA, "cudaHostAlloc failed");
    }
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(MDG_ERR_CUDA, "cudaStreamCreate failed");
    if (cudaMalloc(&s->d_totals, sizeof(mdg::BamTotals)) != cudaSuccess ||
        cudaHostAlloc((void **)&s->h_totals, sizeof(mdg::BamTotals), cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc(&s->d_crc_tables, (256 + 32) * 4) != cudaSuccess)
        return bail(MDG_ERR_CUDA, "device allocation failed");
    {
        uint32_t tables[256 + 32];
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = c & 1 ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            tables[i] = c;
        }
        // x2n[k] = x^(2^k) mod P (reflected: bit 31 is x^0)
        auto mulmod = [](uint32_t a, uint32_t b) {
            uint32_t m = 1u << 31, p = 0;
            for (;;) {
                if (a & m) {
                    p ^= b;
                    if (!(a & (m - 1))) break;
                }
                m >>= 1;
                b = b & 1 ? (b >> 1) ^ 0xEDB88320u : b >> 1;
                if (!m) break;
            }
            return p;
        };
        uint32_t p = 1u << 30;  // x^1
        tables[256] = p;
        for (int k = 1; k < 32; ++k) tables[256 + k] = p = mulmod(p, p);
        if (cudaMemcpy(s->d_crc_tables, tables, sizeof tables, cudaMemcpyHostToDevice) != cudaSuccess) return bail(MDG_ERR_CUDA, "copy failed");
    }
    *out = s;
    return MDG_OK;
}

void mdg_bam_stream_close(mdg_bam_stream *s)
{
    if (!s) return;
    {
        std::lock_guard<std::mutex> lock(s->mutex);
        s->stop = true;
    }
    s->cond.notify_all();
    if (s->reader.joinable()) s->reader.join();
    if (s->decoder.joinable()) s->decoder.join();
    if (s->e_writer.joinable()) s->e_writer.join();
    if (s->ctx) cudaSetDevice(s->ctx->cfg.device);
    if (s->ctx) cudaStreamSynchronize(s->ctx->compute);
    if (s->fd >= 0) close(s->fd);
    for (auto &slab : s->slabs) pinned_pool().give(slab.data, slab.cap);
    for (auto &d : s->decoded) {
        cudaFree(d.stream);
        cudaFree(d.arrays.block);
        cudaFree(d.rec_off);
        cudaFree(d.has_mr);
        cudaFree(d.batch.res_mr);
        cudaFree(d.batch.res_status);
    }
    cudaFree(s->d_comp);
    cudaFree(s->d_meta);
    cudaFree(s->d_scratch);
    cudaFreeHost(s->h_meta);
    cudaFree(s->seg.start);
    cudaFree(s->d_scan);
    cudaFree(s->d_totals);
    cudaFreeHost(s->h_totals);
    cudaFree(s->d_crc_tables);
    cudaFree(s->d_lib_ids);
    cudaFree(s->d_lib_index);
    cudaFree(s->e_out_off);
    cudaFree(s->e_records);
    cudaFree(s->e_blocks);
    cudaFree(s->e_packed);
    cudaFree(s->e_sizes);
    cudaFree(s->e_prefix);
    cudaFree(s->e_scan);
    if (s->e_total_host) cudaFreeHost(s->e_total_host);
    for (int i = 0; i < 2; ++i) pinned_pool().give(s->e_host[i], s->e_host_cap[i]);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
}

const char *mdg_bam_stream_error(const mdg_bam_stream *s) { return s ? s->error.c_str() : g_stream_error.c_str(); }

int mdg_bam_stream_set_libraries(mdg_bam_stream *s, const char *const *read_groups, const uint16_t *library, int32_t n)
{
    if (!s || n < 0 || (n && (!read_groups || !library))) return MDG_ERR_ARGUMENT;
    if (s->reader.joinable()) return sfail(s, MDG_ERR_STATE, "mdg_bam_stream_set_libraries must precede the first mdg_bam_stream_next");
    s->lib_ids.clear();
    s->lib_index.clear();
    for (int32_t i = 0; i < n; ++i) {
        const size_t l = strlen(read_groups[i]);
        s->lib_ids.insert(s->lib_ids.end(), read_groups[i], read_groups[i] + l + 1);
        s->lib_index.push_back(library[i]);
    }
    cudaSetDevice(s->ctx->cfg.device);
    cudaFree(s->d_lib_ids);
    cudaFree(s->d_lib_index);
    s->d_lib_ids = nullptr;
    s->d_lib_index = nullptr;
    if (n) {
        if (cudaMalloc(&s->d_lib_ids, s->lib_ids.size()) != cudaSuccess || cudaMalloc(&s->d_lib_index, (size_t)n * 2) != cudaSuccess ||
            cudaMemcpy(s->d_lib_ids, s->lib_ids.data(), s->lib_ids.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(s->d_lib_index, s->lib_index.data(), (size_t)n * 2, cudaMemcpyHostToDevice) != cudaSuccess)
            return sfail(s, MDG_ERR_CUDA, "mdg_bam_stream_set_libraries: device copy failed");
    }
    return MDG_OK;
}

int64_t mdg_bam_stream_next(mdg_bam_stream *s, uint32_t drop_flags, int32_t with_qual, int32_t want_mr, mdg_dev_batch **out)
{
    if (!s || !out) return MDG_ERR_ARGUMENT;
    *out = nullptr;
    if (s->finished) return 0;
    if (!s->reader.joinable()) {
        // first call: the options are fixed from here on
        s->drop_flags = drop_flags;
        s->with_qual = with_qual != 0;
        s->want_mr = want_mr != 0;
        s->reader = std::thread(bamdev_reader_loop, s);
        s->decoder = std::thread(bamdev_decoder_loop, s);
    } else if (drop_flags != s->drop_flags || (with_qual != 0) != s->with_qual || (want_mr != 0) != s->want_mr) {
        return sfail(s, MDG_ERR_STATE, "mdg_bam_stream_next: options differ from the first call");
    }
    while (true) {
        // the batch handed out before is the caller's no longer: its kernels (on the context's streams) must be done
        // before the decoder overwrites it
        if (s->handed >= 0) {
            cudaSetDevice(s->ctx->cfg.device);
            if (cudaStreamSynchronize(s->ctx->compute) != cudaSuccess) return sfail(s, MDG_ERR_CUDA, "mdg_bam_stream_next: stream sync failed");
            BamDevDecoded &prev = s->decoded[s->handed];
            const bool was_last = prev.last;
            {
                std::lock_guard<std::mutex> lock(s->mutex);
                s->decoded_state[s->handed] = 0;
                s->handed = -1;
            }
            s->cond.notify_all();
            if (was_last) {
                s->finished = true;
                return 0;
            }
        }
        int d;
        const double t0 = stream_now();
        {
            std::unique_lock<std::mutex> lock(s->mutex);
            s->cond.wait(lock, [&] { return s->decoded_state[s->decode_take] == 1; });
            d = s->decode_take;
            s->decode_take = d ^ 1;
            s->decoded_state[d] = 2;
            s->t_wait += stream_now() - t0;
        }
        BamDevDecoded &dec = s->decoded[d];
        s->handed = d;
        if (dec.error) {
            s->finished = true;  // the decoder thread has stopped
            return sfail(s, dec.error, "%s", dec.message.c_str());
        }
        if (dec.n == 0) {
            if (dec.last) {
                std::lock_guard<std::mutex> lock(s->mutex);
                s->decoded_state[d] = 0;
                s->handed = -1;
                s->finished = true;
                return 0;
            }
            continue;  // a slab without a kept record
        }
        dec.batch.arrays = dec.arrays;
        if (dec.batch.res_cap < dec.arrays.cap_reads) dec.batch.res_cap = 0;  // reallocated on demand by mdg_rescale_resident
        *out = &dec.batch;
        return dec.n;
    }
}

/* has_mr flags (rescale.py:277-278) of the batch handed out last, copied to the host */
int mdg_bam_stream_has_mr(mdg_bam_stream *s, uint8_t *has_mr, int64_t n)
{
    if (!s || !has_mr || s->handed < 0) return MDG_ERR_ARGUMENT;
    BamDevDecoded &dec = s->decoded[s->handed];
    if (n != dec.n || !s->want_mr) return sfail(s, MDG_ERR_ARGUMENT, "mdg_bam_stream_has_mr: no such batch");
    cudaSetDevice(s->ctx->cfg.device);
    if (cudaMemcpy(has_mr, dec.has_mr, (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return sfail(s, MDG_ERR_CUDA, "copy failed");
    return MDG_OK;
}

/* counters: records walked (dropped ones included), BGZF blocks inflated on the device / redone on the host,
 * segments whose first-record guess was wrong, seconds spent reading / decoding / waiting for a batch */
int mdg_bam_stream_stats(const mdg_bam_stream *s, int64_t *records_seen, int64_t *blocks_device, int64_t *blocks_host,
                         int64_t *guesses_wrong, double *seconds3)
{
    if (!s) return MDG_ERR_ARGUMENT;
    if (records_seen) *records_seen = s->records_seen;
    if (blocks_device) *blocks_device = s->blocks_done - s->blocks_host;
    if (blocks_host) *blocks_host = s->blocks_host;
    if (guesses_wrong) *guesses_wrong = s->guesses_wrong;
    if (seconds3) {
        seconds3[0] = s->t_read;
        seconds3[1] = s->t_decode;
        seconds3[2] = s->t_wait;
    }
    return MDG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// BAM encoder on the GPU: what pysam.AlignmentFile(path, "wb", template=...).write(read) is to the reference
// (rescale.py:298-299,344).  The records of a decoded slab are re-emitted in input order -- a rescaled one with its new
// qualities and an MR:f tag appended (rescale.py:273-280) -- and the byte stream is cut into BGZF blocks of 0xff00
// bytes, each deflated by one thread block: a dynamic Huffman code over the literals (built by one thread from the
// block's byte histogram, length-limited to 15 bits the way zlib does it), no matches -- sequence and quality bytes
// have none worth finding, and a literal-only block is a valid DEFLATE stream any inflater reads -- or a stored block
// when that is smaller.  CRC32 of every block is computed alongside.  The host writes the finished blocks.
namespace mdg {

constexpr int BGZF_CHUNK = 0xff00;           // uncompressed bytes per block, as htslib cuts them
constexpr int BGZF_SLOT = 65536 + 64;        // bytes reserved per encoded block on the device
constexpr int DEFLATE_THREADS = 256;
constexpr int DEFLATE_PIECE = BGZF_CHUNK / DEFLATE_THREADS;  // 255 bytes per thread

// new size of every record: 4 + block_size (+ 7 for the MR:f tag of a rescaled one)
__global__ void __launch_bounds__(256) bam_emit_totals(const uint8_t *__restrict__ s, const uint64_t *__restrict__ rec_off,
                                                       const uint8_t *__restrict__ status, int64_t n, unsigned long long *totals)
{
    __shared__ uint2 total;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t size = 0;
    if (i < n) size = 4 + bam_ld32(s + rec_off[i]) + ((status && (status[i] & 1)) ? 7 : 0);
    block_exclusive_scan2(size, 0, &total);
    if (threadIdx.x == 0) {
        totals[2 * (size_t)blockIdx.x] = total.x;
        totals[2 * (size_t)blockIdx.x + 1] = 0;
    }
}

__global__ void __launch_bounds__(256) bam_emit_offsets(const uint8_t *__restrict__ s, const uint64_t *__restrict__ rec_off,
                                                        const uint8_t *__restrict__ status, int64_t n,
                                                        const unsigned long long *__restrict__ block_off, uint64_t *out_off)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t size = 0;
    if (i < n) size = 4 + bam_ld32(s + rec_off[i]) + ((status && (status[i] & 1)) ? 7 : 0);
    const uint2 off = block_exclusive_scan2(size, 0, nullptr);
    if (i < n) out_off[i] = block_off[2 * (size_t)blockIdx.x] + off.x;
}

struct BamEmit {
    const uint64_t *rec_off, *out_off;
    const uint32_t *l_seq, *base_off;
    const uint8_t *qual;     // the batch's (rewritten) qualities
    const uint8_t *status;
    const float *mr;
};

// one warp per record: the record as read, new qualities and the MR tag where it was rescaled
__global__ void __launch_bounds__(256) bam_emit_kernel(const uint8_t *__restrict__ s, int64_t n, BamEmit e, uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const uint8_t *rec = s + e.rec_off[i];
        uint8_t *to = out + e.out_off[i];
        const uint32_t bs = bam_ld32(rec);
        const bool rescaled = e.status && (e.status[i] & 1);
        const uint32_t l_name = rec[12], n_cig = bam_ld16(rec + 16), l_seq = bam_ld32(rec + 20);
        const uint32_t q_at = 4 + 32 + l_name + 4 * n_cig + (l_seq + 1) / 2;
        for (uint32_t k = lane; k < 4 + bs; k += 32) {
            uint8_t v = rec[k];
            if (rescaled && k >= q_at && k < q_at + l_seq) v = e.qual[e.base_off[i] + (k - q_at)];
            to[k] = v;
        }
        if (rescaled) {
            __syncwarp();
            if (lane == 0) {
                const uint32_t size = bs + 7;
                to[0] = (uint8_t)size; to[1] = (uint8_t)(size >> 8); to[2] = (uint8_t)(size >> 16); to[3] = (uint8_t)(size >> 24);
                uint8_t *tag = to + 4 + bs;
                const uint32_t bits = __float_as_uint(e.mr[i]);
                tag[0] = 'M'; tag[1] = 'R'; tag[2] = 'f';
                tag[3] = (uint8_t)bits; tag[4] = (uint8_t)(bits >> 8); tag[5] = (uint8_t)(bits >> 16); tag[6] = (uint8_t)(bits >> 24);
            }
        }
    }
}

// appends `len` bits (LSB first) to a stream of 32-bit words; the first and last word of a thread's run may be shared
// with its neighbours, so everything goes through atomicOr on a zeroed buffer
struct BitWriter {
    uint32_t *words;      // output as 32-bit words (zeroed)
    uint64_t acc = 0;     // pending bits, starting at bit `fill_at` of word `at`
    uint32_t at = 0;
    int have = 0;         // bits in acc (including the lead-in offset)
    __device__ void start(uint32_t *w, uint64_t bit_offset)
    {
        words = w;
        at = (uint32_t)(bit_offset >> 5);
        have = (int)(bit_offset & 31);
        acc = 0;
    }
    __device__ __forceinline__ void put(uint32_t bits, int len)
    {
        acc |= (uint64_t)bits << have;
        have += len;
        if (have >= 32) {
            atomicOr(words + at, (uint32_t)acc);
            ++at;
            acc >>= 32;
            have -= 32;
        }
    }
    __device__ void finish()
    {
        if (have > 0) atomicOr(words + at, (uint32_t)acc);
    }
};

// One BGZF block per thread block.  out = slot c of `blocks` (BGZF_SLOT bytes, zeroed); sizes[c] = bytes of the block.
// Dynamic shared memory: the BGZF_CHUNK data bytes.
__global__ void __launch_bounds__(DEFLATE_THREADS) bgzf_deflate_kernel(const uint8_t *__restrict__ in, uint64_t total, uint8_t *__restrict__ blocks,
                                                                       uint32_t *__restrict__ sizes, const uint32_t *__restrict__ crc_table,
                                                                       const uint32_t *__restrict__ x2n)
{
    extern __shared__ uint8_t data[];         // [BGZF_CHUNK]
    __shared__ uint32_t freq[257];
    __shared__ uint32_t code[257];            // code bits (LSB first) | length << 16
    __shared__ uint32_t t_crc[256];
    __shared__ uint16_t order[257];           // used symbols by (frequency, symbol)
    __shared__ uint32_t node_w[513];
    __shared__ uint16_t node_parent[513];
    __shared__ uint8_t node_depth[513];
    __shared__ uint32_t scan_tmp[DEFLATE_THREADS / 32], w_crc[DEFLATE_THREADS / 32], w_len[DEFLATE_THREADS / 32];
    __shared__ uint32_t s_misc[4];            // used symbols, code bits (end-of-block included), stored flag, crc
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const uint64_t lo = chunk * BGZF_CHUNK;
    const uint32_t len = (uint32_t)min((uint64_t)BGZF_CHUNK, total - lo);
    for (int i = tid; i < 257; i += DEFLATE_THREADS) freq[i] = i == 256 ? 1u : 0u;
    t_crc[tid] = crc_table[tid];
    __syncthreads();
    for (uint32_t i = tid; i < len; i += DEFLATE_THREADS) {
        const uint8_t v = in[lo + i];
        data[i] = v;
        atomicAdd(&freq[v], 1u);
    }
    __syncthreads();
    // ---- rank of every used symbol by (frequency, symbol) ----
    for (int sym = tid; sym < 257; sym += DEFLATE_THREADS) {
        const uint32_t f = freq[sym];
        if (!f) continue;
        int rank = 0;
        for (int j = 0; j < 257; ++j) {
            const uint32_t fj = freq[j];
            rank += fj && (fj < f || (fj == f && j < sym));
        }
        order[rank] = (uint16_t)sym;
    }
    if (tid == 32) {
        int n_used = 0;
        for (int j = 0; j < 257; ++j) n_used += freq[j] != 0;
        s_misc[0] = (uint32_t)n_used;
    }
    __syncthreads();
    // ---- Huffman code lengths, one thread: two sorted queues (leaves, internal nodes) ----
    if (tid == 0) {
        const int n = (int)s_misc[0];  // >= 2: a literal and the end-of-block symbol
        for (int i = 0; i < n; ++i) node_w[i] = freq[order[i]];
        int leaf = 0, inner = n, made = n;
        while (made < 2 * n - 1) {
            int pick[2];
            for (int k = 0; k < 2; ++k) {
                if (leaf < n && (inner >= made || node_w[leaf] <= node_w[inner])) pick[k] = leaf++;
                else pick[k] = inner++;
            }
            node_w[made] = node_w[pick[0]] + node_w[pick[1]];
            node_parent[pick[0]] = (uint16_t)made;
            node_parent[pick[1]] = (uint16_t)made;
            ++made;
        }
        const int root = 2 * n - 2;
        node_depth[root] = 0;
        int bl_count[16];
        for (int b = 0; b < 16; ++b) bl_count[b] = 0;
        int overflow = 0;
        for (int i = root - 1; i >= 0; --i) {
            int d = node_depth[node_parent[i]] + 1;
            if (i < n) {
                if (d > 15) { d = 15; ++overflow; }
                ++bl_count[d];
            } else if (d > 200) {
                d = 200;
            }
            node_depth[i] = (uint8_t)d;
        }
        // leaves deeper than 15 were put at 15, which over-subscribes the code: Kraft sum * 2^15 exceeds 2^15.  As
        // zlib's gen_bitlen does, a leaf of depth b < 15 is replaced by an inner node whose two children are that leaf
        // and one of the leaves at 15: -2^(15-b) + 2 * 2^(14-b) - 1, one unit less per step, until the code is complete
        if (overflow > 0) {
            int excess = -(1 << 15);
            for (int bits = 1; bits <= 15; ++bits) excess += bl_count[bits] << (15 - bits);
            while (excess > 0) {
                int bits = 14;
                while (bl_count[bits] == 0) --bits;
                --bl_count[bits];
                bl_count[bits + 1] += 2;
                --bl_count[15];
                --excess;
            }
        }
        // lengths by rank: the rarest symbols get the longest codes
        {
            int i = 0;
            for (int bits = 15; bits >= 1; --bits)
                for (int k = 0; k < bl_count[bits]; ++k) node_depth[i++] = (uint8_t)bits;
        }
        // canonical codes in symbol order (RFC 1951 3.2.2)
        for (int j = 0; j < 257; ++j) code[j] = 0;
        for (int i = 0; i < n; ++i) code[order[i]] = (uint32_t)node_depth[i] << 16;
        int count[16];
        for (int b = 0; b < 16; ++b) count[b] = 0;
        for (int j = 0; j < 257; ++j) ++count[code[j] >> 16];
        count[0] = 0;
        uint32_t next_code[16];
        uint32_t c = 0;
        for (int b = 1; b <= 15; ++b) {
            c = (c + (uint32_t)count[b - 1]) << 1;
            next_code[b] = c;
        }
        uint32_t bits_total = 0;
        for (int j = 0; j < 257; ++j) {
            const uint32_t l = code[j] >> 16;
            if (!l) continue;
            code[j] |= __brev(next_code[l]++) >> (32 - l);
            bits_total += l * freq[j];
        }
        s_misc[1] = bits_total;
        // BFINAL, BTYPE | HLIT, HDIST, HCLEN | 19 code length code lengths | 257 + 2 code lengths of 4 bits
        const uint32_t dyn_bits = 3 + 14 + 57 + 259 * 4 + bits_total;
        s_misc[2] = (dyn_bits + 7) / 8 >= len + 5 ? 1u : 0u;
    }
    __syncthreads();
    uint8_t *const out = blocks + chunk * BGZF_SLOT;
    // the payload starts at byte 18 of the block: bits are ORed into the words from byte 16 on, 16 bits in
    uint32_t *const words = (uint32_t *)(out + 16);
    const uint32_t header_bits = 16 + 3 + 14 + 57 + 259 * 4;
    const bool stored = s_misc[2] != 0;
    // ---- CRC32 of the data (a piece per thread, combined pairwise) and the bit offset of every piece ----
    uint32_t crc = 0xFFFFFFFFu;
    const uint32_t p_lo = min(len, (uint32_t)tid * DEFLATE_PIECE), p_hi = min(len, p_lo + DEFLATE_PIECE);
    uint32_t my_bits = 0;
    for (uint32_t k = p_lo; k < p_hi; ++k) {
        const uint8_t v = data[k];
        crc = t_crc[(crc ^ v) & 0xFF] ^ (crc >> 8);
        my_bits += code[v] >> 16;
    }
    crc ^= 0xFFFFFFFFu;
    uint32_t piece_len = p_hi - p_lo;
    uint32_t inc = my_bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scan_tmp[warp] = inc;
    for (int step = 1; step < 32; step <<= 1) {
        const uint32_t oc = __shfl_down_sync(0xffffffffu, crc, step), ol = __shfl_down_sync(0xffffffffu, piece_len, step);
        if (!(lane & (2 * step - 1))) {
            if (ol) crc = crc_mulmod(crc_x_pow_bytes(ol, x2n), crc) ^ oc;
            piece_len += ol;
        }
    }
    if (lane == 0) { w_crc[warp] = crc; w_len[warp] = piece_len; }
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += scan_tmp[w];
    const uint32_t bit_off = before + inc - my_bits;
    if (tid == 32) {
        uint32_t c = w_crc[0];
        for (int w = 1; w < DEFLATE_THREADS / 32; ++w)
            if (w_len[w]) c = crc_mulmod(crc_x_pow_bytes(w_len[w], x2n), c) ^ w_crc[w];
        s_misc[3] = c;
    }
    // ---- the DEFLATE stream ----
    uint32_t payload;
    if (stored) {
        // BFINAL = 1, BTYPE = 00, then LEN / NLEN and the bytes
        uint8_t *p = out + 18;
        if (tid == 0) {
            p[0] = 1;
            p[1] = (uint8_t)len; p[2] = (uint8_t)(len >> 8);
            p[3] = (uint8_t)~len; p[4] = (uint8_t)(~len >> 8);
        }
        for (uint32_t i = tid; i < len; i += DEFLATE_THREADS) p[5 + i] = data[i];
        payload = 5 + len;
    } else {
        const uint32_t eob_len = code[256] >> 16;
        if (tid == 0) {
            BitWriter bw;
            bw.start(words, 16);
            bw.put(1, 1);        // BFINAL
            bw.put(2, 2);        // BTYPE = dynamic
            bw.put(0, 5);        // HLIT: 257 codes
            bw.put(1, 5);        // HDIST: 2 codes
            bw.put(15, 4);       // HCLEN: 19 code length codes
            // code length alphabet: symbols 0..15 get 4 bits each, 16..18 none; sent in the order 16 17 18 0 8 7 9 ...
            const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            for (int i = 0; i < 19; ++i) bw.put(perm[i] < 16 ? 4u : 0u, 3);
            // the 4-bit code of length l is l itself, most significant bit first
            for (int j = 0; j < 257; ++j) bw.put(__brev(code[j] >> 16) >> 28, 4);
            bw.put(__brev(1u) >> 28, 4);  // two distance codes of one bit each, never used (as zlib sends them)
            bw.put(__brev(1u) >> 28, 4);
            bw.finish();
        } else if (tid == 32) {
            // the end-of-block symbol, behind the last literal
            BitWriter bw;
            bw.start(words, (uint64_t)header_bits + s_misc[1] - eob_len);
            bw.put(code[256] & 0xFFFFu, (int)eob_len);
            bw.finish();
        }
        BitWriter bw;
        bw.start(words, (uint64_t)header_bits + bit_off);
        for (uint32_t k = p_lo; k < p_hi; ++k) {
            const uint32_t c = code[data[k]];
            bw.put(c & 0xFFFFu, (int)(c >> 16));
        }
        bw.finish();
        payload = (header_bits - 16 + s_misc[1] + 7) / 8;
    }
    __syncthreads();
    if (tid == 0) {
        // BGZF / gzip member header (SAM specification 4.1)
        const uint32_t block_size = 18 + payload + 8;
        const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
        for (int i = 0; i < 16; ++i) out[i] = head[i];
        out[16] = (uint8_t)(block_size - 1);
        out[17] = (uint8_t)((block_size - 1) >> 8);
        uint8_t *tail = out + 18 + payload;
        const uint32_t c = s_misc[3];
        tail[0] = (uint8_t)c; tail[1] = (uint8_t)(c >> 8); tail[2] = (uint8_t)(c >> 16); tail[3] = (uint8_t)(c >> 24);
        tail[4] = (uint8_t)len; tail[5] = (uint8_t)(len >> 8); tail[6] = (uint8_t)(len >> 16); tail[7] = (uint8_t)(len >> 24);
        sizes[chunk] = block_size;
    }
}

// packs the encoded blocks back to back: block c goes to offset prefix[c]
__global__ void __launch_bounds__(256) bgzf_pack_kernel(const uint8_t *__restrict__ blocks, const uint32_t *__restrict__ sizes,
                                                        const unsigned long long *__restrict__ prefix, uint8_t *__restrict__ out)
{
    const uint64_t c = blockIdx.x;
    const uint32_t size = sizes[c];
    const uint8_t *from = blocks + c * BGZF_SLOT;
    uint8_t *to = out + prefix[c];
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) to[i] = from[i];
}

// exclusive scan of the block sizes (one thread block; tens of thousands of entries); prefix[n] = total
__global__ void __launch_bounds__(1024) bgzf_prefix_kernel(const uint32_t *__restrict__ sizes, int64_t n, unsigned long long *prefix)
{
    __shared__ unsigned long long carry, warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t k = base + threadIdx.x;
        const unsigned long long v = k < n ? sizes[k] : 0;
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        unsigned long long before = carry;
        for (int w = 0; w < warp; ++w) before += warp_sums[w];
        if (k < n) prefix[k] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[n] = carry;
}

}  // namespace mdg

extern "C" {

// Re-emits the records of the batch handed out last (every record of the slab, in input order: the stream must have
// been opened with drop_flags = 0) as BGZF blocks made on the GPU, and appends them to `writer`.  Where
// mdg_rescale_resident marked a record, it carries the rewritten qualities and an MR:f tag (rescale.py:273-280,344).
// The file write of this batch overlaps the GPU work of the next one; mdg_bam_encode_flush waits for it.
int mdg_bam_encode_batch(mdg_bam_stream *s, mdg_dev_batch *batch, mdg_bam_writer *writer)
{
    if (!s || !batch || !writer) return MDG_ERR_ARGUMENT;
    if (s->handed < 0 || batch != &s->decoded[s->handed].batch) return sfail(s, MDG_ERR_ARGUMENT, "mdg_bam_encode_batch: not the batch handed out last");
    if (s->drop_flags) return sfail(s, MDG_ERR_STATE, "mdg_bam_encode_batch: the stream drops records (drop_flags != 0)");
    BamDevDecoded &dec = s->decoded[s->handed];
    const int64_t n = dec.n;
    if (n == 0) return MDG_OK;
    const double t0 = stream_now();
#define MDG_E_CUDA(call)                                                                                   \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return sfail(s, MDG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
    MDG_E_CUDA(cudaSetDevice(s->ctx->cfg.device));
    cudaStream_t st = s->ctx->compute;  // behind the rescale kernels of this batch
    if (!s->e_attr_set) {
        MDG_E_CUDA(cudaFuncSetAttribute(mdg::bgzf_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mdg::BGZF_CHUNK));
        MDG_E_CUDA(cudaHostAlloc((void **)&s->e_total_host, 16, cudaHostAllocDefault));
        s->e_attr_set = true;
    }
    const uint8_t *status = batch->res_cap >= n ? batch->res_status : nullptr;
    const float *mr = batch->res_cap >= n ? batch->res_mr : nullptr;
    const int64_t n_blocks = (n + 255) / 256;
    if (!bamdev_grow(s->e_out_off, s->e_out_off_cap, (size_t)n) || !bamdev_grow(s->e_scan, s->e_scan_cap, (size_t)(2 * n_blocks + 4)))
        return sfail(s, MDG_ERR_CUDA, "out of device memory (encoder offsets)");
    mdg::bam_emit_totals<<<(unsigned)n_blocks, 256, 0, st>>>(dec.stream, dec.rec_off, status, n, s->e_scan);
    mdg::synth_scan_totals<<<1, 1024, 0, st>>>(s->e_scan, n_blocks);
    mdg::bam_emit_offsets<<<(unsigned)n_blocks, 256, 0, st>>>(dec.stream, dec.rec_off, status, n, s->e_scan, s->e_out_off);
    MDG_E_CUDA(cudaGetLastError());
    MDG_E_CUDA(cudaMemcpyAsync(s->e_total_host, s->e_scan + 2 * n_blocks, 8, cudaMemcpyDeviceToHost, st));
    MDG_E_CUDA(cudaStreamSynchronize(st));
    const uint64_t total = *s->e_total_host;
    const int64_t n_chunks = (int64_t)((total + mdg::BGZF_CHUNK - 1) / mdg::BGZF_CHUNK);
    if (!bamdev_grow(s->e_records, s->e_records_cap, (size_t)total + 64) ||
        !bamdev_grow(s->e_blocks, s->e_blocks_cap, (size_t)n_chunks * mdg::BGZF_SLOT) ||
        !bamdev_grow(s->e_packed, s->e_packed_cap, (size_t)n_chunks * 65536) ||
        !bamdev_grow(s->e_sizes, s->e_sizes_cap, (size_t)n_chunks + 1) || !bamdev_grow(s->e_prefix, s->e_prefix_cap, (size_t)n_chunks + 2))
        return sfail(s, MDG_ERR_CUDA, "out of device memory (encoder buffers)");
    const mdg::DevBatch &v = dec.arrays.view;
    mdg::BamEmit emit{dec.rec_off, s->e_out_off, v.l_seq, v.base_off, v.qual, status, mr};
    const int egrid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)s->ctx->sm_count * 16);
    mdg::bam_emit_kernel<<<egrid, 256, 0, st>>>(dec.stream, n, emit, s->e_records);
    MDG_E_CUDA(cudaMemsetAsync(s->e_blocks, 0, (size_t)n_chunks * mdg::BGZF_SLOT, st));
    mdg::bgzf_deflate_kernel<<<(unsigned)n_chunks, mdg::DEFLATE_THREADS, mdg::BGZF_CHUNK, st>>>(s->e_records, total, s->e_blocks, s->e_sizes,
                                                                                                s->d_crc_tables, s->d_crc_tables + 256);
    mdg::bgzf_prefix_kernel<<<1, 1024, 0, st>>>(s->e_sizes, n_chunks, s->e_prefix);
    mdg::bgzf_pack_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(s->e_blocks, s->e_sizes, s->e_prefix, s->e_packed);
    MDG_E_CUDA(cudaGetLastError());
    MDG_E_CUDA(cudaMemcpyAsync(s->e_total_host + 1, s->e_prefix + n_chunks, 8, cudaMemcpyDeviceToHost, st));
    MDG_E_CUDA(cudaStreamSynchronize(st));
    s->ctx->launches += 7;
    const uint64_t packed = s->e_total_host[1];
    const int turn = s->e_turn;
    if (s->e_host_cap[turn] < packed) {
        pinned_pool().give(s->e_host[turn], s->e_host_cap[turn]);
        s->e_host_cap[turn] = 0;
        s->e_host[turn] = pinned_pool().take((size_t)packed * 9 / 8 + (1u << 20), &s->e_host_cap[turn]);
        if (!s->e_host[turn]) return sfail(s, MDG_ERR_CUDA, "cudaHostAlloc failed (encoder output)");
    }
    MDG_E_CUDA(cudaMemcpyAsync(s->e_host[turn], s->e_packed, (size_t)packed, cudaMemcpyDeviceToHost, st));
    MDG_E_CUDA(cudaStreamSynchronize(st));
#undef MDG_E_CUDA
    s->e_bytes_in += (int64_t)total;
    s->e_bytes_out += (int64_t)packed;
    s->t_encode += stream_now() - t0;
    // the previous batch's blocks must be in the file before these
    const double t1 = stream_now();
    if (s->e_writer.joinable()) s->e_writer.join();
    s->t_write_wait += stream_now() - t1;
    if (s->e_write_rc) return sfail(s, s->e_write_rc, "writing the output BAM failed");
    const uint8_t *data = s->e_host[turn];
    s->e_writer = std::thread([s, writer, data, packed] { s->e_write_rc = mdg_bam_write_raw(writer, data, (int64_t)packed); });
    s->e_turn = turn ^ 1;
    return MDG_OK;
}

// Waits for the file write mdg_bam_encode_batch left running; call before mdg_bam_finish.
int mdg_bam_encode_flush(mdg_bam_stream *s, int64_t *bytes_in, int64_t *bytes_out, double *seconds2)
{
    if (!s) return MDG_ERR_ARGUMENT;
    if (s->e_writer.joinable()) s->e_writer.join();
    if (bytes_in) *bytes_in = s->e_bytes_in;
    if (bytes_out) *bytes_out = s->e_bytes_out;
    if (seconds2) {
        seconds2[0] = s->t_encode;
        seconds2[1] = s->t_write_wait;
    }
    if (s->e_write_rc) return sfail(s, s->e_write_rc, "writing the output BAM failed");
    return MDG_OK;
}

}  // extern "C"
