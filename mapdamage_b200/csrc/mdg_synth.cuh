// Synthetic aDNA alignments generated directly in HBM (benchmark / test input).
//
// Implements the synthetic inputs of SURVEY.md section 8(d): reads drawn from
// the uploaded genome with post-mortem damage in BAM orientation (C->T with
// p = damage0 * decay^i at distance i from the left end of the alignment, G->A
// mirrored from the right end), a uniform sequencing-error rate, and a CIGAR
// mix of plain matches, one 1-3 bp insertion, one 1-3 bp deletion, or 0-10 bp
// soft clips; optionally inward-facing proper pairs (flags 99/147, 163/83).
// Everything is a pure function of (seed, read index), so any sub-range can be
// regenerated; the batch lands in the same SoA layout mdg_batch describes.
// This is input preparation, not part of the counting path.
#pragma once
#include "../../include/mapdamage_b200.h"
#include "mdg_device.cuh"

namespace mdg {

constexpr int SYNTH_DAMAGE_REACH = 24;

struct SynthDev {
    uint64_t seed;
    int64_t n_reads;
    int32_t len_lo, len_hi;
    uint32_t mix_cum[4];  // cumulative weights of plain / insertion / deletion / clipped
    int32_t paired, with_qual, n_lib;
    uint32_t error_u24, read_n_u24, filtered_u24;  // thresholds on a 24-bit uniform
    float damage0, decay;
    uint64_t genome_bases;  // sum of contig lengths
    int32_t sorted;         // 1: positions grow with the read index (a coordinate-sorted file), else uniform at random
};

struct SynthOut {
    uint16_t *flag;
    int32_t *tid, *pos;
    uint16_t *lib;
    uint32_t *l_seq, *base_off, *cigar_off, *cigar;
    uint8_t *seq4, *qual;
    int32_t *tlen, *mtid, *mpos;
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// CIGAR shape of one read: [s1 S] a M [k I|D] b M [s2 S]
struct ReadShape {
    int32_t l_seq, kind, k, s1, s2, a, nq, rspan, n_ops;
};

__device__ __forceinline__ ReadShape shape_of(const SynthDev &p, int64_t i)
{
    uint64_t h = mix64(p.seed ^ mix64((uint64_t)i * 2 + 1));
    ReadShape s;
    s.l_seq = p.len_lo + (int32_t)((h & 0xFFFF) * (uint64_t)(p.len_hi - p.len_lo + 1) >> 16);
    uint32_t pick = (uint32_t)((h >> 16) & 0xFFFF) * p.mix_cum[3] >> 16;
    s.kind = pick < p.mix_cum[0] ? 0 : pick < p.mix_cum[1] ? 1 : pick < p.mix_cum[2] ? 2 : 3;
    if (s.l_seq < 16 && (s.kind == 1 || s.kind == 2)) s.kind = 0;  // no room for an indel with 5 bp anchors
    s.k = (s.kind == 1 || s.kind == 2) ? 1 + (int32_t)((h >> 32) & 0xFF) * 3 / 256 : 0;
    s.s1 = s.s2 = 0;
    if (s.kind == 3) {
        s.s1 = (int32_t)((h >> 40) & 0xFF) * 11 / 256;
        s.s2 = (int32_t)((h >> 48) & 0xFF) * 11 / 256;
        if (s.s1 == 0 && s.s2 == 0) s.s1 = 1;
        if (s.l_seq - s.s1 - s.s2 < 10) {
            s.s1 = s.l_seq > 10 ? 1 : 0;
            s.s2 = 0;
        }
    }
    s.nq = s.l_seq - s.s1 - s.s2;
    if (s.kind == 1 || s.kind == 2) {
        int32_t span = s.kind == 1 ? s.nq - s.k - 10 : s.nq - 10;
        s.a = 5 + (int32_t)(((h >> 56) & 0xFF) * (uint64_t)(span > 1 ? span : 1) >> 8);
    } else {
        s.a = s.nq;
    }
    s.rspan = s.nq - (s.kind == 1 ? s.k : 0) + (s.kind == 2 ? s.k : 0);
    s.n_ops = s.kind == 0 ? 1 : (s.kind == 3 ? 1 + (s.s1 > 0) + (s.s2 > 0) : 3);
    return s;
}

__device__ __forceinline__ uint2 block_exclusive_scan2(uint32_t a, uint32_t b, uint2 *total)
{
    __shared__ uint32_t wa[32], wb[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint32_t ia = warp_inclusive_scan(a, lane), ib = warp_inclusive_scan(b, lane);
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        uint32_t va = lane < n_warps ? wa[lane] : 0, vb = lane < n_warps ? wb[lane] : 0;
        uint32_t sa = warp_inclusive_scan(va, lane), sb = warp_inclusive_scan(vb, lane);
        wa[lane] = sa - va;
        wb[lane] = sb - vb;
        if (lane == 31 && total) { total->x = sa; total->y = sb; }
    }
    __syncthreads();
    uint2 out = make_uint2(wa[warp] + ia - a, wb[warp] + ib - b);
    __syncthreads();
    return out;
}

// per-block totals of (padded bases, CIGAR ops); blockDim = 256
__global__ void __launch_bounds__(256) synth_block_totals(SynthDev p, unsigned long long *totals)
{
    __shared__ uint2 total;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bases = 0, ops = 0;
    if (i < p.n_reads) {
        ReadShape s = shape_of(p, i);
        bases = (uint32_t)(s.l_seq + 1) & ~1u;
        ops = (uint32_t)s.n_ops;
    }
    block_exclusive_scan2(bases, ops, &total);
    if (threadIdx.x == 0) {
        totals[2 * (size_t)blockIdx.x] = total.x;
        totals[2 * (size_t)blockIdx.x + 1] = total.y;
    }
}

// exclusive scan of the block totals in place; entry n_blocks receives the grand totals. One block of 1024.
__global__ void __launch_bounds__(1024) synth_scan_totals(unsigned long long *totals, int64_t n_blocks)
{
    __shared__ unsigned long long carry[2], warp_sums[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t k = base + threadIdx.x;
        unsigned long long v[2], inc[2];
        for (int c = 0; c < 2; ++c) {
            v[c] = k < n_blocks ? totals[2 * k + c] : 0;
            unsigned long long x = v[c];
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += t;
            }
            inc[c] = x;
            if (lane == 31) warp_sums[c][warp] = x;
        }
        __syncthreads();
        unsigned long long before[2] = {carry[0], carry[1]};
        for (int c = 0; c < 2; ++c)
            for (int w = 0; w < warp; ++w) before[c] += warp_sums[c][w];
        if (k < n_blocks)
            for (int c = 0; c < 2; ++c) totals[2 * k + c] = before[c] + inc[c] - v[c];
        __syncthreads();
        if (threadIdx.x == 1023)
            for (int c = 0; c < 2; ++c) carry[c] = before[c] + inc[c];
        __syncthreads();
    }
    if (threadIdx.x < 2) totals[2 * n_blocks + threadIdx.x] = carry[threadIdx.x];
}

// ---- default layout of a batch passed without base_off / cigar_off (mdg_batch: optional arrays) ----
// base_off[i] = sum over k < i of l_seq[k] rounded up to even; cigar_off[i] = i (one op per read)
__global__ void __launch_bounds__(256) layout_block_totals(const uint32_t *__restrict__ l_seq, int64_t n, unsigned long long *totals)
{
    __shared__ uint2 total;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t bases = i < n ? (l_seq[i] + 1) & ~1u : 0;
    block_exclusive_scan2(bases, 0, &total);
    if (threadIdx.x == 0) {
        totals[2 * (size_t)blockIdx.x] = total.x;
        totals[2 * (size_t)blockIdx.x + 1] = 0;
    }
}

__global__ void __launch_bounds__(256) layout_fill(const uint32_t *__restrict__ l_seq, int64_t n, const unsigned long long *block_off,
                                                   uint32_t *base_off, uint32_t *cigar_off, uint64_t n_bases, int32_t *error_flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t len = i < n ? l_seq[i] : 0;
    const uint2 off = block_exclusive_scan2((len + 1) & ~1u, 0, nullptr);
    if (i >= n) return;
    if (cigar_off) {
        cigar_off[i] = (uint32_t)i;
        if (i == n - 1) cigar_off[n] = (uint32_t)n;
    }
    if (base_off) {
        uint64_t at = block_off[2 * (size_t)blockIdx.x] + off.x;
        if (at + len > n_bases) {
            atomicCAS(error_flag, 0, DATA_ERR_LAYOUT);
            at = 0;
        }
        base_off[i] = (uint32_t)at;
    }
}

// a random contig straight into the one-hot genome image: base k of the contig is code mix64(seed, k / 32) bits
__global__ void __launch_bounds__(256) synth_reference_kernel(uint32_t *words, uint64_t n_bases, uint64_t seed)
{
    const uint64_t n_words = (n_bases + 7) / 8;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix64(seed ^ mix64(w >> 2)) >> (16 * (w & 3));
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (8 * w + k < n_bases) out |= (1u << ((h >> (2 * k)) & 3)) << (4 * k);
        words[w] = out;
    }
}

__global__ void __launch_bounds__(256) fill_words(uint32_t *to, int64_t n, uint32_t word)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) to[i] = word;
}

// l_seq of a batch passed without it: the bases its CIGAR consumes from the read (M, I, S, =, X).
// cigar_off == null: one op per read.
__global__ void __launch_bounds__(256) lseq_from_cigar(const uint32_t *__restrict__ cigar, const uint32_t *__restrict__ cigar_off,
                                                       int64_t n, uint32_t *l_seq)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c0 = cigar_off ? cigar_off[i] : (uint32_t)i, c1 = cigar_off ? cigar_off[i + 1] : (uint32_t)i + 1;
    uint32_t len = 0;
    for (uint32_t k = c0; k < c1; ++k) {
        const uint32_t w = cigar[k], op = w & 0xF;
        if (op == OP_M || op == OP_I || op == OP_S || op == OP_EQ || op == OP_X) len += w >> 4;
    }
    l_seq[i] = len;
}

__device__ __forceinline__ void place_read(const SynthDev &p, const DevRef &ref, uint64_t h, int64_t i, int32_t rspan,
                                           int32_t *tid, int64_t *pos)
{
    // contig chosen in proportion to its length: a uniform base of the genome (or, sorted, base i / n of it)
    uint64_t g = p.sorted ? (uint64_t)(((unsigned __int128)(uint64_t)i * p.genome_bases) / (uint64_t)p.n_reads)
                          : (uint64_t)(((unsigned __int128)h * p.genome_bases) >> 64);
    int c = 0;
    while (c + 1 < ref.n_contigs && g >= ref.contig_len[c]) {
        g -= ref.contig_len[c];
        ++c;
    }
    // sorted: one scale for every read of the contig (the longest reference span a shape can have is len_hi + 3,
    // a 3-base deletion), else the order would break near the contig end where the room depends on the read
    int64_t room = (int64_t)ref.contig_len[c] - (p.sorted ? p.len_hi + 3 : rspan);
    if (room < 0) room = 0;
    *tid = c;
    *pos = (int64_t)(g * (uint64_t)(room + 1) / ref.contig_len[c]);
    if (*pos > room) *pos = room;
}

// one thread per read; blockDim = 256
__global__ void __launch_bounds__(256) synth_fill(SynthDev p, DevRef ref, const unsigned long long *block_off, SynthOut o)
{
    __shared__ uint32_t damage_u24[SYNTH_DAMAGE_REACH + 1];
    if (threadIdx.x <= SYNTH_DAMAGE_REACH) {
        float pr = threadIdx.x < SYNTH_DAMAGE_REACH ? p.damage0 * powf(p.decay, (float)threadIdx.x) : 0.f;
        damage_u24[threadIdx.x] = (uint32_t)(pr * 16777216.f);
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < p.n_reads;
    ReadShape s{};
    if (live) s = shape_of(p, i);
    uint2 off = block_exclusive_scan2(live ? (uint32_t)(s.l_seq + 1) & ~1u : 0, live ? (uint32_t)s.n_ops : 0, nullptr);
    const uint64_t boff = block_off[2 * (size_t)blockIdx.x] + off.x;
    const uint64_t coff = block_off[2 * (size_t)blockIdx.x + 1] + off.y;
    if (!live) return;  // no barrier below this point
    if (i == p.n_reads - 1) o.cigar_off[i + 1] = (uint32_t)(coff + s.n_ops);

    int32_t tid;
    int64_t pos;
    uint32_t flag = 0;
    int32_t tlen = 0, mtid = -1, mpos = -1;
    const uint64_t hr = mix64(p.seed ^ mix64((uint64_t)i * 2));
    if (p.paired) {
        // records 2q / 2q+1 are mates on one contig: leftmost forward, rightmost reverse
        const int64_t left_i = i & ~1ll, right_i = left_i + 1;
        const bool is_left = i == left_i;
        ReadShape sl = is_left ? s : shape_of(p, left_i);
        ReadShape sr = right_i < p.n_reads ? (is_left ? shape_of(p, right_i) : s) : sl;
        const uint64_t hp = mix64(p.seed ^ mix64((uint64_t)left_i * 2));
        int64_t pos_l;
        place_read(p, ref, hp, left_i, sl.rspan, &tid, &pos_l);
        const int64_t gap = (int64_t)((mix64(hp) & 0xFFFF) * 301 >> 16);
        int64_t room_r = (int64_t)ref.contig_len[tid] - sr.rspan;
        if (room_r < 0) room_r = 0;
        int64_t pos_r = pos_l + gap < room_r ? pos_l + gap : room_r;
        if (pos_l > pos_r) pos_l = pos_r;
        const bool first_left = (mix64(hp) >> 16) & 1;
        const int64_t end_l = pos_l + sl.rspan, end_r = pos_r + sr.rspan;
        const int64_t frag = (end_r > end_l ? end_r : end_l) - pos_l;
        if (is_left) {
            flag = first_left ? 99 : 163;
            pos = pos_l;
            tlen = (int32_t)frag;
            mpos = (int32_t)pos_r;
        } else {
            flag = first_left ? 147 : 83;
            pos = pos_r;
            tlen = (int32_t)-frag;
            mpos = (int32_t)pos_l;
        }
        mtid = tid;
    } else {
        place_read(p, ref, hr, i, s.rspan, &tid, &pos);
        if ((mix64(hr) >> 8) & 1) flag = 16;
    }
    const uint64_t hx = mix64(hr ^ 0xA5A5A5A5ull);
    if ((uint32_t)(hx & 0xFFFFFF) < p.filtered_u24) flag |= 0x100u << ((hx >> 24) & 3);

    o.flag[i] = (uint16_t)flag;
    o.tid[i] = tid;
    o.pos[i] = (int32_t)pos;
    o.lib[i] = p.n_lib > 1 ? (uint16_t)((hx >> 32) % (uint32_t)p.n_lib) : 0;
    o.l_seq[i] = (uint32_t)s.l_seq;
    o.base_off[i] = (uint32_t)boff;
    o.cigar_off[i] = (uint32_t)coff;
    o.tlen[i] = tlen;
    o.mtid[i] = mtid;
    o.mpos[i] = mpos;
    uint32_t *cig = o.cigar + coff;
    if (s.kind == 0) {
        cig[0] = ((uint32_t)s.nq << 4) | OP_M;
    } else if (s.kind == 3) {
        int w = 0;
        if (s.s1 > 0) cig[w++] = ((uint32_t)s.s1 << 4) | OP_S;
        cig[w++] = ((uint32_t)s.nq << 4) | OP_M;
        if (s.s2 > 0) cig[w++] = ((uint32_t)s.s2 << 4) | OP_S;
    } else {
        const int32_t b = s.nq - s.a - (s.kind == 1 ? s.k : 0);
        cig[0] = ((uint32_t)s.a << 4) | OP_M;
        cig[1] = ((uint32_t)s.k << 4) | (s.kind == 1 ? OP_I : OP_D);
        cig[2] = ((uint32_t)b << 4) | OP_M;
    }

    // damage_u24 is visible here: the block scan above contains barriers
    const uint64_t contig_off = ref.contig_off[tid];
    const uint64_t key = mix64(p.seed ^ ((uint64_t)i << 20) ^ 0x5EEDull);
    uint8_t *seq = o.seq4 + boff / 2;
    uint8_t *qual = p.with_qual ? o.qual + boff : nullptr;
    uint32_t byte = 0;
    for (int32_t j = 0; j < s.l_seq; ++j) {
        const uint64_t h = mix64(key + (uint64_t)j);
        const int32_t jq = j - s.s1;
        bool aligned = jq >= 0 && jq < s.nq;
        int32_t shift = 0;
        if (s.kind == 1) {
            if (jq >= s.a && jq < s.a + s.k) aligned = false;
            else if (jq >= s.a + s.k) shift = -s.k;
        } else if (s.kind == 2 && jq >= s.a) {
            shift = s.k;
        }
        uint32_t base = (uint32_t)(h & 3);
        if (aligned) {
            uint32_t g = ref_code(ref.words, contig_off + (uint64_t)(pos + jq + shift));
            if (g < 4) base = g;
            const uint32_t u = (uint32_t)(h >> 8) & 0xFFFFFF;
            const int32_t d5 = jq < SYNTH_DAMAGE_REACH ? jq : SYNTH_DAMAGE_REACH;
            const int32_t r3 = s.nq - 1 - jq;
            const int32_t d3 = r3 < SYNTH_DAMAGE_REACH ? r3 : SYNTH_DAMAGE_REACH;
            if (base == 1 && u < damage_u24[d5]) base = 3;       // C -> T
            else if (base == 2 && u < damage_u24[d3]) base = 0;  // G -> A
        }
        const uint32_t e = (uint32_t)(h >> 32) & 0xFFFFFF;
        if (e < p.error_u24) base = (base + 1 + (uint32_t)((h >> 2) & 0x3F) * 3 / 64) & 3;
        uint32_t nib = 1u << base;
        if ((uint32_t)(h >> 40) < p.read_n_u24) nib = 15;
        if (j & 1) {
            seq[j >> 1] = (uint8_t)(byte | nib);
        } else {
            byte = nib << 4;
            if (j == s.l_seq - 1) seq[j >> 1] = (uint8_t)byte;
        }
        if (qual) qual[j] = (uint8_t)(2 + (uint32_t)((h >> 56) & 0xFF) * 39 / 256);
    }
    if (qual && (s.l_seq & 1)) qual[s.l_seq] = 0xFF;
}

}  // namespace mdg
