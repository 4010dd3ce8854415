// Counting pass, bit-plane kernel with specialised warps: producer teams parse and stage tiles, consumer warps count.
//
// Same contract and the same arithmetic as count_planes_kernel (mdg_planes.cuh: the loop body of main.py:165-217 for
// reads whose CIGAR is [H][S] M/=/X+ [S][H]; bit planes, vertical carry-save counters, substitutions as events), but the
// three block-wide phases of that kernel (parse | stage | count, a barrier after each) are taken apart:
//
//   * kTeams producer teams of kTeamWarps warps.  A team owns one stage buffer of T = 32 kTeamWarps rows and walks its
//     own tiles of T reads: parse (filter, CIGAR shape, per-read events; the read's row = its rank in its group
//     (library, strand), one warp-aggregated atomic), then -- once the consumers have released the buffer -- the plane
//     words of every read's window(s) into its row.  The thread that parses a read stages it: the record stays in
//     registers.  The tile's slices of the record arrays and its stretch of seq4 arrive by bulk asynchronous copies
//     (cp.async.bulk + mbarrier) issued a tile ahead.  Teams run out of step with each other.
//   * kConsWarps consumer warps hold ALL the counters (planes 0-7 in registers, 4-11 in shared memory).  They wait on a
//     buffer's `full` mbarrier, add its reads into the vertical counters (thread = read slot x window word x reference
//     base; a read slot belongs to one group), and arrive on its `empty` mbarrier.  They alone reduce counters into the
//     64-bit tables.
//
// full / empty are mbarriers with one arrival per thread of the side that signals (release / acquire at CTA scope order
// the staged words); producers synchronise among themselves with a named barrier per team (bar.sync id, T), consumers
// with their own.  There is no block-wide barrier between the first tile and the last.
//
// Substitution events go to a block-wide table indexed by TABLE cell ([library][anchor][strand][class][position]), so that
// teams working in different window layouts can share it; it is drained once, at the end.
//
// Variants (template parameters, chosen by the host): kNL = 2 counts two libraries in one launch; kGather
// has all of a window's genome loads in flight before its first word (genomes that do not fit L2); kIndels stages reads
// with one insertion / deletion itself (stage_indel_word); kQual applies the -Q mask.
#pragma once
#include "mdg_planes.cuh"

namespace mdg {

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity)
{
    // the hardware parks the warp until the phase completes (or the hint, in ns, runs out): a waiting warp takes no issue slots
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity), "r"(1000000u)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");  // ordered after the mbarrier wait it follows
    return v;
}
template <int kCount>
__device__ __forceinline__ void named_barrier(int id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kCount) : "memory");
}

// Planes p and p + 2 (p = 0, 1) of the eight bases of a seq4 word, each as eight bits in bits 24..31 of a word.
// BAM's nibble IS one-hot (A, C, G, T = 1, 2, 4, 8), so plane p of base j is bit p of nibble j: keep bits p and p + 2 of
// every nibble (one AND), put the two nibbles of a byte side by side in base order -- BAM stores the first base of a byte
// in the HIGH nibble -- (two shifts, one OR-AND: the byte then holds {base 2b, base 2b+1} of plane p in bits 0-1 and of
// plane p + 2 in bits 2-3), and gather the four bytes' bit pairs into the top byte with one multiplication per plane (the
// sixteen partial products land on distinct bits: no carries).  The shifts and multiplications issue on the FMA pipe.
__device__ __forceinline__ void ws_plane_bytes(uint32_t w, int p, uint32_t &lo_plane, uint32_t &hi_plane)
{
    const uint32_t y = (p ? shr_fma(w, p) : w) & 0x55555555u;
    const uint32_t t = shr_fma(y, 4) | shl_fma(y, 1);
    lo_plane = (t & 0x03030303u) * 0x01041040u;
    hi_plane = (t & 0x0C0C0C0Cu) * 0x00410410u;
}

// Words of a tile's control block.  Reads are grouped by (library, strand): group = strand + 2 * library.
constexpr int WS_CTL = 16;
constexpr int WS_MAX_LIB = 2;     // libraries counted in one launch (more: one launch per library over an index list)
constexpr int CTL_GROUP = 0;      // [4] gap-free reads of the tile per group
constexpr int CTL_MIN = 4, CTL_MAX = 5, CTL_MODE = 6;  // columns of the gap-free reads, window layout chosen
constexpr int CTL_CX = 8;         // [2] reads for the general kernel, per library
constexpr int CTL_IX = 10;        // [2] one-indel reads left to count_staged_kernel, per library (from the front of the list)
constexpr int CTL_IX4 = 12;       // [2] one-indel reads staged here when the tile has two windows per read, else handed over
                                  //     too (from the back of the list)
constexpr int WS_CAPACITY = 4000;  // reads a counter may see between two reductions (12 bits: planes 0-3 + wide 4-11)

// kGather: the stage issues all of a window's genome loads before it makes the first word (a copy of the word's code per
// window length) -- for genomes that do not fit L2, where the gathers are DRAM accesses; otherwise one rolled loop.
// kIndels: reads with one short insertion / deletion (no clips) are staged here too when their tile has two windows per
// read; otherwise all of them go to count_staged_kernel's list.  A variant of its own because the code it adds costs the
// others 3-5 % (instruction cache): the host switches to it once a batch has shown such reads.
// kQual: -Q (main.py:185-197, align.py:67-71): a column whose read base has a quality below p.min_qual counts for nothing in
// the misincorporation table (the read composition keeps it); reads without qualities are not masked.  One-indel reads
// then go to count_staged_kernel.
template <int kTeams, int kTeamWarps, int kConsWarps, int kNWA, int kNL, bool kGather, bool kIndels, bool kQual>
__global__ void __launch_bounds__((kTeams * kTeamWarps + kConsWarps) * 32, 1)
count_planes_ws_kernel(DevBatch b, DevRef ref, CountParams p, CountTables t, PlaneGeom g, uint32_t *__restrict__ worklist,
                       unsigned long long *__restrict__ work_count, uint32_t *__restrict__ indel_list,
                       unsigned long long *__restrict__ indel_count, SwarSubset sub)
{
    constexpr int T = kTeamWarps * 32;   // threads of a producer team = reads per tile
    constexpr int CT = kConsWarps * 32;  // consumer threads
    constexpr int PRODUCERS = kTeams * T;
    constexpr int NTHREADS = PRODUCERS + CT;
    static_assert(kConsWarps % 4 == 0, "consumer warps come in fours: one per reference base");
    static_assert(!(kQual && kIndels), "one-indel reads under a quality mask are count_staged_kernel's");
    extern __shared__ __align__(16) uint32_t smem[];
    const int L = p.L, A = p.A, LA = L + A;
    // words per anchor window ceil((L + A) / 32), per read, and per staged row (rows land on different banks): compile-time,
    // so that the address arithmetic folds
    constexpr int NWA = kNWA, WPR_MAX = 2 * NWA, ROW = 16 * NWA + 4;
    // ---- shared memory ----
    uint32_t *const s_wide = smem;                                    // [PL_WIDE][PL_CLASSES][CT]
    uint32_t *const s_red = s_wide + PL_WIDE * PL_CLASSES * CT;       // [strand][8 classes][32 WPR_MAX]: a reduction's sums
    // libraries counted by this launch: every read's own (sub.offsets without a list: the lists of left-over reads are kept
    // per library at those offsets), or one (the tables passed are that library's)
    // (kNL > 1 is launched like that only: without a list, with the offsets, p.n_lib == kNL)
    constexpr int NL = kNL, n_groups = 2 * NL;
    static_assert(kNL >= 1 && kNL <= WS_MAX_LIB, "libraries per launch");
    uint32_t *const s_sub = s_red + 2 * 8 * 32 * WPR_MAX;             // [lib][anchor][strand][12][L] substitution events
    uint32_t *const s_lg = s_sub + NL * 4 * 12 * L;                   // [lib][kind][strand][MDG_LG_SMEM_BINS]
    uint32_t *const s_clip = s_lg + NL * 4 * MDG_LG_SMEM_BINS;        // [lib][end][strand][L]
    uint32_t *const s_teams = s_clip + NL * 4 * L;  // every piece above is a multiple of four words: 16-byte aligned, and still a shared-memory pointer
    // per team: stage rows [T][ROW] (forward reads from the front, reverse from the back), two index lists [T], masks
    // [WPR_MAX][2], three control blocks, seq4 stretch, record arrays
    constexpr int SEQ_WORDS = T * 14;  // 56 bytes of seq4 per read: reads of up to about 110 bases on average
    // the record arrays of a tile (flag, lib: 16 bit; tid, pos, l_seq, base_off, cigar_off[T + 1]), bulk-copied a tile ahead
    constexpr int HDR_WORDS = T / 2 + T / 2 + 4 * T + (T + 4);
    constexpr int LISTS = 2 * WS_MAX_LIB * T;  // reads this kernel leaves to others: [general | one indel][lib][T]
    constexpr int team_words = T * ROW + LISTS + ((2 * WPR_MAX + 3) & ~3) + 3 * WS_CTL + SEQ_WORDS + HDR_WORDS;
    __shared__ __align__(8) unsigned long long s_full[kTeams], s_empty[kTeams], s_slab_bar[kTeams], s_hdr_bar[kTeams];
    __shared__ int32_t s_slab[kTeams][2];  // first seq4 word held in the team's copy (may be negative), words (0: no copy)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < PL_WIDE * PL_CLASSES * CT; i += NTHREADS) s_wide[i] = 0;
    for (int i = tid; i < NL * (4 * 12 * L + 4 * MDG_LG_SMEM_BINS + 4 * L); i += NTHREADS) s_sub[i] = 0;
    if (tid < kTeams) {
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_full[tid]), T);
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_empty[tid]), CT);
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_slab_bar[tid]), 1);
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_hdr_bar[tid]), 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        s_slab[tid][0] = 0;
        s_slab[tid][1] = 0;
    }
    for (int i = tid; i < kTeams * 3 * WS_CTL; i += NTHREADS) {
        const int team = i / (3 * WS_CTL), w = i % (3 * WS_CTL);
        (s_teams + (size_t)team * team_words + T * ROW + LISTS + ((2 * WPR_MAX + 3) & ~3))[w] = (w % WS_CTL) == CTL_MIN ? 0xffffffffu : 0u;
    }
    __syncthreads();

    const uint32_t *const subset = sub.list ? sub.list + sub.offsets[sub.lib] : nullptr;
    const int64_t n_todo = sub.list ? (int64_t)(sub.offsets[sub.lib + 1] - sub.offsets[sub.lib]) : b.n_reads;
    const int64_t n_tiles = (n_todo + T - 1) / T;
    constexpr int PAIRS = (CT >> 7) * 32;  // (slot, word) pairs per reference base on the consumer side
    auto words_of = [&](int columns) { return columns ? (columns + 2 * A + 31) / 32 : WPR_MAX; };
    // tile `k` of team `team` of this block
    auto tile_of = [&](int64_t k, int team) { return (k * gridDim.x + blockIdx.x) * kTeams + team; };

    if (warp < kTeams * kTeamWarps) {
        // =========================================== producer team ===========================================
        const int team = warp / kTeamWarps, ptid = tid - team * T, pwarp = ptid >> 5;
        uint32_t *const s_stage = s_teams + (size_t)team * team_words;
        uint32_t *const s_cx = s_stage + T * ROW;                       // [lib][T] reads for the general kernel
        uint32_t *const s_ix = s_cx + WS_MAX_LIB * T;                   // [lib][T] one-indel reads
        uint32_t *const s_mask = s_ix + WS_MAX_LIB * T;                 // [WPR_MAX][2] aligned / flank masks of a typical read
        uint32_t *const s_ctl_base = s_mask + ((2 * WPR_MAX + 3) & ~3);
        uint32_t *const s_seq = s_ctl_base + 3 * WS_CTL;
        const uint16_t *const s_hflag = (const uint16_t *)(s_seq + SEQ_WORDS), *const s_hlib = s_hflag + T;
        const int32_t *const s_htid = (const int32_t *)(s_hlib + T), *const s_hpos = s_htid + T;
        const uint32_t *const s_hlseq = (const uint32_t *)(s_hpos + T), *const s_hboff = s_hlseq + T, *const s_hcoff = s_hboff + T;
        const uint32_t hdr_addr = (uint32_t)__cvta_generic_to_shared(&s_hdr_bar[team]);
        uint32_t hdr_phase = 0;
        const uint32_t full_addr = (uint32_t)__cvta_generic_to_shared(&s_full[team]);
        const uint32_t empty_addr = (uint32_t)__cvta_generic_to_shared(&s_empty[team]);
        const uint32_t slab_addr = (uint32_t)__cvta_generic_to_shared(&s_slab_bar[team]);
        const bool staged_headers = !subset;  // tiles of an index list read their records from global memory
        const uint32_t *__restrict__ seq32 = (const uint32_t *)b.seq4;
        const uint4 *__restrict__ planes = ref.planes;
        uint32_t slab_phase = 0;
        int mode = -1;  // window layout the team's masks were made for

        auto fill_masks = [&](int columns) {
            const int wpr = words_of(columns);
            for (int w = ptid; w < wpr; w += T) {
                uint32_t aligned, flank;
                if (columns) {
                    aligned = bit_range(A - 32 * w, A + columns - 32 * w);
                    flank = bit_range(-32 * w, A - 32 * w) | bit_range(A + columns - 32 * w, 2 * A + columns - 32 * w);
                } else if (w < NWA) {
                    aligned = bit_range(A - 32 * w, A + L - 32 * w);
                    flank = bit_range(-32 * w, A - 32 * w);
                } else {
                    const int k = w - NWA, top = 32 * NWA;
                    aligned = bit_range(top - A - L - 32 * k, top - A - 32 * k);
                    flank = bit_range(top - A - 32 * k, top - 32 * k);
                }
                s_mask[2 * w] = aligned;
                s_mask[2 * w + 1] = flank;
            }
        };

        // ---- stage: the plane words of one window of one read (see count_planes_kernel::stage_window) ----
        auto stage_window = [&](auto nw_tag, const PlaneRecord &rec, uint32_t *row_at, int first_word, int n_words, int c_start, int side,
                                int slab_w0, int slab_words, int rstrand, int libx, uint32_t lead) {
            constexpr int kNW = decltype(nw_tag)::value;  // > 0: words of the window, known at compile time (kGather)
            const int cols = (int)(rec.cols & 0x7FFF);
            const int v = (int)(rec.misc & 0xFFFF);
            const int lf = (int)((rec.cols >> 16) & 0xFF), rf = (int)(rec.cols >> 24);
            const bool typical = lf == A && rf == A && (mode || v == L);
            const int64_t qn = (int64_t)rec.q0 + c_start;
            const int qs = (int)(qn & 7);
            const int64_t qw = qn >> 3, in_slab = qw - slab_w0;
            const bool from_smem = slab_words > 0 && in_slab >= 0 && in_slab + 4 * n_words + 1 <= slab_words;
            // (two loads in two address spaces: written as a value select the compiler makes it ONE generic load)
            const uint32_t qs_addr = (uint32_t)__cvta_generic_to_shared(s_seq) + 4u * (uint32_t)(from_smem ? in_slab : 0);
            const uint32_t *const qg_ptr = seq32 + qw;
            auto seq_words4 = [&](int m, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t &w4) {
                if (from_smem) {
                    w1 = lds_u32(qs_addr + 4 * m); w2 = lds_u32(qs_addr + 4 * m + 4); w3 = lds_u32(qs_addr + 4 * m + 8); w4 = lds_u32(qs_addr + 4 * m + 12);
                } else {
                    w1 = __ldg(qg_ptr + m); w2 = __ldg(qg_ptr + m + 1); w3 = __ldg(qg_ptr + m + 2); w4 = __ldg(qg_ptr + m + 3);
                }
            };
            auto seq_word = [&](int m) { return from_smem ? lds_u32(qs_addr + 4 * m) : __ldg(qg_ptr + m); };
            const int64_t rn = ((int64_t)rec.rg << 5) + (int)((rec.misc >> 16) & 31) + c_start;
            const uint4 *rp = planes + (rn >> 5);
            const int rs = (int)(rn & 31);
            uint32_t *const sub_at = s_sub + ((size_t)libx * 4 + rstrand) * 12 * L;
            // bit i of the word's quality mask: the read base of column c_start + 32 k + i has a quality of at least min_qual.
            // Quality bytes lie at the read's base indices; four at a time: a byte b < 128 is >= q  <=>  bit 7 of (b + 128 - q);
            // the top bit is added apart so that no byte carries into its neighbour (the pad byte behind an odd-length
            // read may hold anything, and sits right below the next read's first quality); the four bits are gathered by
            // a multiplication.
            const uint32_t q_bias = (uint32_t)(128 - min(p.min_qual, 127)) * 0x01010101u;
            bool masked = false;
            const uint32_t *qual_words = nullptr;
            int qual_shift = 0;
            if constexpr (kQual) {
                // (a read without qualities carries 0xFF in its first quality byte: not masked, main.py:185)
                masked = b.qual && __ldg(b.qual + ((int64_t)rec.q0 - (int64_t)lead)) != 0xFF;
                const int64_t at = (int64_t)rec.q0 + c_start;  // byte of the window's bit 0
                qual_words = (const uint32_t *)(b.qual + (at & ~3ll));
                qual_shift = (int)(at & 3) * 8;
            }
            auto quality_mask = [&](int k) {
                if (!masked) return 0xFFFFFFFFu;
                uint32_t w[9], bits = 0;
#pragma unroll
                for (int m = 0; m < 9; ++m) w[m] = __ldg(qual_words + 8 * k + m);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const uint32_t four = __funnelshift_r(w[m], w[m + 1], qual_shift);
                    const uint32_t ge = ((((four & 0x7F7F7F7Fu) + q_bias) | four) >> 7) & 0x01010101u;
                    bits |= ((ge * 0x00204081u >> 21) & 0xFu) << (4 * m);
                }
                return bits;
            };
            auto emit = [&](int k, uint32_t (&q0)[4], uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, const uint4 &g_lo, const uint4 &g_hi) {
                uint32_t xp[4];
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    uint32_t q1[2], q2[2], q3[2], q4[2];
                    ws_plane_bytes(w1, pp, q1[0], q1[1]);
                    ws_plane_bytes(w2, pp, q2[0], q2[1]);
                    ws_plane_bytes(w3, pp, q3[0], q3[1]);
                    ws_plane_bytes(w4, pp, q4[0], q4[1]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int pl = pp + 2 * h;
                        const uint32_t lo = __byte_perm(__byte_perm(q0[pl], q1[h], 0x0073), __byte_perm(q2[h], q3[h], 0x0073), 0x5410);
                        xp[pl] = __funnelshift_r(lo, q4[h] >> 24, qs);
                        q0[pl] = q4[h];
                    }
                }
                const uint32_t xa = xp[0], xc = xp[1], xg = xp[2], xt = xp[3];
                const uint32_t ya = __funnelshift_r(g_lo.x, g_hi.x, rs), yc = __funnelshift_r(g_lo.y, g_hi.y, rs);
                const uint32_t yg = __funnelshift_r(g_lo.z, g_hi.z, rs), yt = __funnelshift_r(g_lo.w, g_hi.w, rs);
                uint32_t aligned, flank;
                if (typical) {
                    aligned = s_mask[2 * (first_word + k)];
                    flank = s_mask[2 * (first_word + k) + 1];
                } else {
                    const int base = c_start + 32 * k;  // column of bit 0
                    if (mode) {
                        aligned = bit_range(-base, cols - base);
                        flank = bit_range(-lf - base, -base) | bit_range(cols - base, cols + rf - base);
                    } else if (side == 0) {
                        aligned = bit_range(-base, v - base);
                        flank = bit_range(-lf - base, -base);
                    } else {
                        aligned = bit_range(cols - v - base, cols - base);
                        flank = bit_range(cols - base, cols + rf - base);
                    }
                }
                // statistics.py:27: a column counts only when the read base is A/C/G/T; flank bits carry the reference base alone
                const uint32_t one = (xa ^ xc ^ xg ^ xt) & ~((xa & xc) | (xg & xt));
                const uint32_t keep = one & aligned, keep_y = (kQual ? keep & quality_mask(k) : keep) | flank;
                uint4 xs, ys;
                xs.x = xa & keep; xs.y = xc & keep; xs.z = xg & keep; xs.w = xt & keep;
                ys.x = ya & keep_y; ys.y = yc & keep_y; ys.z = yg & keep_y; ys.w = yt & keep_y;
                uint4 *out = (uint4 *)(row_at + 8 * (first_word + k));
                out[0] = xs;
                out[1] = ys;
                // substitutions (reference g read as another base): one shared-memory atomic per event, on the table cell
                uint32_t ev = (ys.x | ys.y | ys.z | ys.w) & keep & ~((xs.x & ys.x) | (xs.y & ys.y) | (xs.z & ys.z) | (xs.w & ys.w));
                if (ev) {
                    const uint32_t g1 = ys.y | ys.w, g2 = ys.z | ys.w, r1 = xs.y | xs.w, r2 = xs.z | xs.w;
                    const int col0 = c_start + 32 * k;  // column of bit 0 of this word
                    do {
                        const int j = __ffs(ev) - 1;
                        ev &= ev - 1;
                        const int gb = (int)((g1 >> j) & 1u) + 2 * (int)((g2 >> j) & 1u);
                        int rb = (int)((r1 >> j) & 1u) + 2 * (int)((r2 >> j) & 1u);
                        rb -= rb > gb ? 1 : 0;
                        uint32_t *const cell = sub_at + (3 * gb + rb) * L;
                        const int col = col0 + j, back = cols - 1 - col;  // distance from the left / right end of the alignment
                        if (mode) {
                            if (col < L) atomicAdd(cell + col, 1u);
                            if (back < L) atomicAdd(cell + 2 * 12 * L + back, 1u);
                        } else if (side == 0) {
                            atomicAdd(cell + col, 1u);
                        } else {
                            atomicAdd(cell + 2 * 12 * L + back, 1u);
                        }
                    } while (ev);
                }
            };
            if constexpr (kNW > 0) {
                uint4 gw[kNW + 1];
#pragma unroll
                for (int k = 0; k <= kNW; ++k) gw[k] = __ldg(rp + k);
                uint32_t q0[4];
                {
                    const uint32_t w0 = seq_word(0);
                    ws_plane_bytes(w0, 0, q0[0], q0[2]);
                    ws_plane_bytes(w0, 1, q0[1], q0[3]);
                }
#pragma unroll
                for (int k = 0; k < kNW; ++k) {
                    uint32_t w1, w2, w3, w4;
                    seq_words4(4 * k + 1, w1, w2, w3, w4);
                    emit(k, q0, w1, w2, w3, w4, gw[k], gw[k + 1]);
                }
            } else {
                uint32_t q0[4];
                {
                    const uint32_t w0 = seq_word(0);
                    ws_plane_bytes(w0, 0, q0[0], q0[2]);
                    ws_plane_bytes(w0, 1, q0[1], q0[3]);
                }
                // ONE copy of the word's code (unrolled per window length it is 7 % slower on a genome that sits in L2: the
                // two roles and two teams of an SM run different code at the same time and miss the instruction cache); the
                // genome entry after next is in flight while a word is made
                uint4 g_lo = __ldg(rp), g_hi = __ldg(rp + 1);
#pragma unroll 1
                for (int k = 0; k < n_words; ++k) {
                    uint32_t w1, w2, w3, w4;
                    seq_words4(4 * k + 1, w1, w2, w3, w4);
                    const uint4 g_next = __ldg(rp + min(k + 2, n_words));
                    emit(k, q0, w1, w2, w3, w4, g_lo, g_hi);
                    g_lo = g_hi;
                    g_hi = g_next;
                }
            }
        };

        // ---- stage: ONE window word of a read with one insertion / deletion (two windows per read) ----
        // Columns of the alignment are c = 0 .. cols - 1 with the gap at [a, a + k).  Exactly one of the two sequences is
        // discontinuous there: behind an insertion the reference lags (reference offset c - k for c >= a + k, nothing under
        // the inserted bases), behind a deletion the read does (read base c - k, nothing over the deleted bases).  So the
        // word is made from the planes of that sequence at TWO alignments, blended by column masks, and the other one's
        // at one.  Misincorporation positions are columns; read composition positions are read bases (statistics.py:76-83,
        // SURVEY N1), i.e. the read's planes at the alignment of the window's anchor side.  Gap columns are events
        // (g>- for every deleted reference base, ->b for every inserted read base; align.py:14-88), a few global atomics.
        auto stage_indel_word = [&](const PlaneRecord &rec, uint32_t gapw, uint32_t *row_at, int word, int slab_w0, int slab_words,
                                    int rstrand, int libx) {
            const int side = word >= NWA, kw = side ? word - NWA : word;
            const int cols = (int)(rec.cols & 0x7FFF), lf = (int)((rec.cols >> 16) & 0xFF), rf = (int)(rec.cols >> 24);
            const int a = (int)(gapw & 0x7FFF), k = (int)((gapw >> 15) & 7), del = (int)((gapw >> 18) & 1);
            const int c0 = (side ? cols + A - 32 * NWA : -A) + 32 * kw;  // column of bit 0
            auto range = [&](int lo_col, int hi_col) { return bit_range(lo_col - c0, hi_col - c0); };
            const uint32_t lo_mask = range(-(1 << 20), a), hi_mask = range(a + k, 1 << 20), gap_mask = range(a, a + k);
            // the read's planes of 32 bases from base index `qn` of seq4 (five words, no carry between window words)
            auto x_planes = [&](int64_t qn, uint32_t (&xp)[4]) {
                const int qs = (int)(qn & 7);
                const int64_t qw = qn >> 3, in_slab = qw - slab_w0;
                const bool from_smem = slab_words > 0 && in_slab >= 0 && in_slab + 5 <= slab_words;
                uint32_t w[5];
                if (from_smem) {
                    const uint32_t at = (uint32_t)__cvta_generic_to_shared(s_seq) + 4u * (uint32_t)in_slab;
#pragma unroll
                    for (int m = 0; m < 5; ++m) w[m] = lds_u32(at + 4 * m);
                } else {
#pragma unroll
                    for (int m = 0; m < 5; ++m) w[m] = __ldg(seq32 + qw + m);
                }
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    uint32_t q[5][2];
#pragma unroll
                    for (int m = 0; m < 5; ++m) ws_plane_bytes(w[m], pp, q[m][0], q[m][1]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t lo = __byte_perm(__byte_perm(q[0][h], q[1][h], 0x0073), __byte_perm(q[2][h], q[3][h], 0x0073), 0x5410);
                        xp[pp + 2 * h] = __funnelshift_r(lo, q[4][h] >> 24, qs);
                    }
                }
            };
            auto y_planes = [&](int64_t rn, uint32_t (&yp)[4]) {
                const uint4 g_lo = __ldg(planes + (rn >> 5)), g_hi = __ldg(planes + (rn >> 5) + 1);
                const int rs = (int)(rn & 31);
                yp[0] = __funnelshift_r(g_lo.x, g_hi.x, rs); yp[1] = __funnelshift_r(g_lo.y, g_hi.y, rs);
                yp[2] = __funnelshift_r(g_lo.z, g_hi.z, rs); yp[3] = __funnelshift_r(g_lo.w, g_hi.w, rs);
            };
            const int64_t q_at = (int64_t)rec.q0 + c0, r_at = ((int64_t)rec.rg << 5) + (int)((rec.misc >> 16) & 31) + c0;
            uint32_t x_lo[4], x_hi[4], y_m[4];
            x_planes(q_at, x_lo);
            y_planes(r_at, y_m);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) x_hi[pl] = x_lo[pl];
            if (del) {
                // (the right window's read composition is anchored at the read's last base: that alignment for all of it)
                if (hi_mask || side) x_planes(q_at - k, x_hi);
            } else if (hi_mask) {
                uint32_t y_hi[4];
                y_planes(r_at - k, y_hi);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) y_m[pl] = (y_m[pl] & lo_mask) | (y_hi[pl] & hi_mask);  // nothing under the inserted bases
            }
            uint32_t x_m[4], x_h[4];
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) {
                x_m[pl] = del ? (x_lo[pl] & lo_mask) | (x_hi[pl] & hi_mask) : x_lo[pl];  // by column
                x_h[pl] = side ? x_hi[pl] : x_lo[pl];                                     // by read base, from the window's anchor
            }
            const int v = min(cols, L), vh = min(cols - (del ? k : 0), L);
            const uint32_t m_range = side ? range(cols - v, cols) : range(0, v);
            const uint32_t h_range = side ? range(cols - vh, cols) : range(0, vh);
            const uint32_t flank = side ? range(cols, cols + rf) : range(-lf, 0);
            auto one_hot = [](const uint32_t (&x)[4]) { return (x[0] ^ x[1] ^ x[2] ^ x[3]) & ~((x[0] & x[1]) | (x[2] & x[3])); };
            const uint32_t one_m = one_hot(x_m), keep_m = one_m & m_range & ~gap_mask;
            const uint32_t keep_y = keep_m | (del ? gap_mask & m_range : 0u) | flank;  // a deleted base still counts as reference base
            const uint32_t keep_h = one_hot(x_h) & h_range;
            uint4 xs, ys;
            xs.x = x_h[0] & keep_h; xs.y = x_h[1] & keep_h; xs.z = x_h[2] & keep_h; xs.w = x_h[3] & keep_h;
            ys.x = y_m[0] & keep_y; ys.y = y_m[1] & keep_y; ys.z = y_m[2] & keep_y; ys.w = y_m[3] & keep_y;
            uint4 *out = (uint4 *)(row_at + 8 * word);
            out[0] = xs;
            out[1] = ys;
            // substitutions: aligned columns where both sides are bases and differ
            const uint32_t g1 = y_m[1] | y_m[3], g2 = y_m[2] | y_m[3], r1 = x_m[1] | x_m[3], r2 = x_m[2] | x_m[3];
            uint32_t ev = (y_m[0] | y_m[1] | y_m[2] | y_m[3]) & keep_m & ~((x_m[0] & y_m[0]) | (x_m[1] & y_m[1]) | (x_m[2] & y_m[2]) | (x_m[3] & y_m[3]));
            uint32_t *const sub_at = s_sub + ((size_t)libx * 4 + rstrand) * 12 * L;
            while (ev) {
                const int j = __ffs(ev) - 1;
                ev &= ev - 1;
                const int gb = (int)((g1 >> j) & 1u) + 2 * (int)((g2 >> j) & 1u);
                int rb = (int)((r1 >> j) & 1u) + 2 * (int)((r2 >> j) & 1u);
                rb -= rb > gb ? 1 : 0;
                const int col = c0 + j;
                atomicAdd(sub_at + (3 * gb + rb) * L + (side ? 2 * 12 * L + cols - 1 - col : col), 1u);
            }
            // gap columns inside the table: a reference base over nothing (deletion), a read base under nothing (insertion)
            uint32_t gaps = gap_mask & m_range & (del ? (y_m[0] | y_m[1] | y_m[2] | y_m[3]) : one_m);
            while (gaps) {
                const int j = __ffs(gaps) - 1;
                gaps &= gaps - 1;
                int base = del ? (int)((g1 >> j) & 1u) + 2 * (int)((g2 >> j) & 1u) : (int)((r1 >> j) & 1u) + 2 * (int)((r2 >> j) & 1u);
                if (rstrand) base = 3 - base;
                const int col = c0 + j, pos = side ? cols - 1 - col : col;
                const int es = libx * 4 + (side ^ rstrand) * 2 + rstrand;
                const int cls = del ? 4 + 5 * base + 4 : 4 + 5 * 4 + base;
                atomicAdd(t.misincorp + ((size_t)es * MDG_N_CLASSES + cls) * L + pos, 1ull);
            }
        };

        // ---- parse of one read: filter, classify, per-read events (statistics.py:37-51,117-126) ----
        // kind: 0 nothing to do, 1 gap-free (record made), 2 for the general kernel, 3 one short indel
        // FragmentLengths.update, statistics.py:117-126
        auto count_length = [&](int lkind, int64_t length, int rstrand, int libx) {
            if (length < MDG_LG_SMEM_BINS && length < p.lg_bins) {
                atomicAdd(s_lg + ((libx * 2 + lkind) * 2 + rstrand) * MDG_LG_SMEM_BINS + length, 1u);
            } else if (length < p.lg_bins) {
                atomicAdd(t.lghist + (size_t)((libx * 2 + lkind) * 2 + rstrand) * p.lg_bins + length, 1ull);
            } else {
                const unsigned long long at = atomicAdd(t.lg_overflow_count, 1ull);
                if ((int64_t)at < t.lg_overflow_cap) {
                    int32_t *row = t.lg_overflow_rows + at * 4;
                    row[0] = sub.list ? sub.lib : libx; row[1] = lkind; row[2] = rstrand; row[3] = (int32_t)length;
                }
            }
        };
        // kind: 0 nothing to do, 1 gap-free (record made), 2 for the general kernel, 3 one short indel (for the staged kernel),
        // 4 one short indel, no clips: record made and `gapw` = gap column | length << 15 | deletion << 18 |
        // fragment-length kind << 19 (0: none, 1: single-end = reference span, 2: |tlen| of the first mate of a proper pair)
        auto parse_read = [&](bool live, int64_t r, int q, int &kind, int &rstrand, int &libx, uint32_t &columns, PlaneRecord &rec,
                              uint32_t &gapw, uint32_t &lead_clip) {
            kind = 0;
            gapw = 0;
            lead_clip = 0;
            rstrand = 0;
            libx = 0;
            columns = 0;
            if (!live) return;
            uint32_t flag, lib, l_seq, boff, c0, c1;
            int32_t tid_ref;
            int64_t pos;
            if (staged_headers) {  // read q of the tile, from the team's copy
                flag = s_hflag[q]; lib = s_hlib[q]; tid_ref = s_htid[q]; pos = s_hpos[q];
                l_seq = s_hlseq[q]; boff = s_hboff[q]; c0 = s_hcoff[q]; c1 = s_hcoff[q + 1];
            } else {
                flag = b.flag[r]; lib = b.lib[r]; tid_ref = b.tid[r]; pos = b.pos[r];
                l_seq = b.l_seq[r]; boff = b.base_off[r]; c0 = b.cigar_off[r]; c1 = b.cigar_off[r + 1];
            }
            const uint32_t cig0 = c1 > c0 ? __ldg(b.cigar + c0) : 0;
            if (flag & FILTERED_FLAGS) return;
            if (lib >= (uint32_t)p.n_lib) {
                atomicCAS(t.error_flag, 0, DATA_ERR_LIB);
                return;
            }
            if (subset && lib != (uint32_t)sub.lib) return;  // cannot happen: the list is grouped by library
            libx = NL > 1 ? (int)lib : 0;  // which of this launch's table sets and lists the read belongs to
            if (tid_ref < 0 || tid_ref >= ref.n_contigs) {
                atomicCAS(t.error_flag, 0, DATA_ERR_TID);
                return;
            }
            rstrand = (flag >> 4) & 1;
            uint32_t lead = 0, trail = 0, cols = 0, gap_len = 0, gap_del = 0, gap_at = 0;
            int state = 0, n_lead = 0, n_trail = 0;
            bool simple = c1 > c0;
            if (c1 - c0 == 1 && ((0x181u >> (cig0 & 0xF)) & 1u)) {
                cols = cig0 >> 4;  // one match block (M, = or X): nearly every read of an untrimmed library
                state = 1;
            } else
            for (uint32_t k = c0; k < c1 && simple; ++k) {
                const uint32_t w = k == c0 ? cig0 : __ldg(b.cigar + k), op = w & 0xF, len = w >> 4;
                const bool match = op == OP_M || op == OP_EQ || op == OP_X;
                if (state == 0) {
                    if (op == OP_H) simple = n_lead == 0;
                    else if (op == OP_S) { lead += len; ++n_lead; }
                    else if (match) { cols += len; state = 1; }
                    else simple = false;
                } else if (state == 1) {
                    if (match) cols += len;
                    else if (op == OP_S) { trail += len; ++n_trail; state = 2; }
                    else if (op == OP_H) state = 3;
                    else if ((op == OP_I || op == OP_D) && !gap_len && len >= 1 && len <= 7 && cols >= 1) {
                        gap_len = len; gap_del = op == OP_D;
                        gap_at = cols;
                        cols += len;
                        state = 4;
                    } else simple = false;
                } else if (state == 2) {
                    if (op == OP_S) { trail += len; ++n_trail; }
                    else if (op == OP_H) state = 3;
                    else simple = false;
                } else if (state == 4) {  // the match block after the indel
                    if (match && len >= 1) { cols += len; state = 1; }
                    else simple = false;
                } else {
                    simple = op == OP_H;
                }
            }
            simple = simple && state != 4;
            const uint32_t n_query = cols - (gap_del ? gap_len : 0), ref_span = cols - (gap_len && !gap_del ? gap_len : 0);
            const int64_t contig_len = ref.contig_len[tid_ref];
            const uint64_t ref0 = ref.contig_off[tid_ref] + (uint64_t)(pos > 0 ? pos : 0);
            simple = simple && state >= 1 && cols > 0 && cols < 32768 && n_lead <= 1 && n_trail <= 1 &&
                     (uint64_t)lead + n_query + trail == l_seq && pos >= 0 && pos + (int64_t)ref_span <= contig_len &&
                     ref0 < (1ull << 33);
            if (!simple) {
                kind = 2;
                return;
            }
            if (gap_len && (!kIndels || lead || trail)) {
                kind = 3;  // count_staged_kernel's indel variant parses this read again and does all of its bookkeeping
                return;
            }
            kind = gap_len ? 4 : 1;
            columns = cols;
            lead_clip = lead;
            if constexpr (kQual) {
                if (b.qual) {  // the read's qualities towards L2: the stage reads them a barrier and a buffer hand-over later
                    prefetch_l2(b.qual + boff);
                    prefetch_l2(b.qual + boff + l_seq - 1);
                }
            }
            const int64_t aend = pos + ref_span;
            const uint32_t lf = (uint32_t)min((int64_t)A, pos);
            const uint32_t rf = (uint32_t)min((int64_t)A, contig_len - aend);
            rec.q0 = (uint32_t)((uint64_t)boff + lead);
            rec.rg = (uint32_t)(ref0 >> 5);
            if (g.prefetch_bases & 2) {
                // a genome that does not fit L2: the entries this read's window(s) will gather, pulled towards L2 now -- the
                // stage comes a barrier and a buffer hand-over later.  Measured on 3.1 Gbp: reads in coordinate order (what
                // a sorted BAM holds) 0.546 -> 0.487 ms per 4.17 M reads, reads in random order 0.591 -> 0.611.
                const char *const first = (const char *)(planes + ((int64_t)(ref0 >> 5) - 1));
                prefetch_l2(first);
                prefetch_l2(first + 16 * (((ref0 & 31) + cols + 2 * A + 63) >> 5));
            }
            rec.cols = cols | (lf << 16) | (rf << 24);
            rec.misc = min(cols, (uint32_t)L) | (uint32_t)(ref0 & 31) << 16;
            if (kind == 4) {
                // its fragment length is counted once it is known who counts the read (this kernel: tiles with two windows
                // per read; count_staged_kernel otherwise, which does all of the read's bookkeeping itself)
                const uint32_t lg = (flag & 0x1) ? (((flag & 0x40) && (flag & 0x2)) ? 2u : 0u) : 1u;
                gapw = gap_at | gap_len << 15 | gap_del << 18 | lg << 19;
                return;
            }
            {
                int64_t length = -1;
                int lkind = 0;
                if (flag & 0x1) {
                    if ((flag & 0x40) && (flag & 0x2)) {
                        const int64_t tl = b.tlen[r];
                        length = tl < 0 ? -tl : tl;
                    }
                } else {
                    lkind = 1;
                    length = ref_span;
                }
                if (length >= 0) count_length(lkind, length, rstrand, libx);
            }
            // update_soft_clipping, statistics.py:37-51
            if (lead) {
                const int end = rstrand ? 1 : 0, lim = (int)min(lead, (uint32_t)L);
                for (int i = 0; i < lim; ++i) atomicAdd(s_clip + ((libx * 2 + end) * 2 + rstrand) * L + i, 1u);
            }
            if (trail) {
                const int end = rstrand ? 0 : 1, lim = (int)min(trail, (uint32_t)L);
                for (int i = 0; i < lim; ++i) atomicAdd(s_clip + ((libx * 2 + end) * 2 + rstrand) * L + i, 1u);
            }
        };

        // L2 prefetch of the record arrays of a tile this team will parse later
        auto prefetch_headers = [&](int64_t tile_index) {
            const int64_t r4 = tile_index * T + (int64_t)ptid * 32;
            if (subset) {
                const int64_t at = tile_index * T + ptid;
                if (at < n_todo) {
                    const uint32_t r = subset[at];
                    prefetch_l2(b.flag + r);
                    prefetch_l2(b.tid + r);
                    prefetch_l2(b.pos + r);
                    prefetch_l2(b.l_seq + r);
                    prefetch_l2(b.base_off + r);
                    prefetch_l2(b.cigar_off + r);
                }
            } else if (ptid * 32 < T && r4 < b.n_reads) {
                // the record arrays come by bulk copy; what parse still reads from global memory: the first CIGAR word
                // (one op per read: this place exactly; otherwise a guess) and, for pairs, the template length
                prefetch_l2(b.cigar + r4);
                prefetch_l2(b.tlen + r4);
            }
        };
        // one thread: the record arrays of a tile as seven bulk copies on one mbarrier (sizes rounded up to 16 bytes: the
        // batch's arrays are 256-byte aligned with slack behind them, and a tile starts on a multiple of T records)
        auto issue_headers = [&](int64_t tile_index) {
            if (!staged_headers || tile_index >= n_tiles) return;
            const int64_t r0 = tile_index * T;
            const uint32_t n = (uint32_t)min((int64_t)T, n_todo - r0);
            const uint32_t b16 = (2 * n + 15) & ~15u, b32 = (4 * n + 15) & ~15u, b32p = (4 * (n + 1) + 15) & ~15u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hdr_addr), "r"(2 * b16 + 4 * b32 + b32p) : "memory");
            auto copy = [&](const void *to, const void *from, uint32_t bytes) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(to)),
                             "l"(from), "r"(bytes), "r"(hdr_addr)
                             : "memory");
            };
            copy(s_hflag, b.flag + r0, b16);
            copy(s_hlib, b.lib + r0, b16);
            copy(s_htid, b.tid + r0, b32);
            copy(s_hpos, b.pos + r0, b32);
            copy(s_hlseq, b.l_seq + r0, b32);
            copy(s_hboff, b.base_off + r0, b32);
            copy(s_hcoff, b.cigar_off + r0, b32p);
        };
        // one thread, once the tile's records have landed: the stretch of seq4 its reads occupy (reads are laid out in order;
        // a stage thread checks that its read really lies inside), 32 bytes more in front and 48 behind for the windows'
        // flanks, as one bulk copy
        auto issue_slab = [&](int64_t tile_index) {
            s_slab[team][1] = 0;
            if (subset || tile_index >= n_tiles) return;
            mbar_wait(hdr_addr, hdr_phase & 1u);  // (every thread waits for the same phase again before it parses)
            if (g.seq_words <= 0) return;  // bases from global memory (tests)
            const int n = (int)min((int64_t)T, n_todo - tile_index * T);
            const uint64_t first = s_hboff[0], last = (uint64_t)s_hboff[n - 1] + s_hlseq[n - 1];
            const int64_t lo = ((int64_t)(first >> 1) - 32) & ~15ll, hi = ((int64_t)((last + 1) >> 1) + 48 + 15) & ~15ll;
            const int64_t bytes = hi - lo;
            if (bytes <= 0 || bytes > 4ll * SEQ_WORDS) return;
            s_slab[team][0] = (int32_t)(lo >> 2);
            s_slab[team][1] = (int32_t)(bytes >> 2);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_seq);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(slab_addr), "r"((uint32_t)bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"((const char *)b.seq4 + lo), "r"((uint32_t)bytes), "r"(slab_addr)
                         : "memory");
        };

        prefetch_headers(tile_of(0, team));
        prefetch_headers(tile_of(1, team));
        if (ptid == 0) {
            issue_headers(tile_of(0, team));
            issue_slab(tile_of(0, team));
        }
        named_barrier<T>(1 + team);  // s_slab of the first tile
        for (int64_t k = 0;; ++k) {
            const int64_t tile = tile_of(k, team);
            if (tile >= n_tiles) break;
            uint32_t *const s_ctl = s_ctl_base + WS_CTL * (int)(k % 3);
            uint32_t *const s_ctl_next = s_ctl_base + WS_CTL * (int)((k + 1) % 3);

            // ---- parse: one read per thread ----
            if (staged_headers) {
                mbar_wait(hdr_addr, hdr_phase & 1u);
                ++hdr_phase;
            }
            // the thread that parses a read also stages it: the record stays in its registers
            int kind, rstrand, libx;
            uint32_t rank = 0;  // of a gap-free read: its index among the tile's reads of its group (library, strand)
            int key;            // group of a read with a row; CTL_CX / CTL_IX + library of a read left to another kernel; -1: nothing
            PlaneRecord rec{};
            uint32_t gapw;      // kind 4: where its gap is (parse_read)
            uint32_t lead_clip; // bases soft-clipped in front of the first aligned base (kQual: where the read's qualities start)
            int64_t my_read;
            {
                const int64_t at = tile * T + ptid;
                const bool live = at < n_todo;
                const int64_t r = !live ? 0 : subset ? (int64_t)subset[at] : at;
                my_read = r;
                uint32_t columns;
                parse_read(live, r, ptid, kind, rstrand, libx, columns, rec, gapw, lead_clip);
                if (g.uniform) {
                    const uint32_t lo = __reduce_min_sync(0xffffffffu, kind == 1 ? columns : 0xffffffffu);
                    const uint32_t hi = __reduce_max_sync(0xffffffffu, kind == 1 ? columns : 0u);
                    if (lane == 0 && hi) {
                        atomicMin(s_ctl + CTL_MIN, lo);
                        atomicMax(s_ctl + CTL_MAX, hi);
                    }
                }
                // warp-aggregated appends: the lanes with the same key take consecutive places behind one atomic
                key = (kind == 1 || kind == 4) ? CTL_GROUP + rstrand + 2 * libx : kind == 2 ? CTL_CX + libx : kind == 3 ? CTL_IX + libx : -1;
                if constexpr (kNL == 1) {
                    // four keys: one vote each (cheaper than a match)
                    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
                    for (int which = 0; which < 4; ++which) {
                        const int this_key = which == 2 ? CTL_CX : which == 3 ? CTL_IX : CTL_GROUP + which;
                        const uint32_t m = __ballot_sync(0xffffffffu, key == this_key);
                        if (m) {
                            uint32_t base = 0;
                            if (lane == __ffs(m) - 1) base = atomicAdd(s_ctl + this_key, (uint32_t)__popc(m));
                            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                            if (key == this_key) rank = base + __popc(m & lt);
                        }
                    }
                } else {
                    const uint32_t peers = __match_any_sync(0xffffffffu, key);
                    if (key >= 0) {
                        const int leader = __ffs(peers) - 1;
                        uint32_t base = 0;
                        if (lane == leader) base = atomicAdd(s_ctl + key, (uint32_t)__popc(peers));
                        rank = __shfl_sync(peers, base, leader) + __popc(peers & ((1u << lane) - 1u));
                    }
                }
                if (kind == 2) s_cx[libx * T + rank] = (uint32_t)r;
                else if (kind == 3) s_ix[libx * T + rank] = (uint32_t)r;
                // a one-indel read this kernel may stage has a row (its rank above) AND a place at the back of the library's
                // list of one-indel reads, in case the tile turns out to have one window per read
                const uint32_t m4 = kIndels ? __ballot_sync(0xffffffffu, kind == 4) : 0u;
                if (m4) {
#pragma unroll
                    for (int lib = 0; lib < NL; ++lib) {
                        const uint32_t ml = NL == 1 ? m4 : __ballot_sync(0xffffffffu, kind == 4 && libx == lib);
                        if (!ml) continue;
                        uint32_t base = 0;
                        if (lane == __ffs(ml) - 1) base = atomicAdd(s_ctl + CTL_IX4 + lib, (uint32_t)__popc(ml));
                        base = __shfl_sync(0xffffffffu, base, __ffs(ml) - 1);
                        if (kind == 4 && libx == lib) s_ix[lib * T + T - 1 - (base + __popc(ml & ((1u << lane) - 1u)))] = (uint32_t)r;
                    }
                }
            }
            named_barrier<T>(1 + team);

            // ---- one window per read when every gap-free read of the tile has the same length ----
            int want = 0;
            {
                const uint32_t lo = s_ctl[CTL_MIN], hi = s_ctl[CTL_MAX];
                if (g.uniform && lo == hi && hi > 0) {
                    const int words = ((int)hi + 2 * A + 31) / 32;
                    if (words < WPR_MAX && PAIRS / words >= n_groups) want = (int)hi;
                }
            }
            // ---- reads this kernel does not count go to the two work lists ----
            if (pwarp < 2) {  // warp 0: reads for the general kernel, warp 1: one-indel reads; per library
                for (int lib = 0; lib < NL; ++lib) {
                    const uint32_t n_front = s_ctl[(pwarp ? CTL_IX : CTL_CX) + lib];
                    // the one-indel reads this kernel could stage itself are handed over too when the tile has one window per read
                    const uint32_t n_back = pwarp && want ? s_ctl[CTL_IX4 + lib] : 0u;
                    const uint32_t n = n_front + n_back;
                    // (steers the host's choice of variants: every one-indel read met, whoever counts it)
                    if (pwarp && lane == 0 && g.indel_seen && n_front + s_ctl[CTL_IX4 + lib])
                        atomicAdd(g.indel_seen, (unsigned long long)(n_front + s_ctl[CTL_IX4 + lib]));
                    if (!n) continue;
                    const int at_lib = sub.list ? sub.lib : lib;  // the global lists are kept per library
                    const uint32_t *const from = (pwarp ? s_ix : s_cx) + lib * T;
                    uint32_t *const to = (pwarp ? indel_list : worklist) + (sub.offsets ? sub.offsets[at_lib] : 0);
                    unsigned long long base = 0;
                    if (lane == 0) {
                        base = atomicAdd((pwarp ? indel_count : work_count) + at_lib, (unsigned long long)n);
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (uint32_t i = lane; i < n; i += 32) to[base + i] = from[i < n_front ? i : T - 1 - (i - n_front)];
                }
            }
            // the control block of the tile after this one (last read by the consumers two tiles ago)
            if (ptid < WS_CTL) s_ctl_next[ptid] = ptid == CTL_MIN ? 0xffffffffu : 0u;
            if (ptid == 32) issue_headers(tile_of(k + 1, team));  // everyone is done with this tile's records: the next tile's land while this one is staged
            if (ptid == 0) s_ctl[CTL_MODE] = (uint32_t)want;
            if (want != mode) {
                mode = want;
                fill_masks(want);
                named_barrier<T>(1 + team);
            }
            prefetch_headers(tile_of(k + 2, team));

            // ---- stage: one thread per (read, window), once the consumers are done with the tile before ----
            if (k > 0) mbar_wait(empty_addr, (uint32_t)((k - 1) & 1));
            const int slab_w0 = s_slab[team][0], slab_words = s_slab[team][1];
            if (slab_words) {
                mbar_wait(slab_addr, slab_phase & 1u);
                ++slab_phase;
            }
            int row = (int)rank;  // the groups' rows follow each other
#pragma unroll
            for (int lower = 0; lower < n_groups - 1; ++lower)
                if (lower < key) row += (int)s_ctl[CTL_GROUP + lower];
            uint32_t *const row_at = s_stage + (size_t)row * ROW;
            if (kind == 1) {
                const int n_words = mode ? words_of(mode) : NWA;
#pragma unroll 1
                for (int side = 0; side < (mode ? 1 : 2); ++side) {
                    const int first_word = side ? NWA : 0;
                    const int c_start = side ? (int)(rec.cols & 0x7FFF) + A - 32 * NWA : -A;
                    if (kGather && n_words == 4) stage_window(std::integral_constant<int, 4>{}, rec, row_at, first_word, 4, c_start, side, slab_w0, slab_words, rstrand, libx, lead_clip);
                    else if (kGather && n_words == 3) stage_window(std::integral_constant<int, 3>{}, rec, row_at, first_word, 3, c_start, side, slab_w0, slab_words, rstrand, libx, lead_clip);
                    else stage_window(std::integral_constant<int, 0>{}, rec, row_at, first_word, n_words, c_start, side, slab_w0, slab_words, rstrand, libx, lead_clip);
                }
            }
            // ---- reads with one insertion / deletion ----
            if (const uint32_t m4 = kIndels ? __ballot_sync(0xffffffffu, kind == 4) : 0u) {
                if (mode) {
                    // one window per read: count_staged_kernel has them (above); their rows hold nothing
                    if (kind == 4)
                        for (int w = 0; w < 2 * words_of(mode); ++w) ((uint4 *)row_at)[w] = make_uint4(0u, 0u, 0u, 0u);
                } else {
                    if (kind == 4) {
                        const uint32_t lg = (gapw >> 19) & 3u, cols = rec.cols & 0x7FFFu, glen = (gapw >> 15) & 7u;
                        if (lg == 1) count_length(1, (int64_t)(cols - (((gapw >> 18) & 1u) ? 0u : glen)), rstrand, libx);
                        else if (lg == 2) {
                            const int64_t tl = b.tlen[my_read];
                            count_length(0, tl < 0 ? -tl : tl, rstrand, libx);
                        }
                    }
                    // The window words of the warp's one-indel reads are dealt to ALL its lanes (the reads' records travel by
                    // shuffles): a lane that holds such a read would otherwise make six slow words while 31 wait.
                    const int n_items = __popc(m4) * WPR_MAX;
                    for (int first = 0; first < n_items; first += 32) {
                        const int item = first + lane;
                        const bool live = item < n_items;
                        const int src = live ? (int)__fns(m4, 0, item / WPR_MAX + 1) : 0;
                        PlaneRecord irec;
                        irec.q0 = __shfl_sync(0xffffffffu, rec.q0, src);
                        irec.rg = __shfl_sync(0xffffffffu, rec.rg, src);
                        irec.cols = __shfl_sync(0xffffffffu, rec.cols, src);
                        irec.misc = __shfl_sync(0xffffffffu, rec.misc, src);
                        const uint32_t igap = __shfl_sync(0xffffffffu, gapw, src);
                        const int irow = __shfl_sync(0xffffffffu, row, src);
                        const int iwho = __shfl_sync(0xffffffffu, rstrand | (libx << 1), src);
                        if constexpr (kIndels)
                            if (live) stage_indel_word(irec, igap, s_stage + (size_t)irow * ROW, item % WPR_MAX, slab_w0, slab_words, iwho & 1, iwho >> 1);
                    }
                }
            }
            mbar_arrive(full_addr);  // release: this thread's words of the buffer (and the control block) are visible to the consumers
            named_barrier<T>(1 + team);
            if (ptid == 0) issue_slab(tile_of(k + 1, team));  // lands while the next tile is parsed
            // (the next parse does not read s_slab; the barrier after it orders this write before the stage)
        }
    } else {
        // =============================================== consumers ===============================================
        const int ctid = tid - PRODUCERS, cwarp = ctid >> 5;
        const int group = cwarp & 3;                 // reference base of this thread's classes
        const int pair = (cwarp >> 2) * 32 + lane;   // index among the (slot, word) pairs of its group
        // a read slot belongs to one group (library, strand) and takes every spg-th read of it
        int mode = 0, ws = 0, grp = 0, idx = 0, strand = 0, mylib = 0, spg = 1;
        bool active = false;
        auto set_mode = [&](int columns) {
            mode = columns;
            const int wpr = words_of(columns);
            spg = (PAIRS / wpr) / n_groups;  // slots per group (at least one: the host and the producers see to it)
            active = pair < wpr * spg * n_groups;
            ws = pair % wpr;
            const int slot = pair / wpr;
            grp = slot % n_groups;
            idx = slot / n_groups;
            strand = grp & 1;
            mylib = grp >> 1;
        };
        set_mode(0);
        uint32_t cnt[PL_CLASSES][PL_REG];
#pragma unroll
        for (int c = 0; c < PL_CLASSES; ++c)
#pragma unroll
            for (int k = 0; k < PL_REG; ++k) cnt[c][k] = 0;
        int n_iter = 0;  // eight-read iterations since planes 4 .. 7 were moved up
        uint32_t *const my_wide = s_wide + ctid;  // plane k of class c at my_wide[(k * PL_CLASSES + c) * CT]

        // planes 4 .. 7 of the register counters -> the wide counters in shared memory (a ripple-carry add, plane by plane)
        auto spill = [&]() {
#pragma unroll
            for (int c = 0; c < PL_CLASSES; ++c) {
                uint32_t *w = my_wide + c * CT;
                uint32_t carry = 0;
#pragma unroll
                for (int k = 0; k < PL_REG - 4; ++k) {
                    const uint32_t wk = w[k * PL_CLASSES * CT], v = cnt[c][4 + k];
                    w[k * PL_CLASSES * CT] = wk ^ v ^ carry;
                    carry = (wk & v) | ((wk ^ v) & carry);
                    cnt[c][4 + k] = 0;
                }
                for (int k = PL_REG - 4; carry && k < PL_WIDE; ++k) {
                    const uint32_t wk = w[k * PL_CLASSES * CT];
                    w[k * PL_CLASSES * CT] = wk ^ carry;
                    carry &= wk;
                }
            }
            n_iter = 0;
        };
        auto add_cell = [&](int lib, int canchor, int cstrand, int cls, int pos, unsigned long long sum) {
            // window position `pos` of an anchor -> table cell of the library; classes are complemented on the reverse strand
            const int es = lib * 4 + (canchor ^ cstrand) * 2 + cstrand;
            if (pos >= 0) {
                if (cls < 4) {
                    const int gb = cstrand ? 3 - cls : cls;
                    atomicAdd(t.misincorp + ((size_t)es * MDG_N_CLASSES + gb) * L + pos, sum);
                } else {
                    const int rb = cstrand ? 3 - (cls - 4) : cls - 4;
                    atomicAdd(t.dnacomp + ((size_t)es * 4 + rb) * LA + pos, sum);
                }
            } else if (cls < 4) {
                const int gb = cstrand ? 3 - cls : cls;
                atomicAdd(t.dnacomp + ((size_t)es * 4 + gb) * LA + L - pos - 1, sum);
            }
        };
        // reduces the consumers' counters into the 64-bit tables (end of the kernel, on layout changes, and before a
        // 12-bit counter could overflow)
        auto flush = [&]() {
            spill();
            const int wpr = words_of(mode), bits = 32 * wpr;
            for (int lib = 0; lib < NL; ++lib) {  // one library at a time through the reduction area
                for (int i = ctid; i < 2 * 8 * bits; i += CT) s_red[i] = 0;
                named_barrier<CT>(1 + kTeams);
                if (active && mylib == lib) {
#pragma unroll
                    for (int c = 0; c < PL_CLASSES; ++c) {
                        const int cls = c == 0 ? group : 4 + group;  // class of the tables: R_g, H_g
                        uint32_t pl[4 + PL_WIDE];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            pl[k] = cnt[c][k];
                            cnt[c][k] = 0;
                        }
#pragma unroll
                        for (int k = 0; k < PL_WIDE; ++k) {
                            pl[4 + k] = my_wide[(k * PL_CLASSES + c) * CT];
                            my_wide[(k * PL_CLASSES + c) * CT] = 0;
                        }
                        uint32_t any = 0;
#pragma unroll
                        for (int k = 0; k < 4 + PL_WIDE; ++k) any |= pl[k];
                        uint32_t *const to = s_red + ((size_t)strand * 8 + cls) * bits + 32 * ws;
                        while (any) {
                            const int j = __ffs(any) - 1;
                            any &= any - 1;
                            uint32_t v = 0;
#pragma unroll
                            for (int k = 0; k < 4 + PL_WIDE; ++k) v |= ((pl[k] >> j) & 1u) << k;
                            atomicAdd(to + j, v);
                        }
                    }
                }
                named_barrier<CT>(1 + kTeams);
                for (int cell = ctid; cell < 2 * 8 * bits; cell += CT) {
                    const int bit = cell % bits, cls = (cell / bits) % 8, cstrand = cell / (8 * bits);
                    const unsigned long long sum = s_red[cell];
                    if (!sum) continue;
                    // a window bit feeds the table of the left end (anchor 0), of the right end (anchor 1), or both
                    int pos0 = L, pos1 = L;  // L: no cell
                    if (mode) {
                        const int col = bit - A;
                        if (col < 0) pos0 = col;                      // left flank
                        else if (col >= mode) pos1 = mode - 1 - col;  // right flank at distance col - C + 1
                        else {
                            pos0 = col;
                            pos1 = mode - 1 - col;
                        }
                    } else if (bit < 32 * NWA) {
                        pos0 = bit - A;
                    } else {
                        pos1 = 32 * NWA - A - 1 - (bit - 32 * NWA);  // columns from the right end; negative: flank
                    }
#pragma unroll 1
                    for (int canchor = 0; canchor < 2; ++canchor) {
                        const int pos = canchor ? pos1 : pos0;
                        if (pos < L) add_cell(lib, canchor, cstrand, cls, pos, sum);
                    }
                }
                named_barrier<CT>(1 + kTeams);
            }
        };

        int since_flush = 0, tiles_since_flush = 0;  // reads a counter may have seen / tiles since the last reduction
        int mode_counted = 0;                        // window layout the counters hold
        bool dirty = false;
        for (int64_t k = 0, team = 0;; ++team) {
            if (team == kTeams) {
                team = 0;
                ++k;
            }
            const bool last = tile_of(k, (int)team) >= n_tiles;  // tiles grow with (k, team): nothing behind this one either
            const uint32_t *const s_stage = s_teams + (size_t)team * team_words;
            const uint32_t *const s_ctl = s_stage + T * ROW + LISTS + ((2 * WPR_MAX + 3) & ~3) + WS_CTL * (int)(k % 3);
            int n_most = 0, n_all = 0, n_mine = 0, my_row = 0, want = mode;
            if (!last) {
                mbar_wait((uint32_t)__cvta_generic_to_shared(&s_full[team]), (uint32_t)(k & 1));
                want = (int)s_ctl[CTL_MODE];
                if (want != mode) set_mode(want);  // (the counters are reduced below before they are used in the new layout)
                for (int other = 0; other < n_groups; ++other) {
                    const int n = (int)s_ctl[CTL_GROUP + other];
                    n_most = max(n_most, n);
                    n_all += n;
                    if (other < grp) my_row += n;  // the groups' rows follow each other
                    if (other == grp) n_mine = n;
                }
            }
            const int bound = ((n_most + spg - 1) / spg + 7) & ~7;  // reads a thread adds at most, whole iterations
            if (last || (want != mode_counted && dirty) || since_flush + bound > WS_CAPACITY ||
                (g.flush_tiles > 0 && tiles_since_flush >= g.flush_tiles)) {
                // (reduce in the layout the counters were filled in)
                const int now = mode;
                if (now != mode_counted) set_mode(mode_counted);
                flush();
                if (now != mode_counted) set_mode(now);
                since_flush = 0;
                tiles_since_flush = 0;
                dirty = false;
            }
            if (last) break;
            mode_counted = mode;
            since_flush += bound;
            ++tiles_since_flush;
            dirty = dirty || n_all > 0;
            // ---- count: this thread's window word and reference base, every spg-th read of its group ----
            if (active) {
                const int row_step = spg * ROW;
                const uint32_t *at = s_stage + (size_t)(my_row + idx) * ROW + 8 * ws + group;
                for (int i = idx; i < n_mine; i += 8 * spg) {
                    uint32_t xg[8], y[8];
                    // (one path with tests: a copy without them for full iterations measured the same and is more code)
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool live = i + u * spg < n_mine;
                        xg[u] = live ? at[u * row_step] : 0u;
                        y[u] = live ? at[u * row_step + 4] : 0u;
                    }
                    at += 8 * row_step;
                    MDG_ADD8(cnt[0], y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7])  // R_g
                    MDG_ADD8(cnt[1], xg[0], xg[1], xg[2], xg[3], xg[4], xg[5], xg[6], xg[7])  // H_g
                    if (++n_iter == 30) spill();  // planes 0-3 hold at most 15, thirty more iterations add 240: 255 fits eight planes
                }
            }
            mbar_arrive((uint32_t)__cvta_generic_to_shared(&s_empty[team]));  // the buffer may be staged again
        }
    }

    // ---- everybody: the block's event tables into the 64-bit tables ----
    __syncthreads();
    for (int cell = tid; cell < NL * 4 * 12 * L; cell += NTHREADS) {
        const uint32_t v = s_sub[cell];
        if (!v) continue;
        const int pos = cell % L, cls = (cell / L) % 12, cstrand = (cell / (12 * L)) & 1, canchor = (cell / (24 * L)) & 1, lib = cell / (48 * L);
        int gb = cls / 3, rb = cls % 3;
        rb += rb >= gb ? 1 : 0;
        if (cstrand) { gb = 3 - gb; rb = 3 - rb; }
        const int es = lib * 4 + (canchor ^ cstrand) * 2 + cstrand;
        atomicAdd(t.misincorp + ((size_t)es * MDG_N_CLASSES + 4 + 5 * gb + rb) * L + pos, (unsigned long long)v);
    }
    for (int i = tid; i < NL * 4 * MDG_LG_SMEM_BINS; i += NTHREADS) {  // [lib][kind][strand][bin]: the tables' own order
        const uint32_t v = s_lg[i];
        if (v) atomicAdd(t.lghist + (size_t)(i / MDG_LG_SMEM_BINS) * p.lg_bins + i % MDG_LG_SMEM_BINS, (unsigned long long)v);
    }
    for (int i = tid; i < NL * 4 * L; i += NTHREADS) {  // [lib][end][strand][position]
        const uint32_t v = s_clip[i];
        if (v) atomicAdd(t.misincorp + ((size_t)(i / L) * MDG_N_CLASSES + MDG_CLASS_SOFTCLIP) * L + i % L, (unsigned long long)v);
    }
}

// dynamic shared memory of count_planes_ws_kernel<kTeams, kTeamWarps, kConsWarps, kNWA> (bytes)
inline size_t planes_ws_smem(int teams, int team_warps, int cons_warps, int L, int nw_anchor, int n_lib)
{
    const size_t T = (size_t)team_warps * 32, CT = (size_t)cons_warps * 32, wpr_max = 2 * (size_t)nw_anchor, row = 16 * (size_t)nw_anchor + 4;
    const size_t shared = PL_WIDE * PL_CLASSES * CT + 2 * 8 * 32 * wpr_max + n_lib * (4 * 12 * (size_t)L + 4 * MDG_LG_SMEM_BINS + 4 * (size_t)L);
    const size_t team = T * row + 2 * WS_MAX_LIB * T + ((2 * wpr_max + 3) & ~(size_t)3) + 3 * WS_CTL + T * 14 + 6 * T + 4;
    return (shared + teams * team) * 4;
}

}  // namespace mdg
