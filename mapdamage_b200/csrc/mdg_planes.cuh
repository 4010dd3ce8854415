// Counting pass, bit-plane kernel for gap-free reads: the production counting kernel of round 2.
//
// Same contract as count_staged_kernel (mdg_stage.cuh) -- the loop body of main.py:165-217 for reads whose CIGAR is
// [H][S] M/=/X+ [S][H]: read filter (reader.py:121-132), fragment lengths (statistics.py:117-126), soft clips
// (statistics.py:37-51), both MisincorporationRates.update walks (statistics.py:22-35), read and flank composition
// (statistics.py:75-93) -- and the same idea of turning the reference's loop inside out (a thread owns table
// positions, the reads are its loop), but the data are held as BIT PLANES instead of nibbles:
//
//   a window of 32 table positions of one read is eight 32-bit words: read base is A / C / G / T (X planes) and
//   reference base is A / C / G / T (Y planes), one bit per position.  The genome is kept in that form (DevRef::planes);
//   the read's BAM nibbles are transposed once per read (four delta swaps per eight bases).
//
//   every table class is then ONE word for 32 positions -- R_g = Y_g, H_b = X_b, P_gb = Y_g & X_b -- and is added into
//   vertical counters (bit plane k of a counter word holds bit k of 32 counts) with carry-save adders: eight reads
//   cost 19 three-input logic operations per class, 2.4 per read and 32 positions, against 1.5 per read and EIGHT
//   positions for the nibble counters of mdg_stage.cuh.
//
// A thread of the count phase is (read slot, window word, reference base g) and owns the five classes R_g, H_g and
// P_g* of 32 positions; warps are uniform in g, so the class wiring is compile-time.  Planes 0-5 of its counters live in
// registers, planes 4-13 in shared memory (the top two register planes are moved up every 48 reads); the block
// reduces everything into the 64-bit tables when a counter could overflow, when the window layout changes, and at
// the end.  Window layouts are those of mdg_stage.cuh: one window [-A, C + A) per read while every gap-free read
// of a tile has the same C columns, otherwise a left-anchored and a right-anchored window [-A, L) per read.
//
// Reads with one short insertion / deletion are appended to `indel_list` (count_staged_kernel's three-plane variant
// counts them), everything else this layout cannot express to `worklist` (count_general_kernel).
//
// Tile input arrives by bulk asynchronous copies (cp.async.bulk + mbarrier, one elected thread): the record arrays of
// tile t + 1 land in shared memory while tile t is staged and counted.
#pragma once
#include <type_traits>
#include "mdg_swar.cuh"

namespace mdg {

struct PlaneGeom {
    int32_t threads;      // blockDim.x (a multiple of 128: warps are uniform in the reference base g = warp & 3)
    int32_t tile;         // reads per tile
    int32_t uniform;      // 1: one window per read for tiles of equal-length reads
    int32_t flush_tiles;  // > 0: reduce the counters every so many tiles (tests)
    int32_t nw_anchor;    // words per anchor window: ceil((L + A) / 32)
    int32_t row_words;    // words per staged read: 8 * (2 * nw_anchor) + 4 (rows land on different banks)
    int32_t seq_words;    // shared-memory words for the tile's slab of seq4 (bulk-copied ahead); 0: none
    unsigned long long *indel_seen;  // counts the one-indel reads met (steers the host's choice of variants)
    int32_t prefetch_bases;  // 1: L2 prefetch of the bases of the tile after next even when the bulk copy brings them in
};

constexpr int PL_REG = 8;    // counter planes in registers (counts to 255)
constexpr int PL_WIDE = 8;   // counter planes in shared memory, weights 2^4 .. 2^11
constexpr int PL_CLASSES = 2;  // R_g and H_g; the substitution classes P_g* are rare events and go to a shared table

struct __align__(16) PlaneRecord {
    uint32_t q0;    // base index (nibble) of the first aligned base in seq4: base_off + leading clip
    uint32_t rg;    // 32-base group of the genome holding the first aligned column
    uint32_t cols;  // columns (15 bits) | has_qual << 15 | left flank bases << 16 | right flank bases << 24
    uint32_t misc;  // min(L, columns) | bit of the genome group << 16
};

// eight BAM nibbles (natural order: nibble j = base j) -> byte p = plane p (bit j = bit p of nibble j)
__device__ __forceinline__ uint32_t nibbles_to_planes(uint32_t x)
{
    uint32_t t;
    t = ((x >> 1) ^ x) & 0x22222222u; x ^= t ^ (t << 1);
    t = ((x >> 3) ^ x) & 0x0A0A0A0Au; x ^= t ^ (t << 3);
    t = ((x >> 6) ^ x) & 0x00CC00CCu; x ^= t ^ (t << 6);
    t = ((x >> 12) ^ x) & 0x0000F0F0u; x ^= t ^ (t << 12);
    return x;
}

// The same for a word of seq4 as BAM stores it (first base of a byte in the high nibble).  Logic and funnel shifts
// issue on the ALU pipe, one warp-instruction per two cycles per scheduler, and this kernel is bound by it; a shift by a
// constant is also a multiplication (left: x * 2^d, right: the high half of x * 2^(32 - d)), which issues on the FMA
// pipe next to it.
__device__ __forceinline__ uint32_t shl_fma(uint32_t x, int d) { return x * (1u << d); }
__device__ __forceinline__ uint32_t shr_fma(uint32_t x, int d) { return __umulhi(x, 1u << (32 - d)); }
__device__ __forceinline__ uint32_t bam_word_to_planes(uint32_t w)
{
    uint32_t x = (shl_fma(w & 0x0F0F0F0Fu, 4)) | (shr_fma(w, 4) & 0x0F0F0F0Fu);  // nibble j = base j
    uint32_t t;
    t = (shr_fma(x, 1) ^ x) & 0x22222222u; x ^= t ^ shl_fma(t, 1);
    t = (shr_fma(x, 3) ^ x) & 0x0A0A0A0Au; x ^= t ^ shl_fma(t, 3);
    t = (shr_fma(x, 6) ^ x) & 0x00CC00CCu; x ^= t ^ shl_fma(t, 6);
    t = (shr_fma(x, 12) ^ x) & 0x0000F0F0u; x ^= t ^ shl_fma(t, 12);
    return x;
}

// bits [lo, hi) of a 32-bit word, both clamped to [0, 32]
__device__ __forceinline__ uint32_t bit_range(int lo, int hi)
{
    lo = max(lo, 0);
    hi = min(hi, 32);
    if (hi <= lo) return 0u;
    return (0xFFFFFFFFu >> (32 - (hi - lo))) << lo;
}

// carry-save adder: (h, l) = a + b + c
#define MDG_CSA(h, l, a, b, c)                \
    {                                         \
        const uint32_t u_ = (a) ^ (b);        \
        h = ((a) & (b)) | (u_ & (c));         \
        l = u_ ^ (c);                         \
    }
// adds eight one-bit masks into an eight-plane vertical counter
#define MDG_ADD8(c, m0, m1, m2, m3, m4, m5, m6, m7)                 \
    {                                                               \
        uint32_t a1, b1, c1, d1, a2, b2, a4, k;                     \
        MDG_CSA(a1, c[0], c[0], m0, m1)                             \
        MDG_CSA(b1, c[0], c[0], m2, m3)                             \
        MDG_CSA(c1, c[0], c[0], m4, m5)                             \
        MDG_CSA(d1, c[0], c[0], m6, m7)                             \
        MDG_CSA(a2, c[1], c[1], a1, b1)                             \
        MDG_CSA(b2, c[1], c[1], c1, d1)                             \
        MDG_CSA(a4, c[2], c[2], a2, b2)                             \
        k = c[3] & a4; c[3] ^= a4;                                  \
        a4 = c[4] & k; c[4] ^= k;                                   \
        k = c[5] & a4; c[5] ^= a4;                                  \
        a4 = c[6] & k; c[6] ^= k;                                   \
        c[7] ^= a4;                                                 \
    }

#ifdef MDG_PHASE_CLOCKS
__device__ unsigned int mdg_plane_phase_dump[16];
#endif

template <int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 2 : 1)
count_planes_kernel(DevBatch b, DevRef ref, CountParams p, CountTables t, PlaneGeom g, uint32_t *__restrict__ worklist,
                    unsigned long long *__restrict__ work_count, uint32_t *__restrict__ indel_list,
                    unsigned long long *__restrict__ indel_count, SwarSubset sub)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int nthreads = g.threads, T = g.tile, L = p.L, A = p.A;
    const int NWA = g.nw_anchor, WPR_MAX = 2 * NWA, ROW = g.row_words;
    uint32_t *const s_wide = smem;                                              // [PL_WIDE][PL_CLASSES][nthreads]
    uint32_t *const s_stage = s_wide + PL_WIDE * PL_CLASSES * nthreads;         // [T][ROW]; the reduction table at a flush
    PlaneRecord *const s_rec = (PlaneRecord *)(s_stage + (size_t)T * ROW);      // [T]: forward reads from the front, reverse from the back
    uint32_t *const s_cx = (uint32_t *)(s_rec + T);                             // [T] reads for the general kernel
    uint32_t *const s_ix = s_cx + T;                                            // [T] one-indel reads
    uint32_t *const s_mask = s_ix + T;                                          // [WPR_MAX][2] aligned / flank masks of a typical read
    uint32_t *const s_lg = s_mask + 2 * WPR_MAX;                                // [kind][strand][MDG_LG_SMEM_BINS]
    uint32_t *const s_clip = s_lg + 4 * MDG_LG_SMEM_BINS;                       // [end][strand][L]
    uint32_t *const s_ctl_base = s_clip + 4 * L;                                // two sets of {n_fwd, n_rev, n_cx, min cols, max cols, n_ix, -, -}
    // the tile's stretch of seq4, brought in by one bulk asynchronous copy (cp.async.bulk, completion on an mbarrier)
    // while the tile before is counted
    uint32_t *const s_sub = s_ctl_base + 16;                                    // [strand][12 substitution classes][32 * WPR_MAX window bits]
    uint32_t *const s_seq = s_sub + 2 * 12 * 32 * WPR_MAX;  // every piece above is a multiple of four words: 16-byte aligned (a cast through an integer would make every load from it a generic load)
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int32_t s_slab[2];  // first seq4 word held in s_seq (may be negative: the words in front of the array), words (0: no copy)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < PL_WIDE * PL_CLASSES * nthreads; i += nthreads) s_wide[i] = 0;
    for (int i = tid; i < 4 * MDG_LG_SMEM_BINS + 4 * L; i += nthreads) s_lg[i] = 0;
    for (int i = tid; i < 2 * 12 * 32 * WPR_MAX; i += nthreads) s_sub[i] = 0;

    const uint32_t mbar_addr = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        s_slab[0] = 0;
        s_slab[1] = 0;
    }
    uint32_t slab_phase = 0;  // bulk copies waited for so far (the mbarrier's phase parity)
    const uint32_t *__restrict__ seq32 = (const uint32_t *)b.seq4;
    const uint4 *__restrict__ planes = ref.planes;
    const uint32_t *const subset = sub.list ? sub.list + sub.offsets[sub.lib] : nullptr;
    const int64_t n_todo = sub.list ? (int64_t)(sub.offsets[sub.lib + 1] - sub.offsets[sub.lib]) : b.n_reads;

    // ---- window layout of the block ----
    // mode 0: two windows per read, words [0, NWA) left-anchored (bit i of the window = position i - A from the left
    // end), words [NWA, 2 NWA) right-anchored in memory order (the window ends A bases behind the last column).
    // mode C > 0: every gap-free read of the tile has C columns; one window [-A, C + A).
    int mode = 0;
    auto words_of = [&](int columns) { return columns ? (columns + 2 * A + 31) / 32 : WPR_MAX; };
    const int group = warp & 3;                      // reference base of this thread's classes
    const int pair = (warp >> 2) * 32 + lane;        // index among the (slot, word) pairs of its group
    const int pairs = (nthreads >> 7) * 32;
    auto slots_of = [&](int columns) { return (pairs / words_of(columns)) & ~1; };
    bool active;
    int ws, slot, strand, mode_slots;
    auto set_mode = [&](int columns) {
        mode = columns;
        const int wpr = words_of(columns);
        mode_slots = slots_of(columns);
        active = pair < wpr * mode_slots;
        ws = pair % wpr;
        slot = pair / wpr;
        strand = slot & 1;
        // masks of the typical read (all flank bases on the contig, at least L columns), by window word
        for (int w = tid; w < wpr; w += nthreads) {
            uint32_t aligned, flank;
            if (columns) {
                aligned = bit_range(A - 32 * w, A + columns - 32 * w);
                flank = bit_range(-32 * w, A - 32 * w) | bit_range(A + columns - 32 * w, 2 * A + columns - 32 * w);
            } else if (w < NWA) {
                aligned = bit_range(A - 32 * w, A + L - 32 * w);
                flank = bit_range(-32 * w, A - 32 * w);
            } else {
                // right-anchored: bit J of the window is column C + A - 32 NWA + J
                const int k = w - NWA, top = 32 * NWA;
                aligned = bit_range(top - A - L - 32 * k, top - A - 32 * k);
                flank = bit_range(top - A - 32 * k, top - 32 * k);
            }
            s_mask[2 * w] = aligned;
            s_mask[2 * w + 1] = flank;
        }
    };
    set_mode(0);

    uint32_t cnt[PL_CLASSES][PL_REG];
#pragma unroll
    for (int c = 0; c < PL_CLASSES; ++c)
#pragma unroll
        for (int k = 0; k < PL_REG; ++k) cnt[c][k] = 0;
    int n_iter = 0;  // eight-read iterations since planes 4 .. 7 were moved up
    uint32_t *const my_wide = s_wide + tid;  // plane k of class c at my_wide[(k * PL_CLASSES + c) * nthreads]

    // planes 4 .. 7 of the register counters -> the wide counters in shared memory (a ripple-carry add, plane by plane)
    auto spill = [&]() {
#pragma unroll
        for (int c = 0; c < PL_CLASSES; ++c) {
            uint32_t *w = my_wide + c * nthreads;
            uint32_t carry = 0;
#pragma unroll
            for (int k = 0; k < PL_REG - 4; ++k) {
                const uint32_t wk = w[k * PL_CLASSES * nthreads], v = cnt[c][4 + k];
                w[k * PL_CLASSES * nthreads] = wk ^ v ^ carry;
                carry = (wk & v) | ((wk ^ v) & carry);
                cnt[c][4 + k] = 0;
            }
            for (int k = PL_REG - 4; carry && k < PL_WIDE; ++k) {
                const uint32_t wk = w[k * PL_CLASSES * nthreads];
                w[k * PL_CLASSES * nthreads] = wk ^ carry;
                carry &= wk;
            }
        }
        n_iter = 0;
    };

    const int LA = L + A;
    auto add_cell = [&](int canchor, int cstrand, int cls, int pos, unsigned long long sum) {
        // window position `pos` of an anchor -> table cell; classes are complemented on the reverse strand
        const int es = (canchor ^ cstrand) * 2 + cstrand;
        if (pos >= 0) {
            if (cls < 4) {
                const int gb = cstrand ? 3 - cls : cls;
                atomicAdd(t.misincorp + ((size_t)es * MDG_N_CLASSES + gb) * L + pos, sum);
            } else if (cls < 8) {
                const int rb = cstrand ? 3 - (cls - 4) : cls - 4;
                atomicAdd(t.dnacomp + ((size_t)es * 4 + rb) * LA + pos, sum);
            } else {
                int gb = (cls - 8) / 3, rb = (cls - 8) % 3;
                rb += rb >= gb ? 1 : 0;
                if (cstrand) { gb = 3 - gb; rb = 3 - rb; }
                atomicAdd(t.misincorp + ((size_t)es * MDG_N_CLASSES + 4 + 5 * gb + rb) * L + pos, sum);
            }
        } else if (cls < 4) {
            const int gb = cstrand ? 3 - cls : cls;
            atomicAdd(t.dnacomp + ((size_t)es * 4 + gb) * LA + L - pos - 1, sum);
        }
    };
    // reduces the block's counters into the 64-bit tables (end of the kernel, on mode changes, and before a
    // thread's 12-bit counters could overflow).  The stage area is free at that point: it holds the block's sums,
    // [strand][class 0..19][window bit].
    auto flush_block = [&]() {
        spill();
        __syncthreads();
        const int wpr = words_of(mode), bits = 32 * wpr;
        uint32_t *const red = s_stage;
        for (int i = tid; i < 2 * 20 * bits; i += nthreads) red[i] = 0;
        __syncthreads();
        if (active) {
#pragma unroll
            for (int c = 0; c < PL_CLASSES; ++c) {
                // class of the tables: R_g, H_g
                const int cls = c == 0 ? group : 4 + group;
                uint32_t pl[4 + PL_WIDE];
#pragma unroll
                for (int k = 0; k < 4; ++k) pl[k] = cnt[c][k];
#pragma unroll
                for (int k = 0; k < PL_WIDE; ++k) pl[4 + k] = my_wide[(k * PL_CLASSES + c) * nthreads];
                uint32_t any = 0;
#pragma unroll
                for (int k = 0; k < 4 + PL_WIDE; ++k) any |= pl[k];
                uint32_t *const to = red + ((size_t)strand * 20 + cls) * bits + 32 * ws;
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
#pragma unroll
        for (int c = 0; c < PL_CLASSES; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) cnt[c][k] = 0;
        __syncthreads();
        for (int i = tid; i < PL_WIDE * PL_CLASSES * nthreads; i += nthreads) s_wide[i] = 0;
        for (int cell = tid; cell < 2 * 20 * bits; cell += nthreads) {
            const int bit = cell % bits, cls = (cell / bits) % 20, cstrand = cell / (20 * bits);
            unsigned long long sum = red[cell];
            if (cls >= 8) {  // substitution classes: counted one event at a time
                uint32_t *const from = s_sub + ((size_t)cstrand * 12 + (cls - 8)) * 32 * WPR_MAX + bit;
                sum = *from;
                *from = 0;
            }
            if (!sum) continue;
            if (mode) {
                const int pos = bit - A;  // column
                if (pos < 0) add_cell(0, cstrand, cls, pos, sum);                          // left flank
                else if (pos >= mode) add_cell(1, cstrand, cls, mode - 1 - pos, sum);      // right flank at distance pos - C + 1
                else {
                    if (pos < L) add_cell(0, cstrand, cls, pos, sum);
                    if (mode - 1 - pos < L) add_cell(1, cstrand, cls, mode - 1 - pos, sum);
                }
            } else if (bit < 32 * NWA) {
                const int pos = bit - A;
                if (pos < L) add_cell(0, cstrand, cls, pos, sum);
            } else {
                const int pos = 32 * NWA - A - 1 - (bit - 32 * NWA);  // columns from the right end; negative: flank
                if (pos < L) add_cell(1, cstrand, cls, pos, sum);
            }
        }
        __syncthreads();
    };
    const int min_slots = max(2, (pairs / WPR_MAX) & ~1);
    // a thread counts at most ceil(T / (slots / 2)) reads per tile, rounded up to whole iterations of eight
    const int per_tile = ((T + (min_slots >> 1) - 1) / (min_slots >> 1) + 7) & ~7;
    // twelve counter bits (register planes 0-3, wide planes 4-11): 4095 reads of one class at one position at most
    const int flush_period = g.flush_tiles > 0 ? g.flush_tiles : max(1, 3800 / per_tile);
    int tiles_since_flush = 0;
    bool dirty = false;

    // ---- stage: the plane words of one window of one read ----
    // window bit J (word k = J >> 5) is column c_start + J of the alignment.  kNW > 0: the word count is known at
    // compile time, all loads of the window are issued before the first use; kNW = 0: any count, word by word.
    auto stage_window = [&](auto nw_tag, const PlaneRecord &rec, uint32_t *row_at, int first_word, int n_words, int c_start, int side,
                            int slab_w0, int slab_words, int rstrand) {
        constexpr int kNW = decltype(nw_tag)::value;
        const int cols = (int)(rec.cols & 0x7FFF);
        const int v = (int)(rec.misc & 0xFFFF);
        const int lf = (int)((rec.cols >> 16) & 0xFF), rf = (int)(rec.cols >> 24);
        const bool typical = lf == A && rf == A && (mode || v == L);
        // read: nibble index of window bit 0, eight bases per seq4 word
        const int64_t qn = (int64_t)rec.q0 + c_start;
        const int qs = (int)(qn & 7);
        // from the tile's copy in shared memory when the window's words all lie inside it, else from global memory
        const int64_t qw = qn >> 3, in_slab = qw - slab_w0;
        const bool from_smem = slab_words > 0 && in_slab >= 0 && in_slab + 4 * (kNW > 0 ? kNW : n_words) + 1 <= slab_words;
        const uint32_t *const qs_ptr = s_seq + (from_smem ? in_slab : 0), *const qg_ptr = seq32 + qw;
        auto seq_word = [&](int m) { return from_smem ? qs_ptr[m] : __ldg(qg_ptr + m); };
        // genome: 32 bases per uint4
        const int64_t rn = ((int64_t)rec.rg << 5) + (int)((rec.misc >> 16) & 31) + c_start;
        const uint4 *rp = planes + (rn >> 5);
        const int rs = (int)(rn & 31);
        // one window word from five seq4 words (as BAM stores them) and two genome entries.  Plane p of eight bases:
        // bit p of every nibble (one AND), the two nibbles of a byte side by side in base order (two shifts, one
        // OR-AND), and the four bytes' bit pairs gathered into the top byte by one multiplication -- the shifts and the
        // multiplication issue on the FMA pipe, which this kernel leaves idle, only two operations on the ALU pipe.
        // q4[] carries the fifth word's plane bytes over to the next window word (its first).
        auto plane_byte = [](uint32_t w, int pl) {
            const uint32_t y = (pl ? shr_fma(w, pl) : w) & 0x11111111u;
            return ((shr_fma(y, 4) | shl_fma(y, 1)) & 0x03030303u) * 0x01041040u;  // the plane's eight bits in bits 24..31
        };
        auto emit = [&](int k, uint32_t (&q0)[4], uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, const uint4 &g_lo, const uint4 &g_hi) {
            uint32_t xp[4];
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) {
                const uint32_t q1 = plane_byte(w1, pl), q2 = plane_byte(w2, pl), q3 = plane_byte(w3, pl), q4 = plane_byte(w4, pl);
                const uint32_t lo = __byte_perm(__byte_perm(q0[pl], q1, 0x0073), __byte_perm(q2, q3, 0x0073), 0x5410);
                xp[pl] = __funnelshift_r(lo, q4 >> 24, qs);
                q0[pl] = q4;
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
            // a column counts only when the read base is A/C/G/T (statistics.py:27): exactly one X plane set.  The
            // genome planes are all zero for anything that is not A/C/G/T.  Flank bits carry the reference base alone
            // (statistics.py:85-93).
            const uint32_t one = (xa ^ xc ^ xg ^ xt) & ~((xa & xc) | (xg & xt));
            const uint32_t keep = one & aligned, keep_y = keep | flank;
            uint4 xs, ys;
            xs.x = xa & keep; xs.y = xc & keep; xs.z = xg & keep; xs.w = xt & keep;
            ys.x = ya & keep_y; ys.y = yc & keep_y; ys.z = yg & keep_y; ys.w = yt & keep_y;
            uint4 *out = (uint4 *)(row_at + 8 * (first_word + k));
            out[0] = xs;
            out[1] = ys;
            // substitutions (reference g read as another base b) are rare events: one shared-memory atomic each on the
            // block's table, here where one thread sees all of a read's window word; the counting loop keeps the
            // dense classes R_g and H_g only
            uint32_t ev = (ys.x | ys.y | ys.z | ys.w) & keep & ~((xs.x & ys.x) | (xs.y & ys.y) | (xs.z & ys.z) | (xs.w & ys.w));
            if (ev) {
                // base code of a one-hot quadruple of planes: bit 0 from planes C | T, bit 1 from planes G | T
                const uint32_t g1 = ys.y | ys.w, g2 = ys.z | ys.w, r1 = xs.y | xs.w, r2 = xs.z | xs.w;
                uint32_t *const to = s_sub + (size_t)rstrand * 12 * 32 * WPR_MAX + 32 * (first_word + k);
                do {
                    const int j = __ffs(ev) - 1;
                    ev &= ev - 1;
                    const int gb = (int)((g1 >> j) & 1u) + 2 * (int)((g2 >> j) & 1u);
                    int rb = (int)((r1 >> j) & 1u) + 2 * (int)((r2 >> j) & 1u);
                    rb -= rb > gb ? 1 : 0;
                    atomicAdd(to + (3 * gb + rb) * 32 * WPR_MAX + j, 1u);
                } while (ev);
            }
        };
        if constexpr (kNW > 0) {
            uint32_t w[4 * kNW + 1];
            uint4 gw[kNW + 1];
#pragma unroll
            for (int m = 0; m <= 4 * kNW; ++m) w[m] = seq_word(m);
#pragma unroll
            for (int k = 0; k <= kNW; ++k) gw[k] = __ldg(rp + k);
            uint32_t q0[4];
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) q0[pl] = plane_byte(w[0], pl);
#pragma unroll
            for (int k = 0; k < kNW; ++k) emit(k, q0, w[4 * k + 1], w[4 * k + 2], w[4 * k + 3], w[4 * k + 4], gw[k], gw[k + 1]);
        } else {
            uint32_t q0[4];
            {
                const uint32_t w0 = seq_word(0);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) q0[pl] = plane_byte(w0, pl);
            }
            uint4 g_lo = __ldg(rp);
            for (int k = 0; k < n_words; ++k) {
                const uint32_t w1 = seq_word(4 * k + 1), w2 = seq_word(4 * k + 2), w3 = seq_word(4 * k + 3), w4 = seq_word(4 * k + 4);
                const uint4 g_hi = __ldg(rp + k + 1);
                emit(k, q0, w1, w2, w3, w4, g_lo, g_hi);
                g_lo = g_hi;
            }
        }
    };

    __shared__ uint32_t indel_here;
    if (tid == 0) indel_here = 0;
    struct Header {
        uint32_t index, flag, lib, l_seq, boff, c0, c1, cig0;
        int32_t tid_ref, pos;
        bool live;
    };
    // ---- parse of one read: filter, classify, per-read events (statistics.py:37-51,117-126) ----
    // kind: 0 nothing to do, 1 gap-free (record made), 2 for the general kernel, 3 one short indel
    auto parse_read = [&](const Header &h, int64_t r, int &kind, int &rstrand, uint32_t &columns, PlaneRecord &rec) {
        kind = 0;
        rstrand = 0;
        columns = 0;
        if (!h.live || (h.flag & FILTERED_FLAGS)) return;
        if (h.lib >= (uint32_t)p.n_lib) {
            atomicCAS(t.error_flag, 0, DATA_ERR_LIB);
            return;
        }
        if (subset && h.lib != (uint32_t)sub.lib) return;  // cannot happen: the list is grouped by library
        if (h.tid_ref < 0 || h.tid_ref >= ref.n_contigs) {
            atomicCAS(t.error_flag, 0, DATA_ERR_TID);
            return;
        }
        rstrand = (h.flag >> 4) & 1;
        uint32_t lead = 0, trail = 0, cols = 0, gap_len = 0, gap_del = 0;
        int state = 0, n_lead = 0, n_trail = 0;
        bool simple = h.c1 > h.c0;
        if (h.c1 - h.c0 == 1 && ((0x181u >> (h.cig0 & 0xF)) & 1u)) {
            // one match block (M, = or X): nearly every read of an untrimmed library
            cols = h.cig0 >> 4;
            state = 1;
        } else
        for (uint32_t k = h.c0; k < h.c1 && simple; ++k) {
            const uint32_t w = k == h.c0 ? h.cig0 : __ldg(b.cigar + k), op = w & 0xF, len = w >> 4;
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
        const int64_t pos = h.pos;
        const int64_t contig_len = ref.contig_len[h.tid_ref];
        const uint64_t ref0 = ref.contig_off[h.tid_ref] + (uint64_t)(pos > 0 ? pos : 0);
        simple = simple && state >= 1 && cols > 0 && cols < 32768 && n_lead <= 1 && n_trail <= 1 &&
                 (uint64_t)lead + n_query + trail == h.l_seq && pos >= 0 && pos + (int64_t)ref_span <= contig_len &&
                 ref0 < (1ull << 33);
        if (!simple) {
            kind = 2;
            return;
        }
        if (gap_len) {
            // count_staged_kernel's indel variant parses this read again and does all of its bookkeeping
            kind = 3;
            return;
        }
        kind = 1;
        columns = cols;
        const int64_t aend = pos + ref_span;
        const uint32_t lf = (uint32_t)min((int64_t)A, pos);
        const uint32_t rf = (uint32_t)min((int64_t)A, contig_len - aend);
        const uint64_t q0 = (uint64_t)h.boff + lead;
        rec.q0 = (uint32_t)q0;
        rec.rg = (uint32_t)(ref0 >> 5);
        rec.cols = cols | (lf << 16) | (rf << 24);
        rec.misc = min(cols, (uint32_t)L) | (uint32_t)(ref0 & 31) << 16;
        // FragmentLengths.update, statistics.py:117-126
        int64_t length = -1;
        int lkind = 0;
        if (h.flag & 0x1) {
            if ((h.flag & 0x40) && (h.flag & 0x2)) {
                const int64_t tl = b.tlen[r];
                length = tl < 0 ? -tl : tl;
            }
        } else {
            lkind = 1;
            length = ref_span;
        }
        if (length >= 0) {
            if (length < MDG_LG_SMEM_BINS && length < p.lg_bins) {
                atomicAdd(s_lg + (lkind * 2 + rstrand) * MDG_LG_SMEM_BINS + length, 1u);
            } else if (length < p.lg_bins) {
                atomicAdd(t.lghist + (size_t)(lkind * 2 + rstrand) * p.lg_bins + length, 1ull);
            } else {
                const unsigned long long at = atomicAdd(t.lg_overflow_count, 1ull);
                if ((int64_t)at < t.lg_overflow_cap) {
                    int32_t *row = t.lg_overflow_rows + at * 4;
                    row[0] = sub.list ? sub.lib : 0; row[1] = lkind; row[2] = rstrand; row[3] = (int32_t)length;
                }
            }
        }
        // update_soft_clipping, statistics.py:37-51
        if (lead) {
            const int end = rstrand ? 1 : 0, lim = (int)min(lead, (uint32_t)L);
            for (int i = 0; i < lim; ++i) atomicAdd(s_clip + (end * 2 + rstrand) * L + i, 1u);
        }
        if (trail) {
            const int end = rstrand ? 0 : 1, lim = (int)min(trail, (uint32_t)L);
            for (int i = 0; i < lim; ++i) atomicAdd(s_clip + (end * 2 + rstrand) * L + i, 1u);
        }
    };

    // L2 prefetch of a tile two ahead (the same scheme as count_staged_kernel)
    const int per = (T + nthreads - 1) / nthreads;
    auto prefetch_headers = [&](int64_t tile_index, uint32_t &boff, uint32_t &coff) {
        const int64_t start = tile_index * T;
        const int64_t rn = start + (int64_t)tid * per;
        const bool live = tid * per < T && rn < b.n_reads;
        if (live) {
            boff = b.base_off[rn];
            coff = b.cigar_off[rn];
        }
        const int64_t r4 = start + (int64_t)tid * 32;
        if (tid * 32 < T && r4 < b.n_reads) {
            prefetch_l2(b.tid + r4);
            prefetch_l2(b.pos + r4);
            prefetch_l2(b.l_seq + r4);
            prefetch_l2(b.tlen + r4);
            if (!(tid & 1)) {
                prefetch_l2(b.flag + r4);
                prefetch_l2(b.lib + r4);
            }
        }
        return live;
    };
    auto prefetch_bases = [&](uint32_t boff, uint32_t coff) {
        const char *seq_at = (const char *)b.seq4 + (boff >> 1);
        for (int off = 0; off < per * 80; off += 128) prefetch_l2(seq_at + off);
        prefetch_l2(b.cigar + coff);
    };
    auto prefetch_listed_headers = [&](int64_t tile_index, uint32_t (&boff)[2], uint32_t (&coff)[2]) {
        uint32_t live = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int q = tid + u * nthreads;
            const int64_t at = tile_index * T + q;
            if (q < T && at < n_todo) {
                const uint32_t r = subset[at];
                boff[u] = b.base_off[r];
                coff[u] = b.cigar_off[r];
                prefetch_l2(b.flag + r);
                prefetch_l2(b.tid + r);
                prefetch_l2(b.pos + r);
                prefetch_l2(b.l_seq + r);
                prefetch_l2(b.tlen + r);
                live |= 1u << u;
            }
        }
        return live;
    };
    auto prefetch_listed_bases = [&](uint32_t live, const uint32_t (&boff)[2], const uint32_t (&coff)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!(live >> u & 1)) continue;
            const char *seq_at = (const char *)b.seq4 + (boff[u] >> 1);
            prefetch_l2(seq_at);
            prefetch_l2(seq_at + 127);
            prefetch_l2(b.cigar + coff[u]);
        }
    };

    constexpr int PREP = 2;
    const int64_t n_tiles = (n_todo + T - 1) / T;
    if (!subset) {
        uint32_t boff0 = 0, coff0 = 0, boff1 = 0, coff1 = 0;
        const bool live0 = prefetch_headers(blockIdx.x, boff0, coff0);
        const bool live1 = prefetch_headers(blockIdx.x + (int64_t)gridDim.x, boff1, coff1);
        if (live0) prefetch_bases(boff0, coff0);
        if (live1) prefetch_bases(boff1, coff1);
    } else {
        uint32_t boff1[2] = {0, 0}, coff1[2] = {0, 0};
        const uint32_t live1 = prefetch_listed_headers(blockIdx.x + (int64_t)gridDim.x, boff1, coff1);
        prefetch_listed_bases(live1, boff1, coff1);
    }
    // one thread: the stretch of seq4 the reads of a tile occupy (reads are laid out in order: a stage thread checks
    // that its read really lies inside), 32 bytes more in front and 48 behind for the windows' flanks, as one bulk copy
    auto issue_slab = [&](int64_t tile_index) {
        s_slab[1] = 0;
        if (subset || g.seq_words <= 0 || tile_index >= n_tiles) return;
        const int64_t r0 = tile_index * T, r1 = min(n_todo, r0 + (int64_t)T) - 1;
        const uint64_t first = b.base_off[r0], last = (uint64_t)b.base_off[r1] + b.l_seq[r1];
        const int64_t lo = ((int64_t)(first >> 1) - 32) & ~15ll, hi = ((int64_t)((last + 1) >> 1) + 48 + 15) & ~15ll;
        const int64_t bytes = hi - lo;
        if (bytes <= 0 || bytes > 4ll * g.seq_words) return;
        s_slab[0] = (int32_t)(lo >> 2);
        s_slab[1] = (int32_t)(bytes >> 2);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_seq);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"((uint32_t)bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"((const char *)b.seq4 + lo), "r"((uint32_t)bytes), "r"(mbar_addr)
                     : "memory");
    };
    int tile_parity = 0;
    if (tid < 6) s_ctl_base[tid] = tid == 3 ? 0xffffffffu : 0u;
    __syncthreads();
    if (tid == 0) issue_slab(blockIdx.x);
#ifdef MDG_PHASE_CLOCKS
    __shared__ unsigned int s_pc[8];
    if (tid < 8) s_pc[tid] = 0;
    __syncthreads();
    unsigned int pc_last = (unsigned int)clock();
#define MDG_PPHASE(i)                                    \
    if (tid == 0) {                                      \
        const unsigned int now_ = (unsigned int)clock(); \
        s_pc[i] += now_ - pc_last;                       \
        pc_last = now_;                                  \
    }
#else
#define MDG_PPHASE(i)
#endif
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t *const s_ctl = s_ctl_base + 8 * (tile_parity & 1);
        uint32_t *const s_ctl_next = s_ctl_base + 8 * ((tile_parity & 1) ^ 1);
        ++tile_parity;

        // ---- parse ----
        const int64_t tile_start = tile * T;
        for (int q0 = 0; q0 < T; q0 += nthreads * PREP) {
            Header h[PREP];
#pragma unroll
            for (int u = 0; u < PREP; ++u) {
                const int q = q0 + u * nthreads + tid;
                h[u].live = q < T && tile_start + q < n_todo;
                const int64_t r = !h[u].live ? 0 : subset ? (int64_t)subset[tile_start + q] : tile_start + q;
                h[u].index = (uint32_t)r;
                if (h[u].live) {
                    h[u].flag = b.flag[r];
                    h[u].lib = b.lib[r];
                    h[u].tid_ref = b.tid[r];
                    h[u].pos = b.pos[r];
                    h[u].l_seq = b.l_seq[r];
                    h[u].boff = b.base_off[r];
                    h[u].c0 = b.cigar_off[r];
                    h[u].c1 = b.cigar_off[r + 1];
                }
            }
#pragma unroll
            for (int u = 0; u < PREP; ++u) h[u].cig0 = h[u].live && h[u].c1 > h[u].c0 ? __ldg(b.cigar + h[u].c0) : 0;
#pragma unroll
            for (int u = 0; u < PREP; ++u) {
                if (q0 + u * nthreads >= T) break;  // the same for every thread: a round nobody has a read in
                int kind, rstrand;
                uint32_t columns;
                PlaneRecord rec{};
                parse_read(h[u], h[u].index, kind, rstrand, columns, rec);
                if (g.uniform) {
                    const uint32_t lo = __reduce_min_sync(0xffffffffu, kind == 1 ? columns : 0xffffffffu);
                    const uint32_t hi = __reduce_max_sync(0xffffffffu, kind == 1 ? columns : 0u);
                    if (lane == 0 && hi) {
                        atomicMin(s_ctl + 3, lo);
                        atomicMax(s_ctl + 4, hi);
                    }
                }
                // warp-aggregated appends to the four lists
                const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
                for (int which = 0; which < 4; ++which) {
                    const bool mine = which == 2 ? kind == 2 : which == 3 ? kind == 3 : (kind == 1 && rstrand == which);
                    const uint32_t m = __ballot_sync(0xffffffffu, mine);
                    if (m) {
                        uint32_t base = 0;
                        if (lane == __ffs(m) - 1) base = atomicAdd(s_ctl + (which == 3 ? 5 : which), (uint32_t)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                        if (mine) {
                            const uint32_t at = base + __popc(m & lt);
                            if (which == 2) s_cx[at] = h[u].index;
                            else if (which == 3) s_ix[at] = h[u].index;
                            else s_rec[which == 0 ? at : T - 1 - at] = rec;
                        }
                    }
                }
            }
        }
        MDG_PPHASE(0)
        __syncthreads();
        MDG_PPHASE(1)

        // ---- reads this kernel does not count go to the two work lists ----
        if (tid < 32 && s_ctl[2]) {
            const uint32_t n_cx = s_ctl[2];
            unsigned long long base = 0;
            uint32_t *const wl = sub.list ? worklist + sub.offsets[sub.lib] : worklist;
            if (lane == 0) base = atomicAdd(work_count + (sub.list ? sub.lib : 0), (unsigned long long)n_cx);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t i = lane; i < n_cx; i += 32) wl[base + i] = s_cx[i];
        }
        if (tid >= 32 && tid < 64 && s_ctl[5]) {
            const uint32_t n_ix = s_ctl[5];
            unsigned long long base = 0;
            uint32_t *const il = sub.list ? indel_list + sub.offsets[sub.lib] : indel_list;
            if (lane == 0) {
                base = atomicAdd(indel_count + (sub.list ? sub.lib : 0), (unsigned long long)n_ix);
                if (g.indel_seen) atomicAdd(g.indel_seen, (unsigned long long)n_ix);
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t i = lane; i < n_ix; i += 32) il[base + i] = s_ix[i];
        }

        // ---- one window per read when every gap-free read of the tile has the same length ----
        {
            int want = 0;
            const uint32_t lo = s_ctl[3], hi = s_ctl[4];
            if (g.uniform && lo == hi && hi > 0) {
                const int words = ((int)hi + 2 * A + 31) / 32;
                if (words < WPR_MAX && pairs / words >= 2) want = (int)hi;
            }
            if (want != mode) {
                if (dirty) flush_block();
                set_mode(want);
                __syncthreads();
                dirty = false;
                tiles_since_flush = 0;
            }
            dirty = dirty || s_ctl[0] + s_ctl[1] > 0;
        }

        // ---- pull the tile after next towards L2 ----
        uint32_t ahead_boff = 0, ahead_coff = 0;
        const bool ahead_live = !subset && prefetch_headers(tile + 2 * (int64_t)gridDim.x, ahead_boff, ahead_coff);
        uint32_t listed_boff[2] = {0, 0}, listed_coff[2] = {0, 0};
        const uint32_t listed_live = subset ? prefetch_listed_headers(tile + 2 * (int64_t)gridDim.x, listed_boff, listed_coff) : 0u;

        // ---- stage: one thread per (read, window) ----
        const int n_fwd = (int)s_ctl[0], n_rev = (int)s_ctl[1];
        const int slab_w0 = s_slab[0], slab_words = s_slab[1];
        if (slab_words) {
            // the bulk copy of this tile's seq4 stretch was issued a tile ago: it has long landed
            uint32_t landed;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(landed)
                             : "r"(mbar_addr), "r"(slab_phase & 1u)
                             : "memory");
            } while (!landed);
            ++slab_phase;
        }
        {
            const int n_windows = mode ? 1 : 2, n_items = (n_fwd + n_rev) * n_windows;
            const int wpr = words_of(mode);
            for (int item = tid; item < n_items; item += nthreads) {
                const int li = mode ? item : item >> 1, side = mode ? 0 : item & 1;
                const int row = li < n_fwd ? li : T - 1 - (li - n_fwd);
                const PlaneRecord rec = s_rec[row];
                uint32_t *const row_at = s_stage + (size_t)row * ROW;
                const int n_words = mode ? wpr : NWA, first_word = side ? NWA : 0;
                const int c_start = side ? (int)(rec.cols & 0x7FFF) + A - 32 * NWA : -A;
                if (n_words == 4) stage_window(std::integral_constant<int, 4>{}, rec, row_at, first_word, 4, c_start, side, slab_w0, slab_words, li >= n_fwd);
                else if (n_words == 3) stage_window(std::integral_constant<int, 3>{}, rec, row_at, first_word, 3, c_start, side, slab_w0, slab_words, li >= n_fwd);
                else stage_window(std::integral_constant<int, 0>{}, rec, row_at, first_word, n_words, c_start, side, slab_w0, slab_words, li >= n_fwd);
            }
        }
        if (ahead_live) {
            if (g.prefetch_bases || !slab_words) prefetch_bases(ahead_boff, ahead_coff);
            else prefetch_l2(b.cigar + ahead_coff);
        }
        if (listed_live) prefetch_listed_bases(listed_live, listed_boff, listed_coff);
        MDG_PPHASE(2)
        __syncthreads();
        MDG_PPHASE(3)

        if (tid < 6) s_ctl_next[tid] = tid == 3 ? 0xffffffffu : 0u;
        if (tid == 0) issue_slab(tile + gridDim.x);  // lands while this tile is counted and the next one parsed
        // ---- count: this thread's window word and reference base, every stride-th read of its strand ----
        if (active) {
            const int n_mine = strand ? n_rev : n_fwd;
            const int stride = mode_slots >> 1;
            const int row_step = (strand ? -stride : stride) * ROW;
            const uint32_t *at0 = s_stage + (size_t)(strand ? T - 1 - (slot >> 1) : (slot >> 1)) * ROW + 8 * ws;
            // the class wiring is compile-time: a warp is uniform in the reference base G of its classes
            auto count_tile = [&](auto gtag) {
                constexpr int G = decltype(gtag)::value;
                const uint32_t *at = at0;
                for (int i = slot >> 1; i < n_mine; i += 8 * stride) {
                    uint32_t xg[8], y[8];
                    if (i + 7 * stride < n_mine) {  // eight reads in hand: no tests
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            xg[u] = at[u * row_step + G];
                            y[u] = at[u * row_step + 4 + G];
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const bool live = i + u * stride < n_mine;
                            xg[u] = live ? at[u * row_step + G] : 0u;
                            y[u] = live ? at[u * row_step + 4 + G] : 0u;
                        }
                    }
                    at += 8 * row_step;
                    MDG_ADD8(cnt[0], y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7])  // R_g
                    MDG_ADD8(cnt[1], xg[0], xg[1], xg[2], xg[3], xg[4], xg[5], xg[6], xg[7])  // H_g
                    if (++n_iter == 30) spill();  // planes 0-3 hold at most 15, thirty more iterations add 240: 255 fits eight planes
                }
            };
            switch (group) {
            case 0: count_tile(std::integral_constant<int, 0>{}); break;
            case 1: count_tile(std::integral_constant<int, 1>{}); break;
            case 2: count_tile(std::integral_constant<int, 2>{}); break;
            default: count_tile(std::integral_constant<int, 3>{}); break;
            }
        }
        MDG_PPHASE(4)
        __syncthreads();
        MDG_PPHASE(5)
        if (++tiles_since_flush == flush_period) {
            flush_block();
            tiles_since_flush = 0;
            dirty = false;
        }
    }
#ifdef MDG_PHASE_CLOCKS
    __syncthreads();
    if (blockIdx.x == 3 && tid < 8) mdg_plane_phase_dump[tid] = s_pc[tid];
#endif

    flush_block();
    for (int i = tid; i < 4 * MDG_LG_SMEM_BINS; i += nthreads) {
        const uint32_t v = s_lg[i];
        if (v) atomicAdd(t.lghist + (size_t)(i / MDG_LG_SMEM_BINS) * p.lg_bins + i % MDG_LG_SMEM_BINS, (unsigned long long)v);
    }
    for (int i = tid; i < 4 * L; i += nthreads) {
        const uint32_t v = s_clip[i];
        if (v) atomicAdd(t.misincorp + ((size_t)(i / L) * MDG_N_CLASSES + MDG_CLASS_SOFTCLIP) * L + i % L, (unsigned long long)v);
    }
}

// genome image in plane form: group k (bases 32 k .. 32 k + 31 of the packed stream) = {A, C, G, T} words
__global__ void __launch_bounds__(256) ref_planes_kernel(const uint32_t *__restrict__ one_hot, int64_t n_groups, uint4 *__restrict__ planes)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_groups; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t t0 = nibbles_to_planes(one_hot[4 * k]), t1 = nibbles_to_planes(one_hot[4 * k + 1]);
        const uint32_t t2 = nibbles_to_planes(one_hot[4 * k + 2]), t3 = nibbles_to_planes(one_hot[4 * k + 3]);
        uint4 out;
        out.x = __byte_perm(__byte_perm(t0, t1, 0x0040), __byte_perm(t2, t3, 0x0040), 0x5410);
        out.y = __byte_perm(__byte_perm(t0, t1, 0x0051), __byte_perm(t2, t3, 0x0051), 0x5410);
        out.z = __byte_perm(__byte_perm(t0, t1, 0x0062), __byte_perm(t2, t3, 0x0062), 0x5410);
        out.w = __byte_perm(__byte_perm(t0, t1, 0x0073), __byte_perm(t2, t3, 0x0073), 0x5410);
        planes[k] = out;
    }
}

}  // namespace mdg
