// Counting pass, bit-sliced kernel for gap-free reads ([H][S] M/=/X+ [S][H]).
//
// Same outputs as count_general_kernel (reference main.py:165-217: the read
// filter reader.py:121-132, FragmentLengths.update statistics.py:117-126,
// get_around align.py:22-35, align / align_with_qual align.py:38-73, revcomp
// main.py:200-205, update_soft_clipping statistics.py:37-51, both
// MisincorporationRates.update walks statistics.py:22-35 and
// DNAComposition.update_read / update_reference statistics.py:75-93), but
// organised the other way round: instead of a warp walking the columns of one
// read and firing one atomic per table cell, a THREAD owns eight consecutive
// table positions of one (anchor, strand) and the reads are its loop.
//
//   anchor 0 (left):  position p = alignment column           -> 5' table of forward reads, 3' of reverse reads
//   anchor 1 (right): position p = columns from the right end -> 3' table of forward reads, 5' of reverse reads
//   p < 0: flanking reference base at distance -p (DNAComposition.update_reference)
// The window of an anchor is positions [-A, L), cut into words of eight; every thread runs the same code.
//
// For a gap-free read, column == query index == reference offset, so the
// eight (read, reference) pairs a thread needs are one 32-bit word of the BAM
// 4-bit sequence and one word of the one-hot genome, both shifted into place
// with a funnel shift.  Every table class becomes a bit mask with one bit per
// nibble:
//     R_g  = reference base g         (4 classes)   -> misincorporation columns A,C,G,T / flank composition
//     H_b  = read base b              (4 classes)   -> read composition
//     P_gb = reference g, read b != g (12 classes)  -> substitution columns
// and is added, eight positions at a time, into 4-bit counters (registers),
// spilled every 15 reads into 8-bit counters (registers) and every 255 reads
// into the thread's private 16-bit counters in shared memory.  There is no
// atomic in the loop; the block reduces its private counters once, at the
// end, into the 64-bit tables.  Reads whose CIGAR has I/D/N/P (or anything
// else this layout cannot express) are appended to a work list for
// count_general_kernel.
#pragma once
#include "mdg_device.cuh"

namespace mdg {

constexpr int SWAR_CLASSES = 20;
constexpr int SWAR_L2_WORDS = 4 * SWAR_CLASSES;  // 16-bit counters: 20 classes x 8 positions
constexpr int SWAR_MAX_THREADS = 512;  // largest block any variant is compiled for
constexpr uint32_t K1 = 0x11111111u;

struct SwarGeom {
    int32_t words;         // 32-bit words per anchor window: ceil((A + L) / 8), window positions [-A, 8 words - A)
    int32_t slots;         // read slots per block (even: half per strand)
    int32_t threads;       // blockDim.x
    int32_t work_threads;  // 2 * words * slots
    int32_t tile;          // reads staged per iteration of the block
    int32_t uniform;       // 1: tiles whose gap-free reads all have the same length are counted in one window per read
    int32_t flush_tiles;   // > 0: reduce the private counters every so many tiles (tests); 0: only when they could overflow
};

// One staged read (16 bytes), shared by both anchors.  Base indices are kept as (32-bit word, nibble)
// so that the loop needs no 64-bit arithmetic.
struct __align__(16) SwarRecord {
    uint32_t qi;    // 32-bit word of seq4 holding the first aligned base (base_off + leading clip)
    uint32_t ri;    // 32-bit word of the genome holding the first aligned column
    uint32_t cols;  // columns (15 bits) | has_qual << 15 | left flank bases << 16 | right flank bases << 24
    uint32_t misc;  // min(L, columns) | nibble of the genome word << 16 | nibble of the seq4 word << 20
};

__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// BAM packs the first base of a byte in the high nibble; make nibble i of the word base i
__device__ __forceinline__ uint32_t natural_order(uint32_t w)
{
    return ((w & 0x0F0F0F0Fu) << 4) | ((w >> 4) & 0x0F0F0F0Fu);
}

// bit 4i set iff nibble i has exactly one bit set (A, C, G or T in BAM code)
__device__ __forceinline__ uint32_t one_hot_nibbles(uint32_t x)
{
    const uint32_t b = x >> 1, c = x >> 2, d = x >> 3;
    const uint32_t exactly_one = (x ^ b ^ c) & ~(x & b & c);  // of bits 0..2
    const uint32_t none = ~(x | b | c);
    return ((exactly_one & ~d) | (none & d)) & K1;
}

// the low n nibbles, n = bits / 4 clamped to [0, 8]
__device__ __forceinline__ uint32_t low_nibbles(int bits) { return __funnelshift_lc(0xffffffffu, 0u, max(bits, 0)); }

// With several libraries the kernel runs once per library over that library's reads (partition_by_library):
// the thread-private counters know nothing of libraries, the tables `t` points at are that library's.
struct SwarSubset {
    const uint32_t *list;               // read indices grouped by library, or null: every read of the batch
    const unsigned long long *offsets;  // [n_lib + 1] into list
    int32_t lib;
};

// thread t = ((slot * 2) + anchor) * words + word
// kReads = reads a thread counts per loop iteration: their instruction streams interleave (ILP) and their
// class masks are summed before they touch the counters (one 3-input add per class and pair of reads).
template <bool kQual, int kMaxThreads, int kReads, int kBlocksPerSm = 1>
__global__ void __launch_bounds__(kMaxThreads, kBlocksPerSm)
count_swar_kernel(DevBatch b, DevRef ref, CountParams p, CountTables t, SwarGeom g, uint32_t *__restrict__ worklist,
                  unsigned long long *__restrict__ work_count, SwarSubset sub)
{
    extern __shared__ uint32_t smem[];
    const int nthreads = g.threads, T = g.tile, L = p.L, A = p.A, W = g.words;
    const int l2_words = SWAR_L2_WORDS * nthreads;
    uint32_t *const s_l2 = smem;                                // [SWAR_L2_WORDS][nthreads] private 16-bit counters
    SwarRecord *const s_rec = (SwarRecord *)(s_l2 + l2_words);  // [T]: forward reads from the front, reverse from the back
    uint32_t *const s_cx = (uint32_t *)(s_rec + T);             // [T] complex reads of the tile
    uint32_t *const s_lg = s_cx + T;                            // [kind][strand][MDG_LG_SMEM_BINS]
    uint32_t *const s_clip = s_lg + 4 * MDG_LG_SMEM_BINS;       // [end][strand][L]
    uint32_t *const s_ctl = s_clip + 4 * L;                     // n_fwd, n_rev, n_cx, min / max columns

    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < l2_words; i += nthreads) s_l2[i] = 0;
    for (int i = tid; i < 4 * MDG_LG_SMEM_BINS + 4 * L; i += nthreads) s_lg[i] = 0;

    const uint32_t *__restrict__ seq32 = (const uint32_t *)b.seq4;
    const uint32_t *__restrict__ ref32 = ref.words;
    const uint32_t *const subset = sub.list ? sub.list + sub.offsets[sub.lib] : nullptr;
    const int64_t n_todo = sub.list ? (int64_t)(sub.offsets[sub.lib + 1] - sub.offsets[sub.lib]) : b.n_reads;

    // Thread geometry, a function of the block's mode (set_mode):
    //   mode 0 (two anchors): thread = ((slot * 2) + anchor) * W + word.  The window word covers positions
    //     pbase .. pbase + 7, pbase = 8 word - A; nibble i holds position pbase + i (left anchor) or
    //     pbase + 7 - i (right anchor: memory order runs towards the end).  With z = the nibble index of
    //     position 0 (left) / one past it (right):
    //       left:  aligned nibbles [z, z + v),  flank nibbles [z - f, z)      v = min(L, columns)
    //       right: aligned nibbles [z - v, z),  flank nibbles [z, z + f)      f = flank bases on the contig
    //   mode C > 0 (every gap-free read of the tile has C columns): one window [-A, C + A) per read, thread =
    //     slot * Wu + word; a column feeds the left-anchored table at p = column and the right-anchored one at
    //     p = C - 1 - column when the counters are reduced, so each base is visited once instead of twice.
    int mode = 0;
    auto words_of = [&](int columns) { return columns ? (columns + 2 * A + 7) / 8 : W; };
    auto slots_of = [&](int columns) { return columns ? (nthreads / words_of(columns)) & ~1 : g.slots; };
    bool active;
    int word, anchor, slot, strand, z4, s4, fshift, cbase, amul;
    uint32_t flip_a, side_a;
    auto set_mode = [&](int columns) {
        mode = columns;
        if (columns == 0) {
            active = tid < g.work_threads;
            word = tid % W;
            anchor = (tid / W) & 1;
            slot = tid / (2 * W);
        } else {
            const int mode_words = words_of(columns);
            active = tid < mode_words * slots_of(columns);
            word = tid % mode_words;
            anchor = 0;
            slot = tid / mode_words;
        }
        strand = slot & 1;
        const int pbase = 8 * word - A;
        z4 = 4 * (anchor ? pbase + 8 : -pbase);
        s4 = anchor ? -4 : 4;
        flip_a = anchor ? 0xffffffffu : 0u;
        side_a = anchor ? low_nibbles(z4) : ~low_nibbles(z4);  // the side of z aligned nibbles are on
        fshift = anchor ? 24 : 16;
        // base offset of nibble 0 from the first aligned base: left pbase; right columns - 8 - pbase
        cbase = anchor ? -8 - pbase : pbase;
        amul = anchor ? 1 : 0;
    };
    set_mode(0);

    uint32_t acc0[SWAR_CLASSES];  // 8 x 4-bit counters per class
    uint32_t acc1[16];            // classes 0..7 (reference / read bases): 2 x (4 x 8-bit) counters, even / odd nibbles
#pragma unroll
    for (int c = 0; c < SWAR_CLASSES; ++c) acc0[c] = 0;
#pragma unroll
    for (int w = 0; w < 16; ++w) acc1[w] = 0;
    int n0 = 0, n1 = 0;
    uint32_t *const my_l2 = s_l2 + tid;  // word w at my_l2[w * nthreads]

    auto spill1 = [&]() {  // 8-bit -> private 16-bit counters in shared memory
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            const uint32_t v = acc1[w];
            my_l2[(2 * w) * nthreads] += v & 0x00FF00FFu;
            my_l2[(2 * w + 1) * nthreads] += (v >> 8) & 0x00FF00FFu;
            acc1[w] = 0;
        }
        n1 = 0;
    };
    auto spill0 = [&]() {  // 4-bit counters: base classes -> 8-bit registers; substitution classes (mostly
                           // zero) straight into the 16-bit counters
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            acc1[2 * c] += acc0[c] & 0x0F0F0F0Fu;
            acc1[2 * c + 1] += (acc0[c] >> 4) & 0x0F0F0F0Fu;
            acc0[c] = 0;
        }
#pragma unroll
        for (int c = 8; c < SWAR_CLASSES; ++c) {
            const uint32_t v = acc0[c];
            if (v) {
                // 16-bit lane of (class, nibble): word 4 c + 2 (nibble & 1) + ((nibble >> 1) & 1), half nibble >> 2
                uint32_t *at = my_l2 + (4 * c) * nthreads;
                at[0] += v & 0x000F000Fu;
                at[nthreads] += (v >> 8) & 0x000F000Fu;
                at[2 * nthreads] += (v >> 4) & 0x000F000Fu;
                at[3 * nthreads] += (v >> 12) & 0x000F000Fu;
                acc0[c] = 0;
            }
        }
        n0 = 0;
        if (++n1 == 17) spill1();
    };

    // reduces the block's private counters into the 64-bit tables (end of the kernel, and before a
    // thread's 16-bit counters could overflow)
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
    auto flush_block = [&]() {
        if (n0) spill0();
        if (n1) spill1();
        __syncthreads();
        const int n_anchors = mode ? 1 : 2, mode_words = words_of(mode), mode_slots = slots_of(mode);
        const int n_cells = n_anchors * mode_words * 2 * SWAR_CLASSES * 8;  // anchor, word, strand, class, nibble
        for (int cell = tid; cell < n_cells; cell += nthreads) {
            int rest = cell;
            const int nib = rest & 7; rest >>= 3;
            const int cls = rest % SWAR_CLASSES; rest /= SWAR_CLASSES;
            const int cstrand = rest & 1; rest >>= 1;
            const int cword = rest % mode_words;
            const int canchor = rest / mode_words;
            const int pb = 8 * cword - A;
            const int pos = canchor ? pb + 7 - nib : pb + nib;
            if (pos < -A || (pos < 0 && cls >= 4)) continue;
            if (mode ? pos >= mode + A || (pos >= mode && cls >= 4) : pos >= L) continue;
            // 16-bit lane holding (cls, nib): 8-bit lane bl = nib >> 1 of acc1[2 * cls + (nib & 1)]
            const int w1 = 2 * cls + (nib & 1), bl = nib >> 1;
            const int w2 = 2 * w1 + (bl & 1), half = bl >> 1;
            unsigned long long sum = 0;
            for (int cslot = cstrand; cslot < mode_slots; cslot += 2) {
                const uint32_t v = s_l2[w2 * nthreads + (cslot * n_anchors + canchor) * mode_words + cword];
                sum += half ? v >> 16 : v & 0xFFFFu;
            }
            if (!sum) continue;
            if (!mode) {
                add_cell(canchor, cstrand, cls, pos, sum);
            } else if (pos < 0) {
                add_cell(0, cstrand, cls, pos, sum);  // left flank
            } else if (pos >= mode) {
                add_cell(1, cstrand, cls, mode - 1 - pos, sum);  // right flank at distance pos - C + 1
            } else {
                if (pos < L) add_cell(0, cstrand, cls, pos, sum);
                if (mode - 1 - pos < L) add_cell(1, cstrand, cls, mode - 1 - pos, sum);
            }
        }
        __syncthreads();
        for (int i = tid; i < l2_words; i += nthreads) s_l2[i] = 0;
        __syncthreads();
    };
    // worst case every read of a tile lands on one strand: T / (slots / 2) reads per thread per tile
    // (the uniform mode never has fewer slots than the two-anchor mode it replaces)
    const int flush_period = g.flush_tiles > 0 ? g.flush_tiles : max(1, 60000 / ((T + (g.slots >> 1) - 1) / (g.slots >> 1)));
    int tiles_since_flush = 0;
    bool dirty = false;  // counters hold counts of the current mode

    // ---- per-thread stages of the software pipeline: loads of read i+1 fly while read i is counted ----
    struct Stage {
        uint32_t w0, w1, r0, r1, aligned, flank, sh;
        uint32_t qa, qm, qz;
    };
    auto fetch = [&](const SwarRecord *at, Stage &st) {
        const SwarRecord rec = *at;
        const int v = (int)(rec.misc & 0xFFFF), f = (int)((rec.cols >> fshift) & 0xFF);
        if (mode) {
            // aligned nibbles [z, zr), zr = z + C; left flank nibbles [z - lf, z); right flank nibbles [zr, zr + rf)
            const int zr4 = z4 + 4 * mode;
            const uint32_t below_z = low_nibbles(z4), below_zr = low_nibbles(zr4);
            st.aligned = below_zr & ~below_z;
            st.flank = (below_z & ~low_nibbles(z4 - 4 * f)) | (low_nibbles(zr4 + 4 * (int)(rec.cols >> 24)) & ~below_zr);
        } else {
            st.aligned = (low_nibbles(z4 + s4 * v) ^ flip_a) & side_a;
            st.flank = (low_nibbles(z4 - s4 * f) ^ ~flip_a) & ~side_a;
        }
        st.sh = 0;
        st.w0 = st.w1 = st.r0 = st.r1 = 0;
        if (st.aligned | st.flank) {
            // nibble 0 of this word, relative to the first aligned base
            const int off = cbase + amul * (int)(rec.cols & 0x7FFF);
            const int tr = (int)((rec.misc >> 16) & 7) + off, tq = (int)(rec.misc >> 20) + off;
            const uint32_t *rp = ref32 + ((int)rec.ri + (tr >> 3));
            st.r0 = __ldg(rp);
            st.r1 = __ldg(rp + 1);
            st.sh = ((uint32_t)(tq & 7) << 2) | ((uint32_t)(tr & 7) << 10) | 0x10000u;
            const int qw = (int)rec.qi + (tq >> 3);
            if (st.aligned) {
                const uint32_t *qp = seq32 + qw;
                st.w0 = __ldg(qp);
                st.w1 = __ldg(qp + 1);
                if (kQual) {
                    if (rec.cols & 0x8000u) {
                        // qualities of the window's eight bases: three words from the aligned word below cover them
                        const uint32_t *q32 = (const uint32_t *)b.qual + 2 * (int64_t)qw + ((tq >> 2) & 1);
                        st.qa = __ldg(q32);
                        st.qm = __ldg(q32 + 1);
                        st.qz = __ldg(q32 + 2);
                        st.sh |= 0x20000u | ((uint32_t)(tq & 3) << 27);
                    }
                }
            } else {
                st.w0 = st.w1 = 0;
            }
        }
    };
    // the read / reference words of one staged read, masked down to the nibbles that count
    auto masked_words = [&](const Stage &st, uint32_t &x, uint32_t &xc, uint32_t &y) {
        y = __funnelshift_r(st.r0, st.r1, st.sh >> 8);
        x = __funnelshift_r(natural_order(st.w0), natural_order(st.w1), st.sh);
        // a column counts only when the read base is A/C/G/T (statistics.py:27); the reference side is
        // already 0 for anything that is not A/C/G/T.  Flank nibbles carry the reference base alone
        // (DNAComposition.update_reference, statistics.py:85-93).
        const uint32_t valid = (one_hot_nibbles(x) * 15u) & st.aligned;
        x &= valid;
        y &= valid | st.flank;
        xc = x;  // DNAComposition.update_read ignores the quality mask (statistics.py:75-83)
        if (kQual) {
            if (st.sh & 0x20000u) {
                // align_with_qual, align.py:67-71: bases below --min-basequal become N on both sides
                const int sq = (int)(st.sh >> 24);
                const uint32_t lo = __funnelshift_r(st.qa, st.qm, sq), hi = __funnelshift_r(st.qm, st.qz, sq);
                const uint32_t mq = (uint32_t)p.min_qual * 0x01010101u;
                // bit 7 of a byte of ((q | 0x80) - min_qual) is clear iff q < min_qual
                uint32_t zl = (~((lo | 0x80808080u) - mq) & 0x80808080u) >> 7;
                uint32_t zh = (~((hi | 0x80808080u) - mq) & 0x80808080u) >> 7;
                zl |= zl >> 4;
                zh |= zh >> 4;
                const uint32_t low = ((zl & 0x11u) | ((zl >> 8) & 0x1100u)) | (((zh & 0x11u) | ((zh >> 8) & 0x1100u)) << 16);
                const uint32_t keep = ~(low * 15u) | st.flank;  // there is no read base, hence no quality, on a flank
                x &= keep;
                y &= keep;
            }
        }
    };
    auto count = [&](const Stage (&st)[kReads]) {
        uint32_t x[kReads], xc[kReads], y[kReads];
        bool any = false;
#pragma unroll
        for (int r = 0; r < kReads; ++r) {
            any = any || st[r].sh != 0;
            masked_words(st[r], x[r], xc[r], y[r]);  // a stage with sh == 0 holds zero words and zero masks
        }
        if (!any) return;  // no nibble of this word counts for these reads
#define MDG_CLASS(c, expr)                                     \
        {                                                      \
            uint32_t sum = 0;                                  \
            _Pragma("unroll") for (int r = 0; r < kReads; ++r) \
            {                                                  \
                const uint32_t X = x[r], Y = y[r], XC = xc[r]; \
                (void)X; (void)Y; (void)XC;                    \
                sum += (expr) & K1;                            \
            }                                                  \
            acc0[c] += sum;                                    \
        }
        MDG_CLASS(4, XC)
        MDG_CLASS(5, XC >> 1)
        MDG_CLASS(6, XC >> 2)
        MDG_CLASS(7, XC >> 3)
        MDG_CLASS(0, Y)
        MDG_CLASS(1, Y >> 1)
        MDG_CLASS(2, Y >> 2)
        MDG_CLASS(3, Y >> 3)
        MDG_CLASS(8, Y & (X >> 1))          // A>C
        MDG_CLASS(9, Y & (X >> 2))          // A>G
        MDG_CLASS(10, Y & (X >> 3))         // A>T
        MDG_CLASS(11, (Y >> 1) & X)         // C>A
        MDG_CLASS(12, (Y >> 1) & (X >> 2))  // C>G
        MDG_CLASS(13, (Y >> 1) & (X >> 3))  // C>T
        MDG_CLASS(14, (Y >> 2) & X)         // G>A
        MDG_CLASS(15, (Y >> 2) & (X >> 1))  // G>C
        MDG_CLASS(16, (Y >> 2) & (X >> 3))  // G>T
        MDG_CLASS(17, (Y >> 3) & X)         // T>A
        MDG_CLASS(18, (Y >> 3) & (X >> 1))  // T>C
        MDG_CLASS(19, (Y >> 3) & (X >> 2))  // T>G
#undef MDG_CLASS
        n0 += kReads;
        if (n0 > 15 - kReads) spill0();
    };

    // ---- staging of one read: filter, classify, per-read events (statistics.py:37-51,117-126) ----
    struct Header {
        uint32_t index, flag, lib, l_seq, boff, c0, c1, cig0;
        int32_t tid_ref, pos;
        bool live;
    };
    auto stage_read = [&](const Header &h, int64_t r, int &kind, int &rstrand, uint32_t &columns, SwarRecord &rec) {
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
        uint32_t lead = 0, trail = 0, cols = 0;
        int state = 0, n_lead = 0, n_trail = 0;
        bool simple = h.c1 > h.c0;
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
                else simple = false;
            } else if (state == 2) {
                if (op == OP_S) { trail += len; ++n_trail; }
                else if (op == OP_H) state = 3;
                else simple = false;
            } else {
                simple = op == OP_H;
            }
        }
        const int64_t pos = h.pos;
        const int64_t contig_len = ref.contig_len[h.tid_ref];
        const uint64_t ref0 = ref.contig_off[h.tid_ref] + (uint64_t)(pos > 0 ? pos : 0);
        simple = simple && state >= 1 && cols > 0 && cols < 32768 && n_lead <= 1 && n_trail <= 1 &&
                 (uint64_t)lead + cols + trail == h.l_seq && pos >= 0 && pos + (int64_t)cols <= contig_len &&
                 ref0 < (1ull << 33);
        kind = simple ? 1 : 2;
        if (!simple) return;
        columns = cols;
        const int64_t aend = pos + cols;
        const uint32_t lf = (uint32_t)min((int64_t)A, pos);
        const uint32_t rf = (uint32_t)min((int64_t)A, contig_len - aend);
        uint32_t has_qual = 0;
        if (kQual) has_qual = b.qual[h.boff] != 0xFF;
        const uint64_t q0 = (uint64_t)h.boff + lead;
        rec.qi = (uint32_t)(q0 >> 3);
        rec.ri = (uint32_t)(ref0 >> 3);
        rec.cols = cols | (has_qual << 15) | (lf << 16) | (rf << 24);
        rec.misc = min(cols, (uint32_t)L) | (uint32_t)(ref0 & 7) << 16 | (uint32_t)(q0 & 7) << 20;
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
            length = cols;
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

    // L2 prefetch of one tile, split over the block: the record arrays directly, the sequence (and
    // qualities) through base_off, which is itself a load -- so callers issue the two halves far apart
    const int per = (T + nthreads - 1) / nthreads;  // consecutive reads of a tile covered by one thread
    auto prefetch_headers = [&](int64_t tile_index, uint32_t &boff, uint32_t &coff) {
        const int64_t start = tile_index * T;
        const int64_t rn = start + (int64_t)tid * per;
        const bool live = tid * per < T && rn < b.n_reads;
        if (live) {
            boff = b.base_off[rn];
            coff = b.cigar_off[rn];
        }
        // T entries per array, one 128-byte line per 32 (u32) or 64 (u16) entries
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
        if (kQual) {
            const char *q_at = (const char *)b.qual + boff;
            for (int off = 0; off < per * 160; off += 128) prefetch_l2(q_at + off);
        }
    };

    constexpr int PREP = 4;  // reads staged per thread with their loads in flight together
    const int64_t n_tiles = (n_todo + T - 1) / T;
    if (!subset) {  // the block's first two tiles have nobody to prefetch them
        uint32_t boff0 = 0, coff0 = 0, boff1 = 0, coff1 = 0;
        const bool live0 = prefetch_headers(blockIdx.x, boff0, coff0);
        const bool live1 = prefetch_headers(blockIdx.x + (int64_t)gridDim.x, boff1, coff1);
        if (live0) prefetch_bases(boff0, coff0);
        if (live1) prefetch_bases(boff1, coff1);
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid < 5) s_ctl[tid] = tid == 3 ? 0xffffffffu : 0u;  // n_fwd, n_rev, n_cx, min columns, max columns
        __syncthreads();

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
                int kind, rstrand;
                uint32_t columns;
                SwarRecord rec{};
                stage_read(h[u], h[u].index, kind, rstrand, columns, rec);
                if (g.uniform) {
                    const uint32_t lo = __reduce_min_sync(0xffffffffu, kind == 1 ? columns : 0xffffffffu);
                    const uint32_t hi = __reduce_max_sync(0xffffffffu, kind == 1 ? columns : 0u);
                    if (lane == 0 && hi) {
                        atomicMin(s_ctl + 3, lo);
                        atomicMax(s_ctl + 4, hi);
                    }
                }
                // warp-aggregated appends to the three lists
                const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
                for (int which = 0; which < 3; ++which) {
                    const bool mine = which == 2 ? kind == 2 : (kind == 1 && rstrand == which);
                    const uint32_t m = __ballot_sync(0xffffffffu, mine);
                    if (m) {
                        uint32_t base = 0;
                        if (lane == __ffs(m) - 1) base = atomicAdd(s_ctl + which, (uint32_t)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                        if (mine) {
                            const uint32_t at = base + __popc(m & lt);
                            if (which == 2) s_cx[at] = h[u].index;
                            else s_rec[which == 0 ? at : T - 1 - at] = rec;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- complex reads go to the general kernel's work list ----
        if (tid < 32 && s_ctl[2]) {
            const uint32_t n_cx = s_ctl[2];
            unsigned long long base = 0;
            // several libraries: library l appends at worklist + offsets[l], counted in work_count[l]
            uint32_t *const wl = sub.list ? worklist + sub.offsets[sub.lib] : worklist;
            if (lane == 0) base = atomicAdd(work_count + (sub.list ? sub.lib : 0), (unsigned long long)n_cx);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t i = lane; i < n_cx; i += 32) wl[base + i] = s_cx[i];
        }

        // ---- one window per read when every gap-free read of the tile has the same length ----
        {
            int want = 0;
            const uint32_t lo = s_ctl[3], hi = s_ctl[4];
            if (g.uniform && lo == hi && hi > 0) {
                const int words = ((int)hi + 2 * A + 7) / 8;
                if (words < 2 * W && nthreads / words >= 2) want = (int)hi;
            }
            if (want != mode) {
                if (dirty) flush_block();  // the counters are laid out by mode
                set_mode(want);
                dirty = false;
                tiles_since_flush = 0;
            }
            dirty = dirty || s_ctl[0] + s_ctl[1] > 0;
        }

        // ---- pull the tile after next towards L2 while this one is counted ----
        uint32_t ahead_boff = 0, ahead_coff = 0;
        const bool ahead_live = !subset && prefetch_headers(tile + 2 * (int64_t)gridDim.x, ahead_boff, ahead_coff);

        // ---- the counting loop: two stages that swap roles, so no register copy waits on a load ----
        if (active) {
            // this thread's reads: every stride-th record of its strand's list (forward reads from the
            // front of s_rec, reverse reads from the back).  Two stages swap roles, so that the loads of
            // the next read are in flight while this one is counted and no register copy waits on a load.
            const int n_mine = (int)s_ctl[strand];
            const int stride = slots_of(mode) >> 1;
            int left = n_mine > (slot >> 1) ? (n_mine - (slot >> 1) + stride - 1) / stride : 0;
            const SwarRecord *at = strand ? s_rec + (T - 1 - (slot >> 1)) : s_rec + (slot >> 1);
            const int step = strand ? -stride : stride;
            const Stage nothing{};
            auto fetch_group = [&](Stage (&st)[kReads], int have) {
#pragma unroll
                for (int r = 0; r < kReads; ++r) {
                    if (r < have) fetch(at + r * step, st[r]);
                    else st[r] = nothing;
                }
                at += kReads * step;
            };
            Stage sa[kReads], sb[kReads];
            fetch_group(sa, left);
            while (left > 0) {
                fetch_group(sb, left - kReads);
                count(sa);
                left -= kReads;
                if (left <= 0) break;
                fetch_group(sa, left - kReads);
                count(sb);
                left -= kReads;
            }
        }
        if (ahead_live) prefetch_bases(ahead_boff, ahead_coff);
        __syncthreads();
        if (++tiles_since_flush == flush_period) {
            flush_block();
            tiles_since_flush = 0;
            dirty = false;
        }
    }

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

// ---- reads grouped by library: list[offsets[l] .. offsets[l + 1]) = indices of the reads of library l ----
constexpr int PARTITION_MAX_LIBS = 2048;

__global__ void __launch_bounds__(256) library_count_kernel(DevBatch b, int n_lib, unsigned long long *counts, int32_t *error_flag)
{
    extern __shared__ uint32_t s_hist[];
    for (int i = threadIdx.x; i < n_lib; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < b.n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t lib = b.lib[r];
        if (lib < (uint32_t)n_lib) atomicAdd(s_hist + lib, 1u);
        else atomicCAS(error_flag, 0, DATA_ERR_LIB);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_lib; i += blockDim.x)
        if (s_hist[i]) atomicAdd(counts + i, (unsigned long long)s_hist[i]);
}

// offsets = exclusive scan of counts; cursors start at the offsets.  One block.
__global__ void __launch_bounds__(256) library_offsets_kernel(const unsigned long long *counts, int n_lib, unsigned long long *offsets,
                                                              unsigned long long *cursors)
{
    if (threadIdx.x == 0) {
        unsigned long long at = 0;
        for (int i = 0; i < n_lib; ++i) {
            offsets[i] = cursors[i] = at;
            at += counts[i];
        }
        offsets[n_lib] = at;
    }
}

// each block reserves one range per library for its 1024 reads, then hands out places within it
__global__ void __launch_bounds__(256) library_scatter_kernel(DevBatch b, int n_lib, unsigned long long *cursors, uint32_t *list)
{
    extern __shared__ uint32_t s_part[];  // [n_lib] counts, then [n_lib] bases (low 32 bits suffice: list < 2^31 entries)
    uint32_t *s_count = s_part, *s_base = s_part + n_lib;
    const int64_t first = (int64_t)blockIdx.x * 1024;
    for (int i = threadIdx.x; i < n_lib; i += blockDim.x) s_count[i] = 0;
    __syncthreads();
    uint32_t my_lib[4], my_rank[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t r = first + u * 256 + threadIdx.x;
        my_lib[u] = 0xffffffffu;
        if (r < b.n_reads) {
            const uint32_t lib = b.lib[r];
            if (lib < (uint32_t)n_lib) {
                my_lib[u] = lib;
                my_rank[u] = atomicAdd(s_count + lib, 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_lib; i += blockDim.x)
        s_base[i] = s_count[i] ? (uint32_t)atomicAdd(cursors + i, (unsigned long long)s_count[i]) : 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (my_lib[u] != 0xffffffffu) list[s_base[my_lib[u]] + my_rank[u]] = (uint32_t)(first + u * 256 + threadIdx.x);
}

// Base composition of the whole genome (reference composition.py:6-25 over seqtk.comp, seqtk/seqtk.c:92-110):
// the one-hot image holds one bit per A/C/G/T base and nothing else, so the counts are population counts.
__global__ void __launch_bounds__(256) genome_composition_kernel(const uint32_t *__restrict__ words, int64_t n_words,
                                                                 unsigned long long *counts)
{
    unsigned long long mine[4] = {0, 0, 0, 0};
    const uint4 *quads = (const uint4 *)words;
    const int64_t n_quads = n_words / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(quads + i);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            mine[g] += __popc((q.x >> g) & K1) + __popc((q.y >> g) & K1) + __popc((q.z >> g) & K1) + __popc((q.w >> g) & K1);
    }
    if (blockIdx.x == 0 && threadIdx.x < n_words - 4 * n_quads) {
        const uint32_t w = words[4 * n_quads + threadIdx.x];
#pragma unroll
        for (int g = 0; g < 4; ++g) mine[g] += __popc((w >> g) & K1);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        unsigned long long v = mine[g];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(counts + g, v);
    }
}

// Genome as uploaded (0..3 = A,C,G,T, anything else) -> one-hot nibbles (1,2,4,8; 0 = not a base), in place.
__global__ void ref_to_one_hot_kernel(uint32_t *words, int64_t n_words)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = words[i];
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c = (w >> (4 * k)) & 0xF;
            out |= (c < 4 ? 1u << c : 0u) << (4 * k);
        }
        words[i] = out;
    }
}

}  // namespace mdg
