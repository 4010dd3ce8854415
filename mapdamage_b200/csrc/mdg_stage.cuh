// Counting pass, staged bit-sliced kernel for gap-free reads: the production counting kernel.
//
// Same contract and the same counting idea as count_swar_kernel (mdg_swar.cuh: a thread owns eight table
// positions of one strand, class masks are added into packed counters, no atomics on the hot classes), but the
// per-read work is split in two phases with shared memory in between:
//
//   stage: for every read of the tile and every word of its window(s), the read word (natural base order,
//          non-A/C/G/T columns and low-quality bases zeroed) and the reference word (one-hot, masked to the
//          columns / flank bases that count) are built once -- record decoding, masks, shifts and the
//          validity test are done here, a few words at a time so that neighbouring loads are shared;
//   count: a thread reads its word pair of each read with one shared-memory load and only forms the class masks.
//
// The counting loop is then ~55 instructions per read and word instead of ~145, and it has no global loads.
// Substitution classes (rare events) are spilled from the 4-bit counters straight into a shared table with
// atomics, which leaves 32 words of private 16-bit counters per thread for the base classes.
#pragma once
#include <type_traits>
#include "mdg_swar.cuh"

namespace mdg {

struct StagedGeom {
    int32_t words;        // W: 32-bit words per anchor window, ceil((A + L) / 8)
    int32_t threads;      // blockDim.x
    int32_t tile;         // reads staged per iteration of the block
    int32_t uniform;      // 1: one window per read for tiles of equal-length reads
    int32_t flush_tiles;  // > 0: reduce the counters every so many tiles (tests)
    unsigned long long *indel_seen;  // kIndel == false: counts the one-indel reads handed to the general kernel
};

#ifndef MDG_CHUNK_ANCHOR
#define MDG_CHUNK_ANCHOR 5
#endif
#ifndef MDG_CHUNK_UNIFORM
#define MDG_CHUNK_UNIFORM 4
#endif
// consecutive window words one thread stages at a time: measured, the ten words of an anchor window go best as 5 + 5,
// the one window of an equal-length tile in fours
constexpr int STAGE_CHUNK = MDG_CHUNK_ANCHOR, STAGE_CHUNK_UNIFORM = MDG_CHUNK_UNIFORM;

// kIndel: reads with exactly one insertion or deletion between two match blocks are staged too (a third plane holds
// the read as the composition tables see it); without it they go to the general kernel's work list.
#ifdef MDG_PHASE_CLOCKS
__device__ unsigned int mdg_phase_dump[24];
#endif

template <bool kQual, bool kIndel, int kMaxThreads, int kBlocksPerSm>
__global__ void __launch_bounds__(kMaxThreads, kBlocksPerSm)
count_staged_kernel(DevBatch b, DevRef ref, CountParams p, CountTables t, StagedGeom g, uint32_t *__restrict__ worklist,
                    unsigned long long *__restrict__ work_count, SwarSubset sub)
{
    // words staged per (read, window word): read and reference as the misincorporation tables see them[, the read as the
    // composition tables see it: before the quality mask, and contiguous across a deletion]
    constexpr bool kThree = kQual || kIndel;
    constexpr int NW = kThree ? 3 : 2;
    extern __shared__ uint32_t smem[];
    const int nthreads = g.threads, T = g.tile, L = p.L, A = p.A, W = g.words;
    const int wpr_max = 2 * W;  // window words per read in the two-anchor mode (the uniform mode uses fewer)
    const int l2_words = 32 * nthreads;
    const int sub_words = wpr_max * 2 * 12 * 8;
    uint32_t *const s_l2 = smem;                                 // [32][nthreads] private 16-bit counters, base classes
    uint32_t *const s_sub = s_l2 + l2_words;                     // [window word][strand][12][8] substitution classes
    SwarRecord *const s_rec = (SwarRecord *)(s_sub + sub_words);  // [T]: forward reads from the front, reverse from the back
    const int row_words = wpr_max | 1;                            // odd row stride: rows land on different banks
    const int plane = T * row_words;
    uint32_t *const s_stage = (uint32_t *)(s_rec + T);            // [NW planes: read, reference(, unmasked read)][T][row_words]
    uint32_t *const s_mask = s_stage + (size_t)NW * plane;        // [wpr_max][2]: aligned / flank masks of a typical read
    uint32_t *const s_gap = s_mask + 2 * wpr_max;                 // [T] (kIndel) first block | indel length << 15 | deletion << 19, or 0
    uint16_t *const s_gap_rows = (uint16_t *)(s_gap + (kIndel ? T : 0));  // [T] (kIndel) rows holding a read with an indel
    uint32_t *const s_cx = s_gap + (kIndel ? T + (T + 1) / 2 : 0);  // [T] complex reads of the tile
    uint32_t *const s_lg = s_cx + T;                              // [kind][strand][MDG_LG_SMEM_BINS]
    uint32_t *const s_clip = s_lg + 4 * MDG_LG_SMEM_BINS;         // [end][strand][L]
    uint32_t *const s_ctl_base = s_clip + 4 * L;                  // two sets of {n_fwd, n_rev, n_cx, min / max columns, -, -, -}

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t *const subset = sub.list ? sub.list + sub.offsets[sub.lib] : nullptr;
    const int64_t n_todo = sub.list ? (int64_t)(sub.offsets[sub.lib + 1] - sub.offsets[sub.lib]) : b.n_reads;
    if ((int64_t)blockIdx.x * T >= n_todo) return;  // no tile for this block (an empty list costs a launch, nothing more)
    for (int i = tid; i < l2_words + sub_words; i += nthreads) s_l2[i] = 0;
    for (int i = tid; i < 4 * MDG_LG_SMEM_BINS + 4 * L; i += nthreads) s_lg[i] = 0;

    const uint32_t *__restrict__ seq32 = (const uint32_t *)b.seq4;
    const uint32_t *__restrict__ ref32 = ref.words;

    // Block mode: 0 = two anchors per read (window words 0..W-1 left, W..2W-1 right); C > 0 = every gap-free read
    // of the tile has C columns and one window [-A, C + A) per read serves both tables (see mdg_swar.cuh).
    int mode = 0;
    auto words_of = [&](int columns) { return columns ? (columns + 2 * A + 7) / 8 : wpr_max; };
    auto slots_of = [&](int columns) { return (nthreads / words_of(columns)) & ~1; };
    // counting-phase role of this thread: window word `ws` of read slot `slot`
    bool active;
    int ws, slot, strand;
    // stage-phase role: chunk [st_k0, st_k1) of anchor st_anchor of every st_step-th read from st_first on
    // (st_first < 0: no role; divisions happen here, once per mode, not per read)
    int st_first, st_step, st_anchor, st_k0, st_k1;
    // aligned / flank nibble masks of window word k of an anchor, for a read with v countable columns and
    // lf / rf flank bases.  With z = the nibble index of position 0 (left anchor) or one past it (right anchor):
    //   two anchors, left:  aligned [z, z + v),  flank [z - lf, z);   right: aligned [z - v, z), flank [z, z + rf)
    //   one window (mode C): aligned [z, z + C), flank [z - lf, z) and [z + C, z + C + rf)
    auto window_masks = [&](int anchor, int k, int v, int lf, int rf, uint32_t &aligned, uint32_t &flank) {
        const int pbase = 8 * k - A;
        const int z4 = 4 * (anchor ? pbase + 8 : -pbase);
        const uint32_t below_z = low_nibbles(z4);
        if (mode) {
            const int zr4 = z4 + 4 * mode;
            const uint32_t below_zr = low_nibbles(zr4);
            aligned = below_zr & ~below_z;
            flank = (below_z & ~low_nibbles(z4 - 4 * lf)) | (low_nibbles(zr4 + 4 * rf) & ~below_zr);
        } else if (anchor == 0) {
            aligned = low_nibbles(z4 + 4 * v) & ~below_z;
            flank = below_z & ~low_nibbles(z4 - 4 * lf);
        } else {
            aligned = below_z & ~low_nibbles(z4 - 4 * v);
            flank = low_nibbles(z4 + 4 * rf) & ~below_z;
        }
    };
    auto set_mode = [&](int columns) {
        mode = columns;
        const int wpr = words_of(columns);
        active = tid < wpr * slots_of(columns);
        ws = tid % wpr;
        slot = tid / wpr;
        strand = slot & 1;
        {
            const int n_anchors = columns ? 1 : 2, per_anchor = columns ? wpr : W;
            const int chunk = columns ? STAGE_CHUNK_UNIFORM : STAGE_CHUNK;
            const int chunks = (per_anchor + chunk - 1) / chunk, per_read = n_anchors * chunks;
            st_step = nthreads / per_read;
            const int c = tid % per_read;
            st_first = tid / per_read < st_step ? tid / per_read : -1;
            st_anchor = c / chunks;
            st_k0 = (c - st_anchor * chunks) * chunk;
            st_k1 = min(st_k0 + chunk, per_anchor);
        }
        // masks of the typical read (all flank bases on the contig, at least L columns), by window word
        for (int w = tid; w < wpr; w += nthreads) {
            uint32_t aligned, flank;
            window_masks(columns ? 0 : w / W, columns ? w : w % W, columns ? columns : L, A, A, aligned, flank);
            s_mask[2 * w] = aligned;
            s_mask[2 * w + 1] = flank;
        }
    };
    set_mode(0);

    uint32_t acc0[SWAR_CLASSES];  // 8 x 4-bit counters per class
    uint32_t acc1[16];            // base classes 0..7: 2 x (4 x 8-bit) counters, even / odd nibbles
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
    // 4-bit counters: base classes -> 8-bit registers; substitution classes (mostly zero) -> the block's shared
    // table, one atomic per non-zero counter.  A substitution class whose counters are all 0 or 1 is left where it
    // is unless `all`: 14 more reads cannot overflow it, and most classes never get further between two spills.
    auto spill0 = [&](bool all) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            acc1[2 * c] += acc0[c] & 0x0F0F0F0Fu;
            acc1[2 * c + 1] += (acc0[c] >> 4) & 0x0F0F0F0Fu;
            acc0[c] = 0;
        }
        uint32_t *const mine = s_sub + (ws * 2 + strand) * 96;
#pragma unroll
        for (int c = 8; c < SWAR_CLASSES; ++c) {
            uint32_t v = acc0[c];
            if (!(all ? v : v & 0xEEEEEEEEu)) continue;
            while (v) {
                const int nib = (__ffs(v) - 1) >> 2;
                atomicAdd(mine + (c - 8) * 8 + nib, (v >> (4 * nib)) & 15u);
                v &= ~(15u << (4 * nib));
            }
            acc0[c] = 0;
        }
        n0 = 0;
        if (++n1 == 17) spill1();
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
    // thread's 16-bit counters could overflow)
    auto flush_block = [&]() {
        spill0(true);
        if (n1) spill1();
        __syncthreads();
        const int wpr = words_of(mode), mode_slots = slots_of(mode);
        const int n_cells = wpr * 2 * SWAR_CLASSES * 8;  // window word, strand, class, nibble
        for (int cell = tid; cell < n_cells; cell += nthreads) {
            int rest = cell;
            const int nib = rest & 7; rest >>= 3;
            const int cls = rest % SWAR_CLASSES; rest /= SWAR_CLASSES;
            const int cstrand = rest & 1; rest >>= 1;
            const int cws = rest;
            const int canchor = mode ? 0 : cws / W, cword = mode ? cws : cws % W;
            const int pb = 8 * cword - A;
            const int pos = canchor ? pb + 7 - nib : pb + nib;
            if (pos < -A || (pos < 0 && cls >= 4)) continue;
            if (mode ? pos >= mode + A || (pos >= mode && cls >= 4) : pos >= L) continue;
            unsigned long long sum = 0;
            if (cls < 8) {
                // 16-bit lane holding (cls, nib): 8-bit lane bl = nib >> 1 of acc1[2 * cls + (nib & 1)]
                const int w1 = 2 * cls + (nib & 1), bl = nib >> 1;
                const int w2 = 2 * w1 + (bl & 1), half = bl >> 1;
                for (int cslot = cstrand; cslot < mode_slots; cslot += 2) {
                    const uint32_t v = s_l2[w2 * nthreads + cslot * wpr + cws];
                    sum += half ? v >> 16 : v & 0xFFFFu;
                }
            } else {
                sum = s_sub[((cws * 2 + cstrand) * 12 + (cls - 8)) * 8 + nib];
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
        for (int i = tid; i < l2_words + sub_words; i += nthreads) s_l2[i] = 0;
        __syncthreads();
    };
    // a thread counts at most ceil(T / (slots / 2)) reads per tile (every read on one strand)
    const int min_slots = max(2, (nthreads / wpr_max) & ~1);
    const int flush_period = g.flush_tiles > 0 ? g.flush_tiles : max(1, 60000 / ((T + (min_slots >> 1) - 1) / (min_slots >> 1)));
    int tiles_since_flush = 0;
    bool dirty = false;  // counters hold counts of the current mode

    // ---- stage phase, reads with one indel (kIndel): window words [k0, k1) of one anchor ----
    // Seen from an anchor, the alignment is a near block of n1 columns, the indel's g gap columns, and the far block.
    // Read and reference run contiguously from the anchor through the near block; across the gap the side that
    // has the gap falls g bases behind: after an insertion the reference, after a deletion the read.  Misincorporation
    // positions are columns (statistics.py:26); composition positions are read bases (statistics.py:76-83), so the
    // third plane holds the read contiguously.  The gap columns themselves are rare: their classes (ref>- / ->read)
    // go straight to the tables with atomics.
    auto eight_at = [&](const uint32_t *words32, int word0, int nibble_index, bool bam_order) {
        // eight nibbles starting `nibble_index` nibbles after word word0 (may be negative)
        const uint32_t *at = words32 + (word0 + (nibble_index >> 3));
        uint32_t lo = __ldg(at), hi = __ldg(at + 1);
        if (bam_order) { lo = natural_order(lo); hi = natural_order(hi); }
        return __funnelshift_r(lo, hi, (nibble_index & 7) << 2);
    };
    auto low_quality = [&](int qword0, int nibble_index) {  // nibbles whose base quality is below --min-basequal
        const int tq = nibble_index;
        const uint32_t *q32 = (const uint32_t *)b.qual + 2 * (int64_t)(qword0 + (tq >> 3)) + ((tq >> 2) & 1);
        const uint32_t qa = __ldg(q32), qm = __ldg(q32 + 1), qz = __ldg(q32 + 2);
        const int sq = (tq & 3) << 3;
        const uint32_t lo = __funnelshift_r(qa, qm, sq), hi = __funnelshift_r(qm, qz, sq);
        const uint32_t mq = (uint32_t)p.min_qual * 0x01010101u;
        uint32_t zl = (~((lo | 0x80808080u) - mq) & 0x80808080u) >> 7;
        uint32_t zh = (~((hi | 0x80808080u) - mq) & 0x80808080u) >> 7;
        zl |= zl >> 4;
        zh |= zh >> 4;
        return (((zl & 0x11u) | ((zl >> 8) & 0x1100u)) | (((zh & 0x11u) | ((zh >> 8) & 0x1100u)) << 16)) * 15u;
    };
    auto stage_indel_words = [&](const SwarRecord &rec, uint32_t gi, uint32_t *row, int anchor, int k0, int k1, int rstrand) {
        const int cols = (int)(rec.cols & 0x7FFF);
        const int gap = (int)((gi >> 15) & 15), first = (int)(gi & 0x7FFF);
        const bool deletion = (gi >> 19) & 1;
        const int n_query = cols - (deletion ? gap : 0), ref_span = cols - (deletion ? 0 : gap);
        const int n1 = anchor ? cols - first - gap : first;  // near block as this anchor sees it
        const int lf = (int)((rec.cols >> 16) & 0xFF), rf = (int)(rec.cols >> 24);
        const int v_cols = min(L, cols), v_query = min(L, n_query);
        const int qnib = (int)(rec.misc >> 20), rnib = (int)((rec.misc >> 16) & 7);
        for (int k = k0; k < k1; ++k) {
            const int pbase = 8 * k - A;
            // memory offsets (bases from the first aligned base / reference base) of nibble 0, contiguous from the anchor
            const int q_off = anchor ? n_query - 8 - pbase : pbase, r_off = anchor ? ref_span - 8 - pbase : pbase;
            const int behind = anchor ? gap : -gap;  // where the lagging side sits, in memory order
            uint32_t aligned, flank, comp, unused;
            window_masks(anchor, k, v_cols, lf, rf, aligned, flank);
            window_masks(anchor, k, v_query, 0, 0, comp, unused);
            // near / gap / far columns of this word
            uint32_t near, far;
            if (anchor == 0) {
                near = low_nibbles(4 * (n1 - pbase));            // positions < n1 (flank positions included)
                far = ~low_nibbles(4 * (n1 + gap - pbase));      // positions >= n1 + gap
            } else {
                near = ~low_nibbles(4 * (pbase + 8 - n1));
                far = low_nibbles(4 * (pbase + 8 - n1 - gap));
            }
            const uint32_t gap_cols = ~(near | far) & aligned;
            const uint32_t xq = eight_at(seq32, (int)rec.qi, qnib + q_off, true);
            const uint32_t yr = eight_at(ref32, (int)rec.ri, rnib + r_off, false);
            uint32_t x_col, y_col, low = 0;
            if (deletion) {
                const uint32_t xs = eight_at(seq32, (int)rec.qi, qnib + q_off + behind, true);
                x_col = (xq & near) | (xs & far);
                y_col = yr;
                if (kQual && (rec.cols & 0x8000u))
                    low = (low_quality((int)rec.qi, qnib + q_off) & near) | (low_quality((int)rec.qi, qnib + q_off + behind) & far);
            } else {
                const uint32_t ys = eight_at(ref32, (int)rec.ri, rnib + r_off + behind, false);
                x_col = xq;
                y_col = (yr & near) | (ys & far);
                if (kQual && (rec.cols & 0x8000u)) low = low_quality((int)rec.qi, qnib + q_off);
            }
            // composition: the read base at each query position, whatever its quality (statistics.py:75-83)
            const uint32_t xc = xq & (one_hot_nibbles(xq) * 15u) & comp;
            // misincorporation: columns with an A/C/G/T read base of sufficient quality; a deletion column has no
            // read base but still counts its reference base (statistics.py:27-30)
            const uint32_t valid = (one_hot_nibbles(x_col) * 15u) & aligned & ~low;
            x_col &= valid;
            const uint32_t del_cols = deletion ? gap_cols : 0u;
            y_col &= valid | flank | del_cols;
            uint32_t *at = row + (mode ? 0 : anchor * W) + k;
            at[0] = x_col;
            at[plane] = y_col;
            at[2 * plane] = xc;
            // the gap columns: ->read (insertion, needs a valid read base) or ref>- (deletion, needs a reference base)
            uint32_t events = (deletion ? y_col : x_col) & gap_cols;
            while (events) {
                const int nib = (__ffs(events) - 1) >> 2;
                uint32_t code = 31 - __clz((events >> (4 * nib)) & 15u);  // one-hot nibble -> base 0..3
                events &= ~(15u << (4 * nib));
                const int pos = anchor ? pbase + 7 - nib : pbase + nib;
                if (rstrand) code = 3 - code;
                const int cls = deletion ? 4 + 5 * (int)code + 4 : 4 + 5 * 4 + (int)code;
                const int es = (anchor ^ rstrand) * 2 + rstrand;
                atomicAdd(t.misincorp + ((size_t)es * MDG_N_CLASSES + cls) * L + pos, 1ull);
            }
        }
    };

    // ---- stage phase: the masked word pairs of window words [k0, k1) of one anchor of one read ----
    auto stage_words = [&](auto chunk_tag, const SwarRecord &rec, uint32_t *row, int anchor, int k0, int k1) {
        constexpr int kChunk = decltype(chunk_tag)::value;  // words per call, at most: fixes the unrolling
        const int cols = (int)(rec.cols & 0x7FFF);
        const int v = (int)(rec.misc & 0xFFFF);
        const int lf = (int)((rec.cols >> 16) & 0xFF), rf = (int)(rec.cols >> 24);
        const bool typical = lf == A && rf == A && (mode || v == L);
        // nibble 0 of window word k sits `off` bases from the first aligned base: left anchor 8 k - A, ascending
        // with k; right anchor cols - 8 - (8 k - A), descending with k.  Load the chunk's words in memory order.
        const int n = k1 - k0;
        const int off_lo = anchor ? cols - 8 - (8 * (k1 - 1) - A) : 8 * k0 - A;
        const int tq = (int)(rec.misc >> 20) + off_lo, tr = (int)((rec.misc >> 16) & 7) + off_lo;
        const int sx = (tq & 7) << 2, sy = (tr & 7) << 2;
        const uint32_t *qp = seq32 + ((int)rec.qi + (tq >> 3));
        const uint32_t *rp = ref32 + ((int)rec.ri + (tr >> 3));
        uint32_t wq[kChunk + 1], wr[kChunk + 1];
#pragma unroll
        for (int j = 0; j <= kChunk; ++j) {
            wq[j] = j <= n ? __ldg(qp + j) : 0;
            wr[j] = j <= n ? __ldg(rp + j) : 0;
        }
#pragma unroll
        for (int j = 0; j <= kChunk; ++j) wq[j] = natural_order(wq[j]);
        // memory word j holds window word k0 + j (left anchor) or k1 - 1 - j (right anchor): walk the stage row and
        // the mask table by +-1 from there
        const int wfirst = (mode ? 0 : anchor * W) + (anchor ? k1 - 1 : k0), dir = anchor ? -1 : 1;
        uint32_t *at = row + wfirst;
        const uint32_t *mask_at = s_mask + 2 * wfirst;
        int k = anchor ? k1 - 1 : k0;
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            if (j >= n) break;
            uint32_t aligned, flank;
            if (typical) {
                aligned = mask_at[0];
                flank = mask_at[1];
            } else {
                window_masks(anchor, k, v, lf, rf, aligned, flank);
            }
            uint32_t x = __funnelshift_r(wq[j], wq[j + 1], sx);
            uint32_t y = __funnelshift_r(wr[j], wr[j + 1], sy);
            // a column counts only when the read base is A/C/G/T (statistics.py:27); the reference side is already 0
            // for anything that is not A/C/G/T.  Flank nibbles carry the reference base alone (statistics.py:85-93).
            const uint32_t valid = (one_hot_nibbles(x) * 15u) & aligned;
            x &= valid;
            y &= valid | flank;
            if (kThree) at[2 * plane] = x;  // DNAComposition.update_read ignores the quality mask (statistics.py:75-83)
            if (kQual) {
                if (rec.cols & 0x8000u) {
                    // align_with_qual, align.py:67-71: bases below --min-basequal become N on both sides.  The eight
                    // qualities of memory word j start at byte 8 (word) + shift / 4; three aligned words cover them.
                    const int tqj = tq + 8 * j;
                    const uint32_t *q32 = (const uint32_t *)b.qual + 2 * (int64_t)((int)rec.qi + (tqj >> 3)) + ((tqj >> 2) & 1);
                    const uint32_t qa = __ldg(q32), qm = __ldg(q32 + 1), qz = __ldg(q32 + 2);
                    const int sq = (tqj & 3) << 3;
                    const uint32_t lo = __funnelshift_r(qa, qm, sq), hi = __funnelshift_r(qm, qz, sq);
                    const uint32_t mq = (uint32_t)p.min_qual * 0x01010101u;
                    // bit 7 of a byte of ((q | 0x80) - min_qual) is clear iff q < min_qual
                    uint32_t zl = (~((lo | 0x80808080u) - mq) & 0x80808080u) >> 7;
                    uint32_t zh = (~((hi | 0x80808080u) - mq) & 0x80808080u) >> 7;
                    zl |= zl >> 4;
                    zh |= zh >> 4;
                    const uint32_t low = (((zl & 0x11u) | ((zl >> 8) & 0x1100u)) | (((zh & 0x11u) | ((zh >> 8) & 0x1100u)) << 16)) * 15u;
                    const uint32_t keep = ~low | flank;  // there is no read base, hence no quality, on a flank
                    x &= keep;
                    y &= keep;
                }
            }
            at[0] = x;
            at[plane] = y;
            at += dir;
            mask_at += 2 * dir;
            k += dir;
        }
    };

    // ---- count phase: class masks of the staged word pairs of two reads at a time: the two instruction streams
    // interleave and their masks are summed before they touch the counters (one 3-input add per class) ----
    auto count2 = [&](uint32_t xa, uint32_t ya, uint32_t xca, uint32_t xb, uint32_t yb, uint32_t xcb) {
#define MDG_CLASS(c, ea, eb) acc0[c] += ((ea) & K1) + ((eb) & K1);
        MDG_CLASS(0, ya, yb)
        MDG_CLASS(1, ya >> 1, yb >> 1)
        MDG_CLASS(2, ya >> 2, yb >> 2)
        MDG_CLASS(3, ya >> 3, yb >> 3)
        MDG_CLASS(4, xca, xcb)
        MDG_CLASS(5, xca >> 1, xcb >> 1)
        MDG_CLASS(6, xca >> 2, xcb >> 2)
        MDG_CLASS(7, xca >> 3, xcb >> 3)
        MDG_CLASS(8, ya & (xa >> 1), yb & (xb >> 1))                  // A>C
        MDG_CLASS(9, ya & (xa >> 2), yb & (xb >> 2))                  // A>G
        MDG_CLASS(10, ya & (xa >> 3), yb & (xb >> 3))                 // A>T
        MDG_CLASS(11, (ya >> 1) & xa, (yb >> 1) & xb)                 // C>A
        MDG_CLASS(12, (ya >> 1) & (xa >> 2), (yb >> 1) & (xb >> 2))   // C>G
        MDG_CLASS(13, (ya >> 1) & (xa >> 3), (yb >> 1) & (xb >> 3))   // C>T
        MDG_CLASS(14, (ya >> 2) & xa, (yb >> 2) & xb)                 // G>A
        MDG_CLASS(15, (ya >> 2) & (xa >> 1), (yb >> 2) & (xb >> 1))   // G>C
        MDG_CLASS(16, (ya >> 2) & (xa >> 3), (yb >> 2) & (xb >> 3))   // G>T
        MDG_CLASS(17, (ya >> 3) & xa, (yb >> 3) & xb)                 // T>A
        MDG_CLASS(18, (ya >> 3) & (xa >> 1), (yb >> 3) & (xb >> 1))   // T>C
        MDG_CLASS(19, (ya >> 3) & (xa >> 2), (yb >> 3) & (xb >> 2))   // T>G
#undef MDG_CLASS
        n0 += 2;
        if (n0 >= 14) spill0(false);  // a 4-bit counter holds 15
    };

    __shared__ uint32_t indel_here;  // one-indel reads this block left to the general kernel
    if (tid == 0) indel_here = 0;
    // ---- staging of one read: filter, classify, per-read events (statistics.py:37-51,117-126) ----
    struct Header {
        uint32_t index, flag, lib, l_seq, boff, c0, c1, cig0;
        int32_t tid_ref, pos;
        bool live;
    };
    auto stage_read = [&](const Header &h, int64_t r, int &kind, int &rstrand, uint32_t &columns, uint32_t &gap_info,
                          SwarRecord &rec) {
        kind = 0;
        rstrand = 0;
        columns = 0;
        gap_info = 0;
        bool one_indel = false;
        (void)one_indel;
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
        uint32_t gap_at = 0, gap_len = 0, gap_del = 0;  // one insertion / deletion between two match blocks (kIndel)
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
                else if ((op == OP_I || op == OP_D) && !gap_len && len >= 1 && len <= 7 && cols >= 1) {
                    gap_at = cols; gap_len = len; gap_del = op == OP_D;
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
        if (!kIndel && simple && gap_len) {
            // this variant leaves one-indel reads to the general kernel; tell the host how common they are
            simple = false;
            one_indel = true;
        }
        const uint32_t n_query = cols - (gap_del ? gap_len : 0), ref_span = cols - (gap_len && !gap_del ? gap_len : 0);
        const int64_t pos = h.pos;
        const int64_t contig_len = ref.contig_len[h.tid_ref];
        const uint64_t ref0 = ref.contig_off[h.tid_ref] + (uint64_t)(pos > 0 ? pos : 0);
        simple = simple && state >= 1 && cols > 0 && cols < 32768 && n_lead <= 1 && n_trail <= 1 &&
                 (uint64_t)lead + n_query + trail == h.l_seq && pos >= 0 && pos + (int64_t)ref_span <= contig_len &&
                 ref0 < (1ull << 33);
        kind = simple ? 1 : 2;
        if (!kIndel && one_indel) atomicAdd(&indel_here, 1u);
        if (!simple) return;
        columns = gap_len ? cols | 0x8000u : cols;  // a read with an indel never makes a tile "equal length"
        gap_info = gap_len ? gap_at | gap_len << 15 | gap_del << 19 : 0;
        const int64_t aend = pos + ref_span;
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


    // the same for a tile of a library's index list: the reads are scattered, so each thread pulls the records of
    // (up to two of) the tile's reads itself and later the lines their bases start in
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
            if (kQual) {
                prefetch_l2((const char *)b.qual + boff[u]);
                prefetch_l2((const char *)b.qual + boff[u] + 127);
            }
        }
    };

    constexpr int PREP = 2;  // reads parsed per thread with their loads in flight together
    const int64_t n_tiles = (n_todo + T - 1) / T;
    if (!subset) {  // the block's first two tiles have nobody to prefetch them
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
    int tile_parity = 0;
    if (tid < 6) s_ctl_base[tid] = tid == 3 ? 0xffffffffu : 0u;
    __syncthreads();
#ifdef MDG_PHASE_CLOCKS
    // per-phase cycle counts as the first and the last warp see them (shared memory: no registers held);
    // block 3's land in mdg_phase_dump, which mdg_sync prints
    __shared__ unsigned int s_pc[2][13];
    if (tid < 26) (&s_pc[0][0])[tid] = 0;
    __syncthreads();
    const int warp = tid >> 5;
    const bool pc_warp = warp == 0 || warp == (nthreads >> 5) - 1;
    if (pc_warp && lane == 0) s_pc[warp != 0][12] = (unsigned int)clock();
    __syncwarp();
#define MDG_PHASE(i)                                              \
    if (pc_warp) {                                                \
        const unsigned int pt1 = (unsigned int)clock();           \
        if (lane == 0) {                                          \
            s_pc[warp != 0][i] += pt1 - s_pc[warp != 0][12];      \
            s_pc[warp != 0][12] = pt1;                            \
        }                                                         \
        __syncwarp();                                             \
    }
#else
#define MDG_PHASE(i)
#endif
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // the tile counters alternate between two sets: the other set was reset while the previous tile was counted,
        // behind that tile's barriers, so no barrier is needed before this tile's appends
        uint32_t *const s_ctl = s_ctl_base + 8 * (tile_parity & 1);
        uint32_t *const s_ctl_next = s_ctl_base + 8 * ((tile_parity & 1) ^ 1);
        ++tile_parity;

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
                uint32_t columns, gap_info;
                SwarRecord rec{};
                stage_read(h[u], h[u].index, kind, rstrand, columns, gap_info, rec);
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
                            else {
                                const uint32_t row = which == 0 ? at : T - 1 - at;
                                s_rec[row] = rec;
                                if (kIndel) {
                                    s_gap[row] = gap_info;
                                    if (gap_info) s_gap_rows[atomicAdd(s_ctl + 5, 1u)] = (uint16_t)row;
                                }
                            }
                        }
                    }
                }
            }
        }
        MDG_PHASE(0)
        __syncthreads();
        MDG_PHASE(1)

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

        MDG_PHASE(8)
        // ---- one window per read when every gap-free read of the tile has the same length ----
        {
            int want = 0;
            const uint32_t lo = s_ctl[3], hi = s_ctl[4];
            if (g.uniform && lo == hi && hi > 0) {
                const int words = ((int)hi + 2 * A + 7) / 8;
                if (words < wpr_max && nthreads / words >= 2) want = (int)hi;
            }
            if (want != mode) {
                if (dirty) flush_block();  // the counters are laid out by mode
                set_mode(want);
                __syncthreads();  // the mask table of the new mode is read by every thread
                dirty = false;
                tiles_since_flush = 0;
            }
            dirty = dirty || s_ctl[0] + s_ctl[1] > 0;
        }

        MDG_PHASE(9)
        // ---- pull the tile after next towards L2 ----
        uint32_t ahead_boff = 0, ahead_coff = 0;
        const bool ahead_live = !subset && prefetch_headers(tile + 2 * (int64_t)gridDim.x, ahead_boff, ahead_coff);
        uint32_t listed_boff[2] = {0, 0}, listed_coff[2] = {0, 0};
        const uint32_t listed_live = subset ? prefetch_listed_headers(tile + 2 * (int64_t)gridDim.x, listed_boff, listed_coff) : 0u;

        // ---- stage phase ----
        MDG_PHASE(2)
        const int n_fwd = (int)s_ctl[0], n_rev = (int)s_ctl[1];
        if (st_first >= 0) {
            for (int li = st_first; li < n_fwd + n_rev; li += st_step) {
                const int row = li < n_fwd ? li : T - 1 - (li - n_fwd);
                if (kIndel && s_gap[row]) continue;  // second pass below: whole warps of indel reads
                if (mode)
                    stage_words(std::integral_constant<int, STAGE_CHUNK_UNIFORM>{}, s_rec[row], s_stage + (size_t)row * row_words,
                                st_anchor, st_k0, st_k1);
                else
                    stage_words(std::integral_constant<int, STAGE_CHUNK>{}, s_rec[row], s_stage + (size_t)row * row_words,
                                st_anchor, st_k0, st_k1);
            }
            if (kIndel) {
                const int n_gap = (int)s_ctl[5];
                for (int gi = st_first; gi < n_gap; gi += st_step) {
                    const int row = s_gap_rows[gi];
                    stage_indel_words(s_rec[row], s_gap[row], s_stage + (size_t)row * row_words, st_anchor, st_k0, st_k1,
                                      row >= T - n_rev);
                }
            }
        }
        if (ahead_live) prefetch_bases(ahead_boff, ahead_coff);
        if (listed_live) prefetch_listed_bases(listed_live, listed_boff, listed_coff);
        MDG_PHASE(3)
        __syncthreads();
        MDG_PHASE(4)

        if (tid < 6) s_ctl_next[tid] = tid == 3 ? 0xffffffffu : 0u;  // n_fwd, n_rev, n_cx, min / max columns, indel rows
        // ---- count phase: this thread's window word of every stride-th read of its strand ----
        if (active) {
            const int n_mine = strand ? n_rev : n_fwd;
            const int stride = slots_of(mode) >> 1;
            const int row_step = (strand ? -stride : stride) * row_words;
            const uint32_t *at = s_stage + (size_t)(strand ? T - 1 - (slot >> 1) : (slot >> 1)) * row_words + ws;
            for (int i = slot >> 1; i < n_mine; i += 2 * stride) {
                const bool second = i + stride < n_mine;
                const uint32_t xa = at[0], ya = at[plane];
                const uint32_t xca = kThree ? at[2 * plane] : xa;
                const uint32_t *at2 = second ? at + row_step : at;
                uint32_t xb = at2[0], yb = at2[plane];
                uint32_t xcb = kThree ? at2[2 * plane] : xb;
                if (!second) xb = yb = xcb = 0;
                at += 2 * row_step;
                if (xa | ya | xca | xb | yb | xcb) count2(xa, ya, xca, xb, yb, xcb);
            }
        }
        MDG_PHASE(5)
        __syncthreads();
        MDG_PHASE(6)
        if (++tiles_since_flush == flush_period) {
            flush_block();
            tiles_since_flush = 0;
            dirty = false;
        }
        MDG_PHASE(7)
    }
#ifdef MDG_PHASE_CLOCKS
    __syncthreads();
    if (blockIdx.x == 3 && tid < 24) mdg_phase_dump[tid] = s_pc[tid / 12][tid % 12];
#endif

    flush_block();
    if (!kIndel && tid == 0 && indel_here && g.indel_seen) atomicAdd(g.indel_seen, (unsigned long long)indel_here);
    for (int i = tid; i < 4 * MDG_LG_SMEM_BINS; i += nthreads) {
        const uint32_t v = s_lg[i];
        if (v) atomicAdd(t.lghist + (size_t)(i / MDG_LG_SMEM_BINS) * p.lg_bins + i % MDG_LG_SMEM_BINS, (unsigned long long)v);
    }
    for (int i = tid; i < 4 * L; i += nthreads) {
        const uint32_t v = s_clip[i];
        if (v) atomicAdd(t.misincorp + ((size_t)(i / L) * MDG_N_CLASSES + MDG_CLASS_SOFTCLIP) * L + i % L, (unsigned long long)v);
    }
}

// {first, last} of every library's stretch of the one-indel list (count_planes_kernel appends library l from
// offsets[l] on): bounds[2 l], bounds[2 l + 1]
__global__ void indel_bounds_kernel(const unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ offsets, int n_lib,
                                    unsigned long long *__restrict__ bounds)
{
    for (int l = threadIdx.x; l < n_lib; l += blockDim.x) {
        const unsigned long long first = offsets ? offsets[l] : 0ull;
        bounds[2 * l] = first;
        bounds[2 * l + 1] = first + counts[l];
    }
}

}  // namespace mdg
