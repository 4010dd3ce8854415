// Counting pass, general kernel: one warp per read, any CIGAR.
//
// Replaces the body of the reference's per-read loop, main.py:165-217:
//   reader filter (reader.py:121-132), FragmentLengths.update
//   (statistics.py:117-126), align.get_around (align.py:22-35), the reference
//   fetch (main.py:180), align.align / align_with_qual (align.py:38-73),
//   revcomp (main.py:200-205), update_soft_clipping (statistics.py:37-51),
//   MisincorporationRates.update for both ends (statistics.py:22-35) and
//   DNAComposition.update_read / update_reference (statistics.py:75-93).
//
// No strings are built: for alignment column `col` of a read with C columns,
//   5' index = col (forward) or C-1-col (reverse); 3' index = the other one,
// and bases are complemented on the reverse strand.  The
// two walks use the same reference base unless the CIGAR has N ops, in which
// case the walk anchored at the right end reads the contiguous reference
// string `skipped` columns further on.
#pragma once
#include "mdg_device.cuh"

namespace mdg {

// Count sink: a per-block shared-memory slab of 32-bit counters (one library)
// or the global 64-bit tables (any number of libraries / slab too large).
template <bool kShared>
struct Sink {
    uint32_t *s_mis, *s_comp, *s_lg;
    unsigned long long *g_mis, *g_comp;
    int L, LA;

    __device__ __forceinline__ void mis(int lib, int end, int strand, uint32_t cls, int idx) const
    {
        if (kShared)
            atomicAdd(s_mis + (((end * 2 + strand) * MDG_N_CLASSES + cls) * L + idx), 1u);
        else
            atomicAdd(g_mis + ((((size_t)lib * 2 + end) * 2 + strand) * MDG_N_CLASSES + cls) * L + idx, 1ull);
    }
    __device__ __forceinline__ void comp(int lib, int end, int strand, uint32_t base, int slot) const
    {
        if (kShared)
            atomicAdd(s_comp + (((end * 2 + strand) * 4 + base) * LA + slot), 1u);
        else
            atomicAdd(g_comp + ((((size_t)lib * 2 + end) * 2 + strand) * 4 + base) * LA + slot, 1ull);
    }
};

// statistics.py:27-35: one (read, reference) column into one end's table
template <bool kShared>
__device__ __forceinline__ void count_pair(const Sink<kShared> &sink, int lib, int end, int strand, uint32_t b,
                                           uint32_t g, int idx)
{
    if (b <= CODE_GAP && g <= CODE_GAP) {
        if (g != CODE_GAP) sink.mis(lib, end, strand, g, idx);
        if (g != b) sink.mis(lib, end, strand, 4 + 5 * g + b, idx);
    }
}

// own_lib >= 0: the tables `t` points at are those of library own_lib and every read handed in belongs to it
template <bool kShared>
__device__ void count_read(const DevBatch &b, const DevRef &ref, const CountParams &p, const CountTables &t,
                           const Sink<kShared> &sink, int64_t r, int lane, int own_lib)
{
    const uint32_t flag = b.flag[r];
    if (flag & FILTERED_FLAGS) return;
    const int read_lib = b.lib[r];
    const int tid = b.tid[r];
    if (read_lib >= p.n_lib) {
        if (lane == 0) atomicCAS(t.error_flag, 0, DATA_ERR_LIB);
        return;
    }
    const int lib = own_lib >= 0 ? 0 : read_lib;  // index on the library axis of `t`
    if (tid < 0 || tid >= ref.n_contigs) {
        if (lane == 0) atomicCAS(t.error_flag, 0, DATA_ERR_TID);
        return;
    }
    const int strand = (flag >> 4) & 1;  // statistics.py:23: '-' iff is_reverse
    const int L = p.L, A = p.A;
    const int64_t pos = b.pos[r];
    const uint32_t c0 = b.cigar_off[r], c1 = b.cigar_off[r + 1];
    const int n_cig = (int)(c1 - c0);
    const uint32_t *cigar = b.cigar + c0;
    const CigarTotals ct = cigar_totals(cigar, n_cig, lane);
    const uint32_t l_seq = b.l_seq[r];
    const uint64_t boff = b.base_off[r];
    const uint32_t clips = ct.clip_lead + ct.clip_trail;
    const uint32_t n = l_seq > clips ? l_seq - clips : 0;  // len(read.query)
    const uint32_t C = ct.columns;
    const int64_t aend = pos + ct.ref_span;  // align.py:14-19
    const uint64_t contig_off = ref.contig_off[tid];
    const int64_t contig_len = ref.contig_len[tid];
    const uint64_t qbase = boff + ct.clip_lead;
    const bool use_qual = p.min_qual > 0 && b.qual != nullptr && l_seq > 0 && b.qual[boff] != 0xFF;
    const int end_left = strand ? 1 : 0;   // table fed by the left-anchored walk
    const int end_right = strand ? 0 : 1;

    // FragmentLengths.update, statistics.py:117-126
    if (lane == 0) {
        int64_t length = -1;
        int kind = 0;
        if (flag & 0x1) {
            if ((flag & 0x40) && (flag & 0x2)) {
                int64_t tl = b.tlen[r];
                length = tl < 0 ? -tl : tl;
            }
        } else {
            kind = 1;
            length = aend - pos;
        }
        if (length >= 0) {
            if (kShared && length < MDG_LG_SMEM_BINS) {
                atomicAdd(sink.s_lg + (kind * 2 + strand) * MDG_LG_SMEM_BINS + length, 1u);
            } else if (length < p.lg_bins) {
                atomicAdd(t.lghist + (((size_t)lib * 2 + kind) * 2 + strand) * p.lg_bins + length, 1ull);
            } else {
                unsigned long long slot = atomicAdd(t.lg_overflow_count, 1ull);
                if ((int64_t)slot < t.lg_overflow_cap) {
                    int32_t *row = t.lg_overflow_rows + slot * 4;
                    row[0] = read_lib; row[1] = kind; row[2] = strand; row[3] = (int32_t)length;
                }
            }
        }
    }

    // DNAComposition.update_reference, statistics.py:85-93, on the flanks of
    // align.get_around (align.py:22-35); strands swap and complement (main.py:200-205)
    for (int d = 1 + lane; d <= A; d += 32) {
        int64_t left = pos - d, right = aend - 1 + d;
        if (left >= 0 && left < contig_len) {
            uint32_t g = ref_code(ref.words, contig_off + left);
            if (strand) g = complement(g);
            if (g < 4) sink.comp(lib, end_left, strand, g, L + d - 1);
        }
        if (right >= 0 && right < contig_len) {
            uint32_t g = ref_code(ref.words, contig_off + right);
            if (strand) g = complement(g);
            if (g < 4) sink.comp(lib, end_right, strand, g, L + d - 1);
        }
    }

    // walk the CIGAR 32 ops at a time
    uint32_t col_carry = 0, q_carry = 0;
    uint32_t ins_carry = 0;  // insertion columns before the chunk: reference index = column - insertions
    for (int kb = 0; kb < n_cig; kb += 32) {
        const int k = kb + lane;
        uint32_t op = 0xF, len = 0;
        if (k < n_cig) {
            uint32_t w = __ldg(cigar + k);
            op = w & 0xF;
            len = w >> 4;
        }
        const uint32_t cl = op_in_columns(op) ? len : 0;
        const uint32_t ql = op_has_read(op) ? len : 0;
        const uint32_t il = op == OP_I ? len : 0;
        const uint32_t col_end = col_carry + warp_inclusive_scan(cl, lane);
        const uint32_t col_start = col_end - cl;
        const uint32_t q_start = q_carry + warp_inclusive_scan(ql, lane) - ql;
        const uint32_t ins_before = ins_carry + warp_inclusive_scan(il, lane) - il;
        const uint32_t chunk_end = __shfl_sync(0xffffffffu, col_end, 31);

        // update_soft_clipping, statistics.py:37-51: a clip seen before any
        // alignment column is the left one
        uint32_t clip_mask = __ballot_sync(0xffffffffu, op == OP_S);
        while (clip_mask) {
            int src = __ffs(clip_mask) - 1;
            clip_mask &= clip_mask - 1;
            uint32_t clen = __shfl_sync(0xffffffffu, len, src);
            bool is_left = __shfl_sync(0xffffffffu, col_start, src) == 0;
            int end = is_left ? end_left : end_right;
            int lim = min((int)min(clen, (uint32_t)0x7fffffff), L);
            for (int i = lane; i < lim; i += 32) sink.mis(lib, end, strand, MDG_CLASS_SOFTCLIP, i);
        }

        for (uint32_t base = col_carry; base < chunk_end; base += 32) {
            const uint32_t col = base + lane;
            const bool active = col < chunk_end;
            // the op holding this column: first op of the chunk whose end is past it
            int src = 0;
            if (n_cig > 1) {
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    uint32_t v = __shfl_sync(0xffffffffu, col_end, src + step - 1);
                    if (v <= col) src += step;
                }
                src = min(src, 31);
            }
            const uint32_t o = __shfl_sync(0xffffffffu, op, src);
            const uint32_t o_col = __shfl_sync(0xffffffffu, col_start, src);
            const uint32_t o_q = __shfl_sync(0xffffffffu, q_start, src);
            const uint32_t o_ins = __shfl_sync(0xffffffffu, ins_before, src);
            if (!active) continue;
            const uint32_t d = col - o_col;
            const bool has_read = op_has_read(o), has_ref = op_has_ref(o);
            const uint32_t j = o_q + d;  // index in read.query

            uint32_t rb = CODE_GAP;  // read side of the column (D: gap)
            bool masked = false;
            if (has_read) {
                rb = CODE_OTHER;
                if (j < n) {
                    rb = code_of_nibble(read_nibble(b.seq4, qbase + j));
                    if (use_qual) masked = b.qual[qbase + j] < p.min_qual;  // align.py:67-71
                }
            }
            uint32_t gb = CODE_GAP;  // reference side (I: gap)
            if (has_ref) {
                int64_t gpos = pos + (int64_t)(col - o_ins);
                gb = (gpos >= 0 && gpos < contig_len) ? ref_code(ref.words, contig_off + gpos) : CODE_OTHER;
            }

            // DNAComposition.update_read, statistics.py:75-83: ungapped, unmasked query
            if (has_read && rb < 4) {
                uint32_t cb = strand ? 3 - rb : rb;
                if ((int)j < L) sink.comp(lib, end_left, strand, cb, j);
                if ((int)(n - 1 - j) < L && j < n) sink.comp(lib, end_right, strand, cb, n - 1 - j);
            }

            uint32_t b_left = masked ? CODE_OTHER : rb, g_left = masked ? CODE_OTHER : gb;
            uint32_t b_right = b_left, g_right = g_left;
            if (ct.skipped) {
                // right-anchored walk: gapped reference read `skipped` columns further on
                const uint32_t col2 = col + ct.skipped;
                ColumnSite s = locate_column(cigar, n_cig, col2);
                g_right = CODE_GAP;
                if (s.op != OP_I) {
                    int64_t gpos = pos + (int64_t)s.refidx;
                    g_right = (gpos >= 0 && gpos < contig_len) ? ref_code(ref.words, contig_off + gpos) : CODE_OTHER;
                }
                // the mask of column col2 lands on the reference character at col2 (align.py:69-71)
                if (use_qual && s.op != 0xF && op_has_read(s.op) && s.query < n &&
                    b.qual[qbase + s.query] < p.min_qual)
                    g_right = CODE_OTHER;
            }
            if (strand) {
                b_left = complement(b_left); g_left = complement(g_left);
                b_right = complement(b_right); g_right = complement(g_right);
            }
            if ((int)col < L) count_pair(sink, lib, end_left, strand, b_left, g_left, col);
            const uint32_t from_right = C - 1 - col;
            if ((int)from_right < L) count_pair(sink, lib, end_right, strand, b_right, g_right, from_right);
        }
        col_carry = chunk_end;
        q_carry = __shfl_sync(0xffffffffu, q_start + ql, 31);
        ins_carry = __shfl_sync(0xffffffffu, ins_before + il, 31);
    }
}

// Launch: blockDim = 256; dynamic shared memory = slab_words * 4 when kShared.
// With a work list (made by count_swar_kernel) only the listed reads are counted.
template <bool kShared>
__global__ void __launch_bounds__(256) count_general_kernel(DevBatch b, DevRef ref, CountParams p, CountTables t,
                                                            const uint32_t *__restrict__ worklist,
                                                            const unsigned long long *__restrict__ work_count,
                                                            const unsigned long long *__restrict__ lib_offsets, int own_lib)
{
    // per-library work list: the reads of library own_lib start at worklist + lib_offsets[own_lib], their number
    // is work_count[own_lib]
    if (lib_offsets) {
        worklist += lib_offsets[own_lib];
        work_count += own_lib;
    }
    const int64_t n_work = worklist ? (int64_t)*work_count : b.n_reads;
    if ((int64_t)blockIdx.x * (blockDim.x >> 5) >= n_work) return;  // nothing for this block (uniform)
    extern __shared__ uint32_t smem[];
    const int L = p.L, LA = p.L + p.A;
    const int mis_words = 4 * MDG_N_CLASSES * L, comp_words = 16 * LA, lg_words = 4 * MDG_LG_SMEM_BINS;
    Sink<kShared> sink;
    sink.L = L;
    sink.LA = LA;
    sink.g_mis = t.misincorp;
    sink.g_comp = t.dnacomp;
    sink.s_mis = smem;
    sink.s_comp = smem + mis_words;
    sink.s_lg = smem + mis_words + comp_words;
    if (kShared) {
        for (int i = threadIdx.x; i < mis_words + comp_words + lg_words; i += blockDim.x) smem[i] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * warps_per_block;
    for (int64_t w = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < n_work; w += stride)
        count_read<kShared>(b, ref, p, t, sink, worklist ? (int64_t)worklist[w] : w, lane, lib_offsets ? own_lib : -1);
    if (kShared) {
        // flush the block's slab into the 64-bit tables of library 0
        __syncthreads();
        for (int i = threadIdx.x; i < mis_words; i += blockDim.x)
            if (smem[i]) atomicAdd(t.misincorp + i, (unsigned long long)smem[i]);
        for (int i = threadIdx.x; i < comp_words; i += blockDim.x)
            if (sink.s_comp[i]) atomicAdd(t.dnacomp + i, (unsigned long long)sink.s_comp[i]);
        for (int i = threadIdx.x; i < lg_words; i += blockDim.x) {
            uint32_t v = sink.s_lg[i];
            if (v) {
                int bin = i % MDG_LG_SMEM_BINS, ks = i / MDG_LG_SMEM_BINS;
                if (bin < p.lg_bins) {
                    atomicAdd(t.lghist + (size_t)ks * p.lg_bins + bin, (unsigned long long)v);
                } else {
                    // dense histogram narrower than the shared one: spill to the overflow list
                    for (uint32_t c = 0; c < v; ++c) {
                        unsigned long long slot = atomicAdd(t.lg_overflow_count, 1ull);
                        if ((int64_t)slot < t.lg_overflow_cap) {
                            int32_t *row = t.lg_overflow_rows + slot * 4;
                            row[0] = lib_offsets ? own_lib : 0; row[1] = ks >> 1; row[2] = ks & 1; row[3] = bin;
                        }
                    }
                }
            }
        }
    }
}

}  // namespace mdg
