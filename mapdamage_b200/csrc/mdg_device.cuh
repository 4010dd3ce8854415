// Device-side data structures and helpers shared by the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define MDG_N_CLASSES 30
#define MDG_CLASS_SOFTCLIP 29
#define MDG_LG_SMEM_BINS 512  // fragment lengths below this are histogrammed in shared memory

namespace mdg {

// BAM CIGAR operation codes
enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

// base codes used on the device: 0..3 = A,C,G,T, 4 = alignment gap, 7 = not a base
constexpr uint32_t CODE_GAP = 4;
constexpr uint32_t CODE_OTHER = 7;

// reader.py:9-13,121-132: unmapped | secondary | QC-fail | duplicate | supplementary
constexpr uint32_t FILTERED_FLAGS = 0x4 | 0x100 | 0x200 | 0x400 | 0x800;

struct DevBatch {
    int64_t n_reads;
    const uint16_t *flag;
    const int32_t *tid;
    const int32_t *pos;
    const uint16_t *lib;
    const uint32_t *l_seq;
    const uint32_t *base_off;
    const uint32_t *cigar_off;
    const uint32_t *cigar;
    const uint8_t *seq4;
    const uint8_t *qual;  // may be null
    const int32_t *tlen;
    const int32_t *mtid;
    const int32_t *mpos;
};

struct DevRef {
    const uint4 *planes;         // the same genome as bit planes: 32 bases per entry, {A, C, G, T} words (mdg_planes.cuh)
    const uint32_t *words;       // 8 bases per word, low nibble first, one-hot
    const uint64_t *contig_off;  // first base of each contig in the packed stream
    const uint32_t *contig_len;
    int32_t n_contigs;
};

struct CountTables {
    unsigned long long *misincorp;  // [lib][end][strand][30][L]
    unsigned long long *dnacomp;    // [lib][end][strand][4][L+A]
    unsigned long long *lghist;     // [lib][kind][strand][lg_bins]
    int32_t *lg_overflow_rows;      // [cap][4] = lib, kind, strand, length
    unsigned long long *lg_overflow_count;
    int64_t lg_overflow_cap;
    int32_t *error_flag;            // first data error seen (0 = none)
};

struct CountParams {
    int32_t L, A, min_qual, n_lib, lg_bins;
};

// data-error codes written to CountTables::error_flag
enum : int32_t { DATA_ERR_LIB = 1, DATA_ERR_TID = 2, DATA_ERR_QUAL = 3, DATA_ERR_CLIP = 4, DATA_ERR_LAYOUT = 5 };

// BAM 4-bit nucleotide code (=ACMGRSVTWYHKDBN) -> device base code; only
// A,C,G,T (one bit set) are bases (statistics.py:27)
__device__ __forceinline__ uint32_t code_of_nibble(uint32_t nib)
{
    return (uint32_t)(0x7777777377727107ull >> (nib * 4)) & 0xFu;
}

// seq.py:4 -- complement; gap and non-bases map to themselves
__device__ __forceinline__ uint32_t complement(uint32_t code) { return code < 4 ? 3 - code : code; }

__device__ __forceinline__ uint32_t read_nibble(const uint8_t *__restrict__ seq4, uint64_t base_index)
{
    uint32_t byte = __ldg(seq4 + (base_index >> 1));
    return (base_index & 1) ? (byte & 0xF) : (byte >> 4);
}

// the device genome is one-hot (1,2,4,8 = A,C,G,T; 0 = anything else), the BAM code of the same base
__device__ __forceinline__ uint32_t ref_code(const uint32_t *__restrict__ words, uint64_t base_index)
{
    return code_of_nibble((__ldg(words + (base_index >> 3)) >> ((uint32_t)(base_index & 7) * 4)) & 0xFu);
}

__device__ __forceinline__ bool op_in_columns(uint32_t op)  // align.py:82: M, I, D, =, X
{
    return (0x187u >> op) & 1u;
}
__device__ __forceinline__ bool op_has_read(uint32_t op)  // M, I, =, X
{
    return (0x183u >> op) & 1u;
}
__device__ __forceinline__ bool op_has_ref(uint32_t op)  // M, D, =, X (columns holding a reference base)
{
    return (0x185u >> op) & 1u;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Totals of one CIGAR, computed by a whole warp.
struct CigarTotals {
    uint32_t columns;   // sum of M, I, D, =, X: alignment columns (align.py:82-87)
    uint32_t ins;       // sum of I
    uint32_t ref_span;  // sum of M, D, N, =, X: aend - pos (htslib)
    uint32_t skipped;   // sum of N
    uint32_t clip_lead; // leading soft clip, hard clips skipped (pysam query_alignment_start)
    uint32_t clip_trail;
    uint32_t first_op, last_op;  // first / last op code of the CIGAR
};

__device__ __forceinline__ CigarTotals cigar_totals(const uint32_t *__restrict__ cigar, int n_cig, int lane)
{
    CigarTotals t;
    if (n_cig == 1) {  // the overwhelmingly common shape: one match block
        uint32_t w = __ldg(cigar);
        uint32_t op = w & 0xF, len = w >> 4;
        t.columns = op_in_columns(op) ? len : 0;
        t.ins = op == OP_I ? len : 0;
        t.skipped = op == OP_N ? len : 0;
        t.ref_span = (op_has_ref(op) || op == OP_N) ? len : 0;
        t.clip_lead = op == OP_S ? len : 0;
        t.clip_trail = 0;
        t.first_op = t.last_op = op;
        return t;
    }
    uint32_t cols = 0, ins = 0, span = 0, skip = 0;
    int first = 0x7fffffff, last = -1;
    for (int k = lane; k < n_cig; k += 32) {
        uint32_t w = __ldg(cigar + k);
        uint32_t op = w & 0xF, len = w >> 4;
        if (op_in_columns(op)) cols += len;
        if (op == OP_I) ins += len;
        if (op == OP_N) skip += len;
        if (op_has_ref(op) || op == OP_N) span += len;
        if (op != OP_S && op != OP_H) {
            first = min(first, k);
            last = max(last, k);
        }
    }
    first = warp_min(first);
    last = warp_max(last);
    uint32_t lead = 0, trail = 0;
    for (int k = lane; k < n_cig; k += 32) {
        uint32_t w = __ldg(cigar + k);
        if ((w & 0xF) == OP_S) {
            if (k < first) lead += w >> 4;
            else if (k > last) trail += w >> 4;
        }
    }
    t.columns = warp_sum(cols);
    t.ins = warp_sum(ins);
    t.ref_span = warp_sum(span);
    t.skipped = warp_sum(skip);
    t.clip_lead = warp_sum(lead);
    t.clip_trail = warp_sum(trail);
    t.first_op = __ldg(cigar) & 0xF;
    t.last_op = __ldg(cigar + n_cig - 1) & 0xF;
    return t;
}

// Where alignment column `col` falls, found by a serial walk of the CIGAR
// (used only for reads with N ops, whose 3'-aligned walk is offset: SURVEY N4).
struct ColumnSite {
    uint32_t op;      // op code, or 0xF when the column is past the last op
    uint32_t query;   // index in the clipped read (valid when op_has_read)
    uint32_t refidx;  // index in the contiguous reference string (valid when op_has_ref or past the end)
};

__device__ __forceinline__ ColumnSite locate_column(const uint32_t *__restrict__ cigar, int n_cig, uint32_t col)
{
    uint32_t c = 0, q = 0, ins = 0;
    for (int k = 0; k < n_cig; ++k) {
        uint32_t w = __ldg(cigar + k);
        uint32_t op = w & 0xF, len = w >> 4;
        if (op_in_columns(op)) {
            if (col < c + len) {
                ColumnSite s;
                s.op = op;
                s.query = op_has_read(op) ? q + (col - c) : q;  // D: read bases before the deletion
                s.refidx = col - ins;
                return s;
            }
            c += len;
            if (op_has_read(op)) q += len;
            if (op == OP_I) ins += len;
        }
    }
    ColumnSite s;
    s.op = 0xF;
    s.query = 0;
    s.refidx = col - ins;
    return s;
}

}  // namespace mdg
