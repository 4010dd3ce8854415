// Rescale pass: one warp per record, every record of the batch.
//
// Replaces rescale._rescale_qual_core and _rescale_qual_read
// (rescale.py:195-365).  The new Phred score of a C->T / G->A base is a pure
// function of (type, position slot, old score); the host tabulates it with
// the reference's own expressions (rescale_model.py), so the device only
// indexes `lut` and the rewritten qualities are bit-exact.  The MR tag is the
// sequential fp64 sum of the per-base damage probabilities in 5'->3' order,
// printed with "%.5f" and stored as float32 (rescale.py:244,275,280); both
// the summation order and the decimal rounding are reproduced exactly.
#pragma once
#include "mdg_device.cuh"

namespace mdg {

struct RescaleModel {
    const uint8_t *lut;  // [2][n_slots][94]
    const double *inc;   // [2][n_slots]
    int32_t len5p, len3p, n_slots;
};

struct RescaleOut {
    uint8_t *qual;       // same layout as DevBatch::qual
    float *mr;
    uint8_t *status;
    unsigned long long *stats;  // pairs, improper, without quals, rescaled, too long
    int32_t *error_flag;
    // rescale._record_subs (rescale.py:106-139), integer part: the host turns these into the log summary
    unsigned long long *hist_sub;   // [type C>T, G>A][slot][94]: rescaled columns by position slot and old Phred
    unsigned long long *hist_rev;   // [type T>C, A>G][94]: the reverse transitions by Phred (never rescaled)
    unsigned long long *ref_count;  // [A, C, G, T]: reference bases over all walked columns
};

// float(("%.5f" % x)) narrowed to float32: exact decimal rounding, half to even.
__device__ inline float round_5_decimals(double x)
{
    if (!(x > 0.0)) return 0.0f;
    int e;
    double fr = frexp(x, &e);                      // x = fr * 2^e, fr in [0.5, 1)
    unsigned long long m = (unsigned long long)ldexp(fr, 53);
    e -= 53;                                       // x = m * 2^e exactly
    unsigned __int128 y = (unsigned __int128)m * 100000u;  // x * 1e5 = y * 2^e
    unsigned long long k;
    if (e >= 0) {
        k = (unsigned long long)(y << e);
    } else {
        int s = -e;
        if (s > 72) return 0.0f;                   // y < 2^70 < half
        unsigned __int128 q = y >> s;
        unsigned __int128 rem = y - (q << s);
        unsigned __int128 half = (unsigned __int128)1 << (s - 1);
        if (rem > half || (rem == half && (q & 1))) q += 1;
        k = (unsigned long long)q;
    }
    return (float)((double)k / 100000.0);          // correctly rounded, like strtod
}

__device__ void rescale_read(const DevBatch &b, const DevRef &ref, const RescaleModel &m, const RescaleOut &out,
                             int64_t r, int lane, uint32_t (&ref_seen)[4])
{
    const uint32_t flag = b.flag[r];
    const uint32_t l_seq = b.l_seq[r];
    const uint64_t boff = b.base_off[r];
    // every record is written back, changed or not (rescale.py:344)
    for (uint32_t i = lane; i < l_seq; i += 32) out.qual[boff + i] = b.qual[boff + i];
    if (lane == 0) {
        out.status[r] = 0;
        out.mr[r] = __int_as_float(0x7fc00000);
    }
    if (flag & 0x4) return;  // rescale.py:301
    const bool has_qual = l_seq > 0 && b.qual[boff] != 0xFF;
    if (!has_qual) {  // rescale.py:303-304
        if (lane == 0) atomicAdd(out.stats + 2, 1ull);
        return;
    }
    const int strand = (flag >> 4) & 1;
    const int tid = b.tid[r];
    const int64_t pos = b.pos[r];
    bool both_ends = true;
    if (flag & 0x1) {  // rescale.py:305-340: only inward-facing mates on one contig
        if (lane == 0) atomicAdd(out.stats + 0, 1ull);
        const bool mate_rev = (flag & 0x20) != 0;
        const int64_t mpos = b.mpos[r];
        const bool same = tid == b.mtid[r];
        const bool ok = (!strand && mate_rev && mpos > pos && same) || (strand && !mate_rev && mpos < pos && same);
        if (!ok) {
            if (lane == 0) atomicAdd(out.stats + 1, 1ull);
            return;
        }
        both_ends = false;  // direction="forward"
    }
    const uint32_t c0 = b.cigar_off[r], c1 = b.cigar_off[r + 1];
    const int n_cig = (int)(c1 - c0);
    if (n_cig == 0 || tid < 0 || tid >= ref.n_contigs) {
        if (lane == 0) atomicCAS(out.error_flag, 0, DATA_ERR_TID);
        return;
    }
    const uint32_t *cigar = b.cigar + c0;
    const CigarTotals ct = cigar_totals(cigar, n_cig, lane);
    // rescale.py:266-273: clipped qualities are re-attached only when the outermost
    // op is S; behind a hard clip the reference fails ("quality and sequence mismatch")
    if ((ct.first_op != OP_S && ct.clip_lead) || (ct.last_op != OP_S && ct.clip_trail)) {
        if (lane == 0) atomicCAS(out.error_flag, 0, DATA_ERR_CLIP);
        return;
    }
    const uint32_t clips = ct.clip_lead + ct.clip_trail;
    const uint32_t n = l_seq > clips ? l_seq - clips : 0;
    const uint32_t C = ct.columns;
    const uint64_t contig_off = ref.contig_off[tid];
    const int64_t contig_len = ref.contig_len[tid];
    const uint64_t qbase = boff + ct.clip_lead;
    __syncwarp();  // pass-through copy above is ordered before the rewrites below

    double mr = 0.0;
    // walk the alignment 5'->3': step i is column i (forward) or C-1-i (reverse)
    for (uint32_t base = 0; base < C; base += 32) {
        const uint32_t i = base + lane;
        double add = 0.0;
        bool contributes = false;
        uint32_t ref_here = CODE_OTHER;
        if (i < C) {
            const uint32_t col = strand ? C - 1 - i : i;
            uint32_t op = OP_M, j = col, refidx = col;
            if (n_cig > 1 || !(ct.first_op == OP_M || ct.first_op == OP_EQ || ct.first_op == OP_X)) {
                ColumnSite s = locate_column(cigar, n_cig, col);
                op = s.op; j = s.query; refidx = s.refidx;
            }
            uint32_t gb_counted = CODE_OTHER;  // reference base of this column, if _record_subs sees the column
            const bool read_col = op_has_read(op) && j < n;
            // a deletion column is walked (and its reference base recorded) as long as read bases remain
            // after it in 5'->3' order (rescale.py:249-261); j = read bases to its left
            const bool del_col = op == OP_D && (strand ? j > 0 : j < n);
            if (read_col || del_col) {
                // reference character paired with this column: the reverse-strand zip is
                // anchored at the right end, `skipped` columns further on
                uint32_t gb = CODE_GAP;
                if (strand && ct.skipped) {
                    ColumnSite s2 = locate_column(cigar, n_cig, col + ct.skipped);
                    if (s2.op != OP_I) {
                        int64_t gpos = pos + (int64_t)s2.refidx;
                        gb = (gpos >= 0 && gpos < contig_len) ? ref_code(ref.words, contig_off + gpos) : CODE_OTHER;
                    }
                } else if (op != OP_I) {
                    int64_t gpos = pos + (int64_t)refidx;
                    gb = (gpos >= 0 && gpos < contig_len) ? ref_code(ref.words, contig_off + gpos) : CODE_OTHER;
                }
                if (strand) gb = complement(gb);
                gb_counted = gb;
                if (read_col) {
                    uint32_t rb = code_of_nibble(read_nibble(b.seq4, qbase + j));
                    if (strand) rb = complement(rb);
                    int type = -1;
                    if (rb == 3 && gb == 1) type = 0;       // read T on reference C
                    else if (rb == 0 && gb == 2) type = 1;  // read A on reference G
                    const uint32_t q = b.qual[qbase + j];
                    if (type >= 0) {
                        // _corr_this_base, rescale.py:49-79
                        const int64_t p5 = (int64_t)(strand ? n - 1 - j : j) + 1;
                        const int64_t back = p5 - (int64_t)n - 1;
                        int64_t p = p5;
                        if (both_ends && p5 >= -back) p = back;
                        int slot = 0;
                        if (p > 0 && p <= m.len5p) slot = (int)p;
                        else if (p < 0 && -p <= m.len3p) slot = m.len5p + (int)(-p);
                        if (q > 93) {
                            atomicCAS(out.error_flag, 0, DATA_ERR_QUAL);
                        } else {
                            out.qual[qbase + j] = m.lut[((size_t)type * m.n_slots + slot) * 94 + q];
                            add = m.inc[type * m.n_slots + slot];
                            contributes = slot != 0;  // slot 0 adds exactly 0.0
                            atomicAdd(out.hist_sub + ((size_t)type * m.n_slots + slot) * 94 + q, 1ull);
                        }
                    } else if (q <= 93) {
                        if (rb == 1 && gb == 3) atomicAdd(out.hist_rev + q, 1ull);            // read C on reference T
                        else if (rb == 2 && gb == 0) atomicAdd(out.hist_rev + 94 + q, 1ull);  // read G on reference A
                    }
                }
            }
            if (gb_counted < 4) ref_here = gb_counted;
        }
#pragma unroll
        for (uint32_t g = 0; g < 4; ++g) ref_seen[g] += __popc(__ballot_sync(0xffffffffu, ref_here == g));
        uint32_t mask = __ballot_sync(0xffffffffu, contributes);
        while (mask) {  // sequential fp64 sum in read order (rescale.py:244)
            int src = __ffs(mask) - 1;
            mask &= mask - 1;
            mr += __shfl_sync(0xffffffffu, add, src);
        }
    }
    if (lane == 0) {
        out.status[r] = 1;
        out.mr[r] = round_5_decimals(mr);
        atomicAdd(out.stats + 3, 1ull);
        // trailing gap columns in read order trigger the reference's warning (rescale.py:255-261)
        if (C > 0) {
            ColumnSite s = locate_column(cigar, n_cig, strand ? 0 : C - 1);
            if (s.op == OP_D) atomicAdd(out.stats + 4, 1ull);
        }
    }
}

__global__ void __launch_bounds__(256) rescale_kernel(DevBatch b, DevRef ref, RescaleModel m, RescaleOut out)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * warps_per_block;
    uint32_t ref_seen[4] = {0, 0, 0, 0};  // identical in every lane (ballot counts)
    for (int64_t r = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < b.n_reads; r += stride)
        rescale_read(b, ref, m, out, r, lane, ref_seen);
    const uint32_t mine = lane == 0 ? ref_seen[0] : lane == 1 ? ref_seen[1] : lane == 2 ? ref_seen[2] : ref_seen[3];
    if (lane < 4 && mine) atomicAdd(out.ref_count + lane, (unsigned long long)mine);
}

}  // namespace mdg
