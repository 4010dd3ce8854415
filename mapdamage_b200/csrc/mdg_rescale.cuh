// Rescale pass over every record of the batch: gap-free reads one per thread with word-wide compares
// (rescale_gapfree_kernel), everything else one warp per record (rescale_kernel, fed by a work list).
// The qualities are rewritten in place in the batch's device copy: a record nobody touches is passed through as is
// (rescale.py:344).
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
#include "mdg_swar.cuh"

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
    // optional list of the quality bytes that changed: change_at[k] = index into `qual`, change_q[k] = the new score
    // (what crosses PCIe instead of the whole quality array); entries beyond change_cap are counted, not stored
    uint32_t *change_at;
    uint8_t *change_q;
    unsigned long long *n_changes;
    unsigned long long change_cap;
};

// Appends this lane's pending changes (at most RESCALE_PENDING, packed index << 8 | score is too narrow, so two arrays)
// with one atomic per warp.  Call with the whole warp converged.
constexpr int RESCALE_PENDING = 6;
__device__ __forceinline__ void flush_changes(const RescaleOut &out, const uint32_t (&at)[RESCALE_PENDING],
                                              const uint8_t (&q)[RESCALE_PENDING], int &n_mine, int lane)
{
    const uint32_t any = __ballot_sync(0xffffffffu, n_mine > 0);
    if (!any) return;
    uint32_t inc = (uint32_t)n_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out.n_changes, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0) + (inc - (uint32_t)n_mine);
#pragma unroll
    for (int k = 0; k < RESCALE_PENDING; ++k) {
        if (k < n_mine && base + k < out.change_cap) {
            out.change_at[base + k] = at[k];
            out.change_q[base + k] = q[k];
        }
    }
    n_mine = 0;
}

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
    __syncwarp();  // every lane has looked at the first quality before any is rewritten

    double mr = 0.0;
    // walk the alignment 5'->3': step i is column i (forward) or C-1-i (reverse)
    for (uint32_t base = 0; base < C; base += 32) {
        const uint32_t i = base + lane;
        double add = 0.0;
        bool contributes = false, changed = false;
        uint32_t ref_here = CODE_OTHER, changed_at = 0;
        uint8_t changed_q = 0;
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
                            const uint8_t new_q = m.lut[((size_t)type * m.n_slots + slot) * 94 + q];
                            out.qual[qbase + j] = new_q;
                            if (new_q != q) {
                                changed_at = (uint32_t)(qbase + j);
                                changed_q = new_q;
                                changed = true;
                            }
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
        if (out.change_at) {
            const uint32_t ch = __ballot_sync(0xffffffffu, changed);
            if (ch) {
                unsigned long long base = 0;
                if (lane == __ffs(ch) - 1) base = atomicAdd(out.n_changes, (unsigned long long)__popc(ch));
                base = __shfl_sync(0xffffffffu, base, __ffs(ch) - 1) + __popc(ch & ((1u << lane) - 1u));
                if (changed && base < out.change_cap) {
                    out.change_at[base] = changed_at;
                    out.change_q[base] = changed_q;
                }
            }
        }
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

// `list` / `count`: the records rescale_gapfree_kernel left over, or null: every record of the batch.
__global__ void __launch_bounds__(256) rescale_kernel(DevBatch b, DevRef ref, RescaleModel m, RescaleOut out,
                                                      const uint32_t *__restrict__ list,
                                                      const unsigned long long *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * warps_per_block;
    const int64_t n = list ? (int64_t)*count : b.n_reads;
    uint32_t ref_seen[4] = {0, 0, 0, 0};  // identical in every lane (ballot counts)
    for (int64_t i = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += stride)
        rescale_read(b, ref, m, out, list ? (int64_t)list[i] : i, lane, ref_seen);
    const uint32_t mine = lane == 0 ? ref_seen[0] : lane == 1 ? ref_seen[1] : lane == 2 ? ref_seen[2] : ref_seen[3];
    if (lane < 4 && mine) atomicAdd(out.ref_count + lane, (unsigned long long)mine);
}

// One thread per record, for CIGARs made of match blocks, insertions and deletions with soft clips at the very ends
// (anything a short-read aligner emits for DNA).  Inside a match block column = read base = reference offset, so
// eight columns are one word of the BAM sequence against one word of the one-hot genome, and the four transition
// classes of _record_subs / _rescale_qual_read (rescale.py:106-139,229-247) are bit masks of those words.  Only the
// set bits -- a few per read -- cost a quality load, a table lookup and a store.  Blocks, words and bits are visited
// in 5'->3' order of the read, so the fp64 MR sum keeps the reference's order of additions (rescale.py:244).  An
// insertion has no reference base and changes nothing; a deletion only adds its reference bases to the base counts,
// and only while read bases remain behind it (rescale.py:249-261).  Skips (N), pads, hard clips, reads over a contig
// end and malformed records are appended to `worklist` for rescale_kernel, which spells the reference's semantics
// out column by column.  Histograms live in shared memory (32-bit, flushed once per block) when `hist_words` > 0.
__global__ void __launch_bounds__(256) rescale_gapfree_kernel(DevBatch b, DevRef ref, RescaleModel m, RescaleOut out,
                                                              uint32_t *__restrict__ worklist,
                                                              unsigned long long *__restrict__ work_count, int hist_words)
{
    extern __shared__ uint32_t s_hist[];  // [2][n_slots][94] rescaled | [2][94] reverse transitions
    const int n_sub = 2 * m.n_slots * 94;
    for (int i = threadIdx.x; i < hist_words; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const bool shared_hist = hist_words > 0;
    const int lane = threadIdx.x & 31;
    const uint32_t *const seq32 = (const uint32_t *)b.seq4;
    uint32_t n_pairs = 0, n_improper = 0, n_noqual = 0, n_rescaled = 0, n_too_long = 0;
    uint32_t ref_seen[4] = {0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (b.n_reads + stride - 1) / stride;
    for (int64_t round = 0; round < rounds; ++round) {
        const int64_t r = round * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool complex = false;
        uint32_t pending_at[RESCALE_PENDING];
        uint8_t pending_q[RESCALE_PENDING];
        int n_pending = 0;
        if (r < b.n_reads) do {
            const uint32_t flag = b.flag[r];
            const uint32_t l_seq = b.l_seq[r];
            const uint64_t boff = b.base_off[r];
            out.status[r] = 0;
            out.mr[r] = __int_as_float(0x7fc00000);
            if (flag & 0x4) break;  // rescale.py:301
            if (!(l_seq > 0 && b.qual[boff] != 0xFF)) {  // rescale.py:303-304
                ++n_noqual;
                break;
            }
            const int strand = (flag >> 4) & 1;
            const int tid = b.tid[r];
            const int64_t pos = b.pos[r];
            bool both_ends = true;
            if (flag & 0x1) {  // rescale.py:305-340: only inward-facing mates on one contig
                const bool mate_rev = (flag & 0x20) != 0;
                const int64_t mpos = b.mpos[r];
                const bool same = tid == b.mtid[r];
                if (!((!strand && mate_rev && mpos > pos && same) || (strand && !mate_rev && mpos < pos && same))) {
                    ++n_pairs;
                    ++n_improper;
                    break;
                }
                both_ends = false;  // direction="forward"
            }
            // [S] then M/=/X/I/D in any order, then [S]
            const uint32_t c0 = b.cigar_off[r], c1 = b.cigar_off[r + 1];
            uint32_t lead = 0, trail = 0, n = 0, span = 0, matched = 0;
            bool simple = c1 > c0 && tid >= 0 && tid < ref.n_contigs;
            const uint32_t first_word = simple ? __ldg(b.cigar + c0) : 0;
            for (uint32_t k = c0; k < c1 && simple; ++k) {
                const uint32_t w = k == c0 ? first_word : __ldg(b.cigar + k), op = w & 0xF, len = w >> 4;
                if (op == OP_M || op == OP_EQ || op == OP_X) { n += len; span += len; matched += len; }
                else if (op == OP_I) n += len;
                else if (op == OP_D) span += len;
                else if (op == OP_S && k == c0 && c1 - c0 > 1) lead = len;
                else if (op == OP_S && k + 1 == c1 && k > c0) trail = len;
                else simple = false;
            }
            simple = simple && matched > 0 && (uint64_t)lead + n + trail == l_seq && pos >= 0;
            if (simple) simple = pos + (int64_t)span <= (int64_t)ref.contig_len[tid];
            if (!simple) {
                complex = true;  // rescale_kernel starts over with this record, its statistics included
                break;
            }
            n_pairs += flag & 0x1;
            const uint64_t q0 = boff + lead, ref0 = ref.contig_off[tid] + (uint64_t)pos;
            double mr = 0.0;

            // reference bases of `len` columns from reference offset r_at, as the read's strand sees them
            auto count_reference = [&](uint32_t y) {
                const uint32_t na = __popc(y & K1), nc = __popc(y & (K1 << 1)), ng = __popc(y & (K1 << 2)),
                               nt = __popc(y & (K1 << 3));
                ref_seen[0] += strand ? nt : na;
                ref_seen[1] += strand ? ng : nc;
                ref_seen[2] += strand ? nc : ng;
                ref_seen[3] += strand ? na : nt;
            };
            // one match block: read bases j0 .. j0 + len (0 = first aligned base) on reference offsets r_at ..
            auto match_block = [&](uint32_t j0, uint32_t r_at, uint32_t len) {
                const uint64_t qa = q0 + j0, ra = ref0 + r_at;
                const uint32_t *const qw = seq32 + (qa >> 3), *const rw = ref.words + (ra >> 3);
                const int sq = (int)(qa & 7) << 2, sr = (int)(ra & 7) << 2;
                const int n_words = (int)((len + 7) >> 3);
                // words in 5'->3' order: ascending on the forward strand, descending on the reverse strand; the word
                // pair loaded for one window word is half of the next one's
                const int step = strand ? -1 : 1;
                int w = strand ? n_words - 1 : 0;
                uint32_t q_keep = natural_order(__ldg(qw + w + strand)), r_keep = __ldg(rw + w + strand);
                for (int it = 0; it < n_words; ++it, w += step) {
                    uint32_t q_lo, q_hi, r_lo, r_hi;
                    if (strand) {
                        q_hi = q_keep; r_hi = r_keep;
                        q_lo = natural_order(__ldg(qw + w)); r_lo = __ldg(rw + w);
                        q_keep = q_lo; r_keep = r_lo;
                    } else {
                        q_lo = q_keep; r_lo = r_keep;
                        q_hi = natural_order(__ldg(qw + w + 1)); r_hi = __ldg(rw + w + 1);
                        q_keep = q_hi; r_keep = r_hi;
                    }
                    const uint32_t live = low_nibbles(4 * ((int)len - 8 * w));
                    uint32_t x = __funnelshift_r(q_lo, q_hi, sq) & live;
                    const uint32_t y = __funnelshift_r(r_lo, r_hi, sr) & live;
                    x &= one_hot_nibbles(x) * 15u;  // anything but A/C/G/T pairs with nothing (CODE_OTHER)
                    count_reference(y);  // _record_subs counts the reference base of every walked column
                    const uint32_t xa = x & K1, xc = (x >> 1) & K1, xg = (x >> 2) & K1, xt = (x >> 3) & K1;
                    const uint32_t ya = y & K1, yc = (y >> 1) & K1, yg = (y >> 2) & K1, yt = (y >> 3) & K1;
                    // on the read's own strand: type 0 = T on reference C, type 1 = A on reference G (rescaled);
                    // C on reference T and G on reference A are only counted.  Reverse reads see complements.
                    uint32_t type0 = xt & yc, type1 = xa & yg, back0 = xc & yt, back1 = xg & ya;
                    if (strand) {
                        uint32_t t = type0; type0 = type1; type1 = t;
                        t = back0; back0 = back1; back1 = t;
                    }
                    uint32_t todo = type0 | type1 | back0 | back1;
                    while (todo) {
                        const int bit = strand ? 31 - __clz(todo) : __ffs(todo) - 1;  // bit 4 k of nibble k
                        todo &= ~(1u << bit);
                        const uint32_t j = j0 + 8u * (uint32_t)w + ((uint32_t)bit >> 2);
                        const uint32_t q = b.qual[q0 + j];
                        const uint32_t sel = 1u << bit;
                        if ((type0 | type1) & sel) {
                            const int type = (type1 & sel) ? 1 : 0;
                            // _corr_this_base, rescale.py:49-79
                            const int64_t p5 = (int64_t)(strand ? n - 1 - j : j) + 1;
                            const int64_t back = p5 - (int64_t)n - 1;
                            int64_t pp = p5;
                            if (both_ends && p5 >= -back) pp = back;
                            int slot = 0;
                            if (pp > 0 && pp <= m.len5p) slot = (int)pp;
                            else if (pp < 0 && -pp <= m.len3p) slot = m.len5p + (int)(-pp);
                            if (q > 93) {
                                atomicCAS(out.error_flag, 0, DATA_ERR_QUAL);
                            } else {
                                const int cell = (type * m.n_slots + slot) * 94 + (int)q;
                                const uint8_t new_q = m.lut[cell];
                                out.qual[q0 + j] = new_q;
                                if (new_q != q && out.change_at) {
                                    if (n_pending == RESCALE_PENDING) {
                                        // more changes in one read than a lane holds: this lane appends on its own
                                        const unsigned long long base = atomicAdd(out.n_changes, (unsigned long long)RESCALE_PENDING);
                                        for (int k = 0; k < RESCALE_PENDING; ++k)
                                            if (base + k < out.change_cap) {
                                                out.change_at[base + k] = pending_at[k];
                                                out.change_q[base + k] = pending_q[k];
                                            }
                                        n_pending = 0;
                                    }
#pragma unroll
                                    for (int k = 0; k < RESCALE_PENDING; ++k)
                                        if (k == n_pending) {
                                            pending_at[k] = (uint32_t)(q0 + j);
                                            pending_q[k] = new_q;
                                        }
                                    ++n_pending;
                                }
                                if (slot) mr += m.inc[type * m.n_slots + slot];  // slot 0 adds exactly 0.0
                                if (shared_hist) atomicAdd(s_hist + cell, 1u);
                                else atomicAdd(out.hist_sub + cell, 1ull);
                            }
                        } else if (q <= 93) {
                            const int cell = ((back1 & sel) ? 94 : 0) + (int)q;
                            if (shared_hist) atomicAdd(s_hist + n_sub + cell, 1u);
                            else atomicAdd(out.hist_rev + cell, 1ull);
                        }
                    }
                }
            };
            auto deletion = [&](uint32_t r_at, uint32_t len) {
                const uint64_t ra = ref0 + r_at;
                const uint32_t *const rw = ref.words + (ra >> 3);
                const int sr = (int)(ra & 7) << 2;
                for (uint32_t w = 0; 8 * w < len; ++w)
                    count_reference(__funnelshift_r(__ldg(rw + w), __ldg(rw + w + 1), sr) & low_nibbles(4 * (int)(len - 8 * w)));
            };

            // the operations in 5'->3' order of the read; one loop for both strands keeps a warp's lanes together
            uint32_t j = strand ? n : 0, at = strand ? span : 0;  // read bases / reference bases left of the walk
            uint32_t last_op = OP_M;
            for (uint32_t i = 0; i < c1 - c0; ++i) {
                const uint32_t k = strand ? c1 - 1 - i : c0 + i;
                const uint32_t w = k == c0 ? first_word : __ldg(b.cigar + k), op = w & 0xF, len = w >> 4;
                if (op == OP_S) continue;
                last_op = op;
                if (op == OP_I) {
                    j = strand ? j - len : j + len;
                } else if (op == OP_D) {
                    if (strand) at -= len;
                    if (strand ? j > 0 : j < n) deletion(at, len);
                    if (!strand) at += len;
                } else {
                    if (strand) { j -= len; at -= len; }
                    match_block(j, at, len);
                    if (!strand) { j += len; at += len; }
                }
            }
            if (last_op == OP_D) ++n_too_long;  // the walk ends on gap columns (rescale.py:255-261)
            out.status[r] = 1;
            out.mr[r] = round_5_decimals(mr);
            ++n_rescaled;
        } while (false);
        if (out.change_at) flush_changes(out, pending_at, pending_q, n_pending, lane);
        // warp-aggregated append of the records left to rescale_kernel
        const uint32_t cx = __ballot_sync(0xffffffffu, complex);
        if (cx) {
            unsigned long long base = 0;
            if (lane == __ffs(cx) - 1) base = atomicAdd(work_count, (unsigned long long)__popc(cx));
            base = __shfl_sync(0xffffffffu, base, __ffs(cx) - 1);
            if (complex) worklist[base + __popc(cx & ((1u << lane) - 1u))] = (uint32_t)r;
        }
    }
    auto warp_sum = [](uint32_t v) {
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        return v;
    };
    const uint32_t totals[9] = {warp_sum(n_pairs), warp_sum(n_improper), warp_sum(n_noqual), warp_sum(n_rescaled),
                                warp_sum(n_too_long), warp_sum(ref_seen[0]), warp_sum(ref_seen[1]), warp_sum(ref_seen[2]),
                                warp_sum(ref_seen[3])};
    if (lane < 5 && totals[lane]) atomicAdd(out.stats + lane, (unsigned long long)totals[lane]);
    if (lane >= 5 && lane < 9 && totals[lane]) atomicAdd(out.ref_count + (lane - 5), (unsigned long long)totals[lane]);
    __syncthreads();
    for (int i = threadIdx.x; i < hist_words; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(i < n_sub ? out.hist_sub + i : out.hist_rev + (i - n_sub), (unsigned long long)v);
    }
}

}  // namespace mdg
