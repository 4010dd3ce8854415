/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the mapDamage
 * per-read hot path.  Nothing under mapdamage_b200/ may call into this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or the CPU baseline.
 *
 * It deliberately follows the reference's *string* formulation (build gapped
 * strings, reverse-complement them, zip over them) instead of the index
 * arithmetic the CUDA kernels use, so that the two implementations share no
 * derivation.  Each function cites the reference lines it restates
 * (paths relative to /root/reference/mapdamage/).
 *
 * Pinning: the reference ships no tests or golden vectors (SURVEY.md section
 * 4), so this oracle is pinned against outputs of the UNMODIFIED reference run
 * in the dev container through oracle/pysam_shim.py; the vectors live in
 * tests/golden/ (made by oracle/gen_golden.py) and tests/test_oracle_golden.py
 * checks every one of them.
 *
 * Input: the same SoA batch the product uses (include/mapdamage_b200.h) but
 * the genome as plain ASCII per contig, so that the product's 4-bit genome
 * packing is itself under test.
 *
 * Output slabs (uint64, zero-initialised by the caller, accumulated into):
 *   misincorp [lib][end 5p=0,3p=1][strand +=0,-=1][30 classes][L]
 *       classes: 0..3 ref base A,C,G,T; 4+5*g+b pair ref g -> read b with
 *       g,b in A,C,G,T,gap = 0..4 (g != b); 29 soft clip
 *   dnacomp   [lib][end][strand][base 4][L + A]
 *       slot d < L: read base at distance d from that end;
 *       slot L + d - 1: flanking reference base at distance d = 1..A
 *   lghist    [lib][kind pe=0,se=1][strand][lg_bins]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
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
    const uint8_t *qual; /* may be NULL */
    const int32_t *tlen;
    const int32_t *mtid;
    const int32_t *mpos;
} mdo_batch;

typedef struct {
    int32_t n_contigs;
    const uint8_t *const *seq; /* ASCII, as in the FASTA (any case) */
    const int64_t *len;
} mdo_ref;

#define N_CLASSES 30
#define CLASS_SOFTCLIP 29

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

static const char NIBBLE_CHAR[] = "=ACMGRSVTWYHKDBN";

typedef struct {
    char *s;
    int64_t len, cap;
} str_t;

static void str_reserve(str_t *a, int64_t cap)
{
    if (cap > a->cap) {
        a->cap = cap * 2 + 64;
        a->s = (char *)realloc(a->s, (size_t)a->cap);
    }
}

/* python: lst[idx:idx] = [ch] * n  (idx is clamped to len) */
static void str_insert(str_t *a, int64_t idx, int64_t n, char ch)
{
    if (idx > a->len) idx = a->len;
    str_reserve(a, a->len + n);
    memmove(a->s + idx + n, a->s + idx, (size_t)(a->len - idx));
    memset(a->s + idx, ch, (size_t)n);
    a->len += n;
}

/* ref.fetch(chrom, start, end).upper()  (main.py:180, align.py:32-33) */
static void fetch_upper(const mdo_ref *ref, int tid, int64_t start, int64_t end, str_t *out)
{
    int64_t len = ref->len[tid];
    if (start < 0) start = 0;
    if (end > len) end = len;
    out->len = 0;
    if (end <= start) return;
    str_reserve(out, end - start);
    for (int64_t i = start; i < end; ++i) {
        char c = (char)ref->seq[tid][i];
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
        out->s[out->len++] = c;
    }
}

/* seq.py:4,33-35 -- translate through TABLE, then reverse */
static char complement_char(char c)
{
    static const char *from = "TGCAMRWSYKVHDBtgcamrwsykvhdb";
    static const char *to = "ACGTKYWSRMBDHVacgtkywsrmbdhv";
    const char *p = strchr(from, c);
    return (p && c) ? to[p - from] : c;
}

static void revcomp(str_t *a)
{
    for (int64_t i = 0, j = a->len - 1; i <= j; ++i, --j) {
        char x = complement_char(a->s[i]), y = complement_char(a->s[j]);
        a->s[i] = y;
        a->s[j] = x;
    }
}

static void reverse(str_t *a)
{
    for (int64_t i = 0, j = a->len - 1; i < j; ++i, --j) {
        char x = a->s[i];
        a->s[i] = a->s[j];
        a->s[j] = x;
    }
}

static int base_code(char c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    case '-': return 4;
    default: return -1;
    }
}

typedef struct {
    str_t refseq, seq, qual, before, after, query, newq;
} scratch_t;

/* align.py:38-73: insert gaps per CIGAR; parse_cigar (align.py:76-88) walks
 * the ops accumulating the alignment column over M, I, D, =, X only. */
static void align_gapped(const uint32_t *cig, int n_cig, str_t *seq, str_t *qual, str_t *refseq)
{
    int64_t tlength = 0;
    for (int k = 0; k < n_cig; ++k) { /* parse_cigar(cigarlist, 1): insertions -> gaps in ref */
        int op = cig[k] & 0xF;
        int64_t len = cig[k] >> 4;
        if (op == OP_I) str_insert(refseq, tlength, len, '-');
        if (op == OP_M || op == OP_I || op == OP_D || op == OP_EQ || op == OP_X) tlength += len;
    }
    tlength = 0;
    for (int k = 0; k < n_cig; ++k) { /* parse_cigar(cigarlist, 2): deletions -> gaps in read */
        int op = cig[k] & 0xF;
        int64_t len = cig[k] >> 4;
        if (op == OP_D) {
            str_insert(seq, tlength, len, '-');
            if (qual) str_insert(qual, tlength, len, '-');
        }
        if (op == OP_M || op == OP_I || op == OP_D || op == OP_EQ || op == OP_X) tlength += len;
    }
}

/* htslib: query = SEQ minus leading/trailing soft clips (hard clips skipped) */
static void clip_bounds(const uint32_t *cig, int n_cig, int64_t l_seq, int64_t *start, int64_t *end)
{
    int64_t s = 0, e = l_seq;
    for (int k = 0; k < n_cig; ++k) {
        int op = cig[k] & 0xF;
        if (op == OP_S) s += cig[k] >> 4;
        else if (op != OP_H) break;
    }
    for (int k = n_cig - 1; k >= 0; --k) {
        int op = cig[k] & 0xF;
        if (op == OP_S) e -= cig[k] >> 4;
        else if (op != OP_H) break;
    }
    if (e < s) e = s;
    *start = s;
    *end = e;
}

static int64_t reference_end(const uint32_t *cig, int n_cig, int64_t pos)
{
    for (int k = 0; k < n_cig; ++k) {
        int op = cig[k] & 0xF;
        if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) pos += cig[k] >> 4;
    }
    return pos;
}

static void load_query(const mdo_batch *b, int64_t r, int64_t start, int64_t end, str_t *seq, str_t *qual,
                       int *has_qual)
{
    int64_t off = b->base_off[r];
    str_reserve(seq, end - start + 1);
    seq->len = 0;
    for (int64_t j = start; j < end; ++j) {
        uint8_t byte = b->seq4[(off + j) >> 1];
        int nib = ((off + j) & 1) ? (byte & 0xF) : (byte >> 4);
        seq->s[seq->len++] = NIBBLE_CHAR[nib];
    }
    *has_qual = b->qual && b->l_seq[r] > 0 && b->qual[off] != 0xFF;
    if (qual) {
        qual->len = 0;
        if (*has_qual) {
            str_reserve(qual, end - start + 1);
            for (int64_t j = start; j < end; ++j) qual->s[qual->len++] = (char)(b->qual[off + j] + 33);
        }
    }
}

/* statistics.py:22-35 */
static void misincorp_update(uint64_t *tbl, int L, const str_t *seq, const str_t *ref, int from_end)
{
    int64_t n = seq->len < ref->len ? seq->len : ref->len;
    if (n > L) n = L;
    for (int64_t i = 0; i < n; ++i) {
        char cs = from_end ? seq->s[seq->len - 1 - i] : seq->s[i];
        char cr = from_end ? ref->s[ref->len - 1 - i] : ref->s[i];
        int bs = base_code(cs), br = base_code(cr);
        if (bs < 0 || br < 0) continue;
        if (br != 4) tbl[(int64_t)br * L + i] += 1;
        if (br != bs) tbl[(int64_t)(4 + 5 * br + bs) * L + i] += 1;
    }
}

/*
 * One pass of the counting loop, main.py:165-217, over reads [start, stop).
 * Returns 0, or -1 if a fragment length does not fit lg_bins, -2 on a
 * malformed record.
 */
int mdo_count(const mdo_batch *b, const mdo_ref *ref, int64_t start, int64_t stop, int L, int A, int minqual,
              int n_lib, int lg_bins, uint64_t *misincorp, uint64_t *dnacomp, uint64_t *lghist)
{
    scratch_t sc;
    memset(&sc, 0, sizeof sc);
    int rc = 0;
    for (int64_t r = start; r < stop; ++r) {
        int flag = b->flag[r];
        if (flag & (0x4 | 0x100 | 0x200 | 0x400 | 0x800)) continue; /* reader.py:121-132 */
        int lib = b->lib[r];
        if (lib >= n_lib) { rc = -2; break; }
        int is_rev = (flag & 0x10) != 0;
        const uint32_t *cig = b->cigar + b->cigar_off[r];
        int n_cig = (int)(b->cigar_off[r + 1] - b->cigar_off[r]);
        int tid = b->tid[r];
        if (tid < 0 || tid >= ref->n_contigs) { rc = -2; break; }
        int64_t pos = b->pos[r];
        int64_t aend = reference_end(cig, n_cig, pos); /* align.py:14-19 */

        /* FragmentLengths.update, statistics.py:117-126 */
        {
            int64_t length = -1;
            int kind = 0;
            if (flag & 0x1) {
                if ((flag & 0x40) && (flag & 0x2)) length = llabs((long long)b->tlen[r]);
            } else {
                kind = 1;
                length = aend - pos;
            }
            if (length >= 0) {
                if (length >= lg_bins) { rc = -1; break; }
                lghist[(((int64_t)lib * 2 + kind) * 2 + is_rev) * lg_bins + length] += 1;
            }
        }

        /* align.get_around, align.py:22-35 */
        int64_t pos_before = pos - A > 0 ? pos - A : 0;
        int64_t pos_after = aend + A < ref->len[tid] ? aend + A : ref->len[tid];
        fetch_upper(ref, tid, pos_before, pos, &sc.before);
        fetch_upper(ref, tid, aend, pos_after, &sc.after);
        fetch_upper(ref, tid, pos, aend, &sc.refseq); /* main.py:180 */

        int64_t qs, qe;
        int has_qual;
        clip_bounds(cig, n_cig, b->l_seq[r], &qs, &qe);
        load_query(b, r, qs, qe, &sc.seq, &sc.qual, &has_qual);
        /* the ungapped query, kept for DNAComposition.update_read */
        str_reserve(&sc.query, sc.seq.len + 1);
        memcpy(sc.query.s, sc.seq.s, (size_t)sc.seq.len);
        sc.query.len = sc.seq.len;

        if (!(minqual && has_qual)) { /* main.py:185-197 */
            align_gapped(cig, n_cig, &sc.seq, NULL, &sc.refseq);
        } else {
            align_gapped(cig, n_cig, &sc.seq, &sc.qual, &sc.refseq);
            for (int64_t i = 0; i < sc.qual.len; ++i) { /* align.py:67-71 */
                if ((int)(unsigned char)sc.qual.s[i] - 33 < minqual && sc.seq.s[i] != '-') {
                    if (i >= sc.seq.len || i >= sc.refseq.len) { rc = -2; break; }
                    sc.seq.s[i] = 'N';
                    sc.refseq.s[i] = 'N';
                }
            }
            if (rc) break;
        }

        if (is_rev) { /* main.py:200-205 */
            revcomp(&sc.refseq);
            revcomp(&sc.seq);
            revcomp(&sc.after);
            revcomp(&sc.before);
            str_t t = sc.before;
            sc.before = sc.after;
            sc.after = t;
        }

        uint64_t *mis_lib = misincorp + (int64_t)lib * 2 * 2 * N_CLASSES * L;
#define MIS(end) (mis_lib + ((int64_t)(end) * 2 + is_rev) * N_CLASSES * L)
        /* update_soft_clipping, statistics.py:37-51 */
        {
            int64_t tlength = 0;
            for (int k = 0; k < n_cig; ++k) {
                int op = cig[k] & 0xF;
                int64_t len = cig[k] >> 4;
                if (op == OP_S) {
                    int end = (tlength == 0) ? (is_rev ? 1 : 0) : (is_rev ? 0 : 1);
                    uint64_t *row = MIS(end) + (int64_t)CLASS_SOFTCLIP * L;
                    for (int64_t i = 0; i < (len < L ? len : L); ++i) row[i] += 1;
                }
                if (op == OP_M || op == OP_I || op == OP_D || op == OP_EQ || op == OP_X) tlength += len;
            }
        }
        misincorp_update(MIS(0), L, &sc.seq, &sc.refseq, 0); /* main.py:210 */
        misincorp_update(MIS(1), L, &sc.seq, &sc.refseq, 1); /* main.py:212 */
#undef MIS

        /* DNAComposition.update_read / update_reference, statistics.py:75-103 */
        uint64_t *comp_lib = dnacomp + (int64_t)lib * 2 * 2 * 4 * (L + A);
#define COMP(end) (comp_lib + ((int64_t)(end) * 2 + is_rev) * 4 * (L + A))
        if (is_rev) revcomp(&sc.query);
        for (int64_t j = 0; j < sc.query.len && j < L; ++j) {
            int c = base_code(sc.query.s[j]);
            if (c >= 0 && c < 4) COMP(0)[(int64_t)c * (L + A) + j] += 1;
            c = base_code(sc.query.s[sc.query.len - 1 - j]);
            if (c >= 0 && c < 4) COMP(1)[(int64_t)c * (L + A) + j] += 1;
        }
        for (int64_t t = 0; t < sc.before.len; ++t) { /* index -len+t -> distance len-t */
            int c = base_code(sc.before.s[t]);
            int64_t d = sc.before.len - t;
            if (c >= 0 && c < 4) COMP(0)[(int64_t)c * (L + A) + L + d - 1] += 1;
        }
        for (int64_t t = 0; t < sc.after.len; ++t) { /* index t+1 */
            int c = base_code(sc.after.s[t]);
            if (c >= 0 && c < 4) COMP(1)[(int64_t)c * (L + A) + L + t] += 1;
        }
#undef COMP
    }
    free(sc.refseq.s); free(sc.seq.s); free(sc.qual.s); free(sc.before.s);
    free(sc.after.s); free(sc.query.s); free(sc.newq.s);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Rescale pass, rescale.py:195-365                                    */
/* ------------------------------------------------------------------ */

typedef struct {
    int32_t max_pos;      /* tables cover positions -max_pos..max_pos */
    const double *ct;     /* [2*max_pos+1], index p + max_pos, 0 where absent (corr_prob.get(..., 0)) */
    const double *ga;
} mdo_corr;

typedef struct {
    uint64_t hist[4][2][130]; /* CT, TC, GA, AG x before/after (rescale.py:82-105) */
    uint64_t ref_count[4];    /* A, C, G, T */
    double pvals[6];          /* CT, CT_before, TC, GA, GA_before, AG */
    uint64_t n_pairs, n_improper, n_without_quals, n_rescaled, n_too_long;
} mdo_subs;

static double phred_to_pval(char ch) /* rescale.py:18-20 */
{
    return pow(10.0, -((double)(unsigned char)ch - 33.0) / 10.0);
}

static char pval_to_phred(double pval) /* rescale.py:13-15 */
{
    return (char)((int)rint(-10 * log10(fabs(pval))) + 33);
}

static double corr_this_base(const mdo_corr *corr, char nt_seq, char nt_ref, int64_t pos, int64_t length,
                             int direction) /* rescale.py:49-79; direction 0 = both, 1 = forward */
{
    int64_t back_pos = pos - length - 1;
    if (direction == 0) {
        if (pos >= llabs(back_pos)) pos = back_pos;
    }
    if (pos < -corr->max_pos || pos > corr->max_pos) return 0.0;
    if (nt_ref == 'C' && nt_seq == 'T') return corr->ct[pos + corr->max_pos];
    if (nt_ref == 'G' && nt_seq == 'A') return corr->ga[pos + corr->max_pos];
    return 0.0;
}

static void record_subs(mdo_subs *subs, char nt_seq, char nt_ref, char q, char newq, double prob_corr)
{ /* rescale.py:108-139 */
    int type = -1;
    if (nt_seq == 'T' && nt_ref == 'C') {
        type = 0;
        subs->pvals[0] += prob_corr;
        subs->pvals[1] += 1 - phred_to_pval(q);
    } else if (nt_seq == 'A' && nt_ref == 'G') {
        type = 2;
        subs->pvals[3] += prob_corr;
        subs->pvals[4] += 1 - phred_to_pval(q);
    } else if (nt_seq == 'C' && nt_ref == 'T') {
        type = 1;
        subs->pvals[2] += 1 - phred_to_pval(q);
    } else if (nt_seq == 'G' && nt_ref == 'A') {
        type = 3;
        subs->pvals[5] += 1 - phred_to_pval(q);
    }
    if (type >= 0) {
        int qb = (unsigned char)q - 33, qa = (unsigned char)newq - 33;
        if (qb >= 0 && qb < 130) subs->hist[type][0][qb] += 1;
        if (qa >= 0 && qa < 130) subs->hist[type][1][qa] += 1;
    }
    int c = base_code(nt_ref);
    if (c >= 0 && c < 4) subs->ref_count[c] += 1;
}

/*
 * _rescale_qual_core + _rescale_qual_read.  qual_out must be a copy-sized
 * buffer (same layout as b->qual); mr_out[r] is the MR:f value (NaN when the
 * record was passed through); status_out[r]: 0 passed through, 1 rescaled.
 * Returns 0, or -3 for the "quality and sequence mismatch" failure that makes
 * the reference return 1 (rescale.py:266-273,378-380).
 */
int mdo_rescale(const mdo_batch *b, const mdo_ref *ref, int64_t start, int64_t stop, const mdo_corr *corr,
                uint8_t *qual_out, float *mr_out, uint8_t *status_out, mdo_subs *subs)
{
    scratch_t sc;
    memset(&sc, 0, sizeof sc);
    int rc = 0;
    for (int64_t r = start; r < stop && !rc; ++r) {
        int flag = b->flag[r];
        int64_t off = b->base_off[r];
        int64_t l_seq = b->l_seq[r];
        status_out[r] = 0;
        mr_out[r] = NAN;
        if (b->qual) memcpy(qual_out + off, b->qual + off, (size_t)l_seq);
        int has_qual = b->qual && l_seq > 0 && b->qual[off] != 0xFF;
        int direction;
        if (flag & 0x4) continue;                         /* rescale.py:301 */
        if (!has_qual) { subs->n_without_quals++; continue; } /* :303 */
        int is_rev = (flag & 0x10) != 0;
        if (flag & 0x1) {                                 /* :305-340 */
            subs->n_pairs++;
            int mate_rev = (flag & 0x20) != 0;
            if (!is_rev && mate_rev && b->mpos[r] > b->pos[r] && b->tid[r] == b->mtid[r]) direction = 1;
            else if (is_rev && !mate_rev && b->mpos[r] < b->pos[r] && b->tid[r] == b->mtid[r]) direction = 1;
            else { subs->n_improper++; continue; }
        } else {
            direction = 0;
        }

        const uint32_t *cig = b->cigar + b->cigar_off[r];
        int n_cig = (int)(b->cigar_off[r + 1] - b->cigar_off[r]);
        int tid = b->tid[r];
        if (n_cig == 0 || tid < 0 || tid >= ref->n_contigs) { rc = -2; break; }
        int64_t pos = b->pos[r];
        int64_t aend = reference_end(cig, n_cig, pos);
        fetch_upper(ref, tid, pos, aend, &sc.refseq);     /* rescale.py:210-213 */
        int64_t qs, qe;
        int hq;
        clip_bounds(cig, n_cig, l_seq, &qs, &qe);
        load_query(b, r, qs, qe, &sc.seq, &sc.qual, &hq);
        int64_t length_read = sc.seq.len;
        align_gapped(cig, n_cig, &sc.seq, &sc.qual, &sc.refseq); /* threshold -100 never masks */
        int64_t length_align = sc.seq.len;
        if (is_rev) { revcomp(&sc.refseq); revcomp(&sc.seq); reverse(&sc.qual); }

        str_reserve(&sc.newq, length_read + 1);
        sc.newq.len = length_read;
        memset(sc.newq.s, 0, (size_t)length_read + 1);
        int64_t pos_on_read = 0;
        double n_rescaled = 0.0;
        int64_t n_cols = length_align;
        if (sc.refseq.len < n_cols) n_cols = sc.refseq.len;
        if (sc.qual.len < n_cols) n_cols = sc.qual.len;
        for (int64_t i = 0; i < n_cols; ++i) {            /* rescale.py:228-261 */
            char nt_seq = sc.seq.s[i], nt_ref = sc.refseq.s[i], nt_qual = sc.qual.s[i], newq;
            double newp;
            if ((nt_seq == 'T' && nt_ref == 'C') || (nt_seq == 'A' && nt_ref == 'G')) {
                double pdam = 1 - corr_this_base(corr, nt_seq, nt_ref, pos_on_read + 1, length_read, direction);
                double pseq = 1 - phred_to_pval(nt_qual);
                newp = pdam * pseq;
                newq = pval_to_phred(1 - newp);
                n_rescaled += 1 - pdam;
            } else {
                newp = 1 - phred_to_pval(nt_qual);
                newq = nt_qual;
            }
            if (pos_on_read < length_read) {
                sc.newq.s[pos_on_read] = newq;
                record_subs(subs, nt_seq, nt_ref, nt_qual, newq, newp);
                if (nt_seq != '-') pos_on_read += 1;
            } else {
                subs->n_too_long++;                       /* warning, rescale.py:255-261 */
                break;
            }
        }
        if (is_rev) reverse(&sc.newq);
        /* re-attach soft-clipped qualities only when the outermost op is S (rescale.py:266-271) */
        int64_t lead = ((cig[0] & 0xF) == OP_S) ? (int64_t)(cig[0] >> 4) : 0;
        int64_t trail = ((cig[n_cig - 1] & 0xF) == OP_S) ? (int64_t)(cig[n_cig - 1] >> 4) : 0;
        if (lead > l_seq) lead = l_seq;
        if (trail > l_seq) trail = l_seq;
        if (lead + sc.newq.len + trail != l_seq) { rc = -3; break; } /* pysam: ValueError -> rc 1 */
        for (int64_t j = 0; j < sc.newq.len; ++j) qual_out[off + lead + j] = (uint8_t)((unsigned char)sc.newq.s[j] - 33);
        char text[64];
        snprintf(text, sizeof text, "%.5f", n_rescaled);  /* rescale.py:275 */
        mr_out[r] = (float)strtod(text, NULL);            /* set_tag("MR", v, "f") */
        status_out[r] = 1;
        subs->n_rescaled++;
    }
    free(sc.refseq.s); free(sc.seq.s); free(sc.qual.s); free(sc.before.s);
    free(sc.after.s); free(sc.query.s); free(sc.newq.s);
    return rc;
}

/* The exact value the reference would put in the rescale LUT cell (Q, c):
 * exposed so tests can compare the host LUT builder with libm directly. */
int mdo_rescaled_phred(int q, double corr)
{
    double pdam = 1 - corr;
    double pseq = 1 - phred_to_pval((char)(q + 33));
    return (int)(unsigned char)pval_to_phred(1 - pdam * pseq) - 33;
}

size_t mdo_sizeof_subs(void) { return sizeof(mdo_subs); }
