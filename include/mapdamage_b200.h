/*
 * mapdamage_b200 -- C ABI of the B200-native mapDamage hot path.
 *
 * The reference (ginolhac/mapDamage, pure Python on this path) has no FFI or
 * plugin interface; its seam is two Python call sites.  Every entry point
 * below names the reference code it replaces (paths relative to the
 * reference's mapdamage/ package).  INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Life-cycle, memory, multi-GPU, synthetic-input and measurement entry points have no counterpart in the
 * reference (a single-process Python program); they say so by citing nothing.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success or a negative mdg_status; no function throws, exits or falls back
 * to the CPU.  The message for the last failure on a context is returned by
 * mdg_last_error().  A context is bound to one CUDA device and is not
 * thread-safe; calls on different contexts may run concurrently.
 */
#ifndef MAPDAMAGE_B200_H
#define MAPDAMAGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDG_ABI_VERSION 1

/* Class axis of the misincorporation slab (see mdg_fetch_tables). */
#define MDG_N_CLASSES 30
#define MDG_CLASS_SOFTCLIP 29

typedef enum {
    MDG_OK = 0,
    MDG_ERR_ARGUMENT = -1, /* bad pointer / size / configuration                         */
    MDG_ERR_CUDA = -2,     /* a CUDA runtime call failed (message has the CUDA error)     */
    MDG_ERR_NO_DEVICE = -3,/* no usable CUDA device: there is deliberately no CPU path    */
    MDG_ERR_STATE = -4,    /* call order: e.g. counting before mdg_set_reference          */
    MDG_ERR_CAPACITY = -5, /* batch larger than the configured staging capacity           */
    MDG_ERR_DATA = -6,     /* record the reference would also fail on (see message)       */
    MDG_ERR_NCCL = -7      /* NCCL missing or a collective failed                         */
} mdg_status;

typedef struct mdg_ctx mdg_ctx;
typedef struct mdg_dev_batch mdg_dev_batch;

/*
 * Options of one pass.  length/around/min_qual are the reference's
 * --length/-l, --around/-a, --min-basequal/-Q (config.py:143-166);
 * n_libraries = number of (sample, library) pairs (reader.py:47-50).
 */
typedef struct {
    int32_t device;       /* CUDA device ordinal                                          */
    int32_t length;       /* L: positions tabulated from each read end                    */
    int32_t around;       /* A: flanking reference bases tabulated                        */
    int32_t min_qual;     /* bases with Phred < min_qual are masked (align.py:67-71)      */
    int32_t n_libraries;
    int32_t lg_bins;      /* dense fragment-length bins; longer ones go to an overflow list */
    int32_t n_slots;      /* staging slots for streamed batches (>= 2 = double buffering) */
    int32_t reserved;
    int64_t max_reads;    /* staging capacity of one slot, in reads ...                   */
    int64_t max_cigar_ops;/* ... CIGAR words ...                                          */
    int64_t max_bases;    /* ... and base slots (see mdg_batch.n_bases)                   */
} mdg_config;

/*
 * One struct-of-arrays batch of alignment records: what the reference's loops
 * read from each pysam.AlignedSegment (main.py:165-217, rescale.py:300-344).
 * Read i owns CIGAR words cigar[cigar_off[i] .. cigar_off[i+1]) (BAM encoding
 * len << 4 | op), l_seq[i] bases packed 4 bits each (BAM nibble codes, high
 * nibble first) starting at byte base_off[i] / 2 of seq4, and l_seq[i] raw
 * Phred bytes starting at byte base_off[i] of qual.  base_off[i] is even.
 * qual may be NULL (no read has qualities); a read whose first quality byte
 * is 0xFF has none (BAM convention).  lib[i] indexes the library axis.
 *
 * Optional arrays -- a NULL pointer stands for the common case and saves its
 * bytes on the host->device link (the pass is PCIe-bound end to end):
 *   tid       NULL: every read is on contig 0 (single-contig references)
 *   l_seq     NULL: the bases the read's CIGAR consumes (sum of M, I, S, =, X)
 *   lib       NULL: every read is in library 0
 *   tlen      NULL: 0 (only read for proper-pair first mates, statistics.py:121-124)
 *   mtid/mpos NULL: -1 (only read by the rescale pairing rule, rescale.py:318-337)
 *   cigar_off NULL: every read has exactly one CIGAR op: cigar[i] (n_cigar == n_reads), or the one word cigar[0]
 *                   all reads share (n_cigar == 1: untrimmed reads of one length, all "100M")
 *   base_off  NULL: reads are packed back to back, each starting on an even base:
 *                   base_off[i] = sum over k < i of (l_seq[k] rounded up to even)
 */
typedef struct {
    int64_t n_reads;
    int64_t n_cigar;  /* = cigar_off[n_reads]                                  */
    int64_t n_bases;  /* base slots spanned: seq4 has n_bases / 2 bytes, qual n_bases */
    const uint16_t *flag;
    const int32_t *tid;
    const int32_t *pos;
    const uint16_t *lib;
    const uint32_t *l_seq;
    const uint32_t *base_off;
    const uint32_t *cigar_off; /* n_reads + 1 entries */
    const uint32_t *cigar;
    const uint8_t *seq4;
    const uint8_t *qual;
    const int32_t *tlen;
    const int32_t *mtid;
    const int32_t *mpos;
} mdg_batch;

/* ---- life cycle ------------------------------------------------------ */

/* Creates a context on cfg->device; fails with MDG_ERR_NO_DEVICE when no GPU
 * is present.  Replaces the accumulator construction at main.py:147-155. */
int mdg_create(mdg_ctx **out, const mdg_config *cfg);
void mdg_destroy(mdg_ctx *ctx);
/* Message of the last failure (ctx may be NULL for mdg_create failures). */
const char *mdg_last_error(const mdg_ctx *ctx);
int mdg_abi_version(void);

/* "domain:bus:device.function" of a CUDA device: lets the host side place its threads and pinned
 * buffers on the NUMA node the GPU hangs off (engine.bind_host_to_device). */
int mdg_device_pci_bus_id(int32_t device, char *buf, int32_t cap);

/* Page-locked host memory for batches/results (cudaHostAlloc). */
void *mdg_host_alloc(size_t bytes);
void mdg_host_free(void *ptr);

/*
 * Uploads the genome once: replaces the per-read pysam.FastaFile.fetch calls
 * (main.py:115,180; align.py:32-33; rescale.py:213).  packed = one nibble per
 * base, low nibble first; 0..3 = A,C,G,T (either case), 7 = anything else;
 * contig c starts at base contig_off[c] (a multiple of 8) and has
 * contig_len[c] bases.  n_bytes must cover the last contig rounded up to 8
 * bases.
 */
int mdg_set_reference(mdg_ctx *ctx, const uint8_t *packed, int64_t n_bytes, const uint64_t *contig_off,
                      const uint32_t *contig_len, int32_t n_contigs);

/*
 * A, C, G, T counts over the whole uploaded genome: replaces seqtk.comp
 * (seqtk/seqtk.c:56-143) under composition.write_base_comp
 * (composition.py:6-25), whose dnacomp_genome.csv feeds the Bayesian stage.
 */
int mdg_genome_composition(mdg_ctx *ctx, uint64_t *counts4);

/* ---- counting pass: replaces the loop body main.py:165-217 ------------- */

/*
 * Streams one host batch: asynchronous host->device copy into the next
 * staging slot and the counting kernels on that slot's stream.  Returns once
 * the work is queued; the host arrays must stay valid until mdg_sync() or
 * until n_slots further submits have been made.  Applies the read filter of
 * reader.py:121-132 (flag & 0xF04) on the device.
 */
int mdg_count_submit(mdg_ctx *ctx, const mdg_batch *host);

/* Keeps a batch resident in HBM (benchmarks, repeated passes). */
int mdg_batch_upload(mdg_ctx *ctx, const mdg_batch *host, mdg_dev_batch **out);
int mdg_batch_free(mdg_ctx *ctx, mdg_dev_batch *batch);
/* Counting kernels over a resident batch, on the context's compute stream. */
int mdg_count_resident(mdg_ctx *ctx, const mdg_dev_batch *batch);

/* Waits for all queued work; surfaces asynchronous CUDA errors. */
int mdg_sync(mdg_ctx *ctx);
/* Zeroes the tables: fresh accumulators, as main.py:147-155 builds them per run. */
int mdg_reset_tables(mdg_ctx *ctx);

/*
 * Copies the accumulated tables to the host (implies mdg_sync).  Layouts
 * (uint64, C order) -- the state of MisincorporationRates / DNAComposition /
 * FragmentLengths (statistics.py:9-137):
 *   misincorp [n_libraries][end 5p,3p][strand +,-][MDG_N_CLASSES][length]
 *       class 0..3: reference base A,C,G,T; 4 + 5*g + b: reference g read as b,
 *       g,b in A,C,G,T,gap (g != b); MDG_CLASS_SOFTCLIP: soft-clipped bases
 *   dnacomp   [n_libraries][end][strand][A,C,G,T][length + around]
 *       slot d < length: read base at distance d from that end;
 *       slot length + d - 1: flanking reference base at distance d = 1..around
 *   lghist    [n_libraries][kind pe,se][strand][lg_bins]
 * Any pointer may be NULL to skip that table.
 */
int mdg_fetch_tables(mdg_ctx *ctx, uint64_t *misincorp, uint64_t *dnacomp, uint64_t *lghist);
/* Fragment lengths >= lg_bins (FragmentLengths keeps a dict of any length, statistics.py:117-126):
 * rows of {lib, kind, strand, length}; returns the row count (or < 0); at most max_rows rows are written. */
int64_t mdg_fetch_lg_overflow(mdg_ctx *ctx, int32_t *rows, int64_t max_rows);

/* ---- rescale pass: replaces rescale._rescale_qual_core (rescale.py:285-365) */

/*
 * Correction model built on the host from Stats_out_MCMC_correct_prob.csv
 * (rescale.py:23-46): lut[type C>T,G>A][1 + len5p + len3p][94] = new Phred by
 * old Phred, inc[type][slot] = contribution of one rescaled base to MR.
 * Slot 0 = no entry, slot p = 5' position p, slot len5p + p = 3' position -p.
 */
int mdg_set_rescale_model(mdg_ctx *ctx, const uint8_t *lut, const double *inc, int32_t len5p, int32_t len3p);

/*
 * Rescales one host batch (every record, no read filter: rescale.py:300).
 * qual_out (n_bases bytes, same layout as qual), mr_out[n_reads] (the MR:f
 * tag value) and status_out[n_reads] (0 passed through, 1 rescaled) are
 * filled by asynchronous device->host copies; valid after mdg_sync().
 * stats (optional, 8 uint64: pairs, improper pairs, reads without qualities,
 * rescaled reads, alignments longer than the read, 3 reserved) accumulates.
 */
int mdg_rescale_submit(mdg_ctx *ctx, const mdg_batch *host, uint8_t *qual_out, float *mr_out,
                       uint8_t *status_out);
/*
 * The same pass returning only what changed.  A read has a handful of C->T / G->A bases, so instead of the whole quality
 * array (l_seq bytes per read back over PCIe) the device lists the bytes it rewrote: after mdg_rescale_collect(ticket),
 * qual[change_at[k]] = change_q[k] for k < the returned count patches the caller's own array (indices into the batch's
 * qual array); with patch_qual the library does that itself, on a few threads.  mr_out / status_out as above.  `ticket` names the staging slot; collect it before n_slots further
 * submits.  Returns MDG_ERR_CAPACITY when more than a quarter of all bases changed (use mdg_rescale_submit) or the
 * caller's arrays are too small.
 */
int mdg_rescale_submit_sparse(mdg_ctx *ctx, const mdg_batch *host, float *mr_out, uint8_t *status_out, int32_t *ticket);
int64_t mdg_rescale_collect(mdg_ctx *ctx, int32_t ticket, uint32_t *change_at, uint8_t *change_q, int64_t cap, uint8_t *patch_qual);
/*
 * Rescales a batch resident in HBM in place (its quality array is rewritten on the device), on the compute stream:
 * batches made by mdg_bam_stream_next (file -> device) or mdg_batch_upload.  mr_out / status_out (host, optional) are
 * filled by asynchronous copies; the device copies stay with the batch for mdg_bam_encode_batch.
 */
int mdg_rescale_resident(mdg_ctx *ctx, mdg_dev_batch *batch, float *mr_out, uint8_t *status_out);
/* The counters behind the log lines of rescale.py:303-304,335-343,255-261 (see mdg_rescale_submit). */
int mdg_fetch_rescale_stats(mdg_ctx *ctx, uint64_t *stats8);
/*
 * Integer part of the substitution bookkeeping of rescale._record_subs
 * (rescale.py:106-139), accumulated over every rescale submit since
 * mdg_set_rescale_model; the host derives the log summary of
 * _qual_summary_subs / _print_subs (rescale.py:142-192) from it:
 *   sub [type C>T, G>A][1 + len5p + len3p][94]: rescaled columns by model slot and old Phred
 *   rev [type T>C, A>G][94]: the reverse transitions by Phred (their qualities never change)
 *   ref_count [A, C, G, T]: reference bases over all walked alignment columns
 * Any pointer may be NULL.
 */
int mdg_fetch_rescale_hist(mdg_ctx *ctx, uint64_t *sub, uint64_t *rev, uint64_t *ref_count);

/* ---- multi-GPU: one context per rank, tables summed over ranks ---------- */

/* 128-byte NCCL unique id, made on rank 0 and handed to the other ranks by
 * the caller's launcher (torch.distributed / MPI / a file). */
int mdg_nccl_unique_id(void *id128);
int mdg_nccl_init(mdg_ctx *ctx, const void *id128, int32_t rank, int32_t n_ranks);
/*
 * ncclAllReduce(sum, uint64) over all count tables on the compute stream, OUT OF PLACE: each rank's accumulators keep
 * its own counts, the sums over ranks land in a second buffer, and mdg_fetch_tables returns that buffer until this
 * rank counts again (or resets).  So "count, reduce, count more, reduce" and "reduce twice" both give the sum of what
 * every rank has counted so far -- the tables are sums over reads (main.py:165-217), never sums of sums.
 * Fragment lengths beyond lg_bins (mdg_fetch_lg_overflow) stay per rank; the caller concatenates them.
 */
int mdg_allreduce_tables(mdg_ctx *ctx);

/* ---- synthetic input in HBM (bench.py, tests) ------------------------- */

/*
 * Parameters of the seeded synthetic aDNA generator (mirrors
 * mapdamage_b200/synth.py): reads drawn from the uploaded genome, C->T damage
 * with p = damage0 * damage_decay^i at distance i from the left end of the
 * alignment and G->A mirrored from the right end (BAM orientation), uniform
 * substitution errors, CIGARs mixed by the weights mix[] = {plain match, one
 * 1-3 bp insertion, one 1-3 bp deletion, 0-10 bp soft clips}.
 */
typedef struct {
    uint64_t seed;
    int64_t n_reads;
    int32_t len_lo, len_hi;  /* stored read length, uniform in [len_lo, len_hi]       */
    int32_t mix[4];
    int32_t paired;          /* 1: inward-facing proper pairs, records 2q and 2q+1    */
    int32_t with_qual;       /* 1: base qualities uniform in 2..40                     */
    int32_t n_libraries;     /* lib[] uniform in [0, n_libraries)                      */
    int32_t reserved;
    float error_rate, read_n_rate, filtered_rate;
    float damage0, damage_decay;
    float reserved2;
} mdg_synth_params;

/* A random genome made on the device instead of uploaded (benchmarks with a genome far larger than L2): contig c has
 * contig_len[c] uniform A/C/G/T bases, a function of (seed, c, position).  mdg_reference_download copies the image
 * back (one-hot nibbles 1, 2, 4, 8 = A, C, G, T, low nibble = even base, contigs padded to 8 bases) for the checker. */
int mdg_synth_reference(mdg_ctx *ctx, const uint32_t *contig_len, int32_t n_contigs, uint64_t seed);
int mdg_reference_download(mdg_ctx *ctx, uint8_t *one_hot, int64_t n_bytes);
/* Generates a batch directly in device memory (no host copy); needs mdg_set_reference.  params.reserved = 1: read i
 * sits at base i / n_reads of the genome (a coordinate-sorted file) instead of at a uniformly drawn base. */
int mdg_synth_batch(mdg_ctx *ctx, const mdg_synth_params *params, mdg_dev_batch **out);
int mdg_batch_sizes(mdg_ctx *ctx, const mdg_dev_batch *batch, int64_t *n_reads, int64_t *n_cigar, int64_t *n_bases);
/* Copies a resident batch into caller-allocated host arrays (sized by
 * mdg_batch_sizes; host->qual may be NULL).  The arrays are written through
 * the const pointers of mdg_batch. */
int mdg_batch_download(mdg_ctx *ctx, const mdg_dev_batch *batch, const mdg_batch *host);

/* ---- host input / output path: BGZF + BAM --------------- */

/*
 * What the reference gets from pysam / htslib: iteration over a BAM file
 * (reader.py:38,121-132; rescale.py:298,300) and a BAM writer
 * (rescale.py:299,344).  Decoding is done on n_threads host threads (0 = all)
 * straight into the struct-of-arrays batch; no CUDA call is made here.
 */
typedef struct mdg_bam_reader mdg_bam_reader;
typedef struct mdg_bam_writer mdg_bam_writer;

int mdg_bam_open(const char *path, int32_t n_threads, mdg_bam_reader **out);
void mdg_bam_close(mdg_bam_reader *reader);
/* Message of the last failure (reader may be NULL for mdg_bam_open failures). */
const char *mdg_bam_error(const mdg_bam_reader *reader);
/* SAM header text; returns its length (copies at most cap - 1 bytes when buf is given). */
int64_t mdg_bam_header_text(const mdg_bam_reader *reader, char *buf, int64_t cap);
int32_t mdg_bam_n_references(const mdg_bam_reader *reader);
int mdg_bam_reference(const mdg_bam_reader *reader, int32_t index, char *name, int32_t cap, uint32_t *length);
/*
 * Read group -> library index (reader.py:63-81,98-118).  n = 0 merges all
 * libraries (--merge-libraries); otherwise a read without a listed read group
 * fails the batch with MDG_ERR_DATA, as the reference raises BAMError.
 */
int mdg_bam_set_libraries(mdg_bam_reader *reader, const char *const *read_groups, const uint16_t *library, int32_t n);
/*
 * Decodes up to max_reads records whose flag has none of drop_flags (0xF04 for
 * the counting pass, reader.py:9-13; 0 for the rescale pass) into the caller's
 * arrays (every array of *out but qual is required; written through the const
 * pointers; cigar_off gets n + 1 entries).  Stops early when max_cigar words or
 * max_bases base slots would overflow.  Optionally keeps the records verbatim
 * for mdg_bam_write_batch: raw (raw_cap bytes) and raw_off (n + 1 entries), and
 * flags records that already carry an MR tag (rescale.py:277-278).  Returns the
 * number of records decoded (0 at the end of the file) or a negative status.
 */
int64_t mdg_bam_read_batch(mdg_bam_reader *reader, const mdg_batch *out, int64_t max_reads, int64_t max_cigar,
                           int64_t max_bases, uint32_t drop_flags, uint8_t *raw, int64_t raw_cap, uint64_t *raw_off,
                           uint8_t *has_mr, int64_t *n_cigar, int64_t *n_bases);
/*
 * Reads without a usable read group (reader.py:63-81).  By default the first one fails mdg_bam_read_batch with
 * MDG_ERR_DATA and the reference's BAMError text ("Read 'name' has no read-group. ...").  Lenient: such reads get
 * library 0xFFFF instead and the batch succeeds -- the reference looks the library up only for the reads its
 * down-sampler yields (reader.py:134-164), so a caller that down-samples decides after drawing.
 * mdg_bam_library_failure: k-th failing read of the last batch in batch order (at most 4096 are kept): returns its
 * index in the batch (or -1 past the end) and copies the message.
 */
int mdg_bam_lenient_libraries(mdg_bam_reader *reader, int32_t on);
int64_t mdg_bam_library_failure(const mdg_bam_reader *reader, int64_t k, char *buf, int64_t cap);
/* Records walked so far, dropped ones included. */
int64_t mdg_bam_records_seen(const mdg_bam_reader *reader);
/*
 * The same file decoded on the GPU of a context (csrc/mdg_bamdev.cuh): the host only reads the file, slab by slab, into
 * page-locked memory; BGZF inflate, CRC32 check, record boundaries (guessed per 32 KB segment, then verified against the
 * true chain of length prefixes, so the result is exact), the scatter into the struct-of-arrays batch, read group ->
 * library and the MR-tag test all run on the device, one slab ahead of the caller.  What pysam.AlignmentFile iteration
 * is to the reference (reader.py:38,121-132; rescale.py:298-300).
 *   data_start / n_references: uncompressed offset of the first record and the number of reference sequences, as the
 *     host reader found them in the header (mdg_bam_data_start, mdg_bam_n_references);
 *   slab_bytes: compressed bytes per step (0 = 512 MB, never more than the file).
 * mdg_bam_stream_next returns the number of records of the next batch (0 at the end of the file, < 0 on error) and a
 * batch resident in HBM that belongs to the stream: valid for mdg_count_resident / mdg_rescale_resident until the next
 * call.  Records with flag & drop_flags are skipped (0xF04 for the counting pass, 0 for the rescale pass).  A read
 * without a usable read group fails the batch with the reference's BAMError text (reader.py:67-81).
 */
typedef struct mdg_bam_stream mdg_bam_stream;
uint64_t mdg_bam_data_start(const mdg_bam_reader *reader);
int mdg_bam_stream_open(mdg_ctx *ctx, const char *path, uint64_t data_start, int32_t n_references, int64_t slab_bytes,
                        mdg_bam_stream **out);
void mdg_bam_stream_close(mdg_bam_stream *stream);
const char *mdg_bam_stream_error(const mdg_bam_stream *stream);
int mdg_bam_stream_set_libraries(mdg_bam_stream *stream, const char *const *read_groups, const uint16_t *library, int32_t n);
int64_t mdg_bam_stream_next(mdg_bam_stream *stream, uint32_t drop_flags, int32_t with_qual, int32_t want_mr,
                            mdg_dev_batch **out);
/* "already has an MR tag" flags (rescale.py:277-278) of the batch handed out last (needs want_mr). */
int mdg_bam_stream_has_mr(mdg_bam_stream *stream, uint8_t *has_mr, int64_t n);
/* Records walked (dropped ones included), BGZF blocks inflated on the device / redone on the host, segments whose
 * first-record guess the verification pass corrected, seconds spent {reading, decoding, waiting for a batch}. */
int mdg_bam_stream_stats(const mdg_bam_stream *stream, int64_t *records_seen, int64_t *blocks_device, int64_t *blocks_host,
                         int64_t *guesses_wrong, double *seconds3);

/*
 * Raw DEFLATE (RFC 1951) stream -> out; returns the number of bytes written, or a negative code when the stream is
 * damaged, does not fit out_cap, or uses a form this decoder leaves to zlib.  What mdg_bam_read_batch inflates BGZF
 * blocks with (htslib / zlib under pysam.AlignmentFile in the reference, reader.py:38); every block's CRC32 is
 * checked by the caller, and a block rejected here is given to zlib.
 */
int64_t mdg_inflate_raw(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap);

/*
 * The same on the GPU, for many blocks at once: block i is in[in_off[i] .. + in_len[i]) and inflates to isize[i] bytes at
 * out[out_off[i]]; one thread decodes one block, so the call is worth it from a few thousand blocks up.  `in` and `out`
 * are host buffers (copied whole: in_bytes, out_bytes); status[i] is 0 where block i came out with the right length.
 * The inflater owns its stream and device buffers and may be used from another thread than the mdg_ctx of the same
 * device.  The same kernel inflates the slabs of mdg_bam_stream_next, which keeps the bytes in HBM.
 */
typedef struct mdg_inflater mdg_inflater;
int mdg_inflater_create(int32_t device, mdg_inflater **out);
void mdg_inflater_free(mdg_inflater *inflater);
const char *mdg_inflater_error(const mdg_inflater *inflater);
int mdg_inflate_blocks(mdg_inflater *inflater, const uint8_t *in, int64_t in_bytes, const uint64_t *in_off,
                       const uint32_t *in_len, uint8_t *out, int64_t out_bytes, const uint64_t *out_off,
                       const uint32_t *isize, int32_t n, int32_t *status);

/*
 * Down-sampling of the kept reads (reader.py:134-164).  `mt_state` is the state of CPython's random.Random(seed):
 * 624 MT19937 words followed by the position (getstate()[1]); it is advanced in place, so consecutive calls continue
 * the reference's single stream of draws.
 *   fraction:  keep[i] = (random() < fraction) for the next n kept reads, one draw per read (reader.py:139-142).
 *   reservoir: reads first_index .. first_index + n of the stream; read `index` lands in slot `index` while
 *              index < n_slots, else in slot randint(0, index) when that is < n_slots (reader.py:151-158).
 *              slots[n_slots] holds the index of the read currently in each slot; the caller fills it with -1 first
 *              and selects the reads left in it after the last call.
 */
int mdg_sample_fraction(uint32_t *mt_state, double fraction, int64_t n, uint8_t *keep);
int mdg_sample_reservoir(uint32_t *mt_state, int64_t first_index, int64_t n, int64_t n_slots, int64_t *slots);

/* BAM writer (pysam.AlignmentFile(path, "wb", template=...) at rescale.py:298-299): header as given, BGZF blocks
 * deflated at `level` (0-9, < 0 = 1) on n_threads threads. */
int mdg_bam_create(const char *path, const char *header_text, const char *const *ref_names, const uint32_t *ref_lengths,
                   int32_t n_refs, int32_t n_threads, int32_t level, mdg_bam_writer **out);
const char *mdg_bam_writer_error(const mdg_bam_writer *writer);
/*
 * Appends n records kept by mdg_bam_read_batch, in order (rescale.py:344).
 * Where status[i] & 1, the qualities become qual[base_off[i] .. + l_seq) and
 * an MR:f tag with mr[i] is appended (rescale.py:273-280).
 */
int mdg_bam_write_batch(mdg_bam_writer *writer, const uint8_t *raw, const uint64_t *raw_off, int64_t n, const uint8_t *status,
                        const uint8_t *qual, const uint32_t *base_off, const float *mr);
/*
 * Encodes a struct-of-arrays batch as BAM records (synthetic data, format
 * conversion): names are "<name_prefix><first_index + i>", MAPQ 37, and an
 * RG:Z tag read_group_of_library[lib[i]] when that table is given.
 */
int mdg_bam_write_soa(mdg_bam_writer *writer, const mdg_batch *batch, int64_t first_index, const char *name_prefix,
                      const char *const *read_group_of_library, int32_t n_libraries);
/* Appends finished BGZF blocks behind whatever is pending (mdg_bam_encode_batch uses it). */
int mdg_bam_write_raw(mdg_bam_writer *writer, const uint8_t *blocks, int64_t n_bytes);
/*
 * The writer on the GPU, for batches made by mdg_bam_stream_next with drop_flags = 0: every record of the slab is
 * re-emitted in input order (rescale.py:344), a record mdg_rescale_resident marked with its rewritten qualities and
 * an MR:f tag (rescale.py:273-280); the byte stream is cut into 0xff00-byte BGZF blocks, each deflated by one thread
 * block (literal-only dynamic Huffman, or stored when smaller) with its CRC32, packed and copied to the host.  The
 * file write of one batch overlaps the GPU work of the next; mdg_bam_encode_flush waits for the last one (call it before
 * mdg_bam_finish) and reports uncompressed / compressed bytes and seconds {encoding, waiting for the file}.
 */
int mdg_bam_encode_batch(mdg_bam_stream *stream, mdg_dev_batch *batch, mdg_bam_writer *writer);
int mdg_bam_encode_flush(mdg_bam_stream *stream, int64_t *bytes_in, int64_t *bytes_out, double *seconds2);
/* Flushes, writes the BGZF end-of-file block and closes the file. */
int mdg_bam_finish(mdg_bam_writer *writer);
void mdg_bam_writer_free(mdg_bam_writer *writer);

/* ---- measurement helpers (bench.py) ------------------------------------ */

/* CUDA events on the context's compute stream: which = 0 start, 1 stop. */
int mdg_event_record(mdg_ctx *ctx, int32_t which);
int mdg_event_elapsed_ms(mdg_ctx *ctx, float *ms);
/* Kernels launched by this context so far. */
int64_t mdg_launch_count(const mdg_ctx *ctx);
/* Device duration of the counting kernels of the last pass, in ms (events
 * recorded around every launch; valid after mdg_sync). */
int mdg_last_kernel_ms(mdg_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* MAPDAMAGE_B200_H */
