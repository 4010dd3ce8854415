// Host input / output path: BGZF + BAM decode into the SoA batch, BAM encode for the rescale pass.
//
// The reference leaves this to pysam / htslib (reader.py:38,121-132 iterate an AlignmentFile;
// rescale.py:298-299,344 write one); neither is available here, and per-record Python cannot feed
// the kernels, so the decode is native: BGZF blocks are inflated on a pool of threads, record
// boundaries are found in one serial walk, and the records are copied into the struct-of-arrays
// batch in parallel.  BAM already stores CIGAR as len << 4 | op words and SEQ as 4-bit codes with the
// high nibble first, which is exactly the batch layout, so the copy is memcpy.
//
// Format: SAM/BAM specification v1.6, sections 4.1 (BGZF) and 4.2 (BAM).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/mapdamage_b200.h"

namespace {

uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }
uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

// runs fn(i) for i in [0, n) on up to n_threads threads
template <typename F>
void parallel_for(int64_t n, int n_threads, F fn)
{
    if (n <= 0) return;
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
    if (n_threads == 1) {
        for (int64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&] {
            for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto &th : pool) th.join();
}

struct Block {
    size_t in_off, in_len;  // compressed payload inside `compressed`
    uint32_t isize;         // decompressed size
    size_t out_off;         // where it lands in `stream`
};

}  // namespace

// growable byte buffer without value-initialisation (a vector would zero every refill)
struct Bytes {
    uint8_t *p = nullptr;
    size_t len = 0, cap = 0;
    bool pinned = false;  // page-locked (cudaHostAlloc): what the GPU inflater copies from and to at full speed
    ~Bytes() { release(); }
    void release()
    {
        if (pinned) cudaFreeHost(p);
        else free(p);
        p = nullptr;
        cap = 0;
    }
    uint8_t *data() { return p; }
    const uint8_t *data() const { return p; }
    size_t size() const { return len; }
    bool reserve(size_t want)
    {
        if (want <= cap) return true;
        size_t grown = std::max(want, cap + cap / 2 + (1 << 20));
        uint8_t *q = nullptr;
        if (pinned) {
            if (cudaHostAlloc((void **)&q, grown, cudaHostAllocDefault) != cudaSuccess) return false;
            if (len) memcpy(q, p, len);
            cudaFreeHost(p);
        } else {
            q = (uint8_t *)realloc(p, grown);
            if (!q) return false;
        }
        p = q;
        cap = grown;
        return true;
    }
    // switches the allocator; contents are dropped
    void set_pinned(bool on)
    {
        if (on == pinned) return;
        release();
        len = 0;
        pinned = on;
    }
    void drop_front(size_t n)
    {
        memmove(p, p + n, len - n);
        len -= n;
    }
};

// One unit of read-ahead: a slab of the file, its BGZF blocks, and their inflated bytes.
struct Chunk {
    Bytes compressed, inflated;
    std::vector<Block> blocks;
    bool last = false;  // the file ended in this chunk
    int error = 0;
    std::string message;
};

struct mdg_bam_reader {
    FILE *fp = nullptr;
    int n_threads = 1;
    bool native_inflate = true;  // MDG_BAM_ZLIB=1: zlib only (A/B runs, tests)
    std::string error;
    std::string header_text;
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_lengths;
    std::unordered_map<std::string, int32_t> library_of;  // read group id -> library index
    bool merge_libraries = true;
    // reads of the last batch without a usable read group: (index in the batch, the reference's BAMError text).
    // lenient: they get library 0xFFFF and the caller decides (down-sampling: only drawn reads are looked up, reader.py:139-164)
    bool lenient_libraries = false;
    std::vector<std::pair<int64_t, std::string>> library_failures;
    // decompressed bytes not yet consumed
    Bytes stream;
    size_t stream_pos = 0;  // next unread byte
    size_t keep_from = 0;   // bytes before this may be dropped by the next refill (a batch under construction
                            // keeps its first record here; offsets relative to keep_from survive refills)
    bool eof = false;
    uint64_t data_start = 0;  // uncompressed bytes in front of the first record (magic, header text, reference list)
    int64_t records_seen = 0;
    // read-ahead: a producer thread reads and inflates the next chunk while the caller works on this one
    Chunk chunks[2];
    std::thread producer;
    std::mutex mutex;
    std::condition_variable cond;
    int ready[2] = {0, 0};  // 1 = filled by the producer, waiting for the consumer
    int produce_at = 0, consume_at = 0;
    bool stop = false, producer_done = false;
    Bytes carry;  // bytes of a BGZF block cut by the end of a slab
    size_t slab_bytes = 0;  // MDG_BAM_SLAB: slab size of the host decoders (tests: many slabs from small files)
};

struct mdg_bam_writer {
    FILE *fp = nullptr;
    int n_threads = 1;
    int level = 1;
    std::string error;
    std::vector<uint8_t> pending;  // uncompressed bytes not yet written
};

namespace {

thread_local std::string g_open_error;

int rfail(mdg_bam_reader *r, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (r) r->error = buf;
    else g_open_error = buf;
    return code;
}

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr size_t SLAB_BYTES = 32u << 20;

// Producer side: reads one slab of the file, cuts it into BGZF blocks and inflates them in parallel.
void fill_chunk(mdg_bam_reader *r, Chunk &c)
{
    c.blocks.clear();
    c.error = 0;
    c.last = false;
    c.compressed.len = 0;
    const size_t SLAB = r->slab_bytes;
    static const bool timing = getenv("MDG_BAM_TIMING") != nullptr;  // per-slab stage times on stderr
    const double t0 = now_s();
    if (!c.compressed.reserve(r->carry.len + SLAB)) {
        c.error = MDG_ERR_ARGUMENT;
        c.message = "out of host memory";
        return;
    }
    const size_t base = 0;
    const double t1 = now_s();
    memcpy(c.compressed.p, r->carry.p, r->carry.len);
    const size_t got = fread(c.compressed.p + r->carry.len, 1, SLAB, r->fp);
    c.compressed.len = r->carry.len + got;
    const double t2 = now_s();
    r->carry.len = 0;
    const bool file_done = got < SLAB;
    size_t at = base, out_off = 0;
    const uint8_t *in = c.compressed.p;
    while (true) {
        const size_t left = c.compressed.len - at;
        if (left == 0) break;
        int64_t total = -1;  // whole block size, when its header is complete
        if (left >= 18) {
            if (in[at] != 31 || in[at + 1] != 139 || in[at + 2] != 8 || !(in[at + 3] & 4)) {
                c.error = MDG_ERR_DATA;
                c.message = "not a BGZF block (bad gzip member header)";
                return;
            }
            const uint32_t xlen = le16(in + at + 10);
            if (left >= 12 + (size_t)xlen) {
                int64_t bsize = -1;
                for (size_t x = 0; x + 4 <= xlen;) {
                    const uint32_t slen = le16(in + at + 12 + x + 2);
                    if (in[at + 12 + x] == 'B' && in[at + 12 + x + 1] == 'C' && slen == 2 && x + 6 <= xlen)
                        bsize = le16(in + at + 12 + x + 4);
                    x += 4 + slen;
                }
                if (bsize < 0) {
                    c.error = MDG_ERR_DATA;
                    c.message = "BGZF block without a BC subfield";
                    return;
                }
                total = bsize + 1;
                if (total < 12 + (int64_t)xlen + 8) {
                    c.error = MDG_ERR_DATA;
                    c.message = "BGZF block with an impossible size";
                    return;
                }
                if (left >= (size_t)total) {
                    Block b;
                    b.in_off = at + 12 + xlen;
                    b.in_len = (size_t)total - 12 - xlen - 8;
                    b.isize = le32(in + at + total - 4);
                    if (b.isize > 65536) {  // SAM specification 4.1: a block inflates to at most 64 KB
                        c.error = MDG_ERR_DATA;
                        c.message = "BGZF block claims more than 65536 bytes of data";
                        return;
                    }
                    b.out_off = out_off;
                    out_off += b.isize;
                    c.blocks.push_back(b);
                    at += (size_t)total;
                    continue;
                }
            }
        }
        // an incomplete block at the end of the slab: carry it into the next one
        if (file_done) {
            c.error = MDG_ERR_DATA;
            c.message = "truncated BGZF block";
            return;
        }
        r->carry.reserve(left);
        memcpy(r->carry.p, in + at, left);
        r->carry.len = left;
        break;
    }
    c.last = file_done;
    if (!c.inflated.reserve(out_off + 8)) {
        c.error = MDG_ERR_ARGUMENT;
        c.message = "out of host memory";
        return;
    }
    c.inflated.len = out_off;
    const double t3 = now_s();
    const double t4 = now_s();
    std::atomic<int> bad{0};
    parallel_for((int64_t)c.blocks.size(), r->n_threads, [&](int64_t i) {
        const Block &b = c.blocks[(size_t)i];
        if (!b.isize) return;
        const uint32_t want_crc = le32(c.compressed.p + b.in_off + b.in_len);
        // the native decoder first (mdg_inflate.cpp); zlib for anything it turns down or gets wrong
        if (r->native_inflate &&
            mdg_inflate_raw(c.compressed.p + b.in_off, (int64_t)b.in_len, c.inflated.p + b.out_off, (int64_t)b.isize) == (int64_t)b.isize &&
            (uint32_t)crc32(crc32(0L, Z_NULL, 0), c.inflated.p + b.out_off, b.isize) == want_crc)
            return;
        z_stream z;
        memset(&z, 0, sizeof z);
        if (inflateInit2(&z, -15) != Z_OK) {
            bad = 1;
            return;
        }
        z.next_in = c.compressed.p + b.in_off;
        z.avail_in = (uInt)b.in_len;
        z.next_out = c.inflated.p + b.out_off;
        z.avail_out = b.isize;
        const int rc = inflate(&z, Z_FINISH);
        if (rc != Z_STREAM_END || z.avail_out != 0) bad = 1;
        else if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), c.inflated.p + b.out_off, b.isize) != want_crc)
            bad = 2;
        inflateEnd(&z);
    });
    if (timing)
        fprintf(stderr, "slab: %zu blocks, reserve %.3f s, read %.3f s, scan + reserve %.3f s, inflate + crc %.3f s\n", c.blocks.size(), t1 - t0,
                t2 - t1, t3 - t2, now_s() - t4);
    if (bad) {
        c.error = MDG_ERR_DATA;
        c.message = bad == 2 ? "BGZF block fails its CRC32" : "BGZF block does not inflate";
    }
}

void producer_loop(mdg_bam_reader *r)
{
    while (true) {
        Chunk *c;
        {
            std::unique_lock<std::mutex> lock(r->mutex);
            r->cond.wait(lock, [&] { return r->stop || !r->ready[r->produce_at]; });
            if (r->stop) break;
            c = &r->chunks[r->produce_at];
        }
        fill_chunk(r, *c);
        const bool done = c->last || c->error;
        {
            std::lock_guard<std::mutex> lock(r->mutex);
            r->ready[r->produce_at] = 1;
            r->produce_at ^= 1;
            if (done) r->producer_done = true;
        }
        r->cond.notify_all();
        if (done) break;
    }
}

// Consumer side: appends the next inflated chunk to r->stream.
int refill(mdg_bam_reader *r)
{
    if (r->eof) return MDG_OK;
    // drop what nobody needs any more
    if (r->keep_from) {
        r->stream.drop_front(r->keep_from);
        r->stream_pos -= r->keep_from;
        r->keep_from = 0;
    }
    Chunk *c;
    {
        std::unique_lock<std::mutex> lock(r->mutex);
        r->cond.wait(lock, [&] { return r->ready[r->consume_at] != 0; });
        c = &r->chunks[r->consume_at];
    }
    int rc = MDG_OK;
    if (c->error) {
        rc = rfail(r, c->error, "%s", c->message.c_str());
        r->eof = true;
    } else {
        if (!r->stream.reserve(r->stream.len + c->inflated.len + 8)) return rfail(r, MDG_ERR_ARGUMENT, "out of host memory");
        // the copy is the only serial touch of the decompressed bytes; spread it over the pool
        const size_t piece = 4u << 20, n_pieces = (c->inflated.len + piece - 1) / piece;
        uint8_t *dst = r->stream.p + r->stream.len;
        parallel_for((int64_t)n_pieces, std::min(r->n_threads, 8), [&](int64_t i) {
            const size_t at = (size_t)i * piece;
            memcpy(dst + at, c->inflated.p + at, std::min(piece, c->inflated.len - at));
        });
        r->stream.len += c->inflated.len;
        if (c->last) r->eof = true;
    }
    {
        std::lock_guard<std::mutex> lock(r->mutex);
        r->ready[r->consume_at] = 0;
        r->consume_at ^= 1;
    }
    r->cond.notify_all();
    return rc;
}

// makes at least `need` unread bytes available; returns false at a clean end of file
int ensure(mdg_bam_reader *r, size_t need, bool *ok)
{
    while (r->stream.size() - r->stream_pos < need) {
        if (r->eof) {
            *ok = false;
            return r->stream.size() == r->stream_pos ? MDG_OK : rfail(r, MDG_ERR_DATA, "BAM stream ends inside a record");
        }
        int rc = refill(r);
        if (rc) return rc;
    }
    *ok = true;
    return MDG_OK;
}

void stop_producer(mdg_bam_reader *r)
{
    {
        std::lock_guard<std::mutex> lock(r->mutex);
        r->stop = true;
    }
    r->cond.notify_all();
    if (r->producer.joinable()) r->producer.join();
}

int read_header(mdg_bam_reader *r)
{
    bool ok;
    int rc = ensure(r, 12, &ok);
    if (rc) return rc;
    if (!ok || memcmp(r->stream.data() + r->stream_pos, "BAM\1", 4) != 0) return rfail(r, MDG_ERR_DATA, "not a BAM file (bad magic)");
    const uint32_t l_text = le32(r->stream.data() + r->stream_pos + 4);
    rc = ensure(r, 12 + (size_t)l_text, &ok);
    if (rc || !ok) return rc ? rc : rfail(r, MDG_ERR_DATA, "truncated BAM header");
    r->header_text.assign((const char *)r->stream.data() + r->stream_pos + 8, l_text);
    while (!r->header_text.empty() && r->header_text.back() == '\0') r->header_text.pop_back();
    const uint32_t n_ref = le32(r->stream.data() + r->stream_pos + 8 + l_text);
    r->stream_pos += 12 + (size_t)l_text;
    r->data_start = r->stream_pos;
    r->keep_from = r->stream_pos;
    for (uint32_t i = 0; i < n_ref; ++i) {
        rc = ensure(r, 4, &ok);
        if (rc || !ok) return rc ? rc : rfail(r, MDG_ERR_DATA, "truncated BAM reference list");
        const uint32_t l_name = le32(r->stream.data() + r->stream_pos);
        rc = ensure(r, 8 + (size_t)l_name, &ok);
        if (rc || !ok) return rc ? rc : rfail(r, MDG_ERR_DATA, "truncated BAM reference list");
        const char *name = (const char *)r->stream.data() + r->stream_pos + 4;
        r->ref_names.emplace_back(name, l_name ? l_name - 1 : 0);
        r->ref_lengths.push_back(le32(r->stream.data() + r->stream_pos + 4 + l_name));
        r->stream_pos += 8 + (size_t)l_name;
        r->data_start += r->stream_pos - r->keep_from;
        r->keep_from = r->stream_pos;
    }
    return MDG_OK;
}

// value of the RG:Z tag in the auxiliary fields [aux, end), or null
const char *find_read_group(const uint8_t *aux, const uint8_t *end)
{
    while (aux + 3 <= end) {
        const uint8_t t0 = aux[0], t1 = aux[1], type = aux[2];
        aux += 3;
        size_t skip = 0;
        switch (type) {
        case 'A': case 'c': case 'C': skip = 1; break;
        case 's': case 'S': skip = 2; break;
        case 'i': case 'I': case 'f': skip = 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = aux;
            while (z < end && *z) ++z;
            if (z >= end) return nullptr;
            if (t0 == 'R' && t1 == 'G' && type == 'Z') return (const char *)aux;
            skip = (size_t)(z - aux) + 1;
            break;
        }
        case 'B': {
            if (aux + 5 > end) return nullptr;
            const uint8_t sub = aux[0];
            const size_t count = le32(aux + 1);
            const size_t width = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            skip = 5 + count * width;
            break;
        }
        default: return nullptr;
        }
        if ((size_t)(end - aux) < skip) return nullptr;
        aux += skip;
    }
    return nullptr;
}

bool has_tag(const uint8_t *aux, const uint8_t *end, char a, char b)
{
    while (aux + 3 <= end) {
        if (aux[0] == (uint8_t)a && aux[1] == (uint8_t)b) return true;
        const uint8_t type = aux[2];
        aux += 3;
        size_t skip = 0;
        switch (type) {
        case 'A': case 'c': case 'C': skip = 1; break;
        case 's': case 'S': skip = 2; break;
        case 'i': case 'I': case 'f': skip = 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = aux;
            while (z < end && *z) ++z;
            skip = (size_t)(z - aux) + 1;
            break;
        }
        case 'B': {
            if (aux + 5 > end) return false;
            const uint8_t sub = aux[0];
            const size_t width = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            skip = 5 + (size_t)le32(aux + 1) * width;
            break;
        }
        default: return false;
        }
        if ((size_t)(end - aux) < skip) return false;
        aux += skip;
    }
    return false;
}

}  // namespace

extern "C" {

int mdg_bam_open(const char *path, int32_t n_threads, mdg_bam_reader **out)
{
    if (!path || !out) return rfail(nullptr, MDG_ERR_ARGUMENT, "mdg_bam_open: NULL argument");
    *out = nullptr;
    mdg_bam_reader *r = new (std::nothrow) mdg_bam_reader();
    if (!r) return rfail(nullptr, MDG_ERR_ARGUMENT, "out of host memory");
    r->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    {
        const char *env = getenv("MDG_BAM_ZLIB");
        r->native_inflate = !(env && env[0] == '1');
    }
    r->fp = fopen(path, "rb");
    if (!r->fp) {
        rfail(nullptr, MDG_ERR_ARGUMENT, "cannot open %s", path);
        delete r;
        return MDG_ERR_ARGUMENT;
    }
    setvbuf(r->fp, nullptr, _IONBF, 0);  // slabs are read whole
    {
        const char *env = getenv("MDG_BAM_SLAB");
        const long long want = env ? atoll(env) : 0;
        r->slab_bytes = want >= (1 << 16) ? (size_t)want : SLAB_BYTES;
    }
    r->producer = std::thread(producer_loop, r);
    int rc = read_header(r);
    if (rc) {
        g_open_error = r->error;
        stop_producer(r);
        fclose(r->fp);
        delete r;
        return rc;
    }
    *out = r;
    return MDG_OK;
}

void mdg_bam_close(mdg_bam_reader *r)
{
    if (!r) return;
    stop_producer(r);
    if (r->fp) fclose(r->fp);
    delete r;
}

const char *mdg_bam_error(const mdg_bam_reader *r) { return r ? r->error.c_str() : g_open_error.c_str(); }

int64_t mdg_bam_header_text(const mdg_bam_reader *r, char *buf, int64_t cap)
{
    if (!r) return MDG_ERR_ARGUMENT;
    const int64_t n = (int64_t)r->header_text.size();
    if (buf && cap > 0) {
        const int64_t c = std::min(n, cap - 1);
        memcpy(buf, r->header_text.data(), (size_t)c);
        buf[c] = 0;
    }
    return n;
}

int32_t mdg_bam_n_references(const mdg_bam_reader *r) { return r ? (int32_t)r->ref_names.size() : MDG_ERR_ARGUMENT; }

int mdg_bam_reference(const mdg_bam_reader *r, int32_t index, char *name, int32_t cap, uint32_t *length)
{
    if (!r || index < 0 || index >= (int32_t)r->ref_names.size()) return MDG_ERR_ARGUMENT;
    if (name && cap > 0) snprintf(name, (size_t)cap, "%s", r->ref_names[(size_t)index].c_str());
    if (length) *length = r->ref_lengths[(size_t)index];
    return MDG_OK;
}

int mdg_bam_set_libraries(mdg_bam_reader *r, const char *const *read_groups, const uint16_t *library, int32_t n)
{
    if (!r || n < 0 || (n && (!read_groups || !library))) return MDG_ERR_ARGUMENT;
    r->library_of.clear();
    r->merge_libraries = n == 0;
    for (int32_t i = 0; i < n; ++i) r->library_of[read_groups[i]] = library[i];
    return MDG_OK;
}

int64_t mdg_bam_read_batch(mdg_bam_reader *r, const mdg_batch *out, int64_t max_reads, int64_t max_cigar, int64_t max_bases,
                           uint32_t drop_flags, uint8_t *raw, int64_t raw_cap, uint64_t *raw_off, uint8_t *has_mr,
                           int64_t *n_cigar_out, int64_t *n_bases_out)
{
    if (!r || !out || !out->flag || !out->tid || !out->pos || !out->lib || !out->l_seq || !out->base_off || !out->cigar_off ||
        !out->cigar || !out->seq4 || !out->tlen || !out->mtid || !out->mpos)
        return rfail(r, MDG_ERR_ARGUMENT, "mdg_bam_read_batch: every array but qual must be given");
    struct Rec {
        size_t rel;  // offset of the record's block_size field, relative to r->keep_from
        uint32_t size;
        uint64_t base_off, cigar_off, raw_off;
    };
    std::vector<Rec> recs;
    recs.reserve((size_t)std::min<int64_t>(max_reads, 1 << 20));
    uint64_t bases = 0, cigars = 0, raw_used = 0;
    r->keep_from = r->stream_pos;  // everything from here on stays until the batch has been copied out
    // serial walk over record boundaries: cheap next to the copies
    while ((int64_t)recs.size() < max_reads) {
        bool ok;
        int rc = ensure(r, 4, &ok);
        if (rc) return rc;
        if (!ok) break;
        const uint32_t size = le32(r->stream.data() + r->stream_pos);
        if (size < 32) return rfail(r, MDG_ERR_DATA, "BAM record %lld is shorter than its fixed part", (long long)r->records_seen);
        rc = ensure(r, 4 + (size_t)size, &ok);
        if (rc) return rc;
        if (!ok) return rfail(r, MDG_ERR_DATA, "BAM stream ends inside a record");
        const uint8_t *p = r->stream.data() + r->stream_pos + 4;
        const uint32_t l_name = p[8], n_cig = le16(p + 12), flag = le16(p + 14), l_seq = le32(p + 16);
        if (32 + (uint64_t)l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq > size)
            return rfail(r, MDG_ERR_DATA, "BAM record %lld: fields overrun the record", (long long)r->records_seen);
        if (flag & drop_flags) {
            r->records_seen += 1;
            r->stream_pos += 4 + (size_t)size;
            continue;
        }
        const uint64_t padded = ((uint64_t)l_seq + 1) & ~1ull;
        if ((int64_t)(bases + padded) > max_bases || (int64_t)(cigars + n_cig) > max_cigar ||
            (raw && (int64_t)(raw_used + 4 + size) > raw_cap)) {
            if (recs.empty()) return rfail(r, MDG_ERR_CAPACITY, "one BAM record does not fit the batch arrays");
            break;  // this record opens the next batch
        }
        r->records_seen += 1;
        recs.push_back(Rec{r->stream_pos - r->keep_from, size, bases, cigars, raw_used});
        bases += padded;
        cigars += n_cig;
        raw_used += 4 + size;
        r->stream_pos += 4 + (size_t)size;
    }
    const uint8_t *const origin = r->stream.data() + r->keep_from;
    const int64_t n = (int64_t)recs.size();
    std::mutex failure_mutex;
    r->library_failures.clear();
    const int64_t chunk = 4096, n_chunks = (n + chunk - 1) / chunk;
    parallel_for(n_chunks, r->n_threads, [&](int64_t c) {
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i) {
            const Rec &rec = recs[(size_t)i];
            const uint8_t *p = origin + rec.rel + 4;
            const uint32_t l_name = p[8], n_cig = le16(p + 12), l_seq = le32(p + 16);
            ((int32_t *)out->tid)[i] = (int32_t)le32(p);
            ((int32_t *)out->pos)[i] = (int32_t)le32(p + 4);
            ((uint16_t *)out->flag)[i] = le16(p + 14);
            ((uint32_t *)out->l_seq)[i] = l_seq;
            ((int32_t *)out->mtid)[i] = (int32_t)le32(p + 20);
            ((int32_t *)out->mpos)[i] = (int32_t)le32(p + 24);
            ((int32_t *)out->tlen)[i] = (int32_t)le32(p + 28);
            ((uint32_t *)out->base_off)[i] = (uint32_t)rec.base_off;
            ((uint32_t *)out->cigar_off)[i] = (uint32_t)rec.cigar_off;
            const uint8_t *cig = p + 32 + l_name, *seq = cig + 4ull * n_cig, *qual = seq + (l_seq + 1) / 2;
            memcpy((uint32_t *)out->cigar + rec.cigar_off, cig, 4ull * n_cig);
            memcpy((uint8_t *)out->seq4 + rec.base_off / 2, seq, (l_seq + 1) / 2);
            if (out->qual) {
                memcpy((uint8_t *)out->qual + rec.base_off, qual, l_seq);
                if (l_seq & 1) ((uint8_t *)out->qual)[rec.base_off + l_seq] = 0xFF;
            }
            const uint8_t *aux = qual + l_seq, *end = p + rec.size;
            uint16_t lib = 0;
            if (!r->merge_libraries) {
                // reader.py:63-81: a read without a (known) read group is an error unless libraries are merged
                const char *rg = find_read_group(aux, end);
                auto it = rg ? r->library_of.find(rg) : r->library_of.end();
                if (it == r->library_of.end()) {
                    // reader.py:67-81, with the read's name (and group) as the reference prints them with %r
                    const std::string name((const char *)p + 32, l_name ? l_name - 1 : 0);
                    std::string text = "Read '" + name + "' has ";
                    if (rg) text += std::string("read-group not listed in BAM header ('") + rg + "'); either fix BAM or use --merge-libraries";
                    else text += "no read-group. Either fix BAM or use --merge-libraries";
                    std::lock_guard<std::mutex> lock(failure_mutex);
                    if (r->library_failures.size() < 4096) r->library_failures.emplace_back(i, text);
                    lib = 0xFFFF;
                } else {
                    lib = (uint16_t)it->second;
                }
            }
            ((uint16_t *)out->lib)[i] = lib;
            if (has_mr) has_mr[i] = has_tag(aux, end, 'M', 'R');
            if (raw) {
                memcpy(raw + rec.raw_off, p - 4, 4 + (size_t)rec.size);
                raw_off[i] = rec.raw_off;
            }
        }
    });
    r->keep_from = r->stream_pos;  // the batch has been copied out
    ((uint32_t *)out->cigar_off)[n] = (uint32_t)cigars;
    if (raw) raw_off[n] = raw_used;
    if (n_cigar_out) *n_cigar_out = (int64_t)cigars;
    if (n_bases_out) *n_bases_out = (int64_t)bases;
    if (!r->library_failures.empty()) {
        std::sort(r->library_failures.begin(), r->library_failures.end());
        if (!r->lenient_libraries) return rfail(r, MDG_ERR_DATA, "%s", r->library_failures[0].second.c_str());
    }
    return n;
}

int mdg_bam_lenient_libraries(mdg_bam_reader *r, int32_t on)
{
    if (!r) return MDG_ERR_ARGUMENT;
    r->lenient_libraries = on != 0;
    return MDG_OK;
}

int64_t mdg_bam_library_failure(const mdg_bam_reader *r, int64_t k, char *buf, int64_t cap)
{
    if (!r || k < 0) return MDG_ERR_ARGUMENT;
    if (k >= (int64_t)r->library_failures.size()) return -1;
    if (buf && cap > 0) snprintf(buf, (size_t)cap, "%s", r->library_failures[(size_t)k].second.c_str());
    return r->library_failures[(size_t)k].first;
}

int64_t mdg_bam_records_seen(const mdg_bam_reader *r) { return r ? r->records_seen : 0; }

uint64_t mdg_bam_data_start(const mdg_bam_reader *r) { return r ? r->data_start : 0; }

// ---- writer ---------------------------------------------------------------------------------

static int wfail(mdg_bam_writer *w, int code, const char *msg)
{
    if (w) w->error = msg;
    else g_open_error = msg;
    return code;
}

// compresses `pending` into BGZF blocks (all of it when `all`, else whole blocks only) and writes them
static int flush_blocks(mdg_bam_writer *w, bool all)
{
    const size_t block = 0xff00;
    const size_t n_blocks = all ? (w->pending.size() + block - 1) / block : w->pending.size() / block;
    if (!n_blocks) return MDG_OK;
    std::vector<std::vector<uint8_t>> packed(n_blocks);
    std::atomic<int> bad{0};
    parallel_for((int64_t)n_blocks, w->n_threads, [&](int64_t i) {
        const size_t at = (size_t)i * block, len = std::min(block, w->pending.size() - at);
        std::vector<uint8_t> &dst = packed[(size_t)i];
        dst.resize(18 + compressBound((uLong)len) + 8);
        z_stream z;
        memset(&z, 0, sizeof z);
        if (deflateInit2(&z, w->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
            bad = 1;
            return;
        }
        z.next_in = w->pending.data() + at;
        z.avail_in = (uInt)len;
        z.next_out = dst.data() + 18;
        z.avail_out = (uInt)(dst.size() - 18 - 8);
        if (deflate(&z, Z_FINISH) != Z_STREAM_END) bad = 1;
        const size_t clen = z.total_out;
        deflateEnd(&z);
        static const uint8_t head[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
        memcpy(dst.data(), head, 12);
        dst[12] = 'B'; dst[13] = 'C';
        put16(dst.data() + 14, 2);
        put16(dst.data() + 16, (uint32_t)(18 + clen + 8 - 1));
        put32(dst.data() + 18 + clen, (uint32_t)crc32(crc32(0L, Z_NULL, 0), w->pending.data() + at, (uInt)len));
        put32(dst.data() + 18 + clen + 4, (uint32_t)len);
        dst.resize(18 + clen + 8);
        if (dst.size() > 65536) bad = 1;
    });
    if (bad) return wfail(w, MDG_ERR_DATA, "BGZF compression failed");
    for (auto &b : packed)
        if (fwrite(b.data(), 1, b.size(), w->fp) != b.size()) return wfail(w, MDG_ERR_DATA, "write failed");
    w->pending.erase(w->pending.begin(), w->pending.begin() + (ptrdiff_t)std::min(w->pending.size(), n_blocks * block));
    return MDG_OK;
}

int mdg_bam_create(const char *path, const char *header_text, const char *const *ref_names, const uint32_t *ref_lengths,
                   int32_t n_refs, int32_t n_threads, int32_t level, mdg_bam_writer **out)
{
    if (!path || !out || n_refs < 0 || (n_refs && (!ref_names || !ref_lengths))) return wfail(nullptr, MDG_ERR_ARGUMENT, "mdg_bam_create: bad argument");
    *out = nullptr;
    mdg_bam_writer *w = new (std::nothrow) mdg_bam_writer();
    if (!w) return wfail(nullptr, MDG_ERR_ARGUMENT, "out of host memory");
    w->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    w->level = level < 0 ? 1 : std::min(level, 9);
    w->fp = fopen(path, "wb");
    if (!w->fp) {
        delete w;
        return wfail(nullptr, MDG_ERR_ARGUMENT, "cannot create the output BAM");
    }
    setvbuf(w->fp, nullptr, _IOFBF, 1 << 22);
    const std::string text = header_text ? header_text : "";
    std::vector<uint8_t> &p = w->pending;
    p.insert(p.end(), {'B', 'A', 'M', 1});
    uint8_t tmp[4];
    put32(tmp, (uint32_t)text.size());
    p.insert(p.end(), tmp, tmp + 4);
    p.insert(p.end(), text.begin(), text.end());
    put32(tmp, (uint32_t)n_refs);
    p.insert(p.end(), tmp, tmp + 4);
    for (int32_t i = 0; i < n_refs; ++i) {
        const size_t l = strlen(ref_names[i]) + 1;
        put32(tmp, (uint32_t)l);
        p.insert(p.end(), tmp, tmp + 4);
        p.insert(p.end(), ref_names[i], ref_names[i] + l);
        put32(tmp, ref_lengths[i]);
        p.insert(p.end(), tmp, tmp + 4);
    }
    // htslib starts the alignments on a fresh block; do the same
    int rc = flush_blocks(w, true);
    if (rc) {
        g_open_error = w->error;
        fclose(w->fp);
        delete w;
        return rc;
    }
    *out = w;
    return MDG_OK;
}

const char *mdg_bam_writer_error(const mdg_bam_writer *w) { return w ? w->error.c_str() : g_open_error.c_str(); }

// Appends records: raw[raw_off[i] .. raw_off[i + 1]) is record i as read (block_size included).  Where
// status[i] is set, its qualities are replaced by qual[base_off[i] .. + l_seq) and an MR:f tag is appended
// (rescale.py:273-280).
int mdg_bam_write_batch(mdg_bam_writer *w, const uint8_t *raw, const uint64_t *raw_off, int64_t n, const uint8_t *status,
                        const uint8_t *qual, const uint32_t *base_off, const float *mr)
{
    if (!w || n < 0 || (n && (!raw || !raw_off))) return wfail(w, MDG_ERR_ARGUMENT, "mdg_bam_write_batch: bad argument");
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *rec = raw + raw_off[i];
        const size_t len = (size_t)(raw_off[i + 1] - raw_off[i]);
        const size_t at = w->pending.size();
        const bool rescaled = status && (status[i] & 1);
        w->pending.resize(at + len + (rescaled ? 7 : 0));
        memcpy(w->pending.data() + at, rec, len);
        if (rescaled) {
            uint8_t *p = w->pending.data() + at + 4;
            const uint32_t l_name = p[8], n_cig = le16(p + 12), l_seq = le32(p + 16);
            uint8_t *q = p + 32 + l_name + 4ull * n_cig + (l_seq + 1) / 2;
            if (qual && base_off) memcpy(q, qual + base_off[i], l_seq);
            uint8_t *tag = w->pending.data() + at + len;
            tag[0] = 'M'; tag[1] = 'R'; tag[2] = 'f';
            float value = mr ? mr[i] : 0.f;
            memcpy(tag + 3, &value, 4);
            put32(w->pending.data() + at, (uint32_t)(len - 4 + 7));
        }
    }
    return flush_blocks(w, false);
}

// Encodes the records of a struct-of-arrays batch (synthetic data, format conversion): names are
// "<prefix><first_index + i>", MAPQ 37, optional RG:Z tag per library index.
int mdg_bam_write_soa(mdg_bam_writer *w, const mdg_batch *b, int64_t first_index, const char *name_prefix,
                      const char *const *read_group_of_library, int32_t n_libraries)
{
    if (!w || !b || b->n_reads < 0) return wfail(w, MDG_ERR_ARGUMENT, "mdg_bam_write_soa: bad argument");
    if (b->n_reads && (!b->flag || !b->tid || !b->pos || !b->l_seq || !b->base_off || !b->cigar_off || !b->cigar || !b->seq4))
        return wfail(w, MDG_ERR_ARGUMENT, "mdg_bam_write_soa: flag, tid, pos, l_seq, base_off, cigar_off, cigar and seq4 are required");
    const char *prefix = name_prefix ? name_prefix : "r";
    const int64_t chunk = 1 << 16;
    for (int64_t start = 0; start < b->n_reads; start += chunk) {
        const int64_t stop = std::min(b->n_reads, start + chunk);
        for (int64_t i = start; i < stop; ++i) {
            char name[64];
            const int l_name = snprintf(name, sizeof name, "%s%lld", prefix, (long long)(first_index + i)) + 1;
            const uint32_t c0 = b->cigar_off[i], n_cig = b->cigar_off[i + 1] - c0, l_seq = b->l_seq[i];
            const char *rg = nullptr;
            if (read_group_of_library && b->lib && b->lib[i] < n_libraries) rg = read_group_of_library[b->lib[i]];
            const size_t l_rg = rg ? 3 + strlen(rg) + 1 : 0;
            const size_t size = 32 + (size_t)l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq + l_rg;
            const size_t at = w->pending.size();
            w->pending.resize(at + 4 + size);
            uint8_t *p = w->pending.data() + at;
            put32(p, (uint32_t)size);
            p += 4;
            uint32_t span = 0;
            for (uint32_t k = 0; k < n_cig; ++k) {
                const uint32_t op = b->cigar[c0 + k] & 0xF;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += b->cigar[c0 + k] >> 4;
            }
            const int64_t beg = std::max<int64_t>(b->pos[i], 0), end = beg + (span ? span : 1) - 1;
            uint32_t bin = 0;  // reg2bin, SAM specification 5.3
            if (beg >> 14 == end >> 14) bin = (uint32_t)(4681 + (beg >> 14));
            else if (beg >> 17 == end >> 17) bin = (uint32_t)(585 + (beg >> 17));
            else if (beg >> 20 == end >> 20) bin = (uint32_t)(73 + (beg >> 20));
            else if (beg >> 23 == end >> 23) bin = (uint32_t)(9 + (beg >> 23));
            else if (beg >> 26 == end >> 26) bin = (uint32_t)(1 + (beg >> 26));
            put32(p, (uint32_t)b->tid[i]);
            put32(p + 4, (uint32_t)b->pos[i]);
            p[8] = (uint8_t)l_name;
            p[9] = 37;
            put16(p + 10, bin);
            put16(p + 12, n_cig);
            put16(p + 14, b->flag[i]);
            put32(p + 16, l_seq);
            put32(p + 20, (uint32_t)(b->mtid ? b->mtid[i] : -1));
            put32(p + 24, (uint32_t)(b->mpos ? b->mpos[i] : -1));
            put32(p + 28, (uint32_t)(b->tlen ? b->tlen[i] : 0));
            memcpy(p + 32, name, (size_t)l_name);
            uint8_t *q = p + 32 + l_name;
            memcpy(q, b->cigar + c0, 4ull * n_cig);
            q += 4ull * n_cig;
            memcpy(q, b->seq4 + b->base_off[i] / 2, (l_seq + 1) / 2);
            if (l_seq & 1) q[l_seq / 2] &= 0xF0;
            q += (l_seq + 1) / 2;
            if (b->qual) memcpy(q, b->qual + b->base_off[i], l_seq);
            else memset(q, 0xFF, l_seq);
            q += l_seq;
            if (rg) {
                q[0] = 'R'; q[1] = 'G'; q[2] = 'Z';
                memcpy(q + 3, rg, strlen(rg) + 1);
            }
        }
        int rc = flush_blocks(w, false);
        if (rc) return rc;
    }
    return MDG_OK;
}

// Appends finished BGZF blocks (mdg_bam_encode_batch makes them on the GPU) behind whatever is pending.
int mdg_bam_write_raw(mdg_bam_writer *w, const uint8_t *blocks, int64_t n_bytes)
{
    if (!w || n_bytes < 0 || (n_bytes && !blocks)) return wfail(w, MDG_ERR_ARGUMENT, "mdg_bam_write_raw: bad argument");
    int rc = flush_blocks(w, true);
    if (rc) return rc;
    if (!n_bytes) return MDG_OK;
    // one thread copies into the page cache at 3-4 GB/s, and a slab's blocks are gigabytes: several pwrite calls side by side
    if (n_bytes >= (64 << 20) && fflush(w->fp) == 0) {
        const off_t at = ftello(w->fp);
        const int fd = fileno(w->fp);
        if (at >= 0 && fd >= 0) {
            const int n_parts = 4;
            const int64_t piece = (n_bytes + n_parts - 1) / n_parts;
            std::atomic<int> bad{0};
            std::vector<std::thread> pool;
            for (int t = 0; t < n_parts; ++t)
                pool.emplace_back([&, t] {
                    int64_t lo = (int64_t)t * piece, hi = std::min<int64_t>(n_bytes, lo + piece);
                    while (lo < hi) {
                        const ssize_t k = pwrite(fd, blocks + lo, (size_t)std::min<int64_t>(hi - lo, 256 << 20), at + (off_t)lo);
                        if (k <= 0) {
                            bad = 1;
                            return;
                        }
                        lo += k;
                    }
                });
            for (auto &th : pool) th.join();
            if (bad || fseeko(w->fp, at + (off_t)n_bytes, SEEK_SET) != 0) return wfail(w, MDG_ERR_DATA, "write failed");
            return MDG_OK;
        }
    }
    if (fwrite(blocks, 1, (size_t)n_bytes, w->fp) != (size_t)n_bytes) return wfail(w, MDG_ERR_DATA, "write failed");
    return MDG_OK;
}

int mdg_bam_finish(mdg_bam_writer *w)
{
    if (!w) return MDG_ERR_ARGUMENT;
    int rc = flush_blocks(w, true);
    static const uint8_t eof_block[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (!rc && fwrite(eof_block, 1, 28, w->fp) != 28) rc = wfail(w, MDG_ERR_DATA, "write failed");
    if (w->fp && fclose(w->fp) != 0 && !rc) rc = wfail(w, MDG_ERR_DATA, "close failed");
    w->fp = nullptr;
    return rc;
}

void mdg_bam_writer_free(mdg_bam_writer *w)
{
    if (!w) return;
    if (w->fp) fclose(w->fp);
    delete w;
}

}  // extern "C"
