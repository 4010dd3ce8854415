// Raw DEFLATE (RFC 1951) decoder, shared by the host build (mdg_inflate.cpp) and the device build
// (mdg_inflate_dev.cuh): see mdg_inflate.cpp for what it is for and how it is checked.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define MDG_HD __host__ __device__
#else
#define MDG_HD
#endif
#ifdef __CUDA_ARCH__
#define MDG_CONST __device__ const
#else
#define MDG_CONST const
#endif

namespace mdg_inflate {


// first-level table sizes: on the device a thread's tables have to stay in L1 next to those of its 31 neighbours
// (the layout of InflateScratch differs between the two builds; only device code ever touches a device scratch)
#ifdef __CUDA_ARCH__
constexpr int LITLEN_BITS = 9, OFFSET_BITS = 7, PRECODE_BITS = 7;
#else
constexpr int LITLEN_BITS = 11, OFFSET_BITS = 8, PRECODE_BITS = 7;
#endif
constexpr int LITLEN_CAP = 2048 + 1024, OFFSET_CAP = 256 + 512, PRECODE_CAP = 128;

// table entry: bits 0-4 code length | bits 5-7 kind | bits 8-12 extra bits (or second-level index bits) | 16-31 value
enum Kind : uint32_t { INVALID = 0, LITERAL = 1, BASE = 2, END = 3, SUBTABLE = 4 };
MDG_HD inline uint32_t entry(uint32_t len, Kind kind, uint32_t extra, uint32_t value)
{
    return len | (uint32_t)kind << 5 | extra << 8 | value << 16;
}

MDG_CONST uint16_t LENGTH_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                  67, 83, 99, 115, 131, 163, 195, 227, 258};
MDG_CONST uint8_t LENGTH_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
MDG_CONST uint16_t OFFSET_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                  1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
MDG_CONST uint8_t OFFSET_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
MDG_CONST uint8_t PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Alphabet { LITLEN, OFFSET, PRECODE };

MDG_HD inline uint32_t symbol_entry(Alphabet alphabet, uint32_t sym, uint32_t len)
{
    if (alphabet == PRECODE) return entry(len, LITERAL, 0, sym);
    if (alphabet == OFFSET) return sym < 30 ? entry(len, BASE, OFFSET_EXTRA[sym], OFFSET_BASE[sym]) : entry(len, INVALID, 0, 0);
    if (sym < 256) return entry(len, LITERAL, 0, sym);
    if (sym == 256) return entry(len, END, 0, 0);
    if (sym < 286) return entry(len, BASE, LENGTH_EXTRA[sym - 257], LENGTH_BASE[sym - 257]);
    return entry(len, INVALID, 0, 0);
}

MDG_HD inline uint32_t reverse_bits(uint32_t code, int len)
{
#ifdef __CUDA_ARCH__
    return __brev(code) >> (32 - len);
#else
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
#endif
}

// Canonical Huffman decode table from code lengths (RFC 1951 3.2.2).  false: over-subscribed, or incomplete in a way
// the format does not allow, or larger than the table.
MDG_HD inline bool build_table(const uint8_t *lens, int n_syms, Alphabet alphabet, int table_bits, uint32_t *table, int cap,
                               uint16_t *rev, uint8_t *longest)
{
    int count[16] = {0};
    for (int s = 0; s < n_syms; ++s) ++count[lens[s]];
    count[0] = 0;
    int used = 0;
    int64_t space = 0;  // in units of 2^-15
    for (int l = 1; l <= 15; ++l) {
        used += count[l];
        space += (int64_t)count[l] << (15 - l);
    }
    const int primary = 1 << table_bits;
    for (int i = 0; i < primary; ++i) table[i] = 0;
    if (used == 0) return alphabet == OFFSET;  // a block of literals only
    if (space > (1 << 15)) return false;
    if (space < (1 << 15)) {
        // incomplete: allowed only as one distance code of one bit
        if (!(alphabet == OFFSET && used == 1 && count[1] == 1)) return false;
    }
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
        next_code[l] = code;
        code = (code + (uint32_t)count[l]) << 1;
    }
    // rev[288]; longest[1 << LITLEN_BITS]: per first-level index, the longest code behind it (0: none longer than the table)
    bool any_long = false;
    for (int s = 0; s < n_syms; ++s) {
        const int l = lens[s];
        if (!l) continue;
        rev[s] = (uint16_t)reverse_bits(next_code[l]++, l);
        if (l > table_bits) {
            if (!any_long) {
                for (int i = 0; i < primary; ++i) longest[i] = 0;
                any_long = true;
            }
            uint8_t &m = longest[rev[s] & (primary - 1)];
            if (l > m) m = (uint8_t)l;
        }
    }
    int next_free = primary;
    if (any_long) {
        for (int i = 0; i < primary; ++i) {
            if (!longest[i]) continue;
            const int sub_bits = longest[i] - table_bits;
            if (next_free + (1 << sub_bits) > cap) return false;
            table[i] = entry((uint32_t)table_bits, SUBTABLE, (uint32_t)sub_bits, (uint32_t)next_free);
            for (int k = 0; k < (1 << sub_bits); ++k) table[next_free + k] = 0;
            next_free += 1 << sub_bits;
        }
    }
    for (int s = 0; s < n_syms; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t e = symbol_entry(alphabet, (uint32_t)s, (uint32_t)l);
        if (l <= table_bits) {
            for (int k = rev[s]; k < primary; k += 1 << l) table[k] = e;
        } else {
            const uint32_t p = table[rev[s] & (primary - 1)];
            const int sub_bits = (int)((p >> 8) & 31), start = (int)(p >> 16);
            for (int k = rev[s] >> table_bits; k < (1 << sub_bits); k += 1 << (l - table_bits)) table[start + k] = e;
        }
    }
    return true;
}

struct Tables {
    uint32_t litlen[LITLEN_CAP];
    uint32_t offset[OFFSET_CAP];
};

// everything one decoder instance needs besides the stream: the stack on the host, a slice of global memory per
// thread on the device
struct InflateScratch {
    Tables t;
    uint32_t precode[PRECODE_CAP];
    uint16_t rev[288];
    uint8_t longest[1 << LITLEN_BITS];
    uint8_t lens[288 + 32 + 140];
};

MDG_HD inline uint64_t load64(const uint8_t *p)
{
#ifdef __CUDA_ARCH__
    // eight bytes from any address: two aligned 32-bit words cover four, three cover eight
    const uint32_t *w = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    const int shift = (int)((uintptr_t)p & 3) * 8;
    const uint32_t a = w[0], b = w[1], c = shift ? w[2] : 0;
    const uint64_t lo = (uint64_t)a | (uint64_t)b << 32;
    return shift ? (lo >> shift) | ((uint64_t)c << (64 - shift)) : lo;
#else
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64, aarch64)
#endif
}

struct BitReader {
    const uint8_t *in, *in_end;
    uint64_t bits = 0;
    int cnt = 0;  // valid bits in `bits`
    bool overrun = false;

    MDG_HD void refill_fast()  // needs in + 8 <= in_end
    {
        bits |= load64(in) << cnt;
        in += (63 - cnt) >> 3;
        cnt |= 56;
    }
    MDG_HD void refill()
    {
        if (in_end - in >= 8) {
            refill_fast();
            return;
        }
        while (cnt <= 56 && in < in_end) {
            bits |= (uint64_t)*in++ << cnt;
            cnt += 8;
        }
    }
    MDG_HD uint32_t peek(int n) const { return (uint32_t)(bits & ((1ull << n) - 1)); }
    MDG_HD void consume(int n)
    {
        bits >>= n;
        cnt -= n;
        if (cnt < 0) overrun = true;
    }
    MDG_HD uint32_t take(int n)
    {
        const uint32_t v = peek(n);
        consume(n);
        return v;
    }
};

MDG_HD inline uint32_t lookup(const uint32_t *table, int table_bits, uint64_t bits)
{
    uint32_t e = table[bits & ((1u << table_bits) - 1)];
    if (((e >> 5) & 7) == SUBTABLE) e = table[(e >> 16) + ((bits >> table_bits) & ((1u << ((e >> 8) & 31)) - 1))];
    return e;
}

// One Huffman-coded block.  `fast` regions need 32 bytes of input and 300 bytes of output in hand; the rest is
// decoded with every bound checked.
MDG_HD inline bool inflate_block(BitReader &br, const Tables &t, uint8_t *const out_start, uint8_t *&out_at, uint8_t *const out_end)
{
    uint8_t *out = out_at;
    while (true) {
        // ---- fast: no bound can be hit within one iteration ----
        while (br.in_end - br.in >= 32 && out_end - out >= 300) {
            br.refill_fast();
            uint32_t e = lookup(t.litlen, LITLEN_BITS, br.bits);
            uint32_t kind = (e >> 5) & 7;
            if (kind == LITERAL) {
                br.bits >>= e & 31; br.cnt -= e & 31;
                *out++ = (uint8_t)(e >> 16);
                e = lookup(t.litlen, LITLEN_BITS, br.bits);
                kind = (e >> 5) & 7;
                if (kind == LITERAL) {
                    br.bits >>= e & 31; br.cnt -= e & 31;
                    *out++ = (uint8_t)(e >> 16);
                    e = lookup(t.litlen, LITLEN_BITS, br.bits);
                    kind = (e >> 5) & 7;
                    if (kind == LITERAL) {
                        br.bits >>= e & 31; br.cnt -= e & 31;
                        *out++ = (uint8_t)(e >> 16);
                        continue;
                    }
                }
                br.refill_fast();
            }
            if (kind != BASE) {
                if (kind == END) {
                    br.bits >>= e & 31; br.cnt -= e & 31;
                    out_at = out;
                    return true;
                }
                return false;
            }
            br.bits >>= e & 31; br.cnt -= e & 31;
            const int lx = (int)((e >> 8) & 31);
            const uint32_t length = (e >> 16) + (uint32_t)(br.bits & ((1u << lx) - 1));
            br.bits >>= lx; br.cnt -= lx;
            br.refill_fast();
            const uint32_t d = lookup(t.offset, OFFSET_BITS, br.bits);
            if (((d >> 5) & 7) != BASE) return false;
            br.bits >>= d & 31; br.cnt -= d & 31;
            const int dx = (int)((d >> 8) & 31);
            const uint32_t offset = (d >> 16) + (uint32_t)(br.bits & ((1u << dx) - 1));
            br.bits >>= dx; br.cnt -= dx;
            if (offset > (uint32_t)(out - out_start)) return false;
            const uint8_t *src = out - offset;
            uint8_t *dst = out;
            out += length;
            if (offset >= 8) {
#ifdef __CUDA_ARCH__
                do *dst++ = *src++; while (dst < out);
#else
                do {
                    memcpy(dst, src, 8);
                    dst += 8; src += 8;
                } while (dst < out);
#endif
            } else if (offset == 1) {
                const uint8_t fill = *src;
                do *dst++ = fill; while (dst < out);
            } else {
                do *dst++ = *src++; while (dst < out);
            }
        }
        // ---- careful: one symbol, every bound checked ----
        br.refill();
        uint32_t e = lookup(t.litlen, LITLEN_BITS, br.bits);
        uint32_t kind = (e >> 5) & 7;
        br.consume((int)(e & 31));
        if (br.overrun) return false;
        if (kind == LITERAL) {
            if (out >= out_end) return false;
            *out++ = (uint8_t)(e >> 16);
            continue;
        }
        if (kind == END) {
            out_at = out;
            return true;
        }
        if (kind != BASE) return false;
        const uint32_t length = (e >> 16) + br.take((int)((e >> 8) & 31));
        br.refill();
        const uint32_t d = lookup(t.offset, OFFSET_BITS, br.bits);
        if (((d >> 5) & 7) != BASE) return false;
        br.consume((int)(d & 31));
        const uint32_t offset = (d >> 16) + br.take((int)((d >> 8) & 31));
        if (br.overrun || offset > (uint32_t)(out - out_start) || length > (uint32_t)(out_end - out)) return false;
        const uint8_t *src = out - offset;
        for (uint32_t i = 0; i < length; ++i) out[i] = src[i];
        out += length;
    }
}

MDG_HD inline bool read_dynamic_header(BitReader &br, Tables &t, InflateScratch &s)
{
    br.refill();
    const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t pre_lens[19] = {0};
    for (int i = 0; i < hclen; ++i) {
        br.refill();
        pre_lens[PRECODE_ORDER[i]] = (uint8_t)br.take(3);
    }
    if (br.overrun) return false;
    uint32_t *const precode = s.precode;
    if (!build_table(pre_lens, 19, PRECODE, PRECODE_BITS, precode, PRECODE_CAP, s.rev, s.longest)) return false;
    uint8_t *const lens = s.lens;
    int n = 0;
    while (n < hlit + hdist) {
        br.refill();
        const uint32_t e = precode[br.bits & ((1u << PRECODE_BITS) - 1)];
        if (((e >> 5) & 7) != LITERAL) return false;
        br.consume((int)(e & 31));
        const uint32_t sym = e >> 16;
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
        } else if (sym == 16) {
            if (!n) return false;
            const int rep = 3 + (int)br.take(2);
            for (int i = 0; i < rep; ++i) lens[n + i] = lens[n - 1];
            n += rep;
        } else if (sym == 17) {
            const int rep = 3 + (int)br.take(3);
            for (int i = 0; i < rep; ++i) lens[n + i] = 0;
            n += rep;
        } else {
            const int rep = 11 + (int)br.take(7);
            for (int i = 0; i < rep; ++i) lens[n + i] = 0;
            n += rep;
        }
        if (br.overrun) return false;
    }
    if (n != hlit + hdist || !lens[256]) return false;
    return build_table(lens, hlit, LITLEN, LITLEN_BITS, t.litlen, LITLEN_CAP, s.rev, s.longest) &&
           build_table(lens + hlit, hdist, OFFSET, OFFSET_BITS, t.offset, OFFSET_CAP, s.rev, s.longest);
}

MDG_HD inline bool fixed_tables(Tables &t, InflateScratch &s)
{
    uint8_t *const lens = s.lens;
    int i = 0;
    for (; i < 144; ++i) lens[i] = 8;
    for (; i < 256; ++i) lens[i] = 9;
    for (; i < 280; ++i) lens[i] = 7;
    for (; i < 288; ++i) lens[i] = 8;
    for (; i < 320; ++i) lens[i] = 5;
    return build_table(lens, 288, LITLEN, LITLEN_BITS, t.litlen, LITLEN_CAP, s.rev, s.longest) &&
           build_table(lens + 288, 32, OFFSET, OFFSET_BITS, t.offset, OFFSET_CAP, s.rev, s.longest);
}


// The whole stream; returns the bytes written or a negative number.
MDG_HD inline int64_t inflate_stream(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap, InflateScratch &s)
{
    BitReader br;
    br.in = in;
    br.in_end = in + in_len;
    uint8_t *at = out, *const out_end = out + out_cap;
    Tables &t = s.t;
    while (true) {
        br.refill();
        const uint32_t final_block = br.take(1), type = br.take(2);
        if (br.overrun) return -5;
        if (type == 0) {
            // stored: back to a byte boundary, bytes the bit buffer holds but has not used go back to the input
            br.consume(br.cnt & 7);
            br.in -= br.cnt >> 3;
            br.bits = 0;
            br.cnt = 0;
            if (br.in_end - br.in < 4) return -5;
            const uint32_t len = br.in[0] | (uint32_t)br.in[1] << 8, nlen = br.in[2] | (uint32_t)br.in[3] << 8;
            br.in += 4;
            if ((len ^ nlen) != 0xFFFFu || (int64_t)len > br.in_end - br.in || (int64_t)len > out_end - at) return -5;
            for (uint32_t i = 0; i < len; ++i) at[i] = br.in[i];
            at += len;
            br.in += len;
        } else if (type == 1 || type == 2) {
            if (!(type == 1 ? fixed_tables(t, s) : read_dynamic_header(br, t, s))) return -5;
            if (!inflate_block(br, t, out, at, out_end)) return -5;
        } else {
            return -5;
        }
        if (final_block) break;
    }
    return at - out;
}

}  // namespace mdg_inflate
