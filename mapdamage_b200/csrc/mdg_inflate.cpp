// Raw DEFLATE (RFC 1951) decoder for BGZF blocks.
//
// The BAM path spends its host time inflating (zlib: about 140 MB/s per thread on base / quality data, which is
// mostly literals), so the blocks are decoded here instead: a 64-bit bit buffer refilled without branches, an
// 11-bit first-level table for literal / length codes (8-bit for distances) with second-level tables behind it,
// up to three literals per refill, word-wide match copies.  Written from the RFC; the caller (mdg_bamio.cpp)
// checks every block's CRC32 and hands a block this decoder rejects to zlib, so a stream this code cannot
// handle costs time, never correctness.
#include <cstdint>
#include <cstring>

#include "../../include/mapdamage_b200.h"

#include "mdg_inflate_core.h"

extern "C" int64_t mdg_inflate_raw(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_cap)
{
    if (!in || in_len < 0 || (!out && out_cap > 0) || out_cap < 0) return MDG_ERR_ARGUMENT;
    mdg_inflate::InflateScratch scratch;
    const int64_t n = mdg_inflate::inflate_stream(in, in_len, out, out_cap, scratch);
    return n < 0 ? MDG_ERR_DATA : n;
}
