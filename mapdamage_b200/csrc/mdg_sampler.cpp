// Read down-sampling on the host: the draws of the reference's reader (reader.py:134-164) come from CPython's
// random.Random -- MT19937 (Matsumoto & Nishimura 1998).  The caller hands over the generator state it got from
// random.Random(seed).getstate() (624 words + position), so seeding stays CPython's own; this file restates only the
// published recurrence and the two ways CPython turns words into numbers:
//   random()      = ((a >> 5) * 2^26 + (b >> 6)) / 2^53   from two successive words a, b
//   randint(0, i) = _randbelow(i + 1): k = bit_length(i + 1); draw getrandbits(k) until it is <= i, where
//                   getrandbits(k <= 32) = word >> (32 - k) and wider values take further words, low word first.
#include <cstdint>
#include <cstring>

#include "../../include/mapdamage_b200.h"

namespace {

struct Twister {
    uint32_t *mt;  // caller's 625 words: state[624], position
    uint32_t next() {
        uint32_t pos = mt[624];
        if (pos >= 624) {
            refill();
            pos = 0;
        }
        uint32_t y = mt[pos];
        mt[624] = pos + 1;
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    void refill() {
        const int n = 624, m = 397;
        for (int k = 0; k < n; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % n] & 0x7fffffffu);
            mt[k] = mt[(k + m) % n] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
    }
    double real53() {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    uint64_t bits(int k) {  // 1 <= k <= 64
        if (k <= 32) return next() >> (32 - k);
        const uint64_t low = next();
        return low | ((uint64_t)(next() >> (64 - k)) << 32);
    }
    uint64_t below(uint64_t n) {  // n >= 1
        int k = 64 - __builtin_clzll(n);
        uint64_t r = bits(k);
        while (r >= n) r = bits(k);
        return r;
    }
};

}  // namespace

extern "C" {

int mdg_sample_fraction(uint32_t *mt_state, double fraction, int64_t n, uint8_t *keep) {
    if (!mt_state || (n > 0 && !keep) || n < 0 || !(fraction >= 0.0 && fraction < 1.0) || mt_state[624] > 624) return MDG_ERR_ARGUMENT;
    Twister rng{mt_state};
    for (int64_t i = 0; i < n; ++i) keep[i] = rng.real53() < fraction;
    return MDG_OK;
}

int mdg_sample_reservoir(uint32_t *mt_state, int64_t first_index, int64_t n, int64_t n_slots, int64_t *slots) {
    if (!mt_state || !slots || n < 0 || first_index < 0 || n_slots < 1 || mt_state[624] > 624) return MDG_ERR_ARGUMENT;
    Twister rng{mt_state};
    for (int64_t index = first_index; index < first_index + n; ++index) {
        int64_t slot = index;
        if (index >= n_slots) {
            slot = (int64_t)rng.below((uint64_t)index + 1);
            if (slot >= n_slots) continue;
        }
        slots[slot] = index;
    }
    return MDG_OK;
}

}  // extern "C"
