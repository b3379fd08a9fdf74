// Host bit-string helpers, determinant dictionary and SpookyHash digest for pyci_b200._pyci.
// Semantics follow /root/reference/pyci/src/common.cpp (binomial :41-66, fill_* :68-113, colex
// :114-162, popcnt/ctz :263-282); the code is our own.
#include <algorithm>
#include <cstring>

#include "pyci_host.h"

namespace pyci_host {

static long g_num_threads = 1;

long get_num_threads() { return g_num_threads; }

void set_num_threads(long n) { g_num_threads = std::max(n, 1L); }

// saturating C(n, k): LONG_MAX on overflow (common.cpp:41-66)
long binomial(long n, long k) {
    if (k < 0 || k > n)
        return (k == 0) ? 1 : 0;
    if (k > n - k)
        k = n - k;
    __int128 b = 1;
    const __int128 lim = std::numeric_limits<long>::max();
    for (long d = 1; d <= k; ++d) {
        b = b * (n - k + d) / d;
        if (b >= lim)
            return std::numeric_limits<long>::max();
    }
    return (long)b;
}

long nword_det(long nbasis) { return nbasis / 64 + ((nbasis % 64) ? 1 : 0); }

void fill_det(long nocc, const long *occs, ulong *det) {
    for (long i = 0; i < nocc; ++i)
        det[occs[i] / 64] |= 1UL << (occs[i] % 64);
}

void fill_hartreefock_det(long nocc, ulong *det) {
    long w = 0;
    for (; nocc >= 64; nocc -= 64)
        det[w++] = ~0UL;
    if (nocc)
        det[w] = (1UL << nocc) - 1;
}

void fill_occs(long nword, const ulong *det, long *occs) {
    long j = 0;
    for (long w = 0; w < nword; ++w)
        for (ulong word = det[w]; word; word &= word - 1)
            occs[j++] = __builtin_ctzl(word) + 64 * w;
}

void fill_virs(long nword, long nbasis, const ulong *det, long *virs) {
    long j = 0;
    for (long w = 0; w < nword; ++w) {
        const long left = nbasis - 64 * w;
        const ulong mask = (left >= 64) ? ~0UL : (left > 0 ? ((1UL << left) - 1) : 0UL);
        for (ulong word = det[w] ^ mask; word; word &= word - 1)
            virs[j++] = __builtin_ctzl(word) + 64 * w;
    }
}

void next_colex(long *idx) {
    long i = 0;
    while (idx[i + 1] - idx[i] == 1) {
        idx[i] = i;
        ++i;
    }
    ++idx[i];
}

void unrank_colex(long nbasis, long nocc, long rank, long *occs) {
    // the rank-th nocc-subset of {0..nbasis-1} in colexicographic order
    long n = nbasis;
    for (long j = nocc; j >= 1; --j) {
        long b = binomial(n, j);
        if (b <= rank) { // only when rank is out of range; mirror the reference's early exit
            for (long k = 0; k < j; ++k)
                occs[k] = k;
            return;
        }
        while (b > rank)
            b = binomial(--n, j);
        occs[j - 1] = n;
        rank -= b;
    }
}

long popcnt_det(long nword, const ulong *det) {
    long c = 0;
    for (long i = 0; i < nword; ++i)
        c += __builtin_popcountl(det[i]);
    return c;
}

long ctz_det(long nword, const ulong *det) {
    for (long i = 0; i < nword; ++i)
        if (det[i])
            return __builtin_ctzl(det[i]) + 64 * i;
    return 0;
}

void excite_det(long i, long a, ulong *det) {
    det[i / 64] &= ~(1UL << (i % 64));
    det[a / 64] |= 1UL << (a % 64);
}

// ---- SpookyHash V2, short-message form (messages under 192 bytes) ---------------------------------
namespace {

inline uint64_t rot(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

inline void short_mix(uint64_t &h0, uint64_t &h1, uint64_t &h2, uint64_t &h3) {
    h2 = rot(h2, 50); h2 += h3; h0 ^= h2;
    h3 = rot(h3, 52); h3 += h0; h1 ^= h3;
    h0 = rot(h0, 30); h0 += h1; h2 ^= h0;
    h1 = rot(h1, 41); h1 += h2; h3 ^= h1;
    h2 = rot(h2, 54); h2 += h3; h0 ^= h2;
    h3 = rot(h3, 48); h3 += h0; h1 ^= h3;
    h0 = rot(h0, 38); h0 += h1; h2 ^= h0;
    h1 = rot(h1, 37); h1 += h2; h3 ^= h1;
    h2 = rot(h2, 62); h2 += h3; h0 ^= h2;
    h3 = rot(h3, 34); h3 += h0; h1 ^= h3;
    h0 = rot(h0, 5);  h0 += h1; h2 ^= h0;
    h1 = rot(h1, 36); h1 += h2; h3 ^= h1;
}

inline void short_end(uint64_t &h0, uint64_t &h1, uint64_t &h2, uint64_t &h3) {
    h3 ^= h2; h2 = rot(h2, 15); h3 += h2;
    h0 ^= h3; h3 = rot(h3, 52); h0 += h3;
    h1 ^= h0; h0 = rot(h0, 26); h1 += h0;
    h2 ^= h1; h1 = rot(h1, 51); h2 += h1;
    h3 ^= h2; h2 = rot(h2, 28); h3 += h2;
    h0 ^= h3; h3 = rot(h3, 9);  h0 += h3;
    h1 ^= h0; h0 = rot(h0, 47); h1 += h0;
    h2 ^= h1; h1 = rot(h1, 54); h2 += h1;
    h3 ^= h2; h2 = rot(h2, 32); h3 += h2;
    h0 ^= h3; h3 = rot(h3, 25); h0 += h3;
    h1 ^= h0; h0 = rot(h0, 63); h1 += h0;
}

} // namespace

Hash spooky_rank(const ulong *words, long nwords) {
    const uint64_t sc = 0xdeadbeefdeadbeefULL;
    const size_t length = (size_t)nwords * 8;
    if (length >= 192)
        throw std::runtime_error("rank_det: determinants of 24 or more words are not supported");
    const uint64_t *p = reinterpret_cast<const uint64_t *>(words);
    uint64_t a = 0x23a23cf5033c3c81ULL, b = 0xb3816f6a2c68e530ULL, c = sc, d = sc;
    size_t remainder = length % 32;
    if (length > 15) {
        const uint64_t *end = p + (length / 32) * 4;
        for (; p < end; p += 4) {
            c += p[0];
            d += p[1];
            short_mix(a, b, c, d);
            a += p[2];
            b += p[3];
        }
        if (remainder >= 16) {
            c += p[0];
            d += p[1];
            short_mix(a, b, c, d);
            p += 2;
            remainder -= 16;
        }
    }
    d += ((uint64_t)length) << 56;
    // the message is a whole number of 8-byte words, so the remainder is 0 or 8
    if (remainder == 8) {
        c += p[0];
    } else {
        c += sc;
        d += sc;
    }
    short_end(a, b, c, d);
    return Hash(a, b);
}

// ---- determinant dictionary --------------------------------------------------------------------------

uint64_t DetTable::mix(const ulong *k, long nw) {
    uint64_t h = 0x9e3779b97f4a7c15ULL;
    for (long i = 0; i < nw; ++i) {
        h ^= k[i];
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdULL;
        h ^= h >> 33;
        h *= 0xc4ceb9fe1a85ec53ULL;
        h ^= h >> 33;
    }
    return h;
}

void DetTable::grow(const std::vector<ulong> &dets, long newcap) {
    std::vector<long> old;
    old.swap(slots_);
    slots_.assign((size_t)newcap, -1);
    const uint64_t mask = (uint64_t)newcap - 1;
    for (long s : old) {
        if (s < 0)
            continue;
        uint64_t p = mix(&dets[s * nw_], nw_) & mask;
        while (slots_[p] >= 0)
            p = (p + 1) & mask;
        slots_[p] = s;
    }
}

void DetTable::reserve(long n) {
    long cap = 16;
    while (cap < 2 * n)
        cap <<= 1;
    if (cap > (long)slots_.size()) {
        static const std::vector<ulong> none;
        if (count_ == 0)
            slots_.assign((size_t)cap, -1);
        // with live entries the caller's next assign() grows using the real det array
    }
}

long DetTable::find(const std::vector<ulong> &dets, const ulong *key) const {
    if (slots_.empty())
        return -1;
    const uint64_t mask = slots_.size() - 1;
    uint64_t p = mix(key, nw_) & mask;
    for (;;) {
        const long s = slots_[p];
        if (s < 0)
            return -1;
        if (std::memcmp(&dets[s * nw_], key, sizeof(ulong) * nw_) == 0)
            return s;
        p = (p + 1) & mask;
    }
}

bool DetTable::assign(const std::vector<ulong> &dets, const ulong *key, long idx) {
    if ((count_ + 1) * 2 > (long)slots_.size())
        grow(dets, std::max<long>(16, (long)slots_.size() * 2));
    const uint64_t mask = slots_.size() - 1;
    uint64_t p = mix(key, nw_) & mask;
    for (;;) {
        const long s = slots_[p];
        if (s < 0) {
            slots_[p] = idx;
            ++count_;
            return true;
        }
        if (std::memcmp(&dets[s * nw_], key, sizeof(ulong) * nw_) == 0) {
            slots_[p] = idx;
            return false;
        }
        p = (p + 1) & mask;
    }
}

void DetTable::assign_bulk(const std::vector<ulong> &dets, long first, long n) {
    if (n <= 0)
        return;
    long cap = std::max<long>(16, (long)slots_.size());
    while (cap < 2 * (count_ + n))
        cap <<= 1;
    if (cap != (long)slots_.size())
        grow(dets, cap);
    const uint64_t mask = (uint64_t)cap - 1;
    constexpr long B = 16;
    uint64_t home[B];
    for (long base = 0; base < n; base += B) {
        const long m = std::min(B, n - base);
        for (long j = 0; j < m; ++j) {
            home[j] = mix(&dets[(first + base + j) * nw_], nw_) & mask;
            __builtin_prefetch(&slots_[home[j]], 1);
        }
        for (long j = 0; j < m; ++j) {
            const long idx = first + base + j;
            const ulong *key = &dets[idx * nw_];
            uint64_t p = home[j];
            for (;;) {
                const long s = slots_[p];
                if (s < 0) {
                    slots_[p] = idx;
                    ++count_;
                    break;
                }
                if (std::memcmp(&dets[s * nw_], key, sizeof(ulong) * nw_) == 0) {
                    slots_[p] = idx;
                    break;
                }
                p = (p + 1) & mask;
            }
        }
    }
}

void DetTable::insert_new_bulk(const std::vector<ulong> &dets, long first, long n) {
    if (n <= 0)
        return;
    long cap = std::max<long>(16, (long)slots_.size());
    while (cap < 2 * (count_ + n))
        cap <<= 1;
    if (cap != (long)slots_.size())
        grow(dets, cap); // once, to the final capacity (grow re-inserts the live entries only)
    const uint64_t mask = (uint64_t)cap - 1;
    constexpr long B = 16;
    uint64_t home[B];
    for (long base = 0; base < n; base += B) {
        const long m = std::min(B, n - base);
        for (long j = 0; j < m; ++j) {
            home[j] = mix(&dets[(first + base + j) * nw_], nw_) & mask;
            __builtin_prefetch(&slots_[home[j]], 1);
        }
        for (long j = 0; j < m; ++j) {
            uint64_t p = home[j];
            while (slots_[p] >= 0)
                p = (p + 1) & mask;
            slots_[p] = first + base + j;
        }
    }
    count_ += n;
}

} // namespace pyci_host
