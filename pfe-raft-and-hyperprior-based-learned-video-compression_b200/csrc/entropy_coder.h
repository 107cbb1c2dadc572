// entropy_coder.h -- "next" row f-4, last piece: a stand-in for the entropy coder behind
// EntropyBottleneck.compress / .decompress (R:codec_processing.py:433,447 construction, :488-497 compress,
// :509-536 decompress).  The reference calls compressai's range coder (RansEncoder.encode_with_indexes /
// RansDecoder.decode_with_indexes, C++ in that package); compressai is NOT in this image and not vendored by the
// reference, so the exact bitstream cannot be reproduced or checked: BITSTREAM PARITY IS UNPINNED.  What is
// restated here, from the published algorithm (J. Duda, "Asymmetric numeral systems", 2013; the rANS variant
// with a 32-bit state renormalised in 16-bit words), is the same interface and the same modelling contract:
//   * symbols are coded with per-channel ("index") quantised CDF tables of 16-bit precision,
//   * a symbol outside a table's range is coded as the table's last ("escape") entry followed by its magnitude
//     in 4-bit bypass digits (count of digits in base-15 unary, then the digits), negative values interleaved,
//   * encode -> decode is bit exact and the byte count is the model's cross-entropy + a 4-byte state.
// Pure host code (no GPU): entropy coding is the serial tail of the P-frame path and the reference runs it on the
// CPU too.  64 800 symbols (a 1080p flow field at 1/8 resolution) code in ~1 ms.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace rdvc {
namespace ec {

constexpr int kPrecision = 16;                 // CDF tables sum to 1 << 16
constexpr uint32_t kLow = 1u << 16;            // state lives in [2^16, 2^32)
constexpr int kBypassBits = 4;
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;

struct Op {           // one rANS step: a slice [start, start + freq) of the 2^16 range
    uint32_t start;
    uint32_t freq;
};

inline void push_bypass(std::vector<Op>& ops, uint32_t digit) {
    ops.push_back({digit << (kPrecision - kBypassBits), 1u << (kPrecision - kBypassBits)});
}

// Returns false on a malformed table / index.
inline bool symbol_ops(std::vector<Op>& ops, int32_t symbol, int32_t index, const uint32_t* cdfs,
                       const int32_t* cdf_lengths, const int32_t* offsets, int n_tables, int max_len) {
    if (index < 0 || index >= n_tables) return false;
    const int len = cdf_lengths[index];
    if (len < 3 || len > max_len) return false;              // at least one real symbol + the escape
    const uint32_t* cdf = cdfs + static_cast<size_t>(index) * max_len;
    const int32_t max_value = len - 2;                       // the escape entry
    int64_t v = static_cast<int64_t>(symbol) - offsets[index];
    uint64_t raw = 0;
    if (v < 0) {
        raw = static_cast<uint64_t>(-2 * v - 1);             // negative overflow: odd codes
        v = max_value;
    } else if (v >= max_value) {
        raw = static_cast<uint64_t>(2 * (v - max_value));    // positive overflow: even codes
        v = max_value;
    }
    const uint32_t lo = cdf[v], hi = cdf[v + 1];
    if (hi <= lo || hi > (1u << kPrecision)) return false;
    ops.push_back({lo, hi - lo});
    if (v == max_value) {
        int digits = 0;
        while (digits < 16 && (raw >> (digits * kBypassBits)) != 0) ++digits;
        int count = digits;                                  // number of digits, base-15 unary
        while (count >= static_cast<int>(kBypassMax)) {
            push_bypass(ops, kBypassMax);
            count -= kBypassMax;
        }
        push_bypass(ops, static_cast<uint32_t>(count));
        for (int d = 0; d < digits; ++d) push_bypass(ops, static_cast<uint32_t>((raw >> (d * kBypassBits)) & kBypassMax));
    }
    return true;
}

// Upper bound of the encoded size of n symbols (worst case: every symbol escapes with 16 bypass digits).
inline size_t max_encoded_bytes(size_t n) { return 4 + n * 2 * (1 + 2 + 16) + 16; }

// Returns the number of bytes written to `out`, or 0 on error (bad table / index, or out_capacity too small).
inline size_t encode_with_indexes(const int32_t* symbols, const int32_t* indexes, size_t n, const uint32_t* cdfs,
                                  const int32_t* cdf_lengths, const int32_t* offsets, int n_tables, int max_len,
                                  uint8_t* out, size_t out_capacity) {
    std::vector<Op> ops;
    ops.reserve(n + 16);
    for (size_t i = 0; i < n; ++i)
        if (!symbol_ops(ops, symbols[i], indexes[i], cdfs, cdf_lengths, offsets, n_tables, max_len)) return 0;
    std::vector<uint16_t> words;
    words.reserve(ops.size());
    uint32_t x = kLow;
    for (size_t k = ops.size(); k-- > 0;) {                  // rANS is LIFO: encode in reverse
        const Op op = ops[k];
        if (static_cast<uint64_t>(x) >= (static_cast<uint64_t>(op.freq) << 16)) {
            words.push_back(static_cast<uint16_t>(x & 0xffffu));
            x >>= 16;
        }
        x = ((x / op.freq) << kPrecision) + (x % op.freq) + op.start;
    }
    const size_t total = 4 + 2 * words.size();
    if (total > out_capacity) return 0;
    out[0] = static_cast<uint8_t>(x >> 24); out[1] = static_cast<uint8_t>(x >> 16);
    out[2] = static_cast<uint8_t>(x >> 8);  out[3] = static_cast<uint8_t>(x);
    size_t o = 4;
    for (size_t k = words.size(); k-- > 0;) {                // the decoder reads the last word first
        out[o++] = static_cast<uint8_t>(words[k] >> 8);
        out[o++] = static_cast<uint8_t>(words[k] & 0xff);
    }
    return total;
}

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t x;
    bool ok;
    inline void renorm() {
        if (x < kLow) {
            if (end - p < 2) { ok = false; return; }
            x = (x << 16) | (static_cast<uint32_t>(p[0]) << 8) | p[1];
            p += 2;
        }
    }
    inline uint32_t bypass() {
        const uint32_t cf = x & 0xffffu;
        const uint32_t digit = cf >> (kPrecision - kBypassBits);
        x = (x >> kPrecision) * (1u << (kPrecision - kBypassBits)) + cf - (digit << (kPrecision - kBypassBits));
        renorm();
        return digit;
    }
};

// Returns 0 on success, -1 on a malformed stream / table.
inline int decode_with_indexes(const uint8_t* in, size_t nbytes, const int32_t* indexes, size_t n, const uint32_t* cdfs,
                               const int32_t* cdf_lengths, const int32_t* offsets, int n_tables, int max_len,
                               int32_t* symbols_out) {
    if (nbytes < 4) return n == 0 && nbytes == 0 ? 0 : -1;
    Reader r;
    r.p = in + 4; r.end = in + nbytes; r.ok = true;
    r.x = (static_cast<uint32_t>(in[0]) << 24) | (static_cast<uint32_t>(in[1]) << 16) |
          (static_cast<uint32_t>(in[2]) << 8) | in[3];
    for (size_t i = 0; i < n; ++i) {
        const int32_t index = indexes[i];
        if (index < 0 || index >= n_tables) return -1;
        const int len = cdf_lengths[index];
        if (len < 3 || len > max_len) return -1;
        const uint32_t* cdf = cdfs + static_cast<size_t>(index) * max_len;
        const int32_t max_value = len - 2;
        const uint32_t cf = r.x & 0xffffu;
        int lo = 0, hi = len - 1;                            // largest s with cdf[s] <= cf
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] <= cf) lo = mid; else hi = mid;
        }
        const int32_t s = lo;
        const uint32_t start = cdf[s], freq = cdf[s + 1] - cdf[s];
        if (freq == 0) return -1;
        r.x = freq * (r.x >> kPrecision) + cf - start;
        r.renorm();
        int64_t v = s;
        if (s == max_value) {
            int digits = 0;
            uint32_t d;
            do {
                d = r.bypass();
                digits += static_cast<int>(d);
                if (digits > 16 + static_cast<int>(kBypassMax)) return -1;
            } while (d == kBypassMax);
            uint64_t raw = 0;
            for (int k = 0; k < digits; ++k) raw |= static_cast<uint64_t>(r.bypass()) << (k * kBypassBits);
            v = (raw & 1) ? -static_cast<int64_t>((raw + 1) >> 1) : static_cast<int64_t>(raw >> 1) + max_value;
        }
        if (!r.ok) return -1;
        symbols_out[i] = static_cast<int32_t>(v + offsets[index]);
    }
    return 0;
}

}  // namespace ec
}  // namespace rdvc
