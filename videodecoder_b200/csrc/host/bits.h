// bits.h -- bit-level I/O of the host entropy path: RBSP reader / writer with Exp-Golomb codes, NAL escaping
// (emulation-prevention bytes) and Annex-B start-code scanning.
//
// Replaces, on the reference's side, Bitsbuf (h264/bitsbuf.cpp:45-278: one stdio read per buffer refill, bits handed
// out one at a time through read_bi, emulation-prevention bytes dropped while refilling, h264/bitsbuf.cpp:124-170) and
// the start-code search of Parser::find_nextNAL (h264/Parser.cpp:21-53).  Here a NAL unit is unescaped once into a flat
// RBSP buffer and read through a 64-bit window.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace h264host {

class BitReader {
public:
    BitReader() {}
    BitReader(const uint8_t* p, size_t n) { reset(p, n); }
    void reset(const uint8_t* p, size_t n) { p_ = p; n_ = n; pos_ = 0; }
    size_t bit_pos() const { return pos_; }
    size_t bits_left() const { return n_ * 8 > pos_ ? n_ * 8 - pos_ : 0; }
    bool overrun() const { return pos_ > n_ * 8; }
    uint32_t u(int n)          // n <= 32
    {
        uint32_t v = 0;
        while (n > 0) {
            const size_t byte = pos_ >> 3;
            const int avail = 8 - (int)(pos_ & 7), take = n < avail ? n : avail;
            const uint32_t b = byte < n_ ? p_[byte] : 0u;
            v = (v << take) | ((b >> (avail - take)) & ((1u << take) - 1u));
            pos_ += (size_t)take;
            n -= take;
        }
        return v;
    }
    uint32_t u1() { return u(1); }
    uint32_t ue()
    {
        int zeros = 0;
        while (zeros < 32 && !overrun() && u1() == 0) zeros++;
        if (zeros >= 32) return 0xFFFFFFFFu;
        return zeros ? ((1u << zeros) - 1u + u(zeros)) : 0u;
    }
    int32_t se()
    {
        const uint32_t k = ue();
        return (k & 1u) ? (int32_t)((k + 1u) >> 1) : -(int32_t)(k >> 1);
    }
    void align() { pos_ = (pos_ + 7) & ~(size_t)7; }
    // more_rbsp_data(): something other than the trailing bits is left
    bool more_rbsp_data() const
    {
        if (pos_ >= n_ * 8) return false;
        size_t last = n_;
        while (last > 0 && p_[last - 1] == 0) last--;       // cabac_zero_words / trailing zero bytes
        if (last == 0) return false;
        int stop = 0;
        while (!((p_[last - 1] >> stop) & 1)) stop++;
        const size_t stop_pos = (last - 1) * 8 + (size_t)(7 - stop);
        return pos_ < stop_pos;
    }
    const uint8_t* byte_ptr() const { return p_ + (pos_ >> 3); }
    size_t bytes_left() const { return n_ - (pos_ >> 3); }

private:
    const uint8_t* p_ = nullptr;
    size_t n_ = 0, pos_ = 0;
};

class BitWriter {
public:
    std::vector<uint8_t> bytes;
    void put(uint32_t v, int n)
    {
        for (int i = n - 1; i >= 0; i--) put1((v >> i) & 1u);
    }
    void put1(uint32_t b)
    {
        acc_ = (acc_ << 1) | (b & 1u);
        if (++nbits_ == 8) { bytes.push_back((uint8_t)acc_); acc_ = 0; nbits_ = 0; }
    }
    void ue(uint32_t v)
    {
        const uint32_t x = v + 1;
        int len = 0;
        while ((x >> len) > 1) len++;
        put(0, len);
        put(x, len + 1);
    }
    void se(int32_t v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * (int64_t)v)); }
    void trailing() { put1(1); while (nbits_) put1(0); }
    void align_ones() { while (nbits_) put1(1); }
    void align_zeros() { while (nbits_) put1(0); }
    bool aligned() const { return nbits_ == 0; }

private:
    uint32_t acc_ = 0;
    int nbits_ = 0;
};

// NAL unit payload (after the header byte) -> RBSP: drops the 0x03 of every 00 00 03
inline void nal_unescape(const uint8_t* p, size_t n, std::vector<uint8_t>& out)
{
    out.clear();
    out.reserve(n);
    size_t i = 0;
    int zeros = 0;
    while (i < n) {
        const uint8_t b = p[i++];
        if (zeros >= 2 && b == 3) { zeros = 0; continue; }
        out.push_back(b);
        zeros = b == 0 ? zeros + 1 : 0;
    }
}

// start code + header byte + escaped RBSP, appended to out
inline void nal_emit(int nal_ref_idc, int nal_unit_type, const std::vector<uint8_t>& rbsp, std::vector<uint8_t>& out)
{
    out.push_back(0); out.push_back(0); out.push_back(0); out.push_back(1);
    out.push_back((uint8_t)((nal_ref_idc << 5) | nal_unit_type));
    int zeros = 0;
    for (uint8_t b : rbsp) {
        if (zeros >= 2 && b <= 3) { out.push_back(3); zeros = 0; }
        out.push_back(b);
        zeros = b == 0 ? zeros + 1 : 0;
    }
}

// Annex B: finds the next NAL unit at or after *pos.  Returns false at the end of the stream; else [*start, *start+*len)
// is the unit (header byte first, trailing zero bytes removed) and *pos is where to continue.
inline bool annexb_next(const uint8_t* p, size_t n, size_t* pos, size_t* start, size_t* len)
{
    size_t i = *pos;
    while (i + 3 <= n && !(p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 1)) i++;
    if (i + 3 > n) return false;
    const size_t s = i + 3;
    size_t j = s;
    while (j + 3 <= n && !(p[j] == 0 && p[j + 1] == 0 && (p[j + 2] == 1 || p[j + 2] == 0))) j++;
    size_t e = j + 3 <= n ? j : n;
    while (e > s && p[e - 1] == 0) e--;
    *start = s; *len = e - s; *pos = e;
    return e > s;
}

}  // namespace h264host
