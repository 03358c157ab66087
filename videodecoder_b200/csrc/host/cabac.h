// cabac.h -- CABAC arithmetic coding engines (ITU-T H.264 9.3.1.2, 9.3.3.2, 9.3.4) and the context set of a slice.
//
// Decoder: replaces the reference's cabac::DecodeValueUsingCtxIdx / RenormD / init_variable / init_engine
// (h264/cabac.cpp:81-171, 172-233), which renormalise one bit at a time through Parser::read_bi and look contexts up in
// a heap array2d.  Here codIOffset lives in the upper part of a 64-bit window that is refilled four bytes at a time;
// contexts are one byte each (pStateIdx << 1 | valMPS).
// Encoder: the inverse (9.3.4.2 - 9.3.4.5), used by the stream writer (test / benchmark tooling).
//
// Both engines expose the same three calls -- decision(ctx, bin), bypass(bin), terminate(bin) -- each RETURNING the bin:
// the decoder ignores its argument, the encoder returns it.  The syntax walker of mb_codec.h is written once against
// this interface and instantiated for both directions.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>

#include "cabac_tables.h"

// The decoder's decision() is called from some sixty places of the syntax walker; left to its heuristics the compiler keeps it a
// function with a call per bin.  (Measured on top of this and NOT kept: the engine's registers as a local copy inside the
// residual-block and vector-difference loops -- no change; a branch-free decision with the MPS / LPS outcome as a mask -- 4 %
// slower, twice: the chain range -> LPS range -> comparison -> renormalisation is what a bin costs, not its branches.)
#if defined(__GNUC__)
#define H264HOST_INLINE inline __attribute__((always_inline))
#define H264HOST_NOINLINE __attribute__((noinline))
#else
#define H264HOST_INLINE inline
#define H264HOST_NOINLINE
#endif

namespace h264host {

struct CabacContexts {
    // pStateIdx << 1 | valMPS.  Sixteen bits wide on purpose: a store through a character type may alias anything, the engine's
    // range / offset / bit count included
    uint16_t s[kNumCtx];
    // 9.3.1.1: preCtxState = Clip3(1, 126, ((m * Clip3(0, 51, SliceQPY)) >> 4) + n)
    void init(int slice_qp, bool i_slice, int cabac_init_idc)
    {
        const int8_t (*t)[2] = kCtxInit[i_slice ? 0 : 1 + cabac_init_idc];
        const int qp = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
        for (int i = 0; i < kNumCtx; i++) {
            int pre = ((t[i][0] * qp) >> 4) + t[i][1];
            pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
            s[i] = pre <= 63 ? (uint16_t)((63 - pre) << 1) : (uint16_t)(((pre - 64) << 1) | 1);
        }
    }
};

inline uint16_t cabac_next_state(uint16_t s, bool was_mps)
{
    const int st = s >> 1, mps = s & 1;
    if (was_mps) return (uint16_t)(((st < 62 ? st + 1 : st) << 1) | mps);
    return (uint16_t)((kTransLps[st] << 1) | (st == 0 ? 1 - mps : mps));
}
// ... as a table for the decoder: t[0] after an MPS, t[1] after an LPS
struct CabacNextTable {
    uint16_t t[2][128];
    CabacNextTable()
    {
        for (int s = 0; s < 128; s++) { t[0][s] = cabac_next_state((uint16_t)s, true); t[1][s] = cabac_next_state((uint16_t)s, false); }
    }
};
static const CabacNextTable kCabacNext;

class CabacDecoder {
public:
    static constexpr bool kEncoding = false;
    CabacContexts ctx;
    // data: RBSP bytes starting at the first byte of slice_data's arithmetic code (after cabac_alignment_one_bit)
    void start(const uint8_t* p, size_t n)
    {
        p_ = p; end_ = p + n;
        range_ = 510;
        value_ = 0; bits_ = 0;
        refill();
        // codIOffset = read_bits(9): the window holds codIOffset << bits_
        bits_ -= 9;
        overrun_ = false;
    }
    H264HOST_INLINE int decision(int c, int /*bin*/ = 0)
    {
        uint16_t& s = ctx.s[c];
        const uint32_t lps = kRangeLps[s >> 1][(range_ >> 6) & 3];
        range_ -= lps;
        const uint64_t scaled = (uint64_t)range_ << bits_;
        int bin;
        if (value_ < scaled) {
            bin = s & 1;
            s = kCabacNext.t[0][s];
            if (range_ >= 256) return bin;
        } else {
            value_ -= scaled;
            range_ = lps;
            bin = (s & 1) ^ 1;
            s = kCabacNext.t[1][s];
        }
        renorm();
        return bin;
    }
    H264HOST_INLINE int bypass(int /*bin*/ = 0)
    {
        if (bits_ < 1) refill();
        bits_ -= 1;
        const uint64_t scaled = (uint64_t)range_ << bits_;
        if (value_ >= scaled) { value_ -= scaled; return 1; }
        return 0;
    }
    int terminate(int /*bin*/ = 0)
    {
        range_ -= 2;
        const uint64_t scaled = (uint64_t)range_ << bits_;
        if (value_ >= scaled) return 1;
        renorm();
        return 0;
    }
    bool overrun() const { return overrun_; }
    // after terminate() returned 1 for I_PCM: not supported here (no pcm alignment restart)

private:
    H264HOST_INLINE void renorm()
    {
        if (range_ >= 256) return;
        const int n = __builtin_clz(range_) - 23;      // range_ < 256: shift that brings bit 8 up
        range_ <<= n;
        bits_ -= n;
        if (bits_ < 0) refill();
    }
    H264HOST_NOINLINE void refill()
    {
        // append 32 bits below the window
        uint32_t w = 0;
        for (int i = 0; i < 4; i++) {
            w <<= 8;
            if (p_ < end_) w |= *p_++; else { pad_++; if (pad_ > 8) overrun_ = true; }
        }
        value_ = (value_ << 32) | w;
        bits_ += 32;
    }
    const uint8_t* p_ = nullptr;
    const uint8_t* end_ = nullptr;
    uint32_t range_ = 510;
    uint64_t value_ = 0;     // codIOffset << bits_ | the next bits_ bits of the stream
    int bits_ = 0;
    int pad_ = 0;
    bool overrun_ = false;
};

class CabacEncoder {
public:
    static constexpr bool kEncoding = true;
    CabacContexts ctx;
    std::vector<uint8_t> out;     // arithmetic code bytes (byte aligned at start)
    void start()
    {
        low_ = 0; range_ = 510; first_ = true; outstanding_ = 0;
        acc_ = 0; nbits_ = 0;
        out.clear();
    }
    int decision(int c, int bin)
    {
        uint16_t& s = ctx.s[c];
        const uint32_t lps = kRangeLps[s >> 1][(range_ >> 6) & 3];
        range_ -= lps;
        if (bin != (s & 1)) {
            low_ += range_;
            range_ = lps;
            s = cabac_next_state(s, false);
        } else s = cabac_next_state(s, true);
        renorm();
        return bin;
    }
    int bypass(int bin)
    {
        low_ <<= 1;
        if (bin) low_ += range_;
        if (low_ >= 1024) { put_bit(1); low_ -= 1024; }
        else if (low_ < 512) put_bit(0);
        else { low_ -= 512; outstanding_++; }
        return bin;
    }
    int terminate(int bin)
    {
        range_ -= 2;
        if (bin) {
            low_ += range_;
            // EncodeFlush (9.3.4.5); the last bit written is the rbsp_stop_one_bit
            range_ = 2;
            renorm();
            put_bit((low_ >> 9) & 1);
            write_bits(((low_ >> 7) & 3) | 1, 2);
            while (nbits_) write_bits(0, 1);      // rbsp_alignment_zero_bit
        } else renorm();
        return bin;
    }

private:
    void renorm()
    {
        while (range_ < 256) {
            if (low_ < 256) put_bit(0);
            else if (low_ >= 512) { low_ -= 512; put_bit(1); }
            else { low_ -= 256; outstanding_++; }
            range_ <<= 1;
            low_ <<= 1;
        }
    }
    void put_bit(uint32_t b)
    {
        if (first_) first_ = false; else write_bits(b, 1);
        while (outstanding_ > 0) { write_bits(1 - b, 1); outstanding_--; }
    }
    void write_bits(uint32_t v, int n)
    {
        for (int i = n - 1; i >= 0; i--) {
            acc_ = (acc_ << 1) | ((v >> i) & 1u);
            if (++nbits_ == 8) { out.push_back((uint8_t)acc_); acc_ = 0; nbits_ = 0; }
        }
    }
    uint32_t low_ = 0, range_ = 510;
    bool first_ = true;
    int outstanding_ = 0;
    uint32_t acc_ = 0;
    int nbits_ = 0;
};

}  // namespace h264host
