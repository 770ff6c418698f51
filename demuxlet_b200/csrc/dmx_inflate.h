// dmx_inflate.h -- a raw-DEFLATE decoder for whole BGZF blocks (RFC 1951), used by the byte stream in dmx_bgzf.h.
//
// zlib inflates ~280 MB/s per thread on the build container and is 85 % of the CPU time of the BAM sweep; a BGZF block
// is the easy case for a faster decoder: input and output sizes are known up front (BSIZE, ISIZE), both buffers are
// whole, and a CRC-32 of the output follows -- so this decoder needs no streaming state, can read and write in 8-byte
// steps away from the buffer ends, and every block it produces is checked (length, end of stream, CRC-32) by the caller,
// which falls back to zlib on ANY failure.  Design: a 64-bit bit buffer refilled with one unaligned load per symbol
// (branch-free: the bytes that fit are counted, the rest of the word is OR-ed in again next time), two-level decode
// tables (11 bits for literal / length codes, 8 bits for distance codes, second level sized by the longest code) whose
// entries carry base value, extra-bit count and code length, and 8-byte match copies where source and destination are
// at least 8 bytes apart.
//
// Contract: `in` must be readable for in_n + 8 bytes (a BGZF block is followed by its 8-byte trailer); at most out_n
// bytes are written.  Returns true when the stream decodes to exactly out_n bytes and ends with a final block inside
// the in_n bytes.
#pragma once
#include <stdint.h>
#include <string.h>

namespace dmx_io {

class FastInflate {
 public:
  bool run(const uint8_t* in, size_t in_n, uint8_t* out, size_t out_n) {
    const uint8_t* const in0 = in;
    const uint8_t* const in_end = in + in_n;
    uint8_t* const out0 = out;
    uint8_t* const out_end = out + out_n;
    uint64_t bitbuf = 0;
    unsigned bitsleft = 0;
    // One unaligned 8-byte load; afterwards 56 <= bitsleft <= 63 and `in` stands behind the last byte counted into the
    // bit buffer (so the stream position is in * 8 - bitsleft).  Within 8 bytes of the end the load reaches into the
    // bytes behind in_end (see the contract); once `in` is behind in_end there is nothing left to load, and the same
    // bookkeeping goes on over whatever the buffer holds -- bits a valid stream never consumes; the final check tells.
    auto fill = [&]() -> bool {
      if (in <= in_end) {
        uint64_t w;
        memcpy(&w, in, 8);
        bitbuf |= w << bitsleft;
      } else if (in > in_end + 16) {
        return false;
      }
      in += (63 - bitsleft) >> 3;
      bitsleft |= 56;
      return true;
    };
    auto take = [&](unsigned n) -> uint32_t {
      const uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1));
      bitbuf >>= n;
      bitsleft -= n;
      return v;
    };
    for (bool final_block = false; !final_block;) {
      if (!fill()) return false;
      final_block = take(1) != 0;
      const uint32_t type = take(2);
      if (type == 0) {                           // stored: LEN, ~LEN on the next byte boundary, then the bytes
        take(bitsleft & 7);
        in -= bitsleft >> 3;                     // give back the whole bytes still in the bit buffer
        bitbuf = 0;
        bitsleft = 0;
        if (in + 4 > in_end) return false;
        const uint32_t len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
        in += 4;
        if ((len ^ nlen) != 0xffffu || in + len > in_end || out + len > out_end) return false;
        memcpy(out, in, len);
        in += len;
        out += len;
        continue;
      }
      if (type == 3) return false;
      const uint32_t *lt, *dt;
      if (type == 1) {
        if (!fixed_ready_) {
          uint8_t lens[288 + 32];
          for (int i = 0; i < 144; ++i) lens[i] = 8;
          for (int i = 144; i < 256; ++i) lens[i] = 9;
          for (int i = 256; i < 280; ++i) lens[i] = 7;
          for (int i = 280; i < 288; ++i) lens[i] = 8;
          for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
          if (!build(lens, 288, kLitBits, fixed_lit_, kLitCap, true) || !build(lens + 288, 32, kDistBits, fixed_dist_, kDistCap, false))
            return false;
          fixed_ready_ = true;
        }
        lt = fixed_lit_;
        dt = fixed_dist_;
      } else {
        // dynamic: code lengths of the code-length alphabet, then the two alphabets' lengths run-length coded with it
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const uint32_t hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t pre[19];
        memset(pre, 0, sizeof(pre));
        if (!fill()) return false;
        for (uint32_t i = 0; i < hclen; ++i) {
          pre[order[i]] = (uint8_t)take(3);
          if (bitsleft < 16 && !fill()) return false;
        }
        if (!build(pre, 19, kPreBits, pre_tab_, 1 << kPreBits, false)) return false;
        uint8_t lens[286 + 30 + 140];
        uint32_t n = 0;
        while (n < hlit + hdist) {
          if (!fill()) return false;
          const uint32_t e = pre_tab_[bitbuf & ((1u << kPreBits) - 1)];
          if ((e >> 5 & 7) == kInvalid) return false;
          take(e & 31);
          const uint32_t sym = e >> 16;
          if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
          uint32_t rep, val = 0;
          if (sym == 16) {
            if (n == 0) return false;
            val = lens[n - 1];
            rep = 3 + take(2);
          } else if (sym == 17) {
            rep = 3 + take(3);
          } else {
            rep = 11 + take(7);
          }
          if (n + rep > hlit + hdist) return false;
          memset(lens + n, (int)val, rep);
          n += rep;
        }
        if (lens[256] == 0) return false;        // no end-of-block code
        if (!build(lens, (int)hlit, kLitBits, lit_, kLitCap, true) || !build(lens + hlit, (int)hdist, kDistBits, dist_, kDistCap, false))
          return false;
        lt = lit_;
        dt = dist_;
      }
      // ---- the symbols of the block -----------------------------------------------------------------------------
      for (;;) {
        if (!fill()) return false;               // (the 8 bytes behind in_end may be read, never consumed: checked below)
        uint32_t e = lt[bitbuf & ((1u << kLitBits) - 1)];
        if ((e >> 5 & 7) == kPtr) e = lt[(e >> 16) + ((uint32_t)(bitbuf >> kLitBits) & ((1u << (e >> 8 & 31)) - 1))];
        bitbuf >>= (e & 31);
        bitsleft -= (e & 31);
        const uint32_t t = e >> 5 & 7;
        if (t == kLit) {
          if (out >= out_end) return false;
          *out++ = (uint8_t)(e >> 16);
          // up to two more literals out of the same refill (>= 41, then >= 26 bits are left)
          for (int more = 0; more < 2; ++more) {
            const uint32_t e2 = lt[bitbuf & ((1u << kLitBits) - 1)];
            if ((e2 >> 5 & 7) != kLit || out >= out_end) break;
            bitbuf >>= (e2 & 31);
            bitsleft -= (e2 & 31);
            *out++ = (uint8_t)(e2 >> 16);
          }
          continue;
        }
        if (t == kLen) {
          const uint32_t len = (e >> 16) + take(e >> 8 & 31);
          uint32_t d = dt[bitbuf & ((1u << kDistBits) - 1)];
          if ((d >> 5 & 7) == kPtr) d = dt[(d >> 16) + ((uint32_t)(bitbuf >> kDistBits) & ((1u << (d >> 8 & 31)) - 1))];
          if ((d >> 5 & 7) != kLen) return false;
          bitbuf >>= (d & 31);
          bitsleft -= (d & 31);
          const uint32_t dist = (d >> 16) + take(d >> 8 & 31);
          if (dist > (size_t)(out - out0) || len > (size_t)(out_end - out)) return false;
          const uint8_t* src = out - dist;
          if (dist >= 8 && (size_t)(out_end - out) >= len + 8) {      // 8 bytes at a time, may run up to 7 bytes over
            uint8_t* dst = out;
            uint8_t* const stop = out + len;
            do {
              uint64_t w;
              memcpy(&w, src, 8);
              memcpy(dst, &w, 8);
              src += 8;
              dst += 8;
            } while (dst < stop);
          } else if (dist == 1) {
            memset(out, *src, len);
          } else {
            for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
          }
          out += len;
          continue;
        }
        if (t == kEob) break;
        return false;
      }
    }
    // every bit consumed lies inside the in_n bytes, and the output is complete
    return in - (bitsleft >> 3) <= in_end && in >= in0 && out == out_end;
  }

 private:
  enum : uint32_t { kLit = 0, kLen = 1, kEob = 2, kPtr = 3, kInvalid = 4 };
  static constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
  static constexpr int kLitCap = (1 << kLitBits) + 288 * 16, kDistCap = (1 << kDistBits) + 32 * 128;
  // entry: bits 0-4 code length, 5-7 kind, 8-12 extra bits (kLen) / second-level index bits (kPtr), 16-31 value
  static uint32_t entry(uint32_t kind, uint32_t len, uint32_t extra, uint32_t value) {
    return len | (kind << 5) | (extra << 8) | (value << 16);
  }
  // canonical Huffman code of `n` code lengths -> two-level table indexed by the next bits of the stream (LSB first)
  static bool build(const uint8_t* lens, int n, int tbits, uint32_t* table, int cap, bool litlen) {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    uint32_t count[16] = {0}, next[16];
    int maxlen = 0;
    for (int i = 0; i < n; ++i) {
      if (lens[i] > 15) return false;
      ++count[lens[i]];
      if (lens[i] > maxlen) maxlen = lens[i];
    }
    count[0] = 0;
    uint32_t code = 0, kraft = 0;
    for (int l = 1; l <= 15; ++l) {
      code = (code + count[l - 1]) << 1;
      next[l] = code;
      kraft += count[l] << (15 - l);
    }
    if (kraft > (1u << 15)) return false;                       // over-subscribed
    const uint32_t invalid = entry(kInvalid, 0, 0, 0);
    for (int i = 0; i < (1 << tbits); ++i) table[i] = invalid;
    int next_free = 1 << tbits;
    const int sub_bits = maxlen > tbits ? maxlen - tbits : 0;
    for (int sym = 0; sym < n; ++sym) {
      const int L = lens[sym];
      if (!L) continue;
      uint32_t c = next[L]++, rev = 0;
      for (int b = 0; b < L; ++b) { rev = (rev << 1) | (c & 1); c >>= 1; }
      uint32_t e;
      if (tbits == kPreBits) e = entry(kLit, (uint32_t)L, 0, (uint32_t)sym);
      else if (!litlen) {
        if (sym >= 30) continue;                                  // codes 30, 31 never occur in valid data
        e = entry(kLen, (uint32_t)L, dist_extra[sym], dist_base[sym]);
      } else if (sym < 256) e = entry(kLit, (uint32_t)L, 0, (uint32_t)sym);
      else if (sym == 256) e = entry(kEob, (uint32_t)L, 0, 0);
      else if (sym < 286) e = entry(kLen, (uint32_t)L, len_extra[sym - 257], len_base[sym - 257]);
      else continue;
      if (L <= tbits) {
        for (uint32_t i = rev; i < (1u << tbits); i += 1u << L) table[i] = e;
      } else {
        const uint32_t prefix = rev & ((1u << tbits) - 1);
        uint32_t p = table[prefix];
        if ((p >> 5 & 7) != kPtr) {
          if (next_free + (1 << sub_bits) > cap) return false;
          p = entry(kPtr, 0, (uint32_t)sub_bits, (uint32_t)next_free);
          table[prefix] = p;
          for (int i = 0; i < (1 << sub_bits); ++i) table[next_free + i] = invalid;
          next_free += 1 << sub_bits;
        }
        const uint32_t start = p >> 16;
        for (uint32_t i = rev >> tbits; i < (1u << sub_bits); i += 1u << (L - tbits)) table[start + i] = e;
      }
    }
    return true;
  }
  bool fixed_ready_ = false;
  uint32_t lit_[kLitCap], dist_[kDistCap], fixed_lit_[kLitCap], fixed_dist_[kDistCap], pre_tab_[1 << kPreBits];
};

}  // namespace dmx_io
