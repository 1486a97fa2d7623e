// inflate_fast.h — a raw-DEFLATE (RFC 1951) decoder for BGZF blocks, host code.
//
// The native BAM ingest (ingest.cpp) spends most of its time inflating the ~64 KB BGZF blocks of a locus window
// (the reference leaves this to htslib/zlib inside pysam: bam_parser.py:206, 226, 333, 406).  A BGZF block is a
// complete DEFLATE stream whose inflated size and CRC-32 are known in advance, which allows a decoder without
// zlib's streaming state machine: whole-buffer input, 64-bit bit buffer refilled with one unaligned load,
// table-driven Huffman decoding (11-bit litlen / 8-bit distance root tables with sub-tables), word-wise match
// copies.  The caller verifies the CRC-32 of every block and falls back to zlib if this decoder refuses a
// stream or the checksum differs, so a defect here can cost time but never correctness.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stddef.h>

namespace tredsw_inflate {

class FastInflater {
public:
    // Inflate the raw DEFLATE stream in[0, in_len) into out[0, out_len); the buffer behind `out` must have at
    // least out_len + SLACK writable bytes.  True iff the stream ended exactly at out_len bytes.
    static constexpr size_t SLACK = 16;
    bool inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

private:
    enum Kind : uint32_t { INVALID = 0, LITERAL = 1, LENGTH = 2, END = 3, SUBTABLE = 4, SYMBOL = 5 };
    static constexpr int LIT_BITS = 11, DIST_BITS = 8, PRE_BITS = 7;
    static constexpr int LIT_CAP = 4096, DIST_CAP = 1024, PRE_CAP = 128;
    // entry: base (16) | extra bits (4) << 16 | code bits to consume (4) << 20 | kind (3) << 24
    static uint32_t entry(uint32_t kind, uint32_t base, uint32_t extra, uint32_t bits) {
        return base | (extra << 16) | (bits << 20) | (kind << 24);
    }
    static uint32_t e_base(uint32_t e) { return e & 0xffffu; }
    static uint32_t e_extra(uint32_t e) { return (e >> 16) & 15u; }
    static uint32_t e_bits(uint32_t e) { return (e >> 20) & 15u; }
    static uint32_t e_kind(uint32_t e) { return e >> 24; }

    uint32_t lit_[LIT_CAP], dist_[DIST_CAP], pre_[PRE_CAP];

    // which = 0 literal/length alphabet, 1 distance alphabet, 2 code-length alphabet
    static bool build(const uint8_t *lens, int nsyms, int which, uint32_t *table, int root_bits, int cap);
};

inline bool FastInflater::build(const uint8_t *lens, int nsyms, int which, uint32_t *table, int root_bits, int cap) {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                          67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                           1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    int count[16] = {0};
    for (int s = 0; s < nsyms; ++s) { if (lens[s] > 15) return false; ++count[lens[s]]; }
    count[0] = 0;
    int left = 1;                                   // over-subscribed codes are rejected, incomplete ones allowed
    for (int l = 1; l <= 15; ++l) { left <<= 1; left -= count[l]; if (left < 0) return false; }
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
    const int root_size = 1 << root_bits;
    for (int i = 0; i < root_size; ++i) table[i] = 0;   // INVALID
    // pass 1: canonical codes, bit-reversed (the stream is read LSB first); longest code per root prefix
    uint16_t rev_of[320];
    uint8_t sub_need[1 << LIT_BITS];
    if (nsyms > 320) return false;
    memset(sub_need, 0, (size_t)root_size);
    for (int s = 0; s < nsyms; ++s) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t c = next_code[l]++, r = 0;
        for (int b = 0; b < l; ++b) { r = (r << 1) | (c & 1u); c >>= 1; }
        rev_of[s] = (uint16_t)r;
        if (l > root_bits) {
            const uint32_t prefix = r & (uint32_t)(root_size - 1);
            if (l - root_bits > sub_need[prefix]) sub_need[prefix] = (uint8_t)(l - root_bits);
        }
    }
    // sub-tables behind the root table
    int top = root_size;
    for (int prefix = 0; prefix < root_size; ++prefix) {
        if (!sub_need[prefix]) continue;
        const int size = 1 << sub_need[prefix];
        if (top + size > cap) return false;
        table[prefix] = entry(SUBTABLE, (uint32_t)top, sub_need[prefix], (uint32_t)root_bits);
        for (int i = 0; i < size; ++i) table[top + i] = 0;
        top += size;
    }
    // pass 2: fill
    for (int s = 0; s < nsyms; ++s) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t e;
        auto make = [&](uint32_t bits) -> uint32_t {
            if (which == 0) {
                if (s < 256) return entry(LITERAL, (uint32_t)s, 0, bits);
                if (s == 256) return entry(END, 0, 0, bits);
                if (s <= 285) return entry(LENGTH, len_base[s - 257], len_extra[s - 257], bits);
                return 0;                            // 286, 287: never valid in a stream
            }
            if (which == 1) return s < 30 ? entry(SYMBOL, dist_base[s], dist_extra[s], bits) : 0;
            return entry(SYMBOL, (uint32_t)s, 0, bits);
        };
        const uint32_t r = rev_of[s];
        if (l <= root_bits) {
            e = make((uint32_t)l);
            for (uint32_t i = r; i < (uint32_t)root_size; i += 1u << l) table[i] = e;
        } else {
            const uint32_t prefix = r & (uint32_t)(root_size - 1);
            const uint32_t sub = table[prefix];
            if (e_kind(sub) != SUBTABLE) return false;
            const uint32_t base = e_base(sub), sub_bits = e_extra(sub), rem = (uint32_t)(l - root_bits);
            e = make(rem);
            for (uint32_t i = r >> root_bits; i < (1u << sub_bits); i += 1u << rem) table[base + i] = e;
        }
    }
    return true;
}

inline bool FastInflater::inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    const uint8_t *p = in, *const end = in + in_len;
    uint8_t *op = out, *const oend = out + out_len;
    uint64_t bitbuf = 0;
    unsigned bitcnt = 0;
    size_t zero_fill = 0;                             // bytes of zero padding consumed past the end of the input
    auto refill = [&]() {
        if (end - p >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);                         // little-endian host (x86-64 / aarch64)
            bitbuf |= w << bitcnt;
            p += (63 - bitcnt) >> 3;
            bitcnt |= 56;
        } else {
            while (bitcnt <= 56) {
                if (p < end) bitbuf |= (uint64_t)*p++ << bitcnt; else ++zero_fill;
                bitcnt += 8;
            }
        }
    };
    auto take = [&](unsigned n) -> uint32_t {         // n <= 32, caller has refilled
        const uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1ull));
        bitbuf >>= n; bitcnt -= n;
        return v;
    };
    for (;;) {
        refill();
        const uint32_t bfinal = take(1), btype = take(2);
        if (btype == 0) {
            // stored: skip to the byte boundary, LEN / NLEN, then raw bytes
            take(bitcnt & 7u);
            refill();
            const uint32_t len = take(16), nlen = take(16);
            if ((len ^ 0xffffu) != nlen) return false;
            const size_t in_buf = bitcnt >> 3;        // whole bytes still in the bit buffer: give the real ones back
            if (zero_fill > in_buf) return false;
            p -= in_buf - zero_fill;
            zero_fill = 0; bitbuf = 0; bitcnt = 0;
            if ((size_t)(end - p) < len || (size_t)(oend - op) < len) return false;
            memcpy(op, p, len);
            op += len; p += len;
        } else if (btype == 1 || btype == 2) {
            uint8_t lens[320];
            int nlit, ndist;
            if (btype == 1) {
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 288; i < 320; ++i) lens[i] = 5;
                nlit = 288; ndist = 32;
            } else {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                nlit = (int)take(5) + 257; ndist = (int)take(5) + 1;
                const int npre = (int)take(4) + 4;
                if (nlit > 286 || ndist > 30) return false;
                uint8_t plens[19] = {0};
                refill();
                for (int i = 0; i < npre; ++i) {
                    if (bitcnt < 3) refill();
                    plens[order[i]] = (uint8_t)take(3);
                }
                if (!build(plens, 19, 2, pre_, PRE_BITS, PRE_CAP)) return false;
                int i = 0;
                while (i < nlit + ndist) {
                    refill();
                    const uint32_t e = pre_[bitbuf & ((1u << PRE_BITS) - 1u)];
                    if (e_kind(e) != SYMBOL) return false;
                    take(e_bits(e));
                    const uint32_t sym = e_base(e);
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    int rep; uint8_t v = 0;
                    if (sym == 16) { if (i == 0) return false; v = lens[i - 1]; rep = 3 + (int)take(2); }
                    else if (sym == 17) rep = 3 + (int)take(3);
                    else rep = 11 + (int)take(7);
                    if (i + rep > nlit + ndist) return false;
                    while (rep--) lens[i++] = v;
                }
                if (lens[256] == 0) return false;     // no end-of-block code
                memmove(lens + 288, lens + nlit, (size_t)ndist);         // distance lengths behind a fixed offset
                for (int k = nlit; k < 288; ++k) lens[k] = 0;
                for (int k = 288 + ndist; k < 320; ++k) lens[k] = 0;
                nlit = 288; ndist = 32;
            }
            if (!build(lens, nlit, 0, lit_, LIT_BITS, LIT_CAP)) return false;
            if (!build(lens + 288, ndist, 1, dist_, DIST_BITS, DIST_CAP)) return false;
            // ---- symbols -----------------------------------------------------------------------------------
            for (;;) {
                refill();                              // >= 56 bits: litlen 15 + 5 extra + distance 15 + 13 extra = 48
                uint32_t e = lit_[bitbuf & ((1u << LIT_BITS) - 1u)];
                if (e_kind(e) == SUBTABLE) {
                    take(e_bits(e));
                    e = lit_[e_base(e) + (uint32_t)(bitbuf & ((1ull << e_extra(e)) - 1ull))];
                }
                take(e_bits(e));
                const uint32_t kind = e_kind(e);
                if (kind == LITERAL) {
                    if (op >= oend) return false;
                    *op++ = (uint8_t)e_base(e);
                    // a second and third literal often follow and still fit the bits at hand (3 x 15 <= 48)
                    e = lit_[bitbuf & ((1u << LIT_BITS) - 1u)];
                    if (e_kind(e) == LITERAL && op < oend) {
                        take(e_bits(e));
                        *op++ = (uint8_t)e_base(e);
                        e = lit_[bitbuf & ((1u << LIT_BITS) - 1u)];
                        if (e_kind(e) == LITERAL && op < oend) { take(e_bits(e)); *op++ = (uint8_t)e_base(e); }
                    }
                    continue;
                }
                if (kind == END) break;
                if (kind != LENGTH) return false;
                const uint32_t length = e_base(e) + take(e_extra(e));
                uint32_t d = dist_[bitbuf & ((1u << DIST_BITS) - 1u)];
                if (e_kind(d) == SUBTABLE) {
                    take(e_bits(d));
                    d = dist_[e_base(d) + (uint32_t)(bitbuf & ((1ull << e_extra(d)) - 1ull))];
                }
                if (e_kind(d) != SYMBOL) return false;
                take(e_bits(d));
                const uint32_t dist = e_base(d) + take(e_extra(d));
                if (dist > (size_t)(op - out) || length > (size_t)(oend - op)) return false;
                const uint8_t *src = op - dist;
                uint8_t *const stop = op + length;
                if (dist >= 16) {                      // the common case (BAM: 98 % of matches, 88 % of them <= 16 bytes):
                    memcpy(op, src, 16);               // one 16-byte move, more only for long matches; writes up to
                    for (uint32_t k = 16; k < length; k += 16) memcpy(op + k, src + k, 16);   // 15 bytes past `stop` (SLACK)
                    op = stop;
                } else if (dist >= 8) {                // word copies of an overlapping match
                    do { uint64_t w; memcpy(&w, src, 8); memcpy(op, &w, 8); src += 8; op += 8; } while (op < stop);
                    op = stop;
                } else if (dist == 1) {
                    memset(op, *src, length);
                    op = stop;
                } else {
                    while (op < stop) *op++ = *src++;
                }
            }
            if (zero_fill > 8) return false;          // ran well past the end of the input
        } else {
            return false;
        }
        if (bfinal) break;
    }
    return op == oend && zero_fill <= 8;
}

}  // namespace tredsw_inflate
