// pb_inflate.cpp — raw DEFLATE (RFC 1951) decoder for BGZF members, written for the BAM -> SoA decoder.
//
// The reference leaves decompression to htslib/zlib behind pysam (plastid/genomics/genome_array.py:660,
// 800-809); inflating is the end-to-end bottleneck of the host side.  A BGZF member is at
// most 64 KiB of output whose size is known up front (ISIZE), so this decoder works buffer to buffer, with no
// streaming state: a 64-bit bit buffer refilled 7-8 bytes at a time, one table lookup per symbol (11-bit
// primary table for literals/lengths, 8-bit for distances, second-level tables for the longer codes), up to
// three lookups per refill (each one or two literals), and word-wise match copies while at least 320 bytes of output room remain — a
// careful byte-exact loop finishes each member, so nothing is ever written past `out + out_len` (the next
// member's thread owns those bytes).  Checked against zlib on every member shape in tests/test_bam_io.py.
#include <cstdint>
#include <cstring>

#include "plastid_b200.h"

namespace {

typedef uint64_t u64;
typedef uint32_t u32;

enum { T_INVALID = 0, T_LITERAL = 1, T_LENGTH = 2, T_END = 3, T_SUB = 4 };
#ifndef PB_LIT_BITS
#define PB_LIT_BITS 11
#endif
enum { LIT_BITS = PB_LIT_BITS, OFF_BITS = 8, PRE_BITS = 7, LIT_TABLE = 2400, OFF_TABLE = 512, PRE_TABLE = 128 };

// table entry: bits 0-7 stream bits to consume (the shift count as it is; for lengths and distances the code AND
// its extra bits, so that one shift per symbol is all the bit buffer's dependency chain sees), 8-11 extra bits (or
// sub-table index bits; literals: how many, 1 or 2), 12-15 type, 16-31 value (literal, or two: first | second << 8;
// base length / distance; sub-table start)
inline u32 entry(u32 type, u32 extra, u32 value) { return (extra << 8) | (type << 12) | (value << 16); }
inline u32 e_type(u32 e) { return (e >> 12) & 15; }
inline u32 e_bits(u32 e) { return e & 63; }
inline u32 e_extra(u32 e) { return (e >> 8) & 15; }
inline u32 e_value(u32 e) { return e >> 16; }
// the extra bits of a length / distance entry, out of the bit buffer as it was before the entry's bits were dropped
inline u32 e_extra_value(u64 saved, u32 e) { return (u32)(saved >> (e_bits(e) - e_extra(e))) & ((1u << e_extra(e)) - 1); }
inline u32 with_bits(u32 e, u32 code_bits) { return e | (code_bits + (e_type(e) == T_LENGTH ? e_extra(e) : 0)); }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kOffBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kOffExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPrecodeOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Rev8 {
    uint8_t v[256];
    constexpr Rev8() : v()
    {
        for (int i = 0; i < 256; ++i) {
            int r = 0;
            for (int b = 0; b < 8; ++b) r |= ((i >> b) & 1) << (7 - b);
            v[i] = (uint8_t)r;
        }
    }
    constexpr uint8_t operator[](u32 i) const { return v[i]; }
};
constexpr Rev8 kRev8;

enum Kind { K_PRECODE, K_LITLEN, K_OFFSET };

inline u32 symbol_entry(Kind kind, u32 sym)
{
    if (kind == K_PRECODE) return entry(T_LITERAL, 1, sym);
    if (kind == K_LITLEN) {
        if (sym < 256) return entry(T_LITERAL, 1, sym);
        if (sym == 256) return entry(T_END, 0, 0);
        if (sym < 286) return entry(T_LENGTH, kLenExtra[sym - 257], kLenBase[sym - 257]);
        return 0;
    }
    return sym < 30 ? entry(T_LENGTH, kOffExtra[sym], kOffBase[sym]) : 0;
}

// Two literals on one lookup: where a literal's code leaves enough index bits to determine the next symbol and
// that symbol is a literal too, the entry carries both (value = first | second << 8, bits = both codes).
void pair_literals(u32 *table, u32 tbits)
{
#ifdef PB_INFLATE_NO_PAIRS      // A/B only
    return;
#endif
    // descending and in place: entry i looks at entry i >> (its code length), which lies below i and is still single;
    // written without a branch (the pattern of literal / non-literal entries is not predictable)
    for (u32 i = (1u << tbits) - 1; i > 0; --i) {
        const u32 a = table[i];
        const u32 la = e_bits(a), b = table[i >> la];
        const u32 lit = (u32)T_LITERAL << 12;
        const u32 not_both = (((a ^ lit) | (b ^ lit)) & 0xF000u) | ((tbits - la - e_bits(b)) >> 31);   // 0 = pair them
        const u32 keep = 0u - (u32)(not_both != 0);
        const u32 pair = entry(T_LITERAL, 2, e_value(a) | ((b >> 8) & 0xFF00)) | (la + e_bits(b));
        table[i] = (a & keep) | (pair & ~keep);
    }
    const u32 a = table[0];                               // index 0 pairs with itself
    if (e_type(a) == T_LITERAL && e_extra(a) == 1 && 2 * e_bits(a) <= tbits)
        table[0] = entry(T_LITERAL, 2, e_value(a) * 0x101) | (2 * e_bits(a));
}

// Canonical Huffman code of `lens[0..n)` -> lookup table indexed by the next `tbits` stream bits (LSB first).
// Over-subscribed codes are refused; an incomplete code only in the form zlib accepts too (one code of length 1).
bool build_table(u32 *table, u32 table_cap, u32 tbits, const uint8_t *lens, u32 n, Kind kind)
{
    u32 count[16] = {0};
    for (u32 s = 0; s < n; ++s) count[lens[s]]++;
    const u32 tsize = 1u << tbits;
#ifdef PB_INFLATE_POISON        // tests: whatever an earlier block left in the table must not matter
    for (u32 i = 0; i < table_cap; ++i) table[i] = entry(T_LITERAL, 1, 0x5A) | (1 + i % 3);
#endif
    if (count[0] == n) {                                    // no codes at all: fine for distances (literals only)
        memset(table, 0, tsize * sizeof(u32));
        return kind == K_OFFSET;
    }
    int left = 1;
    u32 max_len = 0;
    for (u32 len = 1; len <= 15; ++len) {
        left = (left << 1) - (int)count[len];
        if (left < 0) return false;
        if (count[len]) max_len = len;
    }
    if (left > 0 && (kind == K_PRECODE || max_len != 1)) return false;
    if (left > 0) memset(table, 0, tsize * sizeof(u32));    // a complete code fills every entry, sub-tables included
    u32 next_code[16], code = 0;
    count[0] = 0;
    for (u32 len = 1; len <= 15; ++len) { code = (code + count[len - 1]) << 1; next_code[len] = code; }
    uint16_t rev[320], long_syms[320];
    uint8_t sub_max[1 << LIT_BITS];                       // per primary index: longest code below it, bit 7 = allocated
    u32 n_long = 0;
    if (max_len > tbits) memset(sub_max, 0, tsize);
    for (u32 s = 0; s < n; ++s) {
        const u32 len = lens[s];
        if (!len) continue;
        const u32 c = next_code[len]++;
        const u32 r = (((u32)kRev8[c & 255] << 8) | kRev8[c >> 8]) >> (16 - len);
        rev[s] = (uint16_t)r;
        if (len > tbits) {
            long_syms[n_long++] = (uint16_t)s;
            uint8_t &m = sub_max[r & (tsize - 1)];
            if (len > m) m = (uint8_t)len;
        } else {
            const u32 e = with_bits(symbol_entry(kind, s), len);
            for (u32 i = r; i < tsize; i += 1u << len) table[i] = e;
        }
    }
    u32 next_free = tsize;
    for (u32 k = 0; k < n_long; ++k) {                    // one sub-table per primary index that long codes share
        const u32 p = rev[long_syms[k]] & (tsize - 1);
        if (sub_max[p] & 0x80) continue;
        const u32 sbits = sub_max[p] - tbits;
        sub_max[p] |= 0x80;
        if (next_free + (1u << sbits) > table_cap) return false;
        table[p] = entry(T_SUB, sbits, next_free) | tbits;
        next_free += 1u << sbits;
    }
    if (kind == K_LITLEN) pair_literals(table, tbits);    // only now is every primary entry this code's own
    for (u32 k = 0; k < n_long; ++k) {
        const u32 s = long_syms[k], len = lens[s];
        const u32 r = rev[s], sub = table[r & (tsize - 1)];
        const u32 e = with_bits(symbol_entry(kind, s), len - tbits);
        u32 *t = table + e_value(sub);
        for (u32 i = r >> tbits; i < (1u << e_extra(sub)); i += 1u << (len - tbits)) t[i] = e;
    }
    return true;
}

struct Tables {
    u32 lit[LIT_TABLE];
    u32 off[OFF_TABLE];
};

const Tables &fixed_tables()
{
    static const Tables fixed = [] {
        Tables t;
        uint8_t lens[288];
        for (int i = 0; i < 144; ++i) lens[i] = 8;
        for (int i = 144; i < 256; ++i) lens[i] = 9;
        for (int i = 256; i < 280; ++i) lens[i] = 7;
        for (int i = 280; i < 288; ++i) lens[i] = 8;
        build_table(t.lit, LIT_TABLE, LIT_BITS, lens, 288, K_LITLEN);
        for (int i = 0; i < 32; ++i) lens[i] = 5;
        build_table(t.off, OFF_TABLE, OFF_BITS, lens, 32, K_OFFSET);
        return t;
    }();
    return fixed;
}

inline void copy8(uint8_t *dst, const uint8_t *src) { u64 v; memcpy(&v, src, 8); memcpy(dst, &v, 8); }
inline void store16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
inline u64 load64(const uint8_t *p) { u64 v; memcpy(&v, p, 8); return v; }      // little-endian hosts only (x86-64, aarch64)

struct Stream {
    const uint8_t *in_next, *in_end;
    u64 bitbuf = 0;
    u32 bitcnt = 0;            // valid bits in bitbuf (bits above them are either zero or the stream's own next bits)
    u32 overread = 0;          // zero bytes supplied past in_end
    void refill()              // careful: byte by byte, phantom zero bytes past the end are counted
    {
        while (bitcnt <= 56) {
            if (in_next < in_end) bitbuf |= (u64)*in_next++ << bitcnt;
            else ++overread;
            bitcnt += 8;
        }
    }
    void refill_fast()         // needs 8 readable bytes at in_next; leaves 56..63 valid bits
    {
        bitbuf |= load64(in_next) << bitcnt;
        in_next += (63 - bitcnt) >> 3;
        bitcnt |= 56;
    }
    u32 peek(u32 n) const { return (u32)bitbuf & ((1u << n) - 1); }
    void drop(u32 n) { bitbuf >>= n; bitcnt -= n; }
    u32 take(u32 n) { const u32 v = peek(n); drop(n); return v; }
};

// one table lookup, sub-table included; consumes the entry's bits (code + extra bits), returns the entry and the
// bit buffer the extra bits can be read from (e_extra_value)
inline u32 decode(Stream &s, const u32 *table, u32 tbits, u64 &saved)
{
    u32 e = table[s.peek(tbits)];
    if (e_type(e) == T_SUB) {
        s.drop(tbits);
        e = table[e_value(e) + s.peek(e_extra(e))];
    }
    saved = s.bitbuf;
    s.drop(e_bits(e));
    return e;
}

bool read_dynamic_header(Stream &stream, Tables &t)
{
    Stream s = stream;                                // locals, as in inflate_block
    s.refill();
    const u32 hlit = s.take(5) + 257, hdist = s.take(5) + 1, hclen = s.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t pre_lens[19] = {0};
    for (u32 i = 0; i < hclen; ++i) {
        if (s.bitcnt < 3) s.refill();
        pre_lens[kPrecodeOrder[i]] = (uint8_t)s.take(3);
    }
    u32 pre[PRE_TABLE];
    if (!build_table(pre, PRE_TABLE, PRE_BITS, pre_lens, 19, K_PRECODE)) return false;
    uint8_t lens[286 + 30 + 138];
    u32 i = 0;
    const u32 total = hlit + hdist;
    while (i < total) {
        if (s.bitcnt < 14) s.refill();                // 7 code bits + 7 extra bits at most
        u64 saved;
        const u32 e = decode(s, pre, PRE_BITS, saved);
        if (e_type(e) != T_LITERAL) return false;
        const u32 sym = e_value(e);
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        u32 rep, val = 0;
        if (sym == 16) {
            if (i == 0) return false;
            val = lens[i - 1];
            rep = 3 + s.take(2);
        } else if (sym == 17) {
            rep = 3 + s.take(3);
        } else {
            rep = 11 + s.take(7);
        }
        if (i + rep > total) return false;
        memset(lens + i, (int)val, rep);
        i += rep;
    }
    if (lens[256] == 0) return false;                 // a block has to be able to end
    stream = s;
    return build_table(t.lit, LIT_TABLE, LIT_BITS, lens, hlit, K_LITLEN) &&
           build_table(t.off, OFF_TABLE, OFF_BITS, lens + hlit, hdist, K_OFFSET);
}

enum { FAST_IN_MARGIN = 32, FAST_OUT_MARGIN = 320 };    // per iteration: 2 refills; 6 literals + 258 + 15 bytes of overrun

// Symbols of one compressed block.  Returns 0 when the end-of-block code was consumed, -1 on bad data.
int inflate_block(Stream &stream, const Tables &t, uint8_t *out_begin, uint8_t *&out_ref, uint8_t *out_end)
{
    // the bit buffer lives in locals: stores through `out` (bytes) could alias a Stream reached by reference
    Stream s = stream;
    uint8_t *out = out_ref;
#define PB_RETURN(rc) do { stream = s; out_ref = out; return (rc); } while (0)
    // fast loop: no bounds checks inside, the margins cover the worst case of one iteration.  The table entry of
    // the next symbol is fetched right after each refill, ahead of the match copy, so that its latency is hidden.
    if ((size_t)(s.in_end - s.in_next) >= FAST_IN_MARGIN && (size_t)(out_end - out) >= FAST_OUT_MARGIN) {
        s.refill_fast();
        u32 e = t.lit[s.peek(LIT_BITS)];
        u64 saved;
        for (;;) {
            // here: at least 56 valid bits, `e` = primary entry of the next symbol (nothing consumed for it yet)
#define PB_RESOLVE(e)                                                                                   \
            do {                                                                                        \
                if (e_type(e) == T_SUB) { s.drop(LIT_BITS); e = t.lit[e_value(e) + s.peek(e_extra(e))]; } \
                saved = s.bitbuf;                                                                       \
                s.drop(e_bits(e));                                                                      \
            } while (0)
            PB_RESOLVE(e);
            // up to three lookups on one refill (3 x 15 bits), each one or two literals: two bytes are stored either
            // way (the second is overwritten by whatever comes next), the entry says how far to advance
            if (e_type(e) == T_LITERAL) {
                store16(out, (uint16_t)e_value(e)); out += e_extra(e);
                e = t.lit[s.peek(LIT_BITS)];
                PB_RESOLVE(e);
                if (e_type(e) == T_LITERAL) {
                    store16(out, (uint16_t)e_value(e)); out += e_extra(e);
                    e = t.lit[s.peek(LIT_BITS)];
                    PB_RESOLVE(e);
                    if (e_type(e) == T_LITERAL) {
                        store16(out, (uint16_t)e_value(e)); out += e_extra(e);
                        if ((size_t)(s.in_end - s.in_next) < FAST_IN_MARGIN || (size_t)(out_end - out) < FAST_OUT_MARGIN) break;
                        s.refill_fast();
                        e = t.lit[s.peek(LIT_BITS)];
                        continue;
                    }
                }
            }
#undef PB_RESOLVE
            if (e_type(e) != T_LENGTH) PB_RETURN(e_type(e) == T_END ? 0 : -1);
            // a length first thing after a refill leaves 56 - 20 bits, enough for any distance (15 + 13); after
            // literals (15 + 15 + 20 bits at most, out of 56) the buffer is topped up first
            const u32 length = e_value(e) + e_extra_value(saved, e);
            if (s.bitcnt < 28) s.refill_fast();
            e = decode(s, t.off, OFF_BITS, saved);
            if (e_type(e) != T_LENGTH) PB_RETURN(-1);
            const u32 offset = e_value(e) + e_extra_value(saved, e);
            if (offset > (size_t)(out - out_begin)) PB_RETURN(-1);
            const uint8_t *src = out - offset;
            uint8_t *dst = out;
            out += length;
            const bool more = (size_t)(s.in_end - s.in_next) >= FAST_IN_MARGIN && (size_t)(out_end - out) >= FAST_OUT_MARGIN;
            if (more) {
                s.refill_fast();
                e = t.lit[s.peek(LIT_BITS)];
            }
            // 16 bytes unconditionally (most matches are shorter; the margin covers the overrun), a loop for the rest
            if (offset >= 8) {
                copy8(dst, src); copy8(dst + 8, src + 8);
                if (length > 16) {
                    dst += 16; src += 16;
                    do { copy8(dst, src); dst += 8; src += 8; } while (dst < out);
                }
            } else {
                // period < 8: every 8-byte copy is right in its first (dst - src) bytes, then the distance doubles
                do { copy8(dst, src); dst += dst - src; } while (dst - src < 8 && dst < out);
                while (dst < out) { copy8(dst, src); dst += 8; src += 8; }
            }
            if (!more) break;
        }
    }
    // careful loop: every read and write checked
    for (;;) {
        s.refill();
        u64 saved;
        u32 e = decode(s, t.lit, LIT_BITS, saved);
        if (e_type(e) == T_LITERAL) {
            if ((size_t)(out_end - out) < e_extra(e)) PB_RETURN(-1);
            *out++ = (uint8_t)e_value(e);
            if (e_extra(e) == 2) *out++ = (uint8_t)(e_value(e) >> 8);
            continue;
        }
        if (e_type(e) != T_LENGTH) {
            PB_RETURN(e_type(e) == T_END ? 0 : -1);
        }
        const u32 length = e_value(e) + e_extra_value(saved, e);
        s.refill();
        e = decode(s, t.off, OFF_BITS, saved);
        if (e_type(e) != T_LENGTH) PB_RETURN(-1);
        const u32 offset = e_value(e) + e_extra_value(saved, e);
        if (offset > (size_t)(out - out_begin) || length > (size_t)(out_end - out)) PB_RETURN(-1);
        const uint8_t *src = out - offset;
        for (u32 k = 0; k < length; ++k) out[k] = src[k];
        out += length;
        if (s.overread > 8) PB_RETURN(-1);
    }
#undef PB_RETURN
}

}  // namespace

// Inflates one raw DEFLATE stream of `in_len` bytes into exactly `out_len` bytes.  0 on success; -1 when the data
// is malformed, ends early, or does not produce exactly `out_len` bytes.  Never writes outside [out, out+out_len).
extern "C" int pb_inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    if ((!in && in_len) || (!out && out_len)) return -1;
    Stream s;
    s.in_next = in;
    s.in_end = in + in_len;
    uint8_t *out_next = out, *const out_end = out + out_len;
    Tables dyn;
    for (;;) {
        s.refill();
        const u32 final_block = s.take(1), type = s.take(2);
        if (type == 0) {                               // stored: back to a byte boundary, LEN, ~LEN, bytes
            s.drop(s.bitcnt & 7);
            u32 unread = s.bitcnt >> 3;
            if (s.overread > unread) return -1;
            unread -= s.overread;
            s.in_next -= unread;
            s.bitbuf = 0; s.bitcnt = 0; s.overread = 0;
            if (s.in_end - s.in_next < 4) return -1;
            const u32 len = s.in_next[0] | (s.in_next[1] << 8), nlen = s.in_next[2] | (s.in_next[3] << 8);
            s.in_next += 4;
            if ((len ^ nlen) != 0xffff) return -1;
            if (len > (size_t)(s.in_end - s.in_next) || len > (size_t)(out_end - out_next)) return -1;
            memcpy(out_next, s.in_next, len);
            s.in_next += len;
            out_next += len;
        } else if (type == 1) {
            if (inflate_block(s, fixed_tables(), out, out_next, out_end)) return -1;
        } else if (type == 2) {
            if (!read_dynamic_header(s, dyn)) return -1;
            if (inflate_block(s, dyn, out, out_next, out_end)) return -1;
        } else {
            return -1;
        }
        if (s.overread > (s.bitcnt >> 3)) return -1;   // bits were taken from beyond the end of the input
        if (final_block) break;
    }
    return out_next == out_end ? 0 : -1;
}
