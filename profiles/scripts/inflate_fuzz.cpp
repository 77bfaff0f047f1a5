// Sanitizer fuzz of pb_inflate_raw against zlib (host only).  From the repo root:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -DPB_INFLATE_POISON -Iinclude \
//       profiles/scripts/inflate_fuzz.cpp plastid_b200/csrc/pb_inflate.cpp -o /tmp/inflate_fuzz -lz && /tmp/inflate_fuzz [streams]
// 6000 streams by default (all levels / strategies / flush kinds, exact-size heap buffers) + 6 truncated or corrupted
// variants of each, every one of them also read by zlib: both have to accept the same streams with the same bytes.
// PB_INFLATE_POISON fills the decode tables with plausible-looking stale entries before every
// build: nothing a previous block left behind may leak into the next one (tests/test_bam_io.py runs this too).
#include <zlib.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <random>
#include <vector>
extern "C" int pb_inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);
int main(int argc, char **argv) {
    const int n_iter = argc > 1 ? atoi(argv[1]) : 6000;
    std::mt19937_64 rng(7);
    size_t ok = 0, rejected = 0, accepted_bad = 0, diffs = 0;
    for (int it = 0; it < n_iter; ++it) {
        size_t n = (it % 5 == 0) ? rng() % 65281 : (size_t[]){0, 1, 3, 100, 319, 320, 321, 4000, 65280}[rng() % 9];
        std::vector<uint8_t> data(n);
        int kind = rng() % 7;
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
            case 0: data[i] = rng(); break;
            case 1: data[i] = rng() % 4; break;
            case 2: data[i] = 0; break;
            case 3: data[i] = i >= 50 ? data[i - 50] : rng(); break;
            case 4: { int p = 1 + it % 7; data[i] = i >= (size_t)p ? data[i - p] : rng(); break; }
            case 5: { int k = 0; while ((rng() & 1) && k < 30) ++k; data[i] = k + 40 * (rng() & 1); break; }
            default: data[i] = (rng() % 10 < 7) ? 'I' - (rng() % 3) : rng(); break;
            }
        }
        z_stream zs; memset(&zs, 0, sizeof zs);
        int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
        deflateInit2(&zs, rng() % 10, Z_DEFLATED, -15, 1 + rng() % 9, strategies[rng() % 5]);
        std::vector<uint8_t> comp(deflateBound(&zs, n) + 64);
        zs.next_in = data.data(); zs.avail_in = n; zs.next_out = comp.data(); zs.avail_out = comp.size();
        if (n > 10 && rng() % 3 == 0) { size_t k = 1 + rng() % (n - 1); zs.avail_in = k; deflate(&zs, (rng() & 1) ? Z_SYNC_FLUSH : Z_FULL_FLUSH); zs.avail_in = n - k; }
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) return 1;
        size_t clen = zs.total_out; deflateEnd(&zs);
        // exact-size heap buffers so that ASAN sees any access outside them
        uint8_t *in = (uint8_t *)malloc(clen ? clen : 1); if (clen) memcpy(in, comp.data(), clen);
        uint8_t *out = (uint8_t *)malloc(n ? n : 1);
        if (pb_inflate_raw(in, clen, out, n) != 0 || (n && memcmp(out, data.data(), n))) { printf("FAIL it=%d n=%zu kind=%d\n", it, n, kind); return 2; }
        ++ok;
        for (int c = 0; c < 6; ++c) {      // corrupted / truncated variants: no crash, no out-of-bounds, never more than n bytes
            size_t l2 = (c < 2 && clen > 1) ? rng() % clen : clen;
            uint8_t *bad = (uint8_t *)malloc(l2 ? l2 : 1); if (l2) memcpy(bad, in, l2);
            if (c >= 2 && l2) for (int f = 0; f < 1 + c; ++f) bad[rng() % l2] ^= 1u << (rng() % 8);
            if (c == 5 && l2) for (size_t i = 0; i < l2; ++i) bad[i] = rng();
            size_t n2 = (c == 4 && n) ? n - 1 - rng() % n : n;
            int rc = pb_inflate_raw(bad, l2, out, n2);
            if (rc == 0) ++accepted_bad; else ++rejected;
            {   // differential: zlib reads the same damaged stream; both have to agree on acceptance and on the bytes
                std::vector<uint8_t> zout(n2 + 16);
                z_stream zi; memset(&zi, 0, sizeof zi); inflateInit2(&zi, -15);
                zi.next_in = bad; zi.avail_in = l2; zi.next_out = zout.data(); zi.avail_out = n2 + 16;
                int zr = inflate(&zi, Z_FINISH);
                bool zok = zr == Z_STREAM_END && zi.total_out == n2;
                inflateEnd(&zi);
                if (zok != (rc == 0) || (zok && n2 && memcmp(zout.data(), out, n2))) {
                    printf("DIFF it=%d c=%d n2=%zu l2=%zu pb=%d zlib=%d total_out=%lu\n", it, c, n2, l2, rc, zr, zi.total_out); ++diffs;
                }
            }
            free(bad);
        }
        free(in); free(out);
    }
    printf("ok %zu, corrupted: rejected %zu, accepted %zu, disagreements with zlib %zu\n", ok, rejected, accepted_bad, diffs);
    if (diffs) return 5;
}
