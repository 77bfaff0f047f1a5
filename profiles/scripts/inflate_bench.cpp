// One-thread inflate time of every BGZF member of a BAM file, zlib vs pb_inflate_raw, min of 25 interleaved passes.
//   g++ -O2 -o /tmp/inflate_bench profiles/scripts/inflate_bench.cpp plastid_b200/csrc/pb_inflate.o -lz && /tmp/inflate_bench x.bam
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
extern "C" int pb_inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);
struct M { size_t src, csize, dst, usize; };
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> comp(n + 64); if (fread(comp.data(), 1, n, f) != n) return 1; fclose(f);
    std::vector<M> ms; size_t off = 0, dst = 0;
    while (off + 18 <= n) {
        const uint8_t *p = comp.data() + off;
        uint32_t xlen = p[10] | (p[11] << 8), bsize = (p[16] | (p[17] << 8)) + 1u;
        uint32_t isize; memcpy(&isize, p + bsize - 4, 4);
        ms.push_back({off + 12 + xlen, bsize - xlen - 20, dst, isize}); dst += isize; off += bsize;
    }
    std::vector<uint8_t> a(dst + 64), b(dst + 64);
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double best_z = 1e9, best_p = 1e9;
    for (int rep = 0; rep < 25; ++rep) {
        double t = now();
        for (auto &m : ms) { if (!m.usize) continue; z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
            zs.next_in = comp.data() + m.src; zs.avail_in = m.csize; zs.next_out = a.data() + m.dst; zs.avail_out = m.usize;
            if (inflate(&zs, Z_FINISH) != Z_STREAM_END) return 2; inflateEnd(&zs); }
        double tz = now() - t; if (tz < best_z) best_z = tz;
        t = now();
        for (auto &m : ms) { if (!m.usize) continue; if (pb_inflate_raw(comp.data() + m.src, m.csize, b.data() + m.dst, m.usize)) return 3; }
        double tp = now() - t; if (tp < best_p) best_p = tp;
    }
    if (memcmp(a.data(), b.data(), dst)) { printf("MISMATCH\n"); return 4; }
    printf("%s: %zu members, %.1f MB -> %.1f MB; zlib %.3f s (%.0f MB/s), pb %.3f s (%.0f MB/s), x%.2f\n", argv[1], ms.size(), n / 1e6, dst / 1e6,
           best_z, dst / 1e6 / best_z, best_p, dst / 1e6 / best_p, best_z / best_p);
}
