// pb_bam.cpp — host BAM -> SoA alignment batch decoder (BGZF + BAM records on zlib, multithreaded).
//
// Replaces what the reference delegates to pysam [3rd-party] on the way into the hot path:
// AlignmentFile open / references / lengths / mapped (plastid/genomics/genome_array.py:660-690),
// the per-region fetch (:800-809) and AlignedSegment.positions / is_reverse
// (plastid/genomics/map_factories.pyx:243,349,448,629).  Instead of seeking per region, the whole
// coordinate-sorted file is streamed once: BGZF blocks are inflated in parallel, records are walked,
// and every mapped record becomes one row of the packed batch (include/plastid_b200.h): CIGAR M/=/X
// runs are merged into aligned blocks (D/N split them, I/S/H/P do not), L = aligned bases.
// Like the reference, no flag other than "reverse" is interpreted (secondary / duplicate / QC-fail
// records count); records without a reference or with the unmapped flag have no positions and are
// skipped.  File format: SAM/BAM specification v1 §4 (BGZF §4.1, BAM §4.2); cross-checked against
// the reference's vendored htslib 1.3 through tests/golden/*.bam (see oracle/Makefile).
// Not resolved: CIGARs of more than 65535 operations kept in a CG:B,I tag behind a "<l>S<n>N" placeholder (SAM spec
// 4.2.2, long reads) — the record then has no aligned block and maps nowhere, as with the reference's htslib 1.3,
// which predates the tag.
#include <climits>
#include <cstdint>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "plastid_b200.h"

void pb_set_error(const char *fmt, ...);

namespace {

struct Block { size_t src, csize, dst, usize; uint32_t crc; };   // one BGZF member inside the current window

struct Decoded {            // per-chunk output of the record conversion
    std::vector<int32_t> start;
    std::vector<uint32_t> meta;
    std::vector<uint32_t> nlisted;   // listed block rows per read (0 for single-block reads)
    std::vector<int32_t> blk;        // {rel_start, len} pairs
    std::vector<int32_t> tid;
    int64_t mapped = 0, skipped = 0;
    int max_span = 1;
    std::string err;
};

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// CIGAR consume table, SAM spec §1.4.6 == kent/src/htslib/htslib/sam.h:79-104: ops M(0) =(7) X(8) emit
// reference positions; D(2) N(3) advance the reference only; I S H P B neither.
bool convert_record(const uint8_t *rec, uint32_t size, Decoded &out)
{
    if (size < 32) { out.err = "truncated BAM record"; return false; }
    const int32_t tid = (int32_t)rd32(rec), pos = (int32_t)rd32(rec + 4);
    const uint32_t l_read_name = rec[8];
    const uint32_t n_cigar = rd16(rec + 12), flag = rd16(rec + 14);
    if (tid < 0 || (flag & 0x4)) { out.skipped++; return true; }
    // `bamfile.mapped` (genome_array.py:690) is the index statistic: every placed record without the unmapped flag,
    // whatever its CIGAR (htslib hts_idx_push counts by flag) — counted here before records without positions are skipped
    out.mapped++;
    if (n_cigar == 0) { out.skipped++; return true; }
    if (32 + l_read_name + 4ull * n_cigar > size) { out.err = "BAM record shorter than its CIGAR"; return false; }
    const uint8_t *cig = rec + 32 + l_read_name;
    if (pos < 0) { out.err = "mapped BAM record with a negative position"; return false; }
    int32_t ref = 0, blocks[2 * PB_MAX_BLOCKS];
    int nb = 0;
    int64_t L = 0, span64 = 0;
    for (uint32_t k = 0; k < n_cigar; ++k) {
        const uint32_t c = rd32(cig + 4 * k), op = c & 0xf, len = c >> 4;
        if (op == 0 || op == 7 || op == 8 || op == 2 || op == 3) {
            span64 += len;                                   // reference bases consumed so far
            if ((int64_t)pos + span64 > INT32_MAX) { out.err = "alignment reaching beyond 2^31 reference positions"; return false; }
        }
        if (op == 0 || op == 7 || op == 8) {
            if (len == 0) continue;
            if (nb && blocks[2 * nb - 2] + blocks[2 * nb - 1] == ref) blocks[2 * nb - 1] += (int32_t)len;
            else {
                if (nb == PB_MAX_BLOCKS) { out.err = "alignment with more than 255 aligned blocks"; return false; }
                blocks[2 * nb] = ref; blocks[2 * nb + 1] = (int32_t)len; ++nb;
            }
            ref += (int32_t)len; L += len;
        } else if (op == 2 || op == 3) {
            ref += (int32_t)len;
        }
    }
    if (L > 0xFFFF) { out.err = "alignment with more than 65535 aligned bases"; return false; }
    // no aligned base at all (a CIGAR of S / I / H / P only): len(read.positions) == 0 — no rule can place it and a
    // batch row without blocks would read as filler; counted with the records that carry no positions
    if (L == 0) { out.skipped++; return true; }
    int32_t start = pos;
    if (nb && blocks[0] != 0) {            // leading D/N: positions start after it
        const int32_t shift = blocks[0];
        start += shift;
        for (int j = 0; j < nb; ++j) blocks[2 * j] -= shift;
    }
    out.tid.push_back(tid);
    out.start.push_back(start);
    out.meta.push_back((uint32_t)L | ((flag & 0x10) ? (1u << 16) : 0u) | ((uint32_t)nb << 24));
    const int span = nb ? blocks[2 * nb - 2] + blocks[2 * nb - 1] : 1;
    if (span > out.max_span) out.max_span = span;
    if (nb > 1) {
        out.nlisted.push_back((uint32_t)nb);
        out.blk.insert(out.blk.end(), blocks, blocks + 2 * nb);
    } else {
        out.nlisted.push_back(0);
    }
    return true;
}

// growable byte buffer that does not zero-fill what inflate is about to overwrite
struct RawBuf {
    uint8_t *p = nullptr;
    size_t cap = 0, len = 0;
    ~RawBuf() { free(p); }
    uint8_t *data() { return p; }
    size_t size() const { return len; }
    bool resize(size_t n)
    {
        if (n > cap) {
            const size_t want = n + n / 4 + 4096;
            uint8_t *q = (uint8_t *)realloc(p, want);
            if (!q) return false;
            p = q; cap = want;
        }
        len = n;
        return true;
    }
};

template <typename F> void parallel_for(int n_threads, size_t n, F fn)
{
    if (n_threads <= 1 || n < 2) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)n_threads, n);
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (auto &th : pool) th.join();
}

}  // namespace

struct pb_bam {
    std::string path;
    std::vector<std::string> ref_name;
    std::vector<int64_t> ref_len;
    std::vector<int32_t> start;
    std::vector<uint32_t> meta, blk_off;
    std::vector<int32_t> blk;
    std::vector<int64_t> chrom_read_off;
    int64_t mapped = 0, skipped = 0;
    int max_span = 1;
    bool decoded = false;
};

extern "C" int pb_bam_open(const char *path, pb_bam **out)
{
    if (!path || !out) { pb_set_error("pb_bam_open: null argument"); return PB_EINVAL; }
    FILE *fh = fopen(path, "rb");
    if (!fh) { pb_set_error("pb_bam_open: cannot open %s", path); return PB_EINVAL; }
    fclose(fh);
    pb_bam *h = new pb_bam();
    h->path = path;
    *out = h;
    return PB_OK;
}

extern "C" void pb_bam_close(pb_bam *h) { delete h; }

// Stream the file window by window: read whole BGZF members, inflate them in parallel behind the
// carried-over tail of the previous window, walk the records, convert them in parallel.
extern "C" int pb_bam_decode(pb_bam *h, int n_threads)
{
    if (!h) { pb_set_error("pb_bam_decode: null handle"); return PB_EINVAL; }
    if (h->decoded) return PB_OK;
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    FILE *fh = fopen(h->path.c_str(), "rb");
    if (!fh) { pb_set_error("pb_bam_decode: cannot open %s", h->path.c_str()); return PB_EINVAL; }
    size_t kWindow = (size_t)64 << 20;               // compressed bytes per window
    if (const char *w = getenv("PB_BAM_WINDOW")) {     // tests shrink it to exercise the carry-over paths
        const long v = atol(w);
        if (v >= (1 << 16)) kWindow = (size_t)v;
    }
    size_t kWalkMin = (size_t)4 << 20;                 // plain bytes below which the record walk stays serial
    if (const char *w = getenv("PB_BAM_WALK_MIN")) kWalkMin = (size_t)std::max(0l, atol(w));
    // input: the file mapped read-only where that works (members are inflated straight out of the page cache by the
    // worker threads, the kernel reads ahead), else read() into a window buffer (pipes; PB_BAM_NOMMAP=1 for tests)
    const uint8_t *map = nullptr;
    size_t map_size = 0, map_pos = 0;
    {
        struct stat st;
        if (!getenv("PB_BAM_NOMMAP") && fstat(fileno(fh), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(fh), 0);
            if (m != MAP_FAILED) {
                map = (const uint8_t *)m;
                map_size = (size_t)st.st_size;
                madvise(m, map_size, MADV_SEQUENTIAL);
            }
        }
    }
    std::vector<uint8_t> comp(map ? 0 : kWindow + (1 << 16));
    const uint8_t *cbase = comp.data();
    RawBuf plain;
    size_t comp_have = 0, carry = 0;
    bool eof = false, header_done = false;
    int32_t last_tid = -1;
    std::vector<uint32_t> nlisted_all;
    std::vector<int32_t> tid_all;
    int rc = PB_OK;
    std::string err;

    const bool dbg = getenv("PB_BAM_DEBUG") != nullptr;
    const bool check_crc = getenv("PB_BAM_NOCRC") == nullptr;      // every member's CRC32 is verified (zlib's crc32)
    const bool use_zlib = getenv("PB_BAM_ZLIB") != nullptr;       // A/B and cross-checks; default: pb_inflate_raw
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_read = 0, t_inflate = 0, t_walk = 0, t_conv = 0, t_app = 0, t0 = now();
    size_t n_chains_all = 0, n_chains_rewalked = 0;        // speculative record walk: stretches / stretches walked twice
    while (rc == PB_OK && !(eof && comp_have == 0)) {
        double ta = now();
        if (map) {
            cbase = map + map_pos;
            comp_have = std::min(kWindow, map_size - map_pos);
            eof = map_pos + comp_have == map_size;
            if (!eof) {                                    // the window after this one: start reading it now
                const size_t next = (map_pos + comp_have) & ~(size_t)4095;
                madvise((void *)(map + next), std::min(kWindow, map_size - next), MADV_WILLNEED);
            }
        } else if (!eof && comp_have < kWindow) {
            const size_t got = fread(comp.data() + comp_have, 1, kWindow - comp_have, fh);
            comp_have += got;
            if (got == 0) eof = true;
        }
        // split the window into whole BGZF members
        std::vector<Block> blocks;
        size_t off = 0, udst = carry;
        while (off + 18 <= comp_have) {
            const uint8_t *p = cbase + off;
            if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) { err = "not a BGZF file (bad member header)"; rc = PB_EINVAL; break; }
            const uint32_t xlen = rd16(p + 10);
            if (off + 12 + xlen > comp_have) break;
            uint32_t bsize = 0;
            for (uint32_t x = 0; x + 4 <= xlen;) {           // find the 'BC' subfield
                const uint8_t *sf = p + 12 + x;
                const uint32_t slen = rd16(sf + 2);
                if (sf[0] == 'B' && sf[1] == 'C' && slen == 2) bsize = rd16(sf + 4) + 1u;
                x += 4 + slen;
            }
            if (!bsize) { err = "BGZF member without a BC subfield"; rc = PB_EINVAL; break; }
            // BSIZE covers the 12-byte header, the extra field and the 8-byte trailer at least: anything smaller would
            // put the CRC / ISIZE words before the member and make the compressed size wrap
            if (bsize < 12 + xlen + 8) { err = "corrupt BGZF member header (BSIZE smaller than the header it sits in)"; rc = PB_EINVAL; break; }
            if (off + bsize > comp_have) break;               // incomplete member: wait for more input
            const uint32_t isize = rd32(p + bsize - 4);
            if (isize > 65536) { err = "corrupt BGZF member header (ISIZE above the 64 KiB the format allows)"; rc = PB_EINVAL; break; }
            blocks.push_back({off + 12 + xlen, bsize - xlen - 20, udst, isize, rd32(p + bsize - 8)});
            udst += isize;
            off += bsize;
        }
        if (rc) break;
        if (blocks.empty()) {
            if (eof) { if (comp_have) { err = "truncated BGZF member at end of file"; rc = PB_EINVAL; } break; }
            if (comp_have >= kWindow) { err = "BGZF member larger than the read window"; rc = PB_EINVAL; break; }
            continue;
        }
        t_read += now() - ta; ta = now();
        if (!plain.resize(udst)) { err = "out of memory"; rc = PB_EINVAL; break; }
        std::atomic<int> bad{0};
        parallel_for(n_threads, blocks.size(), [&](size_t i) {
            const Block &b = blocks[i];
            if (b.usize == 0) return;
            if (!use_zlib) {
                if (pb_inflate_raw(cbase + b.src, b.csize, plain.data() + b.dst, b.usize) != 0) bad = 1;
                else if (check_crc && (uint32_t)crc32(crc32(0L, Z_NULL, 0), plain.data() + b.dst, (uInt)b.usize) != b.crc) bad = 2;
                return;
            }
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            zs.next_in = (Bytef *)(cbase + b.src); zs.avail_in = (uInt)b.csize;
            zs.next_out = plain.data() + b.dst; zs.avail_out = (uInt)b.usize;
            const int zr = inflate(&zs, Z_FINISH);
            if (zr != Z_STREAM_END || zs.total_out != b.usize) bad = 1;
            else if (check_crc && (uint32_t)crc32(crc32(0L, Z_NULL, 0), plain.data() + b.dst, (uInt)b.usize) != b.crc) bad = 2;
            inflateEnd(&zs);
        });
        if (bad) { err = bad == 2 ? "corrupt BGZF member (CRC32 mismatch)" : "corrupt BGZF member (inflate failed)"; rc = PB_EINVAL; break; }
        if (map) map_pos += off;
        else memmove(comp.data(), comp.data() + off, comp_have - off);
        comp_have -= off;

        t_inflate += now() - ta; ta = now();
        // BAM header (once)
        size_t cur = 0;
        const size_t have = plain.size();
        if (!header_done) {
            if (have < 12) { if (eof && comp_have == 0) { err = "truncated BAM header"; rc = PB_EINVAL; } carry = have; continue; }
            if (memcmp(plain.data(), "BAM\1", 4) != 0) { err = "not a BAM file (bad magic)"; rc = PB_EINVAL; break; }
            const uint32_t l_text = rd32(plain.data() + 4);
            size_t p = 8 + (size_t)l_text;
            if (p + 4 > have) { carry = have; continue; }
            const uint32_t n_ref = rd32(plain.data() + p);
            p += 4;
            std::vector<std::string> names;
            std::vector<int64_t> lens;
            bool complete = true;
            for (uint32_t i = 0; i < n_ref; ++i) {
                if (p + 4 > have) { complete = false; break; }
                const uint32_t l_name = rd32(plain.data() + p);
                if (p + 4 + l_name + 4 > have) { complete = false; break; }
                names.emplace_back((const char *)plain.data() + p + 4, l_name ? l_name - 1 : 0);
                lens.push_back((int32_t)rd32(plain.data() + p + 4 + l_name));
                p += 8 + l_name;
            }
            if (!complete) { carry = have; continue; }
            h->ref_name = names; h->ref_len = lens;
            header_done = true;
            cur = p;
        }
        // record boundaries in this window.  A record's size is only known from its own header, so the walk is a
        // dependent chain of 4-byte loads; it is run speculatively in parallel: the window is cut at BGZF member
        // boundaries (htslib starts a new member rather than splitting a record, so a member boundary is almost always
        // a record boundary), every thread walks from its guess to the next cut, and a chain is accepted only when
        // the walk before it arrives exactly at its first byte — otherwise that stretch is walked again serially
        // from the true position.  Only the first record of every 65536 is noted; the conversion threads walk
        // their own chunk again.
        const size_t chunk = 1 << 16;
        struct Chain { size_t begin = 0, limit = 0, end = 0; std::vector<std::pair<size_t, size_t>> chunks; };
        auto walk = [&](Chain &c) {            // records starting in [begin, limit); 'end' = first byte not consumed
            size_t at = c.begin, n_here = 0;
            c.chunks.clear();
            while (at < c.limit && at + 4 <= have) {
                const uint32_t bs = rd32(plain.data() + at);
                if (at + 4 + bs > have) break;
                if ((n_here & (chunk - 1)) == 0) c.chunks.emplace_back(at, 0);
                ++n_here;
                c.chunks.back().second++;
                at += 4 + (size_t)bs;
            }
            c.end = at;
        };
        std::vector<Chain> chains;
        {
            const size_t n_parts = (n_threads > 1 && have - cur >= kWalkMin) ? (size_t)n_threads : 1;
            std::vector<size_t> cuts{cur};
            for (size_t k = 1; k < n_parts; ++k) {
                const size_t g = blocks[k * blocks.size() / n_parts].dst;
                if (g > cuts.back()) cuts.push_back(g);
            }
            chains.resize(cuts.size());
            for (size_t k = 0; k < cuts.size(); ++k) {
                chains[k].begin = cuts[k];
                chains[k].limit = k + 1 < cuts.size() ? cuts[k + 1] : have;
            }
            parallel_for(n_threads, chains.size(), [&](size_t k) { walk(chains[k]); });
        }
        std::vector<std::pair<size_t, size_t>> chunk_first;    // (first byte, records) per conversion chunk
        size_t n_recs = 0;
        for (size_t k = 0; k < chains.size(); ++k) {
            Chain &c = chains[k];
            ++n_chains_all;
            if (c.begin != cur) {                                // guess missed: walk this stretch from the true position
                ++n_chains_rewalked;
                if (cur >= c.limit) continue;                    // (a record reaching past the whole stretch)
                c.begin = cur;
                walk(c);
            }
            chunk_first.insert(chunk_first.end(), c.chunks.begin(), c.chunks.end());
            for (const auto &ch : c.chunks) n_recs += ch.second;
            cur = c.end;
            if (cur < c.limit) break;                            // stopped at an incomplete record: the window ends here
        }
        t_walk += now() - ta; ta = now();
        // convert in parallel, chunk by chunk
        const size_t n_chunks = chunk_first.size();
        std::vector<Decoded> parts(n_chunks);
        parallel_for(n_threads, n_chunks, [&](size_t ci) {
            Decoded &d = parts[ci];
            const size_t n_here = chunk_first[ci].second;
            d.start.reserve(n_here); d.meta.reserve(n_here); d.nlisted.reserve(n_here); d.tid.reserve(n_here);
            size_t at = chunk_first[ci].first;
            for (size_t i = 0; i < n_here && d.err.empty(); ++i) {
                const uint32_t bs = rd32(plain.data() + at);
                convert_record(plain.data() + at + 4, bs, d);
                at += 4 + (size_t)bs;
            }
            // coordinate order is a precondition (samtools sort): reference ids never decrease.  (A leading
            // deletion can move a START past its successor; that is tolerated here and repaired below.)
            for (size_t i = 1; i < d.tid.size() && d.err.empty(); ++i)
                if (d.tid[i] < d.tid[i - 1]) d.err = "BAM file is not coordinate-sorted";
        });
        t_conv += now() - ta; ta = now();
        // append in order: sizes -> offsets -> parallel copies into the grown arrays
        std::vector<size_t> at_read(n_chunks + 1, h->start.size()), at_blk(n_chunks + 1, h->blk.size());
        for (size_t ci = 0; ci < n_chunks && rc == PB_OK; ++ci) {
            Decoded &d = parts[ci];
            if (!d.err.empty()) { err = d.err; rc = PB_EINVAL; break; }
            if (!d.tid.empty()) {
                if (d.tid.front() < last_tid) { err = "BAM file is not coordinate-sorted"; rc = PB_EINVAL; break; }
                last_tid = d.tid.back();
            }
            at_read[ci + 1] = at_read[ci] + d.start.size();
            at_blk[ci + 1] = at_blk[ci] + d.blk.size();
            h->mapped += d.mapped; h->skipped += d.skipped;
            h->max_span = std::max(h->max_span, d.max_span);
        }
        if (rc) break;
        h->start.resize(at_read[n_chunks]); h->meta.resize(at_read[n_chunks]);
        nlisted_all.resize(at_read[n_chunks]); tid_all.resize(at_read[n_chunks]);
        h->blk.resize(at_blk[n_chunks]);
        parallel_for(n_threads, n_chunks, [&](size_t ci) {
            const Decoded &d = parts[ci];
            const size_t n_here = d.start.size();
            if (n_here) {
                memcpy(h->start.data() + at_read[ci], d.start.data(), n_here * sizeof(int32_t));
                memcpy(h->meta.data() + at_read[ci], d.meta.data(), n_here * sizeof(uint32_t));
                memcpy(nlisted_all.data() + at_read[ci], d.nlisted.data(), n_here * sizeof(uint32_t));
                memcpy(tid_all.data() + at_read[ci], d.tid.data(), n_here * sizeof(int32_t));
            }
            if (!d.blk.empty()) memcpy(h->blk.data() + at_blk[ci], d.blk.data(), d.blk.size() * sizeof(int32_t));
        });
        carry = have - cur;
        memmove(plain.data(), plain.data() + cur, carry);
        plain.resize(carry);
        t_app += now() - ta;
    }
    if (map) munmap((void *)map, map_size);
    fclose(fh);
    if (dbg) fprintf(stderr, "pb_bam_decode: read %.3f inflate %.3f walk %.3f convert %.3f append %.3f total %.3f s; "
                             "record walk: %zu stretches, %zu walked again\n",
                     t_read, t_inflate, t_walk, t_conv, t_app, now() - t0, n_chains_all, n_chains_rewalked);
    if (rc == PB_OK && !header_done) { err = "empty or truncated BAM file"; rc = PB_EINVAL; }
    if (rc == PB_OK && carry != 0) { err = "truncated BAM record at end of file"; rc = PB_EINVAL; }
    if (rc) { pb_set_error("pb_bam_decode(%s): %s", h->path.c_str(), err.c_str()); return rc; }

    const size_t n = h->start.size(), n_ref = h->ref_name.size();
    // reference ids never decrease (checked above), so the per-chromosome offsets are n_ref binary searches
    h->chrom_read_off.assign(n_ref + 1, 0);
    if (n && (size_t)tid_all.back() >= n_ref) {
        pb_set_error("pb_bam_decode: record refers to reference %d of %zu", tid_all.back(), n_ref);
        return PB_EINVAL;
    }
    for (size_t c = 0; c <= n_ref; ++c)
        h->chrom_read_off[c] = std::lower_bound(tid_all.begin(), tid_all.end(), (int32_t)c) - tid_all.begin();
    // starts must be non-decreasing per chromosome; a (rare) leading-deletion shift is repaired by a
    // stable sort of that chromosome's rows
    bool any_multi = !h->blk.empty();
    std::atomic<int> unsorted{0};
    const size_t piece = (size_t)1 << 20;
    parallel_for(n_threads, (n + piece - 1) / piece, [&](size_t k) {
        const size_t a = std::max<size_t>(k * piece, 1), e = std::min(n, (k + 1) * piece);
        for (size_t i = a; i < e; ++i)
            if (h->start[i] < h->start[i - 1] && tid_all[i] == tid_all[i - 1]) { unsorted = 1; return; }
    });
    const bool sorted = unsorted == 0;
    if (!sorted) {
        std::vector<uint64_t> blk_start(n + 1, 0);
        for (size_t i = 0; i < n; ++i) blk_start[i + 1] = blk_start[i] + nlisted_all[i];
        std::vector<size_t> order(n);
        for (size_t i = 0; i < n; ++i) order[i] = i;
        for (size_t c = 0; c < n_ref; ++c)
            std::stable_sort(order.begin() + h->chrom_read_off[c], order.begin() + h->chrom_read_off[c + 1],
                             [&](size_t a, size_t b) { return h->start[a] < h->start[b]; });
        std::vector<int32_t> s2(n), b2;
        std::vector<uint32_t> m2(n), nl2(n);
        b2.reserve(h->blk.size());
        for (size_t i = 0; i < n; ++i) {
            const size_t j = order[i];
            s2[i] = h->start[j]; m2[i] = h->meta[j]; nl2[i] = nlisted_all[j];
            b2.insert(b2.end(), h->blk.begin() + 2 * blk_start[j], h->blk.begin() + 2 * blk_start[j + 1]);
        }
        h->start.swap(s2); h->meta.swap(m2); h->blk.swap(b2); nlisted_all.swap(nl2);
    }
    h->blk_off.clear();
    if (any_multi) {
        h->blk_off.resize(n + 1);
        uint64_t run = 0;
        for (size_t i = 0; i < n; ++i) { h->blk_off[i] = (uint32_t)run; run += nlisted_all[i]; }
        if (run > 0xffffffffull) { pb_set_error("pb_bam_decode: more than 2^32-1 block rows"); return PB_EINVAL; }
        h->blk_off[n] = (uint32_t)run;
    }
    h->decoded = true;
    return PB_OK;
}

extern "C" int pb_bam_n_ref(const pb_bam *h) { return h ? (int)h->ref_name.size() : 0; }
extern "C" const char *pb_bam_ref_name(const pb_bam *h, int i) { return (h && i >= 0 && (size_t)i < h->ref_name.size()) ? h->ref_name[i].c_str() : ""; }
extern "C" int64_t pb_bam_ref_len(const pb_bam *h, int i) { return (h && i >= 0 && (size_t)i < h->ref_len.size()) ? h->ref_len[i] : -1; }
extern "C" int64_t pb_bam_n_reads(const pb_bam *h) { return h ? (int64_t)h->start.size() : 0; }
extern "C" int64_t pb_bam_n_blk(const pb_bam *h) { return h ? (int64_t)h->blk.size() / 2 : 0; }
extern "C" int64_t pb_bam_n_mapped(const pb_bam *h) { return h ? h->mapped : 0; }
extern "C" int64_t pb_bam_n_skipped(const pb_bam *h) { return h ? h->skipped : 0; }
extern "C" int32_t pb_bam_max_span(const pb_bam *h) { return h ? h->max_span : 1; }

extern "C" int pb_bam_copy(const pb_bam *h, int32_t *ref_start, uint32_t *meta, uint32_t *blk_off, int32_t *blk,
                           int64_t *chrom_read_off)
{
    if (!h || !h->decoded) { pb_set_error("pb_bam_copy: decode the file first"); return PB_EINVAL; }
    const size_t n = h->start.size();
    if (n && (!ref_start || !meta)) { pb_set_error("pb_bam_copy: null destination"); return PB_EINVAL; }
    if (!chrom_read_off) { pb_set_error("pb_bam_copy: null chrom_read_off"); return PB_EINVAL; }
    if (n) { memcpy(ref_start, h->start.data(), n * 4); memcpy(meta, h->meta.data(), n * 4); }
    memcpy(chrom_read_off, h->chrom_read_off.data(), h->chrom_read_off.size() * 8);
    if (!h->blk.empty()) {
        if (!blk_off || !blk) { pb_set_error("pb_bam_copy: file has multi-block reads, blk_off/blk needed"); return PB_EINVAL; }
        memcpy(blk_off, h->blk_off.data(), h->blk_off.size() * 4);
        memcpy(blk, h->blk.data(), h->blk.size() * 4);
    }
    return PB_OK;
}
