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

// ---------------------------------------------------------------------------------------------------------------
// Indexed region access: the .bai index and `fetch(reference, start, end)`
// ---------------------------------------------------------------------------------------------------------------
// Replaces pysam's `AlignmentFile.fetch(chrom, start, end)` as BAMGenomeArray.get_reads_and_counts calls it
// (plastid/genomics/genome_array.py:800-809) and `bamfile.mapped` (:690, the index statistic) for a file that has a
// .bai beside it: only the BGZF members the index points at are read and inflated, so a single-region query does not
// pay for the whole file.  Format: SAM/BAM specification v1 §5.2 (binning scheme, linear index, metadata pseudo-bin
// 37450); the same arithmetic as the reference's vendored htslib (kent/src/htslib/hts.c: reg2bins, hts_itr_query).
#include <unordered_map>
#include <unistd.h>
#include <fcntl.h>

struct pb_bai {
    struct Ref {
        std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
        std::vector<uint64_t> ioffset;
        int64_t mapped = -1, unmapped = -1;
    };
    std::vector<Ref> refs;
    int64_t no_coor = -1;
};

namespace {

inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// One BGZF member at a time, by file offset (pread): inflate + CRC32.
struct MemberReader {
    int fd = -1;
    std::vector<uint8_t> comp, plain;
    uint64_t coff = 0, next = 0;         // file offset of the member in `plain`, and of the one after it
    std::string err;
    ~MemberReader() { if (fd >= 0) close(fd); }
    bool open(const char *path) { fd = ::open(path, O_RDONLY); return fd >= 0; }
    // -> 1 member loaded, 0 end of file, -1 error
    int load(uint64_t at)
    {
        uint8_t head[18];
        const ssize_t got = pread(fd, head, 18, (off_t)at);
        if (got == 0) return 0;
        if (got != 18) { err = "truncated BGZF member header"; return -1; }
        if (head[0] != 31 || head[1] != 139 || head[2] != 8 || !(head[3] & 4)) { err = "not a BGZF member (bad virtual offset in the index?)"; return -1; }
        const uint32_t xlen = rd16(head + 10);
        std::vector<uint8_t> extra(xlen);
        if (xlen && pread(fd, extra.data(), xlen, (off_t)at + 12) != (ssize_t)xlen) { err = "truncated BGZF extra field"; return -1; }
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint32_t slen = rd16(extra.data() + x + 2);
            if (extra[x] == 'B' && extra[x + 1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = rd16(extra.data() + x + 4) + 1u;
            x += 4 + slen;
        }
        if (!bsize || bsize < 12 + xlen + 8) { err = "corrupt BGZF member header"; return -1; }
        comp.resize(bsize);
        if (pread(fd, comp.data(), bsize, (off_t)at) != (ssize_t)bsize) { err = "truncated BGZF member"; return -1; }
        const uint32_t isize = rd32(comp.data() + bsize - 4), crc = rd32(comp.data() + bsize - 8);
        if (isize > 65536) { err = "corrupt BGZF member header (ISIZE above 64 KiB)"; return -1; }
        plain.resize(isize);
        if (isize) {
            if (pb_inflate_raw(comp.data() + 12 + xlen, bsize - xlen - 20, plain.data(), isize) != 0) { err = "corrupt BGZF member (inflate failed)"; return -1; }
            if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), plain.data(), (uInt)isize) != crc) { err = "corrupt BGZF member (CRC32 mismatch)"; return -1; }
        }
        coff = at;
        next = at + bsize;
        return 1;
    }
};

// A byte stream over consecutive members that knows the virtual offset of its read position.
struct VStream {
    MemberReader mr;
    size_t upos = 0;
    bool loaded = false, eof = false;
    bool seek(uint64_t voff)
    {
        const int r = mr.load(voff >> 16);
        if (r < 0) return false;
        eof = r == 0;
        loaded = true;
        upos = (size_t)(voff & 0xffff);
        return eof || upos <= mr.plain.size();
    }
    uint64_t tell() const { return (mr.coff << 16) | (uint64_t)upos; }
    // -> bytes read (short only at end of file), -1 on error
    long read(uint8_t *dst, size_t n)
    {
        size_t done = 0;
        while (done < n && !eof) {
            if (upos >= mr.plain.size()) {
                const int r = mr.load(mr.next);
                if (r < 0) return -1;
                if (r == 0) { eof = true; break; }
                upos = 0;
                continue;
            }
            const size_t take = std::min(n - done, mr.plain.size() - upos);
            memcpy(dst + done, mr.plain.data() + upos, take);
            upos += take; done += take;
        }
        return (long)done;
    }
    // a position at the very end of a member is the start of the next one (what htslib's bgzf_tell reports)
    void normalise()
    {
        while (!eof && upos >= mr.plain.size()) {
            const int r = mr.load(mr.next);
            if (r <= 0) { eof = true; break; }
            upos = 0;
        }
    }
};

int read_header_into(pb_bam *h, VStream &vs)
{
    if (!vs.seek(0) || vs.eof) { pb_set_error("%s: %s", h->path.c_str(), vs.mr.err.empty() ? "empty BAM file" : vs.mr.err.c_str()); return PB_EINVAL; }
    uint8_t w[8];
    if (vs.read(w, 8) != 8 || memcmp(w, "BAM\1", 4) != 0) { pb_set_error("%s: not a BAM file (bad magic)", h->path.c_str()); return PB_EINVAL; }
    std::vector<uint8_t> text(rd32(w + 4));
    if (vs.read(text.data(), text.size()) != (long)text.size() || vs.read(w, 4) != 4) { pb_set_error("%s: truncated BAM header", h->path.c_str()); return PB_EINVAL; }
    const uint32_t n_ref = rd32(w);
    h->ref_name.clear(); h->ref_len.clear();
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (vs.read(w, 4) != 4) { pb_set_error("%s: truncated BAM header", h->path.c_str()); return PB_EINVAL; }
        const uint32_t l_name = rd32(w);
        std::vector<uint8_t> nm(l_name + 4);
        if (l_name > (1u << 20) || vs.read(nm.data(), nm.size()) != (long)nm.size()) { pb_set_error("%s: truncated BAM header", h->path.c_str()); return PB_EINVAL; }
        h->ref_name.emplace_back((const char *)nm.data(), l_name ? l_name - 1 : 0);
        h->ref_len.push_back((int32_t)rd32(nm.data() + l_name));
    }
    return PB_OK;
}

}  // namespace

extern "C" int pb_bai_open(const char *path, pb_bai **out)
{
    if (!path || !out) { pb_set_error("pb_bai_open: null argument"); return PB_EINVAL; }
    FILE *fh = fopen(path, "rb");
    if (!fh) { pb_set_error("pb_bai_open: cannot open %s", path); return PB_EINVAL; }
    std::vector<uint8_t> buf;
    {
        uint8_t tmp[1 << 16];
        size_t got;
        while ((got = fread(tmp, 1, sizeof(tmp), fh)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    }
    fclose(fh);
    size_t p = 0;
    auto need = [&](size_t n) { return p + n <= buf.size(); };
    if (!need(8) || memcmp(buf.data(), "BAI\1", 4) != 0) { pb_set_error("pb_bai_open(%s): not a BAI index (bad magic)", path); return PB_EINVAL; }
    const int32_t n_ref = (int32_t)rd32(buf.data() + 4);
    p = 8;
    if (n_ref < 0) { pb_set_error("pb_bai_open(%s): negative reference count", path); return PB_EINVAL; }
    if ((size_t)n_ref > (buf.size() - 8) / 8) {       // every reference takes at least its two counts
        pb_set_error("pb_bai_open(%s): truncated index (%d references do not fit in %zu bytes)", path, n_ref, buf.size());
        return PB_EINVAL;
    }
    pb_bai *idx = new pb_bai();
    idx->refs.resize((size_t)n_ref);
    bool ok = true;
    for (int32_t r = 0; r < n_ref && ok; ++r) {
        pb_bai::Ref &ref = idx->refs[(size_t)r];
        if (!need(4)) { ok = false; break; }
        const int32_t n_bin = (int32_t)rd32(buf.data() + p); p += 4;
        for (int32_t b = 0; b < n_bin && ok; ++b) {
            if (!need(8)) { ok = false; break; }
            const uint32_t bin = rd32(buf.data() + p);
            const int32_t n_chunk = (int32_t)rd32(buf.data() + p + 4); p += 8;
            if (n_chunk < 0 || !need((size_t)n_chunk * 16)) { ok = false; break; }
            if (bin == 37450) {                      // metadata pseudo-bin: (ref_beg, ref_end), (n_mapped, n_unmapped)
                if (n_chunk >= 2) { ref.mapped = (int64_t)rd64(buf.data() + p + 16); ref.unmapped = (int64_t)rd64(buf.data() + p + 24); }
            } else {
                auto &v = ref.bins[bin];
                for (int32_t c = 0; c < n_chunk; ++c) v.emplace_back(rd64(buf.data() + p + 16 * (size_t)c), rd64(buf.data() + p + 16 * (size_t)c + 8));
            }
            p += (size_t)n_chunk * 16;
        }
        if (!ok || !need(4)) { ok = false; break; }
        const int32_t n_intv = (int32_t)rd32(buf.data() + p); p += 4;
        if (n_intv < 0 || !need((size_t)n_intv * 8)) { ok = false; break; }
        ref.ioffset.resize((size_t)n_intv);
        for (int32_t i = 0; i < n_intv; ++i) ref.ioffset[(size_t)i] = rd64(buf.data() + p + 8 * (size_t)i);
        p += (size_t)n_intv * 8;
    }
    if (!ok) { delete idx; pb_set_error("pb_bai_open(%s): truncated index", path); return PB_EINVAL; }
    if (need(8)) idx->no_coor = (int64_t)rd64(buf.data() + p);
    *out = idx;
    return PB_OK;
}

extern "C" void pb_bai_close(pb_bai *idx) { delete idx; }
extern "C" int pb_bai_n_ref(const pb_bai *idx) { return idx ? (int)idx->refs.size() : 0; }

// ref >= 0: mapped records of that reference; ref < 0: of the whole file.  -1 where the index carries no statistics.
extern "C" int64_t pb_bai_mapped(const pb_bai *idx, int ref)
{
    if (!idx) return -1;
    if (ref >= 0) {
        if ((size_t)ref >= idx->refs.size()) return -1;
        const pb_bai::Ref &r = idx->refs[(size_t)ref];
        return r.mapped >= 0 ? r.mapped : (r.bins.empty() ? 0 : -1);      // a reference without records has no pseudo-bin
    }
    int64_t total = 0;
    bool any = false;
    for (const auto &r : idx->refs) {
        if (r.mapped >= 0) { total += r.mapped; any = true; }
        else if (!r.bins.empty()) return -1;                               // records, but no statistics
    }
    return any || idx->refs.empty() ? total : -1;
}

extern "C" int pb_bam_read_header(pb_bam *h)
{
    if (!h) { pb_set_error("pb_bam_read_header: null handle"); return PB_EINVAL; }
    VStream vs;
    if (!vs.mr.open(h->path.c_str())) { pb_set_error("pb_bam_read_header: cannot open %s", h->path.c_str()); return PB_EINVAL; }
    return read_header_into(h, vs);
}

// The records of reference `tid` whose span [pos, end) overlaps [beg, end) — end = pos + reference bases consumed
// (one base for a record that consumes none), htslib's bam_endpos — become the handle's batch (rows of that
// reference only).  Replaces whatever the handle held.
extern "C" int pb_bam_fetch(pb_bam *h, const pb_bai *idx, int tid, int64_t beg, int64_t end)
{
    if (!h || !idx) { pb_set_error("pb_bam_fetch: null argument"); return PB_EINVAL; }
    VStream vs;
    if (!vs.mr.open(h->path.c_str())) { pb_set_error("pb_bam_fetch: cannot open %s", h->path.c_str()); return PB_EINVAL; }
    if (h->ref_name.empty()) {
        const int rc = read_header_into(h, vs);
        if (rc) return rc;
    }
    const size_t n_ref = h->ref_name.size();
    h->start.clear(); h->meta.clear(); h->blk_off.clear(); h->blk.clear();
    h->chrom_read_off.assign(n_ref + 1, 0);
    h->mapped = 0; h->skipped = 0; h->max_span = 1; h->decoded = true;
    if (tid < 0 || (size_t)tid >= n_ref || (size_t)tid >= idx->refs.size()) return PB_OK;
    if (beg < 0) beg = 0;
    if (end > h->ref_len[(size_t)tid]) end = h->ref_len[(size_t)tid];
    if (end > ((int64_t)1 << 29)) end = (int64_t)1 << 29;          // the binning scheme's reach
    if (beg >= end) return PB_OK;
    const pb_bai::Ref &ref = idx->refs[(size_t)tid];
    // candidate chunks: every bin overlapping the region (SAM spec 5.3, reg2bins), trimmed by the linear index
    uint64_t min_off = 0;
    if (!ref.ioffset.empty()) {
        const size_t w = (size_t)(beg >> 14);
        min_off = w < ref.ioffset.size() ? ref.ioffset[w] : ref.ioffset.back();
    }
    std::vector<std::pair<uint64_t, uint64_t>> chunks;
    {
        const int64_t e = end - 1;
        auto take = [&](uint32_t bin) {
            auto it = ref.bins.find(bin);
            if (it == ref.bins.end()) return;
            for (const auto &c : it->second) if (c.second > min_off) chunks.push_back(c);
        };
        take(0);
        for (int64_t k = 1 + (beg >> 26); k <= 1 + (e >> 26); ++k) take((uint32_t)k);
        for (int64_t k = 9 + (beg >> 23); k <= 9 + (e >> 23); ++k) take((uint32_t)k);
        for (int64_t k = 73 + (beg >> 20); k <= 73 + (e >> 20); ++k) take((uint32_t)k);
        for (int64_t k = 585 + (beg >> 17); k <= 585 + (e >> 17); ++k) take((uint32_t)k);
        for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (e >> 14); ++k) take((uint32_t)k);
    }
    std::sort(chunks.begin(), chunks.end());
    std::vector<std::pair<uint64_t, uint64_t>> merged;
    for (const auto &c : chunks) {
        if (!merged.empty() && c.first <= merged.back().second) merged.back().second = std::max(merged.back().second, c.second);
        else merged.push_back(c);
    }
    Decoded d;
    std::vector<uint8_t> rec;
    bool past = false;
    for (size_t ci = 0; ci < merged.size() && !past; ++ci) {
        uint64_t from = merged[ci].first;
        if (from < min_off) from = min_off;
        if (!vs.seek(from)) { pb_set_error("pb_bam_fetch(%s): %s", h->path.c_str(), vs.mr.err.c_str()); return PB_EINVAL; }
        for (;;) {
            vs.normalise();
            if (vs.eof || vs.tell() >= merged[ci].second) break;
            uint8_t w[4];
            const long got = vs.read(w, 4);
            if (got == 0) break;
            if (got != 4) { pb_set_error("pb_bam_fetch(%s): truncated BAM record", h->path.c_str()); return PB_EINVAL; }
            const uint32_t size = rd32(w);
            if (size < 32 || size > (1u << 29)) { pb_set_error("pb_bam_fetch(%s): implausible BAM record size (bad index?)", h->path.c_str()); return PB_EINVAL; }
            rec.resize(size);
            if (vs.read(rec.data(), size) != (long)size) {
                pb_set_error("pb_bam_fetch(%s): %s", h->path.c_str(), vs.mr.err.empty() ? "truncated BAM record" : vs.mr.err.c_str());
                return PB_EINVAL;
            }
            const int32_t rtid = (int32_t)rd32(rec.data()), pos = (int32_t)rd32(rec.data() + 4);
            if (rtid != tid) { if (rtid > tid || rtid < 0) { past = true; break; } continue; }
            if ((int64_t)pos >= end) { past = true; break; }
            const uint32_t l_read_name = rec[8], n_cigar = rd16(rec.data() + 12);
            if (32 + l_read_name + 4ull * n_cigar > size) { pb_set_error("pb_bam_fetch(%s): BAM record shorter than its CIGAR", h->path.c_str()); return PB_EINVAL; }
            int64_t rlen = 0;
            for (uint32_t k = 0; k < n_cigar; ++k) {
                const uint32_t c = rd32(rec.data() + 32 + l_read_name + 4 * k), op = c & 0xf;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
            }
            if ((int64_t)pos + (rlen ? rlen : 1) <= beg) continue;
            if (!convert_record(rec.data(), size, d)) { pb_set_error("pb_bam_fetch(%s): %s", h->path.c_str(), d.err.c_str()); return PB_EINVAL; }
        }
    }
    const size_t n = d.start.size();
    // rows in file order; a leading deletion can shift a start past its successor's: stable sort by start
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = i;
    bool sorted = true;
    for (size_t i = 1; i < n; ++i) if (d.start[i] < d.start[i - 1]) { sorted = false; break; }
    if (!sorted) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return d.start[a] < d.start[b]; });
    std::vector<uint64_t> blk_start(n + 1, 0);
    for (size_t i = 0; i < n; ++i) blk_start[i + 1] = blk_start[i] + d.nlisted[i];
    h->start.resize(n); h->meta.resize(n);
    const bool any_multi = !d.blk.empty();
    if (any_multi) h->blk_off.resize(n + 1);
    uint64_t run = 0;
    for (size_t i = 0; i < n; ++i) {
        const size_t j = order[i];
        h->start[i] = d.start[j]; h->meta[i] = d.meta[j];
        if (any_multi) {
            h->blk_off[i] = (uint32_t)run;
            h->blk.insert(h->blk.end(), d.blk.begin() + 2 * blk_start[j], d.blk.begin() + 2 * blk_start[j + 1]);
            run += d.nlisted[j];
        }
    }
    if (any_multi) h->blk_off[n] = (uint32_t)run;
    for (size_t c = (size_t)tid + 1; c <= n_ref; ++c) h->chrom_read_off[c] = (int64_t)n;
    h->mapped = d.mapped; h->skipped = d.skipped; h->max_span = d.max_span;
    return PB_OK;
}

// `samtools index`: writes the .bai of a coordinate-sorted BAM (SAM/BAM specification 5.2: one chunk list per bin of
// the UCSC binning scheme, a linear index over 16 kb windows, the metadata pseudo-bin with the mapped / unmapped
// counts that `bamfile.mapped` reads).  Records are walked once, one BGZF member at a time.
extern "C" int pb_bam_build_index(const char *bam_path, const char *bai_path)
{
    if (!bam_path || !bai_path) { pb_set_error("pb_bam_build_index: null argument"); return PB_EINVAL; }
    pb_bam hdr;
    hdr.path = bam_path;
    VStream vs;
    if (!vs.mr.open(bam_path)) { pb_set_error("pb_bam_build_index: cannot open %s", bam_path); return PB_EINVAL; }
    int rc = read_header_into(&hdr, vs);
    if (rc) return rc;
    const size_t n_ref = hdr.ref_name.size();
    struct RefIdx {
        std::vector<std::pair<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>>> bins;   // in order of first use
        std::unordered_map<uint32_t, size_t> where;
        std::vector<uint64_t> ioffset;
        uint64_t off_beg = 0, off_end = 0, mapped = 0, unmapped = 0;
        bool seen = false;
    };
    std::vector<RefIdx> refs(n_ref);
    uint64_t no_coor = 0;
    auto reg2bin = [](int64_t beg, int64_t end) -> uint32_t {
        --end;
        if (beg >> 14 == end >> 14) return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14));
        if (beg >> 17 == end >> 17) return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17));
        if (beg >> 20 == end >> 20) return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20));
        if (beg >> 23 == end >> 23) return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23));
        if (beg >> 26 == end >> 26) return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26));
        return 0;
    };
    std::vector<uint8_t> rec;
    int32_t last_tid = -1, last_pos = -1;
    for (;;) {
        vs.normalise();
        if (vs.eof) break;
        const uint64_t v0 = vs.tell();
        uint8_t w[4];
        const long got = vs.read(w, 4);
        if (got == 0) break;
        if (got < 0) { pb_set_error("pb_bam_build_index(%s): %s", bam_path, vs.mr.err.c_str()); return PB_EINVAL; }
        const uint32_t size = rd32(w);
        if (got != 4 || size < 32 || size > (1u << 29)) { pb_set_error("pb_bam_build_index(%s): truncated or implausible BAM record", bam_path); return PB_EINVAL; }
        rec.resize(size);
        if (vs.read(rec.data(), size) != (long)size) { pb_set_error("pb_bam_build_index(%s): truncated BAM record", bam_path); return PB_EINVAL; }
        vs.normalise();
        const uint64_t v1 = vs.eof ? ((vs.mr.next << 16)) : vs.tell();
        const int32_t tid = (int32_t)rd32(rec.data()), pos = (int32_t)rd32(rec.data() + 4);
        const uint32_t flag = rd16(rec.data() + 14);
        if (tid < 0) { ++no_coor; continue; }
        if ((size_t)tid >= n_ref) { pb_set_error("pb_bam_build_index(%s): record refers to reference %d of %zu", bam_path, tid, n_ref); return PB_EINVAL; }
        if (tid < last_tid || (tid == last_tid && pos < last_pos)) { pb_set_error("pb_bam_build_index(%s): file is not coordinate-sorted", bam_path); return PB_EINVAL; }
        last_tid = tid; last_pos = pos;
        const uint32_t l_read_name = rec[8], n_cigar = rd16(rec.data() + 12);
        if (32 + l_read_name + 4ull * n_cigar > size) { pb_set_error("pb_bam_build_index(%s): BAM record shorter than its CIGAR", bam_path); return PB_EINVAL; }
        int64_t rlen = 0;
        for (uint32_t k = 0; k < n_cigar; ++k) {
            const uint32_t c = rd32(rec.data() + 32 + l_read_name + 4 * k), op = c & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
        }
        const int64_t beg = pos < 0 ? 0 : pos, end = beg + (rlen ? rlen : 1);
        RefIdx &r = refs[(size_t)tid];
        if (!r.seen) { r.seen = true; r.off_beg = v0; }
        r.off_end = v1;
        if (flag & 0x4) ++r.unmapped; else ++r.mapped;
        const uint32_t bin = reg2bin(beg, end > ((int64_t)1 << 29) ? ((int64_t)1 << 29) : end);
        auto it = r.where.find(bin);
        if (it == r.where.end()) {
            r.where[bin] = r.bins.size();
            r.bins.push_back({bin, {{v0, v1}}});
        } else {
            auto &chunks = r.bins[it->second].second;
            if (chunks.back().second == v0) chunks.back().second = v1; else chunks.emplace_back(v0, v1);
        }
        const size_t w0 = (size_t)(beg >> 14), w1 = (size_t)((end - 1) >> 14);
        if (r.ioffset.size() <= w1) r.ioffset.resize(w1 + 1, 0);
        for (size_t k = w0; k <= w1; ++k) if (r.ioffset[k] == 0) r.ioffset[k] = v0;
    }
    FILE *out = fopen(bai_path, "wb");
    if (!out) { pb_set_error("pb_bam_build_index: cannot write %s", bai_path); return PB_EINVAL; }
    auto w32 = [&](uint32_t v) { uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; fwrite(b, 1, 4, out); };
    auto w64 = [&](uint64_t v) { w32((uint32_t)v); w32((uint32_t)(v >> 32)); };
    fwrite("BAI\1", 1, 4, out);
    w32((uint32_t)n_ref);
    for (RefIdx &r : refs) {
        w32((uint32_t)(r.bins.size() + (r.seen ? 1 : 0)));
        for (const auto &b : r.bins) {
            w32(b.first);
            w32((uint32_t)b.second.size());
            for (const auto &c : b.second) { w64(c.first); w64(c.second); }
        }
        if (r.seen) { w32(37450); w32(2); w64(r.off_beg); w64(r.off_end); w64(r.mapped); w64(r.unmapped); }
        for (size_t k = 1; k < r.ioffset.size(); ++k) if (r.ioffset[k] == 0) r.ioffset[k] = r.ioffset[k - 1];   // windows no record starts in
        w32((uint32_t)r.ioffset.size());
        for (uint64_t v : r.ioffset) w64(v);
    }
    w64(no_coor);
    const bool bad = ferror(out) != 0;
    if (fclose(out) != 0 || bad) { pb_set_error("pb_bam_build_index: write to %s failed", bai_path); return PB_EINVAL; }
    return PB_OK;
}
