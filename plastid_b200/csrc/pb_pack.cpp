// pb_pack.cpp — host encoder of the delta3 transfer format (no CUDA): what the decoder side hands to
// PCIe.  Format: include/plastid_b200.h (pb_unpack_delta3).  Multithreaded over groups of 128-read
// blocks: (1) meta-word histogram -> 31-entry dictionary, (2) per-block counts of wide deltas and
// exceptions, (3) prefix sums, (4) fill.  Produces exactly the streams plastid_b200/batch.py's numpy
// encoder (Delta3Batch.from_batch) produces; tests compare them byte for byte.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

#include "plastid_b200.h"

void pb_set_error(const char *fmt, ...);

namespace {

template <typename F> void pack_parallel_for(int n_threads, size_t n, F fn)
{
    if (n_threads <= 1 || n < 2) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)n_threads, n);
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (auto &th : pool) th.join();
}

constexpr int64_t kBlock = 128;        // reads per block
constexpr int64_t kGroup = 4096;       // blocks per parallel work item

}  // namespace

extern "C" int pb_pack_delta3(const int32_t *ref_start, const uint32_t *meta, const int64_t *chrom_read_off,
                              int32_t n_chrom, int64_t n_reads, int n_threads,
                              uint8_t *packed, uint8_t *wide, int32_t *blk_base, uint32_t *blk_wide_off,
                              uint32_t *blk_exc_off, int32_t *exc_start, uint32_t *exc_meta, uint32_t *dict32,
                              int64_t *n_wide_out, int64_t *n_exc_out)
{
    if (n_reads < 0 || n_chrom < 0 || !chrom_read_off || !packed || !wide || !blk_base || !blk_wide_off || !blk_exc_off ||
        !exc_start || !exc_meta || !dict32 || !n_wide_out || !n_exc_out || (n_reads > 0 && (!ref_start || !meta))) {
        pb_set_error("pb_pack_delta3: null argument"); return PB_EINVAL;
    }
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t n_blk = (n_reads + kBlock - 1) / kBlock;
    const int64_t n_grp = (n_blk + kGroup - 1) / kGroup;

    // (1) dictionary: the 31 most frequent meta words (ties: smaller word first), stored sorted
    std::vector<std::unordered_map<uint32_t, int64_t>> part((size_t)n_grp);
    pack_parallel_for(n_threads, (size_t)n_grp, [&](size_t g) {
        const int64_t a = (int64_t)g * kGroup * kBlock, e = std::min(n_reads, a + kGroup * kBlock);
        auto &h = part[g];
        // a batch has a few dozen distinct words: a 256-slot open-addressing table in front of the map
        uint32_t key[256]; int64_t cnt[256]; bool used[256];
        memset(used, 0, sizeof(used));
        for (int64_t i = a; i < e; ++i) {
            const uint32_t m = meta[i];
            uint32_t s0 = (m * 2654435761u) >> 24;
            int probes = 0;
            while (used[s0] && key[s0] != m && probes < 8) { s0 = (s0 + 1) & 255u; ++probes; }
            if (used[s0] && key[s0] == m) { cnt[s0]++; continue; }
            if (!used[s0]) { used[s0] = true; key[s0] = m; cnt[s0] = 1; continue; }
            h[m] += 1;                                        // table neighbourhood full: straight to the map
        }
        for (int k = 0; k < 256; ++k) if (used[k]) h[key[k]] += cnt[k];
    });
    std::unordered_map<uint32_t, int64_t> all;
    for (auto &h : part) for (auto &kv : h) all[kv.first] += kv.second;
    std::vector<std::pair<uint32_t, int64_t>> words(all.begin(), all.end());
    std::sort(words.begin(), words.end(), [](const std::pair<uint32_t, int64_t> &x, const std::pair<uint32_t, int64_t> &y) {
        return x.second != y.second ? x.second > y.second : x.first < y.first;
    });
    uint32_t dict[32];
    const int n_dict = (int)std::min<size_t>(31, words.size());
    for (int k = 0; k < n_dict; ++k) dict[k] = words[k].first;
    std::sort(dict, dict + n_dict);
    memset(dict32, 0, 32 * sizeof(uint32_t));
    memcpy(dict32, dict, (size_t)n_dict * sizeof(uint32_t));

    // first read of every chromosome (reads of chromosome c are [off[c], off[c+1]))
    auto is_first_of_chrom = [&](int64_t i, int &c) {      // c: cursor, only ever moves forward within a work item
        while (c + 1 < n_chrom && i >= chrom_read_off[c + 1]) ++c;
        return i == chrom_read_off[c] || i == 0;
    };
    auto chrom_cursor = [&](int64_t i) {
        int lo = 0, hi = n_chrom;                          // last c with off[c] <= i
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (chrom_read_off[mid] <= i) lo = mid; else hi = mid; }
        return lo;
    };
    // classification of read i (shared by the counting and the filling pass)
    // direct-mapped table of dictionary codes for the low 18 bits (aligned length, strand, drop flag) + a
    // check of the full word: almost every lookup is one load
    auto code_of = [&](uint32_t m) -> int {
        for (int k = 0; k < n_dict; ++k) if (dict[k] == m) return k;      // only reached for cache misses
        return -1;
    };
    constexpr uint32_t kLow = (1u << 18) - 1u;
    std::vector<int16_t> low_code(kLow + 1, -1);
    std::vector<uint8_t> low_unique(kLow + 1, 0);        // 1: exactly one dictionary word has these low bits
    {
        std::vector<int> cnt(kLow + 1, 0);
        for (int k = 0; k < n_dict; ++k) cnt[dict[k] & kLow]++;
        for (int k = 0; k < n_dict; ++k)
            if (cnt[dict[k] & kLow] == 1) { low_unique[dict[k] & kLow] = 1; low_code[dict[k] & kLow] = (int16_t)k; }
    }
    auto classify = [&](int64_t i, int &c, int64_t &delta, int &code, bool &exc, bool &widef) {
        const bool first_blk = (i % kBlock) == 0;
        delta = (i == 0 || first_blk) ? 0 : (int64_t)ref_start[i] - (int64_t)ref_start[i - 1];
        const bool first_chr = is_first_of_chrom(i, c);
        const uint32_t m = meta[i];
        const uint32_t lowbits = m & kLow;
        if (low_unique[lowbits]) code = (dict[low_code[lowbits]] == m) ? low_code[lowbits] : -1;
        else code = code_of(m);
        exc = delta > 7 + 254 || delta < 0 || (first_chr && !first_blk) || code < 0;
        widef = exc || delta >= 7;
    };

    // (2) per-block counts
    std::vector<uint32_t> n_w((size_t)n_blk + 1, 0), n_e((size_t)n_blk + 1, 0);
    pack_parallel_for(n_threads, (size_t)n_grp, [&](size_t g) {
        const int64_t b0 = (int64_t)g * kGroup, b1 = std::min(n_blk, b0 + kGroup);
        int c = n_chrom > 0 ? chrom_cursor(b0 * kBlock) : 0;
        for (int64_t B = b0; B < b1; ++B) {
            uint32_t w = 0, x = 0;
            const int64_t e = std::min(n_reads, (B + 1) * kBlock);
            for (int64_t i = B * kBlock; i < e; ++i) {
                int64_t delta; int code; bool exc, widef;
                classify(i, c, delta, code, exc, widef);
                w += widef; x += exc;
            }
            n_w[B] = w; n_e[B] = x;
        }
    });
    // (3) exclusive prefix sums
    uint64_t rw = 0, re = 0;
    for (int64_t B = 0; B < n_blk; ++B) {
        blk_wide_off[B] = (uint32_t)rw; blk_exc_off[B] = (uint32_t)re;
        rw += n_w[B]; re += n_e[B];
    }
    if (rw > 0xffffffffull || re > 0xffffffffull) { pb_set_error("pb_pack_delta3: too many escapes"); return PB_EINVAL; }
    blk_wide_off[n_blk] = (uint32_t)rw; blk_exc_off[n_blk] = (uint32_t)re;
    *n_wide_out = (int64_t)rw; *n_exc_out = (int64_t)re;

    // (4) fill
    pack_parallel_for(n_threads, (size_t)n_grp, [&](size_t g) {
        const int64_t b0 = (int64_t)g * kGroup, b1 = std::min(n_blk, b0 + kGroup);
        int c = n_chrom > 0 ? chrom_cursor(b0 * kBlock) : 0;
        for (int64_t B = b0; B < b1; ++B) {
            uint64_t wi = blk_wide_off[B], ei = blk_exc_off[B];
            blk_base[B] = ref_start[B * kBlock];
            const int64_t e = std::min(n_reads, (B + 1) * kBlock);
            for (int64_t i = B * kBlock; i < e; ++i) {
                int64_t delta; int code; bool exc, widef;
                classify(i, c, delta, code, exc, widef);
                packed[i] = (uint8_t)((widef ? 7 : (int)delta) | ((exc ? 31 : code) << 3));
                if (widef) wide[wi++] = exc ? 255 : (uint8_t)(delta - 7);
                if (exc) { exc_start[ei] = ref_start[i]; exc_meta[ei] = meta[i]; ++ei; }
            }
            for (int64_t i = e; i < (B + 1) * kBlock; ++i) packed[i] = 0;      // padding of the last block
        }
    });
    return PB_OK;
}

// Reads per aligned length (meta bits 0-15) of a host meta array, reads with the drop bit (17) left out: the batch
// metadata the Center rule derives its tables of map lengths from (len(read.positions) bucketing, psite.py:187-188).
// One 65536-bin histogram per thread, added up at the end; numpy needed 0.26 s for 20 M reads (mask, gather, astype,
// bincount), 80 % of packing a batch.
extern "C" int pb_meta_length_hist(const uint32_t *meta, int64_t n_reads, int n_threads, int64_t *hist)
{
    if (n_reads < 0 || !hist || (n_reads > 0 && !meta)) { pb_set_error("pb_meta_length_hist: null argument"); return PB_EINVAL; }
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    constexpr int64_t kPiece = 1 << 20;
    const int64_t n_piece = (n_reads + kPiece - 1) / kPiece;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_piece));
    std::vector<std::vector<int64_t>> part((size_t)nt, std::vector<int64_t>(65536, 0));
    std::atomic<int64_t> next{0};
    auto work = [&](int t) {
        int64_t *h = part[(size_t)t].data();
        for (int64_t p; (p = next.fetch_add(1)) < n_piece;) {
            const int64_t a = p * kPiece, e = std::min(n_reads, a + kPiece);
            for (int64_t i = a; i < e; ++i) {
                const uint32_t m = meta[i];
                h[m & 0xFFFFu] += ((m >> 17) & 1u) ^ 1u;
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    for (int b = 0; b < 65536; ++b) {
        int64_t v = 0;
        for (int t = 0; t < nt; ++t) v += part[(size_t)t][(size_t)b];
        hist[b] = v;
    }
    return PB_OK;
}

