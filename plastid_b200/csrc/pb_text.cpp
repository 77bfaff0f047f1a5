// pb_text.cpp — text of a wiggle / bedGraph track from the records pb_export_runs leaves (host side, no CUDA).
//
// The reference writes one Python-formatted line per non-zero position or run:
// `"%s\t%s\n" % (genomic_x + 1, val)` (plastid/genomics/genome_array.py:1030-1037, to_variable_step) and
// `"%s\t%s\t%s\t%s\n" % (chrom, start, end, val)` (:1100-1111, to_bedgraph), val being a numpy scalar.  With the
// run-length compaction done on the device, that formatting loop is what a whole-genome export waits for
// (0.35-1.6 us per line in CPython); here it is one native pass, split over threads.  Numbers come out exactly
// as `str()` of a Python int / numpy float64 gives them: shortest round-trip digits (std::to_chars), fixed
// notation for 1e-4 <= |x| < 1e16 with ".0" added to integers, otherwise d.ddde+XX with at least two exponent digits.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "plastid_b200.h"

void pb_set_error(const char *fmt, ...);

namespace {

inline char *put_int(char *p, int64_t v)
{
    return std::to_chars(p, p + 24, v).ptr;
}

// repr(float) of CPython / str(numpy.float64): Python/pystrtod.c format_float_short, mode 'r' with ADD_DOT_0
char *put_float(char *p, double x)
{
    if (std::isnan(x)) { memcpy(p, "nan", 3); return p + 3; }
    if (std::signbit(x)) { *p++ = '-'; x = -x; }
    if (std::isinf(x)) { memcpy(p, "inf", 3); return p + 3; }
    if (x == 0.0) { memcpy(p, "0.0", 3); return p + 3; }
    char sci[40];
    char *end = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific).ptr;   // d[.ddd]e[+-]XX, shortest
    char digits[24];
    int nd = 0;
    const char *q = sci;
    for (; q < end && *q != 'e'; ++q)
        if (*q != '.') digits[nd++] = *q;
    int exp10 = 0;
    std::from_chars(q + 1 + (q[1] == '+'), end, exp10);
    const int decpt = exp10 + 1;                        // digits are 0.DDDD x 10^decpt
    if (decpt <= -4 || decpt > 16) {                    // exponent form
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        int e = decpt - 1;
        *p++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        if (e < 10) *p++ = '0';
        return std::to_chars(p, p + 8, e).ptr;
    }
    if (decpt <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int k = 0; k < -decpt; ++k) *p++ = '0';
        memcpy(p, digits, nd);
        return p + nd;
    }
    if (decpt >= nd) {
        memcpy(p, digits, nd); p += nd;
        for (int k = nd; k < decpt; ++k) *p++ = '0';
        *p++ = '.'; *p++ = '0';
        return p;
    }
    memcpy(p, digits, decpt); p += decpt;
    *p++ = '.';
    memcpy(p, digits + decpt, nd - decpt);
    return p + (nd - decpt);
}

}  // namespace

extern "C" int64_t pb_format_track_bound(int kind, const char *chrom, int64_t n)
{
    const int64_t per_line = (kind == 1 ? (int64_t)strlen(chrom ? chrom : "") + 1 + 21 + 21 : 21) + 26 + 1;
    return n * per_line + 1;
}

extern "C" int64_t pb_format_track(int kind, const char *chrom, const int64_t *start, const int64_t *end, const void *values,
                                   int values_are_float, int64_t n, char *out, int64_t cap, int n_threads)
{
    if (n < 0 || (kind != 0 && kind != 1) || (n && (!start || !values || !out)) || (kind == 1 && n && (!end || !chrom))) {
        pb_set_error("pb_format_track: bad argument");
        return -1;
    }
    if (cap < pb_format_track_bound(kind, chrom, n)) {
        pb_set_error("pb_format_track: buffer of %lld bytes, %lld needed (pb_format_track_bound)", (long long)cap,
                     (long long)pb_format_track_bound(kind, chrom, n));
        return -1;
    }
    if (n == 0) return 0;
    const size_t chrom_len = kind == 1 ? strlen(chrom) : 0;
    const int64_t per_line = pb_format_track_bound(kind, chrom, 1) - 1;
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int nt = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n / 65536));
    std::vector<int64_t> written(nt, 0);
    auto work = [&](int t) {                            // thread t formats its share at the worst-case position
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        char *p = out + a * per_line;
        const int64_t *vi = (const int64_t *)values;
        const double *vf = (const double *)values;
        for (int64_t i = a; i < b; ++i) {
            if (kind == 1) {
                memcpy(p, chrom, chrom_len); p += chrom_len;
                *p++ = '\t';
                p = put_int(p, start[i]);
                *p++ = '\t';
                p = put_int(p, end[i]);
            } else {
                p = put_int(p, start[i] + 1);           // wiggle positions are 1-based
            }
            *p++ = '\t';
            p = values_are_float ? put_float(p, vf[i]) : put_int(p, vi[i]);
            *p++ = '\n';
        }
        written[t] = p - (out + a * per_line);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    int64_t total = written[0];                         // close the gaps between the threads' pieces
    for (int t = 1; t < nt; ++t) {
        memmove(out + total, out + (n * t / nt) * per_line, (size_t)written[t]);
        total += written[t];
    }
    return total;
}
