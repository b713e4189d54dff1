// aes_json.cpp -- host-side serialisation of the file route's reply (reference
// src/audioblocks/engine.py:107-123).
//
// The reference answers a `process_file` request with one JSON message that carries both signals
// as float lists: `mono.flatten().tolist()` and `processed.mean(axis=1).flatten().tolist()` pushed
// through `json.dumps`.  For the shipped 18.6 s clip that is 1.8 M Python float objects and
// 2.05 s of the 2.3 s request -- 15x the DSP.  These entry points write the very same
// text straight from the float32 buffers, in parallel:
//   * every value is widened to double (what `.tolist()` does) and printed with the shortest
//     digit string that round-trips (what `float.__repr__` does; std::to_chars is specified the
//     same way), laid out by CPython's rules: fixed notation while -4 < decimal exponent <= 16,
//     otherwise d.ddde+XX with at least two exponent digits; `NaN`, `Infinity`, `-Infinity` as
//     json.dumps(allow_nan=True) spells them;
//   * elements are separated by ", " inside "[" ... "]", json.dumps' default separators;
//   * the stereo variant first takes the float32 mean of each frame exactly as numpy does for a
//     (N, 2) float32 array: (l + r) rounded to float32, then halved.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/aesim.h"
#include "aes_common.h"

namespace {

constexpr int kMaxChars = 26;       // "-1.2345678901234567e-308" is 24; ", " adds 2

// repr(float(v)) into p; returns the number of characters written
inline int py_float_repr(double v, char *p)
{
    char *const p0 = p;
    if (std::isnan(v)) { memcpy(p, "NaN", 3); return 3; }
    if (std::isinf(v)) {
        if (v < 0) { memcpy(p, "-Infinity", 9); return 9; }
        memcpy(p, "Infinity", 8);
        return 8;
    }
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (v == 0.0) { memcpy(p, "0.0", 3); return (int)(p - p0) + 3; }
    char sci[40];
    const auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);
    // sci = d[.ddd]e(+|-)XX : collect the digits and the exponent
    char dig[24];
    int nd = 0;
    const char *q = sci;
    while (q < r.ptr && *q != 'e') {
        if (*q != '.') dig[nd++] = *q;
        ++q;
    }
    ++q;                                                // 'e'
    const bool eneg = *q == '-';
    ++q;
    int ex = 0;
    while (q < r.ptr) ex = ex * 10 + (*q++ - '0');
    if (eneg) ex = -ex;
    const int decpt = ex + 1;                           // value = 0.d1d2... * 10^decpt
    if (decpt > -4 && decpt <= 16) {                    // fixed notation (CPython format_float_short, 'r')
        if (decpt <= 0) {
            *p++ = '0'; *p++ = '.';
            for (int i = 0; i < -decpt; ++i) *p++ = '0';
            memcpy(p, dig, nd); p += nd;
        } else if (decpt < nd) {
            memcpy(p, dig, decpt); p += decpt;
            *p++ = '.';
            memcpy(p, dig + decpt, nd - decpt); p += nd - decpt;
        } else {
            memcpy(p, dig, nd); p += nd;
            for (int i = nd; i < decpt; ++i) *p++ = '0';
            *p++ = '.'; *p++ = '0';
        }
    } else {
        *p++ = dig[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, dig + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        int e = decpt - 1;
        if (e < 0) { *p++ = '-'; e = -e; } else *p++ = '+';
        if (e >= 100) { *p++ = (char)('0' + e / 100); e %= 100; *p++ = (char)('0' + e / 10); *p++ = (char)('0' + e % 10); }
        else { *p++ = (char)('0' + e / 10); *p++ = (char)('0' + e % 10); }
    }
    return (int)(p - p0);
}

// Formats values [lo, hi) of the (possibly frame-averaged) signal, ", "-separated, into buf
template <bool STEREO_MEAN>
int64_t format_range(const float *x, int64_t lo, int64_t hi, char *buf)
{
    char *p = buf;
    for (int64_t i = lo; i < hi; ++i) {
        float v;
        if (STEREO_MEAN) {
            volatile float s = x[2 * i] + x[2 * i + 1];     // rounded to float32 before halving, as numpy's float32 add.reduce
            v = s * 0.5f;
        } else {
            v = x[i];
        }
        if (i != lo) { *p++ = ','; *p++ = ' '; }
        p += py_float_repr((double)v, p);
    }
    return p - buf;
}

template <bool STEREO_MEAN>
int64_t json_list(const float *x, int64_t n, char *out, int64_t cap, int threads)
{
    AES_REQUIRE(n >= 0 && (n == 0 || x != nullptr) && out != nullptr, "json list: bad arguments");
    AES_REQUIRE(cap >= 2 + n * kMaxChars, "json list: output buffer smaller than aes_json_float_list_bound(n)");
    if (n == 0) { out[0] = '['; out[1] = ']'; return 2; }
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((int64_t)T > (n + 4095) / 4096) T = (int)((n + 4095) / 4096);
    // Each worker formats its slice in place at the slice's worst-case position, then the slices
    // are closed up front to back (the gaps only ever move text towards the start).
    std::vector<int64_t> len(T), lo(T + 1);
    for (int t = 0; t <= T; ++t) lo[t] = n * t / T;
    auto work = [&](int t) { len[t] = format_range<STEREO_MEAN>(x, lo[t], lo[t + 1], out + 1 + lo[t] * kMaxChars); };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    char *p = out;
    *p++ = '[';
    p += len[0];
    for (int t = 1; t < T; ++t) {
        *p++ = ','; *p++ = ' ';
        memmove(p, out + 1 + lo[t] * kMaxChars, (size_t)len[t]);
        p += len[t];
    }
    *p++ = ']';
    return p - out;
}

}  // namespace

AES_EXPORT int64_t aes_json_float_list_bound(int64_t n_values) { return 2 + (n_values > 0 ? n_values : 0) * kMaxChars; }

AES_EXPORT int64_t aes_json_float_list(const float *x, int64_t n_values, char *out, int64_t cap, int threads)
{
    return json_list<false>(x, n_values, out, cap, threads);
}

AES_EXPORT int64_t aes_json_stereo_mean_list(const float *xy, int64_t n_frames, char *out, int64_t cap, int threads)
{
    return json_list<true>(xy, n_frames, out, cap, threads);
}
