// aes_convreverb.cuh -- IR-convolution reverb (BASELINE configs[3]); device code.
//
// The reference has no convolution reverb (its reverb is the Schroeder network,
// reverb.py:72-277): this is a NEW operator in the style of the other blocks,
//     out[:, c] = clip(mix_dry * x[:, c] + mix_wet * (x[:, c] * h[:, c]), -1, 1),
// with `*` the causal linear convolution with a stereo impulse response h (n_taps, 2).
// Parity is checked against a float64 numpy restatement in the test tree (parity unpinned by the
// reference).
//
// Uniformly partitioned overlap-save in the frequency domain, blocks of BK = N/2 frames, FFTs of
// N = R*R*R/2 points (R = 32: 16384, R = 16: 2048, R = 8: 256):
//   K1  fwd   : z = xL + i*xR -- an interleaved stereo frame IS a complex number -- so ONE complex
//               FFT per block carries both channels; the two real spectra are separated before the
//               store:  S[k] = XL[k], S[N-k] = XR[k] (0 < k < N/2), natural bin order; the four real
//               values XL[0], XL[N/2], XR[0], XR[N/2] go to S[0], S[N/2], S[N], S[N+1] (imaginary part
//               0), so every bin of the product is one ordinary complex multiplication;
//               rows of NS = N + 32 entries, [clip][block][NS]
//   K2  mac   : per bin a FIR over the block index, W_j = sum_p H_p * S_{j-p}, with the IR
//               partitions laid out like S -- ONE complex multiply-add per bin, partition and block
//               (the packed form A_p Z + B_p conj(Z[N-k]) of the first revision needed two and read
//               every spectrum twice).  A thread owns one bin: the PC partition values of its bin sit
//               in registers, PC accumulators rotate over the output blocks, so every spectrum is
//               read once and every product spectrum written once.  The IR tiles ([PC][256 bins]) are
//               staged by TMA (cp.async.bulk + mbarrier), double-buffered across the work items of a
//               persistent CTA.
//   K3  inv   : Y[k] = YL + i*YR rebuilt in registers, inverse FFT, keep the last BK samples
//               (overlap-save), real = yL, imag = yR, dry/wet mix and clip, written as stereo frames.
//
// The FFT: N = R1*R2*R3 with R1 = R2 = R, R3 = H = R/2, n = n1*R*H + n2*H + n3, k = k1 + R*k2 + R*R*k3,
// three passes of register-resident radix-R / radix-H DFTs, NT = R*H threads (512 for 16384 points):
//   forward   pass 1  thread (n2,n3): DFT over n1, input straight from global memory, times W_N^(t*k1)
//             pass 2  thread (k1,n3): DFT over n2, times W_N^(R*n3*k2)
//             pass 3  thread kl = k1+R*k2 AND its mirror R*R-kl: two DFTs over n3; a bin k and its
//                     mirror N-k then sit in the same thread, the L/R separation needs no exchange,
//                     results go straight to global memory
//   inverse   the same three passes backwards (pass A from global memory, pass C to global memory).
// Shared memory is touched twice per transform (two writes + two reads of the 128 KB frame) instead of
// twelve times in the radix-8 version; both hand-off layouts are padded so that every access of a
// half-warp covers 32 distinct banks (layout 2: row per k1, one pad every 16; layout 1: row per kl of
// H+1 entries).
#pragma once
#include <type_traits>
#include "aes_plan.h"
#include "aes_fast_kernel.cuh"          // cp.async.bulk / mbarrier helpers

#define AESC_MT 256         // bins per MAC tile = threads of the MAC kernel

struct alignas(8) cpx { float x, y; };      // one 64-bit load / store per element

__device__ __forceinline__ cpx c_mul(cpx a, cpx b) { cpx r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
// complex add / subtract as ONE packed instruction (sm_100 FADD2; each half rounds like the scalar add)
__device__ __forceinline__ cpx c_add(cpx a, cpx b) { const float2 t = aes_add2(make_float2(a.x, a.y), make_float2(b.x, b.y)); cpx r; r.x = t.x; r.y = t.y; return r; }
__device__ __forceinline__ cpx c_sub(cpx a, cpx b) { const float2 t = aes_add2(make_float2(a.x, a.y), make_float2(-b.x, -b.y)); cpx r; r.x = t.x; r.y = t.y; return r; }
__device__ __forceinline__ cpx c_sq(cpx a) { cpx r; r.x = a.x * a.x - a.y * a.y; r.y = 2.0f * a.x * a.y; return r; }
__device__ __forceinline__ cpx c_conj(cpx a) { cpx r; r.x = a.x; r.y = -a.y; return r; }
__device__ __forceinline__ cpx c_mul_mi(cpx a) { cpx r; r.x = a.y; r.y = -a.x; return r; }     // a * (-i)
__device__ __forceinline__ cpx c_mul_pi(cpx a) { cpx r; r.x = -a.y; r.y = a.x; return r; }     // a * (+i)

// compile-time loop: f(std::integral_constant<int, I>) for I = 0 .. N-1
template <int I, int N, class F>
__device__ __forceinline__ void aesc_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        aesc_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr int aesc_log2(int n) { return n <= 1 ? 0 : 1 + aesc_log2(n / 2); }
__host__ __device__ constexpr int aesc_brev(int k, int bits) { return bits == 0 ? 0 : ((k & 1) << (bits - 1)) | aesc_brev(k >> 1, bits - 1); }
// cos(2 pi e / 32), 0 <= e <= 16
__host__ __device__ constexpr float aesc_c32(int e)
{
    constexpr float C[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                            0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f, 0.0f};
    return e <= 8 ? C[e] : -C[16 - e];
}

// d * exp(-2 pi i E / 32) (INV: the conjugate), 0 <= E < 16; every constant is an immediate
template <int E, bool INV>
__device__ __forceinline__ cpx aesc_mul_w32(cpx d)
{
    if constexpr (E == 0) return d;
    else if constexpr (E == 8) return INV ? c_mul_pi(d) : c_mul_mi(d);
    else {
        constexpr float c = aesc_c32(E), s0 = aesc_c32(E <= 8 ? 8 - E : E - 8);      // cos, sin of 2 pi E / 32
        constexpr float s = INV ? s0 : -s0;
        cpx r;
        r.x = d.x * c - d.y * s;
        r.y = d.x * s + d.y * c;
        return r;
    }
}

// R-point DFT in registers (R <= 32, a power of two): radix-2 decimation in frequency, natural order
// in, element k of the result at v[aesc_brev(k)]
template <int R, int HALF, bool INV>
__device__ __forceinline__ void aesc_dft(cpx (&v)[R])
{
    aesc_for<0, R / 2>([&](auto ic) {
        constexpr int b = decltype(ic)::value, j = b % HALF, g = (b / HALF) * 2 * HALF, e = j * (16 / HALF);
        const cpx p = v[g + j], q = v[g + j + HALF];
        v[g + j] = c_add(p, q);
        v[g + j + HALF] = aesc_mul_w32<e, INV>(c_sub(p, q));
    });
    if constexpr (HALF > 1) aesc_dft<R, HALF / 2, INV>(v);
}

// f(k, X[k] * w^k) for k = 0 .. R-1 in natural order; powers of w as four interleaved chains over w^4
// (at most R/4 + 2 roundings deep)
template <int R, class F>
__device__ __forceinline__ void aesc_emit_tw(const cpx (&v)[R], cpx w, F &&f)
{
    constexpr int LG = aesc_log2(R);
    const cpx w2 = c_sq(w), w3 = c_mul(w2, w), w4 = c_sq(w2);
    cpx pw[4];
    pw[0] = w4; pw[1] = w; pw[2] = w2; pw[3] = w3;
    aesc_for<0, R>([&](auto ic) {
        constexpr int k = decltype(ic)::value, src = aesc_brev(k, LG);
        if constexpr (k == 0) f(ic, v[src]);
        else {
            if constexpr (k > 4) pw[k & 3] = c_mul(pw[k & 3], w4);
            f(ic, c_mul(v[src], pw[k & 3]));
        }
    });
}

// exp(-2 pi i q / N): from the table, or computed when the caller has none (tw == nullptr)
__device__ __forceinline__ cpx aesc_tw(const cpx *__restrict__ tw, int q, int N)
{
    if (tw) return tw[q];
    cpx w;
#ifdef AES_CPU_EMU
    const double ang = -2.0 * M_PI * (double)q / (double)N;
    w.x = (float)cos(ang); w.y = (float)sin(ang);
#else
    sincospif(-2.0f * (float)q / (float)N, &w.y, &w.x);       // the argument is exact in f32
#endif
    return w;
}

template <int R> struct AescGeo {
    static constexpr int H = R / 2, N = R * R * H, NT = R * H, BK = N / 2, R2 = R * R;
    static constexpr int NS = N + 32;                                       // row of separated spectra: N bins + XR[0], XR[N/2] + pad
    static constexpr int S2 = NT + NT / 16;                                 // layout 2: one row per k1, (n2,n3) padded one per 16
    static constexpr int L1 = R2 * (H + 1), L2 = R * S2;
    static constexpr int SMEM_CPX = L1 > L2 ? L1 : L2;
};
__device__ __forceinline__ int aesc_pad16(int i) { return i + (i >> 4); }

// L/R separation of a mirror pair: lo = Z[k], hi = Z[N-k] (0 < k < N/2) -> XL[k], XR[k]
__device__ __forceinline__ void aesc_split(cpx lo, cpx hi, float sc, cpx &xl, cpx &xr)
{
    xl.x = sc * (lo.x + hi.x); xl.y = sc * (lo.y - hi.y);
    xr.x = sc * (lo.y + hi.y); xr.y = sc * (hi.x - lo.x);
}
// ... and back: YL, YR -> Y[k] = YL + i YR, Y[N-k] = conj(YL) + i conj(YR)
__device__ __forceinline__ void aesc_join(cpx yl, cpx yr, cpx &lo, cpx &hi)
{
    lo.x = yl.x - yr.y; lo.y = yl.y + yr.x;
    hi.x = yl.x + yr.y; hi.y = yr.x - yl.y;
}

// Forward transform of one frame: ld(n) -> z[n] (natural order), st(g, value) receives the separated
// spectra S[g] (scaled by 2*sc: pass sc = 0.5 for plain spectra).  All NT threads call it.
template <int R, class LD, class ST>
__device__ __forceinline__ void aesc_fwd(cpx *s, const cpx *__restrict__ tw, int t, float sc, LD &&ld, ST &&st)
{
    using G = AescGeo<R>;
    constexpr int H = G::H, LGH = aesc_log2(H);
    {   // pass 1: over n1 (stride NT); thread t = n2*H + n3
        cpx v[R];
        aesc_for<0, R>([&](auto ic) { v[decltype(ic)::value] = ld(decltype(ic)::value * G::NT + t); });
        aesc_dft<R, R / 2, false>(v);
        const int col = aesc_pad16(t);
        aesc_emit_tw<R>(v, aesc_tw(tw, t, G::N), [&](auto kc, cpx val) { s[decltype(kc)::value * G::S2 + col] = val; });
    }
    __syncthreads();
    {   // pass 2: over n2 (stride H); thread = (k1, n3)
        const int k1 = t / H, n3 = t % H;
        cpx v[R];
        aesc_for<0, R>([&](auto ic) { v[decltype(ic)::value] = s[k1 * G::S2 + aesc_pad16(decltype(ic)::value * H + n3)]; });
        __syncthreads();                                    // the layout changes: every read before any write
        aesc_dft<R, R / 2, false>(v);
        aesc_emit_tw<R>(v, aesc_tw(tw, R * n3, G::N), [&](auto kc, cpx val) { s[(k1 + R * decltype(kc)::value) * (H + 1) + n3] = val; });
    }
    __syncthreads();
    {   // pass 3: over n3; rows kl = t and its mirror
        const int kla = t, klb = t ? G::R2 - t : G::R2 / 2;
        cpx a[H], b[H];
        aesc_for<0, H>([&](auto ic) {
            a[decltype(ic)::value] = s[kla * (H + 1) + decltype(ic)::value];
            b[decltype(ic)::value] = s[klb * (H + 1) + decltype(ic)::value];
        });
        aesc_dft<H, H / 2, false>(a);
        aesc_dft<H, H / 2, false>(b);
        if (t != 0) {
            // bin k = t + R2*j of row a pairs with bin N-k = (R2-t) + R2*(H-1-j) of row b
            aesc_for<0, H>([&](auto jc) {
                constexpr int j = decltype(jc)::value, jm = H - 1 - j;
                const cpx za = a[aesc_brev(j, LGH)], zb = b[aesc_brev(jm, LGH)];
                cpx xl, xr;
                if constexpr (j < H / 2) { aesc_split(za, zb, sc, xl, xr); st(j * G::R2 + kla, xl); st(jm * G::R2 + klb, xr); }
                else { aesc_split(zb, za, sc, xl, xr); st(jm * G::R2 + klb, xl); st(j * G::R2 + kla, xr); }
            });
        } else {
            // row 0: bins R2*j, mirror R2*(H-j); DC and Nyquist are real in both channels.  Row R2/2: bins
            // R2/2 + R2*j, mirror R2/2 + R2*(H-1-j)
            const cpx dc = a[0], ny = a[aesc_brev(H / 2, LGH)];
            cpx r;
            r.y = 0.f;
            r.x = 2.0f * sc * dc.x; st(0, r);
            r.x = 2.0f * sc * dc.y; st(G::N, r);
            r.x = 2.0f * sc * ny.x; st(G::N / 2, r);
            r.x = 2.0f * sc * ny.y; st(G::N + 1, r);
            aesc_for<1, H / 2>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                cpx xl, xr;
                aesc_split(a[aesc_brev(j, LGH)], a[aesc_brev(H - j, LGH)], sc, xl, xr);
                st(j * G::R2, xl); st((H - j) * G::R2, xr);
            });
            aesc_for<0, H / 2>([&](auto jc) {
                constexpr int j = decltype(jc)::value, jm = H - 1 - j;
                cpx xl, xr;
                aesc_split(b[aesc_brev(j, LGH)], b[aesc_brev(jm, LGH)], sc, xl, xr);
                st(j * G::R2 + klb, xl); st(jm * G::R2 + klb, xr);
            });
        }
    }
}

// Inverse transform (unnormalised) of separated spectra: ld(g) -> S[g], st(n, y[n]) is called for the
// second half of the frame only (n >= N/2: overlap-save keeps those); pf(i, n) is called for the same n
// (i = 0 .. R/2-1) before the last pass, so the caller can have its own operands of st in flight
template <int R, class LD, class PF, class ST>
__device__ __forceinline__ void aesc_inv(cpx *s, const cpx *__restrict__ tw, int t, LD &&ld, PF &&pf, ST &&st)
{
    using G = AescGeo<R>;
    constexpr int H = G::H;
    {   // pass A: over k3; rows kl = t and its mirror, packed spectrum rebuilt on the way in
        const int kla = t, klb = t ? G::R2 - t : G::R2 / 2;
        cpx sa[H], sb[H], a[H], b[H];
        aesc_for<0, H>([&](auto ic) {
            sa[decltype(ic)::value] = ld(decltype(ic)::value * G::R2 + kla);
            sb[decltype(ic)::value] = ld(decltype(ic)::value * G::R2 + klb);
        });
        if (t != 0) {
            aesc_for<0, H / 2>([&](auto jc) {
                constexpr int j = decltype(jc)::value, jm = H - 1 - j;
                aesc_join(sa[j], sb[jm], a[j], b[jm]);          // lower bin in row a
                aesc_join(sb[j], sa[jm], b[j], a[jm]);          // lower bin in row b
            });
        } else {
            a[0].x = sa[0].x; a[0].y = ld(G::N).x;
            a[H / 2].x = sa[H / 2].x; a[H / 2].y = ld(G::N + 1).x;
            aesc_for<1, H / 2>([&](auto jc) { constexpr int j = decltype(jc)::value; aesc_join(sa[j], sa[H - j], a[j], a[H - j]); });
            aesc_for<0, H / 2>([&](auto jc) { constexpr int j = decltype(jc)::value; aesc_join(sb[j], sb[H - 1 - j], b[j], b[H - 1 - j]); });
        }
        aesc_dft<H, H / 2, true>(a);
        aesc_dft<H, H / 2, true>(b);
        aesc_emit_tw<H>(a, c_conj(aesc_tw(tw, kla, G::N)), [&](auto nc, cpx val) { s[kla * (H + 1) + decltype(nc)::value] = val; });
        aesc_emit_tw<H>(b, c_conj(aesc_tw(tw, klb, G::N)), [&](auto nc, cpx val) { s[klb * (H + 1) + decltype(nc)::value] = val; });
    }
    __syncthreads();
    {   // pass B: over k2; thread = (k1, n3)
        const int k1 = t / H, n3 = t % H;
        cpx v[R];
        aesc_for<0, R>([&](auto ic) { v[decltype(ic)::value] = s[(k1 + R * decltype(ic)::value) * (H + 1) + n3]; });
        __syncthreads();
        aesc_dft<R, R / 2, true>(v);
        aesc_emit_tw<R>(v, c_conj(aesc_tw(tw, H * k1, G::N)), [&](auto nc, cpx val) { s[k1 * G::S2 + aesc_pad16(decltype(nc)::value * H + n3)] = val; });
    }
    __syncthreads();
    {   // pass C: over k1; thread t = n2*H + n3
        constexpr int LG = aesc_log2(R);
        cpx v[R];
        const int col = aesc_pad16(t);
        aesc_for<R / 2, R>([&](auto nc) { constexpr int n1 = decltype(nc)::value; pf(std::integral_constant<int, n1 - R / 2>{}, n1 * G::NT + t); });
        aesc_for<0, R>([&](auto ic) { v[decltype(ic)::value] = s[decltype(ic)::value * G::S2 + col]; });
        aesc_dft<R, R / 2, true>(v);
        aesc_for<R / 2, R>([&](auto nc) {
            constexpr int n1 = decltype(nc)::value;
            st(std::integral_constant<int, n1 - R / 2>{}, n1 * G::NT + t, v[aesc_brev(n1, LG)]);
        });
    }
}

struct ConvArgs {
    const float *x;         // (B, Nf, 2) f32
    float *y;               // (B, Nf, 2) f32
    cpx *Z, *W;             // [B][nblk][NS] separated spectra, natural bin order
    const cpx *H;           // [P][NS] IR partition spectra in the same layout, 1/N folded in; P a multiple of the MAC's PC
    const cpx *tw;          // [N/2]  exp(-2 pi i q / N)
    long long B, Nf;
    int nblk, P;
    float dry, wet;
};

// K1: forward FFT of the overlap-save frame [(j-1)*BK, (j+1)*BK) of every clip
template <int R>
__device__ __forceinline__ void aesc_fwd_body(const ConvArgs &a)
{
    using G = AescGeo<R>;
    AES_DYN_SMEM(cpx, s);
    const long long blk = blockIdx.x;
    const long long clip = blk / a.nblk;
    const int j = (int)(blk % a.nblk);
    const cpx *xc = reinterpret_cast<const cpx *>(a.x) + clip * a.Nf;
    const long long f0 = (long long)(j - 1) * G::BK;
    cpx *zo = a.Z + ((size_t)clip * a.nblk + j) * G::NS;
    const long long Nf = a.Nf;
    aesc_fwd<R>(s, a.tw, (int)threadIdx.x, 0.5f,
        [&](int n) { const long long f = f0 + n; cpx v; v.x = 0.f; v.y = 0.f; if (f >= 0 && f < Nf) v = xc[f]; return v; },
        [&](int g, cpx v) { zo[g] = v; });
}

// IR preparation: partition p of (hL + i*hR), zero-padded to N, through the same transform; 1/N folded in
template <int R>
__device__ __forceinline__ void aesc_ir_prep_body(const float *ir, int n_taps, cpx *Hout, const cpx *tw)
{
    using G = AescGeo<R>;
    AES_DYN_SMEM(cpx, s);
    const int p = blockIdx.x;
    cpx *ho = Hout + (size_t)p * G::NS;
    aesc_fwd<R>(s, tw, (int)threadIdx.x, 0.5f / (float)G::N,
        [&](int n) {
            const long long tt = (long long)p * G::BK + n;
            cpx v; v.x = 0.f; v.y = 0.f;
            if (n < G::BK && tt < n_taps) { v.x = ir[2 * tt]; v.y = ir[2 * tt + 1]; }
            return v;
        },
        [&](int g, cpx v) { ho[g] = v; });
}

// K3: inverse FFT, overlap-save (keep the last BK samples), dry/wet mix and clip
template <int R>
__device__ __forceinline__ void aesc_inv_body(const ConvArgs &a)
{
    using G = AescGeo<R>;
    AES_DYN_SMEM(cpx, s);
    const long long blk = blockIdx.x;
    const long long clip = blk / a.nblk;
    const int j = (int)(blk % a.nblk);
    const cpx *wi = a.W + ((size_t)clip * a.nblk + j) * G::NS;
    const cpx *xc = reinterpret_cast<const cpx *>(a.x) + clip * a.Nf;
    cpx *yc = reinterpret_cast<cpx *>(a.y) + clip * a.Nf;
    const long long f0 = (long long)j * G::BK - G::BK, Nf = a.Nf;
    const float dry = a.dry, wet = a.wet;
    cpx dv[R / 2];                                      // the dry frames of this thread's outputs, loaded ahead of the last pass
    aesc_inv<R>(s, a.tw, (int)threadIdx.x,
        [&](int g) { return wi[g]; },
        [&](auto ic, int n) {
            const long long f = f0 + n;
            cpx v; v.x = 0.f; v.y = 0.f;
            if (f < Nf) v = xc[f];
            dv[decltype(ic)::value] = v;
        },
        [&](auto ic, int n, cpx wetv) {
            const long long f = f0 + n;
            if (f < Nf) {
                const cpx dryv = dv[decltype(ic)::value];
                cpx o;
                o.x = fminf(fmaxf(__fadd_rn(__fmul_rn(dry, dryv.x), __fmul_rn(wet, wetv.x)), -1.0f), 1.0f);
                o.y = fminf(fmaxf(__fadd_rn(__fmul_rn(dry, dryv.y), __fmul_rn(wet, wetv.y)), -1.0f), 1.0f);
                yc[f] = o;
            }
        });
}

// K2: W[j] (+)= sum_{p < PC} H[p0 + p] * S[j - p0 - p] for one bin per thread (NS = row length of the spectra).
// Persistent CTAs over work items (bin tile, clip group); the item's IR tile [PC][AESC_MT] arrives by TMA in
// one of two shared buffers while the previous item computes.
template <int PC>
__device__ __forceinline__ void aesc_mac_clip(const cpx *__restrict__ zc, cpx *__restrict__ wc, size_t NS, int nq,
                                              const cpx (&h)[PC], int accumulate)
{
    cpx acc[PC];
#pragma unroll
    for (int p = 0; p < PC; ++p) { acc[p].x = 0.f; acc[p].y = 0.f; }
    // input block q = q0 + r adds H[p] * S[q] to output q + p, kept in accumulator (r + p) % PC; output q is
    // complete after input q
    auto mac = [&](auto rc, cpx z) {
        constexpr int r = decltype(rc)::value;
        aesc_for<0, PC>([&](auto pc) {
            constexpr int p = decltype(pc)::value, u = (r + p) % PC;
            acc[u].x = fmaf(-h[p].y, z.y, fmaf(h[p].x, z.x, acc[u].x));
            acc[u].y = fmaf(h[p].y, z.x, fmaf(h[p].x, z.y, acc[u].y));
        });
    };
    auto out = [&](auto rc, int q) {
        constexpr int r = decltype(rc)::value;
        cpx v = acc[r];
        cpx *dst = wc + (size_t)q * NS;
        if (accumulate) { const cpx o = *dst; v.x += o.x; v.y += o.y; }
        *dst = v;
        acc[r].x = 0.f; acc[r].y = 0.f;
    };
    // the spectra stream through in sub-batches of SB blocks: the loads of the next sub-batch are in flight
    // while this one is multiplied (indices past the end are clamped, their values unused)
    constexpr int SB = PC % 3 == 0 ? PC / 3 : PC / 2;
    const int qlast = nq - 1;
    cpx zn[SB];
#pragma unroll
    for (int r = 0; r < SB; ++r) zn[r] = zc[(size_t)(r < qlast ? r : qlast) * NS];
    for (int q0 = 0; q0 < nq; q0 += PC) {
        aesc_for<0, PC / SB>([&](auto sc) {
            constexpr int r0 = decltype(sc)::value * SB;
            cpx z[SB];
#pragma unroll
            for (int r = 0; r < SB; ++r) z[r] = zn[r];
#pragma unroll
            for (int r = 0; r < SB; ++r) {
                const int qn = q0 + r0 + SB + r;
                zn[r] = zc[(size_t)(qn < qlast ? qn : qlast) * NS];
            }
            aesc_for<0, SB>([&](auto rc) {
                constexpr int r = r0 + decltype(rc)::value;
                if (q0 + r < nq) { mac(std::integral_constant<int, r>{}, z[decltype(rc)::value]); out(std::integral_constant<int, r>{}, q0 + r); }
            });
        });
    }
}

template <int PC>
__device__ __forceinline__ void aesc_mac_body(const ConvArgs &a, int NS, int p0, int clips_per_item, int accumulate)
{
    AES_DYN_SMEM(cpx, s);                               // 2 x [PC][MT] IR tiles | 2 mbarriers
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(s + 2 * PC * AESC_MT);
    const int tid = threadIdx.x;
    const int ntiles = (NS + AESC_MT - 1) / AESC_MT;
    const long long ngroups = (a.B + clips_per_item - 1) / clips_per_item;
    const long long nitems = (long long)ntiles * ngroups;
    if (tid == 0) { aes_mbar_init(bars, 1); aes_mbar_init(bars + 1, 1); aes_mbar_init_fence(); }
    __syncthreads();
    auto issue = [&](long long item, int buf) {
        const int tile = (int)(item % ntiles);
        aes_fence_proxy_async_smem();                   // the buffer was read with ordinary loads before
        aes_mbar_expect(bars + buf, (unsigned)(PC * AESC_MT * sizeof(cpx)));
        const unsigned long long keep = aes_policy_evict_last();
        for (int p = 0; p < PC; ++p)                    // (the last tile runs past its row: H has a tile of slack)
            aes_bulk_g2s(s + (buf * PC + p) * AESC_MT, a.H + (size_t)(p0 + p) * NS + tile * AESC_MT,
                         (unsigned)(AESC_MT * sizeof(cpx)), bars + buf, keep);
#ifdef AES_CPU_EMU
        aes_mbar_complete_emu(bars + buf);
#endif
    };
    long long item = blockIdx.x;
    if (tid == 0 && item < nitems) issue(item, 0);
    const int nq = a.nblk - p0;                         // input blocks that reach an output through this chunk
    for (unsigned it = 0; item < nitems; item += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && item + gridDim.x < nitems) issue(item + gridDim.x, buf ^ 1);
        aes_mbar_wait(bars + buf, it >> 1);
        cpx h[PC];
#pragma unroll
        for (int p = 0; p < PC; ++p) h[p] = s[(buf * PC + p) * AESC_MT + tid];
        __syncthreads();                                // every thread holds its values: the buffer may be refilled
        const int tile = (int)(item % ntiles);
        const long long grp = item / ntiles;
        const int g = tile * AESC_MT + tid;
        long long c1 = (grp + 1) * clips_per_item;
        if (c1 > a.B) c1 = a.B;
        if (g < NS)
            for (long long clip = grp * clips_per_item; clip < c1; ++clip)
                aesc_mac_clip<PC>(a.Z + (size_t)clip * a.nblk * NS + g, a.W + ((size_t)clip * a.nblk + p0) * NS + g,
                                  (size_t)NS, nq, h, accumulate);
    }
}
