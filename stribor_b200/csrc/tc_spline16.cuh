// Register-resident spline arithmetic shared by the tensor-core kernels (tc_layer.cu, tc_wide.cu):
// the 48 network outputs of one (row, transformed dim) element arrive from TMEM, 16 bins.
// Reference semantics restated here: util/rational_quadratic_spline.py, util/cubic_spline.py,
// util/search_sorted.py, flow.py:42-47.
#pragma once
#include "common.cuh"
#include "stb_math.cuh"
#include "stb_grad.cuh"
#include "tc_common.cuh"

#ifndef STB_TC_EXP
#define STB_TC_EXP 0                   // experiment bits (profiling builds only): 1 no finish, 2 no ex2, 4 no acc_full wait
#endif

namespace stb {
namespace sp16 {
using namespace tc;

constexpr int kBins = 16;

// -----------------------------------------------------------------------------------------------
// spline evaluation with the 48 parameters of one element in registers
// -----------------------------------------------------------------------------------------------
// exp through ex2.approx: with the maximum subtracted the arguments are <= 0, so the absolute
// error of every term is <= ~2.5 ulp of the LARGEST term (= 1), i.e. the same absolute accuracy on
// the bin sizes as a 2-ulp expf.  Measured on the GPU against the fp64 oracle
// (profiles/r01_tc_accuracy.txt): accurate expf, IEEE division and a true e_i / sum quotient change
// the mean log-det error by < 7 %; only compensated cumulative sums help (12 %) and cost ~90
// instructions per element.
__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// tanh to ~3e-7 ABSOLUTE error: (1 - e^-2|v|) / (1 + e^-2|v|).  The result feeds a contraction
// with O(0.1) weights, where only the absolute error matters (tanhf costs ~3x the instructions).
__device__ __forceinline__ float tanh_fast(float v);

// a / b to ~1 ulp: MUFU.RCP + one residual correction (4 instructions instead of ~10 for the
// IEEE sequence; the operands here are never subnormal / huge)
__device__ __forceinline__ float fdiv(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    const float q = a * r;
    return fmaf(fmaf(-q, b, a), r, q);
}
// sqrt(x), x >= 0, to ~1 ulp: MUFU.RSQ + one Newton step
__device__ __forceinline__ float fsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float s = x * r;
    const float h = 0.5f * r;
    return (x > 0.f) ? fmaf(fmaf(-s, s, x), h, s) : 0.f;
}
// F.softplus (beta 1, threshold 20); exp through ex2.approx (relative error ~|v| * 1e-7)
// softplus(v) = max(v, 0) + log1p(e), e = exp(-|v|) in (0, 1];  log1p(e) = 2 atanh(z), z = e / (2 + e)
// in (0, 1/3]: odd series to z^13 (relative error < 3e-8 over the range) -- about half the
// instructions of log1pf(expf(v)), same accuracy class.
__device__ __forceinline__ float softplus_fast(float v) {
    const float e = ex2_approx(-1.4426950408889634f * fabsf(v));
    const float z = fdiv(e, 2.f + e);
    const float z2 = z * z;
    float p = fmaf(z2, 1.f / 13.f, 1.f / 11.f);
    p = fmaf(p, z2, 1.f / 9.f);
    p = fmaf(p, z2, 1.f / 7.f);
    p = fmaf(p, z2, 1.f / 5.f);
    p = fmaf(p, z2, 1.f / 3.f);
    p = fmaf(p, z2, 1.f);
    return fmaxf(v, 0.f) + 2.f * z * p;
}

__device__ __forceinline__ float tanh_fast(float v) {
    if (STB_TC_EXP & 32) return v * 0.125f;
    const float t = ex2_approx(-2.885390081777927f * fabsf(v));
    return copysignf(fdiv(1.f - t, 1.f + t), v);
}

// hidden activation Sigmoid on the tensor-core paths (the other bounded one; unbounded activations never reach
// these kernels, see tc_layer_supported): clamped so that ex2 cannot overflow into rcp(inf) * inf = NaN
__device__ __forceinline__ float sigmoid_act(float v) {
    const float e = ex2_approx(-1.4426950408889634f * fmaxf(v, -80.f));
    return fdiv(1.f, 1.f + e);
}

struct RqsBin16 {
    float xk, wk, yk, hk, delta, d0, d1;
};

// log f'(x) for x in bin b: log(delta^2 (d1 th^2 + 2 delta th(1-th) + d0 (1-th)^2) / den^2)
// (rational_quadratic_spline.py:245-248, the two logs merged into one)
// Evaluated as ln2 (lg2 dnum - 2 lg2 den) on MUFU.LG2 (absolute error ~2^-22 per term: the same
// order as one fp32 rounding of a log-derivative of O(1), and 6 instructions instead of ~35 for
// an IEEE division + logf).
__device__ __forceinline__ float lg2_approx(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rqs16_log_deriv(const RqsBin16& b, float theta, float tt, float den) {
    const float omt = 1.f - theta;
    const float dnum = b.delta * b.delta * (b.d1 * theta * theta + 2.f * b.delta * tt + b.d0 * omt * omt);
    return 0.69314718055994530942f * fmaf(-2.f, lg2_approx(den), lg2_approx(dnum));
}

// ONE deliberate simplification of the coupling semantics: in the inverse direction the forward
// log-derivative at the recovered point is evaluated in the bin the inverse search found.  The
// reference re-searches (flow.py:42-47 -> coupling.py:84-95) and can land in the neighbouring bin
// only when rounding puts the recovered point within an ulp of a knot, where the spline is C1, so
// the two evaluations agree to O(ulp).
struct CubSel {
    float wp, wk, wn, hp, hk, hn, cw, ch;
    int k;
};

// Divisions and square roots below use the ~1 ulp MUFU + one-Newton-step forms (fdiv / fsqrt): bin sizes are
// >= 1e-2 and O(1), never subnormal or huge.
// (clamped: 2^(1.44 * 80) is finite, so a very negative argument gives ~0 instead of rcp(inf) -> NaN)
__device__ __forceinline__ float sigmoid_fast(float v) { return fdiv(1.f, 1.f + ex2_approx(-1.4426950408889634f * fmaxf(v, -80.f))); }

__device__ __forceinline__ CubBin cubic16_bin(const CubSel& s, float ul, float ur, int K = kBins) {
    float inv_wk;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_wk) : "f"(s.wk));
    inv_wk = fmaf(fmaf(-s.wk, inv_wk, 1.f), inv_wk, inv_wk);         // 1 / w_k to ~1 ulp, used three times
    const float sk = s.hk * inv_wk;
    float dl, dr;
    if (s.k == 0) {
        dl = sigmoid_fast(ul) * 3.f * sk;
    } else {
        const float sp = fdiv(s.hp, s.wp);
        dl = fminf(fminf(fabsf(sp), fabsf(sk)), fdiv(0.5f * (s.wk * sp + s.wp * sk), s.wp + s.wk)) *
             (sign_f(sp) + sign_f(sk));
    }
    if (s.k == K - 1) {
        dr = sigmoid_fast(ur) * 3.f * sk;
    } else {
        const float sn = fdiv(s.hn, s.wn);
        dr = fminf(fminf(fabsf(sk), fabsf(sn)), fdiv(0.5f * (s.wn * sk + s.wk * sn), s.wk + s.wn)) *
             (sign_f(sk) + sign_f(sn));
    }
    CubBin r;
    r.a = (dl + dr - 2.f * sk) * inv_wk * inv_wk;
    r.b = (3.f * sk - 2.f * dl - dr) * inv_wk;
    r.c = dl;
    r.d = s.ch;
    r.xl = s.cw;
    r.xr = (s.k == K - 1) ? 1.f : s.cw + s.wk;
    return r;
}

// ---- split-phase variants used by the kernel: the bin is located from the 32 softmax columns
// first; the remaining parameter columns are pulled from TMEM afterwards ------------------------
//
// The 32 softmax columns of a dim arrive INTERLEAVED, (w_i, h_i) in adjacent TMEM columns ==
// adjacent registers, so everything that treats the two axes alike (bias, shift, sums, cumulative
// walk, normalisation) runs on Blackwell's packed fp32 pairs (FFMA2 / FADD2): half the issue slots
// of the scalar form -- the epilogue is issue-bound (profiles/r01_tc_ncu_summary_16warp.txt).
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

// in place: t[i] = (2^(w_i - mw), 2^(h_i - mh)).  `shift` is warp-uniform: false when the pack
// step proved |logit| <= 100 for this dim (bounded activation, row-wise L1 bound on the last
// Linear), so 2^logit neither overflows nor flushes and the maximum need not be found at all.
__device__ __forceinline__ void softmax16_num2(float2* t, bool shift) {
    if (shift) {
        float mw = t[0].x, mh = t[0].y;
#pragma unroll
        for (int i = 1; i < kBins; i += 2) {
            mw = fmaxf(fmaxf(mw, t[i].x), (i + 1 < kBins) ? t[i + 1].x : t[i].x);
            mh = fmaxf(fmaxf(mh, t[i].y), (i + 1 < kBins) ? t[i + 1].y : t[i].y);
        }
        const float2 nm = f2(-mw, -mh);
#pragma unroll
        for (int i = 0; i < kBins; ++i) t[i] = __fadd2_rn(t[i], nm);
    }
#pragma unroll
    for (int i = 0; i < kBins; ++i) {
        if (STB_TC_EXP & 2) { t[i] = __ffma2_rn(t[i], t[i], f2(1.f)); continue; }
        t[i].x = ex2_approx(t[i].x); t[i].y = ex2_approx(t[i].y);
    }
}

__device__ __forceinline__ float2 sel2(bool p, float2 a, float2 b) { return p ? a : b; }

// Bin search over the softmax numerators e_i = t[i] (pairs: widths, heights) WITHOUT forming the 16
// cumulative sums.  The bin sizes are w_i = min + c e_i, c = (1 - 16 min) / sum, so knot i sits at
// lo + span (i min + c E_i) with E_i = e_0 + .. + e_(i-1), and
//     key >= knot_i   <=>   E_i <= kq - i step        (kq, step: the key and min in units of c)
// is a prefix property of i (search_sorted.py:3-5: bin = #(key >= knot_i) - 1; the nudged last
// knot is never reached by an inside key).  The pairwise sum tree that yields the normaliser also
// yields E_i along a binary descent (E_8 = first half, E_8 +- quarter, ...): 4 compares instead of
// 15, and the descent leaves E_k of BOTH axes, the numerators of bin k and k itself.  The issue
// slots go into ~50 selects instead of ~140 walk instructions (the epilogue is issue-bound).
struct BinSearch16 {
    float2 S;            // sums of the numerators
    float2 c;            // (1 - 16 min) / S
    float2 Ek;           // numerators summed before bin k
    int k;
    bool p3, p2, p1, p0; // bits of k
};

// FULL = false: K <= 16 real bins, the numerators of bins K .. 15 are exactly 0 (their packed biases are -inf), so
// every sum below is already right; the only changes are the scale 1 - K min and that a padded knot index (>= K)
// must never compare true (for an inside key it could only at key == upper, where the reference's nudged last knot
// keeps the last REAL bin).
template <bool ON_H, bool FULL = true>
__device__ __forceinline__ BinSearch16 bin_search16(const float2* t, float min_size, float key01, float2& e_even,
                                                    float2& e_odd, int K = kBins) {
    float2 s8[8], s4[4], s2[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) s8[i] = __fadd2_rn(t[2 * i], t[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s4[i] = __fadd2_rn(s8[2 * i], s8[2 * i + 1]);
    s2[0] = __fadd2_rn(s4[0], s4[1]);
    s2[1] = __fadd2_rn(s4[2], s4[3]);
    BinSearch16 r;
    r.S = __fadd2_rn(s2[0], s2[1]);
    const float scale = 1.f - min_size * (float)(FULL ? kBins : K);
    r.c = f2(fdiv(scale, r.S.x), fdiv(scale, r.S.y));
    const float inv_c = (ON_H ? r.S.y : r.S.x) * (1.f / scale);
    const float step = min_size * inv_c;
    float base = key01 * inv_c;                    // threshold of index i: base - (i - position) step
#define STB_SEARCHED(v) (ON_H ? (v).y : (v).x)
    r.p3 = (STB_SEARCHED(s2[0]) <= fmaf(-8.f, step, base)) && (FULL || 8 < K);
    float2 E = sel2(r.p3, s2[0], f2(0.f));
    base = r.p3 ? fmaf(-8.f, step, base) : base;
    float2 cand = __fadd2_rn(E, sel2(r.p3, s4[2], s4[0]));
    r.p2 = (STB_SEARCHED(cand) <= fmaf(-4.f, step, base)) && (FULL || (r.p3 ? 12 : 4) < K);
    E = sel2(r.p2, cand, E);
    base = r.p2 ? fmaf(-4.f, step, base) : base;
    cand = __fadd2_rn(E, sel2(r.p3, sel2(r.p2, s8[6], s8[4]), sel2(r.p2, s8[2], s8[0])));
    r.p1 = (STB_SEARCHED(cand) <= fmaf(-2.f, step, base)) && (FULL || (r.p3 ? 8 : 0) + (r.p2 ? 4 : 0) + 2 < K);
    E = sel2(r.p1, cand, E);
    base = r.p1 ? fmaf(-2.f, step, base) : base;
    // numerators of the pair 4 p3 + 2 p2 + p1: selected p3 first (known earliest)
    float2 a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = sel2(r.p3, t[8 + i], t[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = sel2(r.p2, a[4 + i], a[i]);
    e_even = sel2(r.p1, b[2], b[0]);
    e_odd = sel2(r.p1, b[3], b[1]);
    cand = __fadd2_rn(E, e_even);
    r.p0 = (STB_SEARCHED(cand) <= base - step) && (FULL || (r.p3 ? 8 : 0) + (r.p2 ? 4 : 0) + (r.p1 ? 2 : 0) + 1 < K);
    r.Ek = sel2(r.p0, cand, E);
#undef STB_SEARCHED
    r.k = (r.p3 ? 8 : 0) + (r.p2 ? 4 : 0) + (r.p1 ? 2 : 0) + (r.p0 ? 1 : 0);
    return r;
}

struct RqsLoc {
    float2 Ek, Ek1;      // (widths, heights): sums of the softmax numerators before bin k / through bin k
    float2 c;            // (1 - 16 min) / sum, per axis
    int k;
};

// t[0..16): log2(e)-scaled raw (width, height) pairs (destroyed).  Leaves E_k and E_(k+1) of both
// axes: the knots of bin k, from which its width / height are re-derived as the reference does
// (rational_quadratic_spline.py:180-192).
template <bool ON_H, bool FULL = true>
__device__ __forceinline__ RqsLoc rqs16_locate(float2* t, bool shift, float lo, float inv_span, float key, int K = kBins) {
    softmax16_num2(t, shift);
    float2 ee, eo;
    const BinSearch16 bs = bin_search16<ON_H, FULL>(t, STB_RQS_MIN, (key - lo) * inv_span, ee, eo, K);
    RqsLoc r;
    r.c = bs.c;
    r.k = bs.k;
    r.Ek = bs.Ek;
    r.Ek1 = __fadd2_rn(bs.Ek, sel2(bs.p0, eo, ee));
    return r;
}

// (v[k - 1], v[k]) for k in [0, 16) from a register array v[0..15), out-of-range entries 0: a
// select tree over a sliding window (19 selects for the pair)
__device__ __forceinline__ void pick_pair16(const float* v, int k, float& lo_v, float& hi_v) {
    float a[9], b[5], c[3];
#pragma unroll
    for (int j = 0; j < 9; ++j) {              // window over V[m] = v[m - 1], V[0] = V[16] = 0
        const float x0 = (j == 0) ? 0.f : v[j - 1];
        const float x1 = (j + 8 == 16) ? 0.f : v[j + 7];
        a[j] = (k & 8) ? x1 : x0;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) b[j] = (k & 4) ? a[j + 4] : a[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) c[j] = (k & 2) ? b[j + 2] : b[j];
    lo_v = (k & 1) ? c[1] : c[0];
    hi_v = (k & 1) ? c[2] : c[1];
}

// softplus_fast on a pair
__device__ __forceinline__ float2 softplus_fast2(float2 v) {
    const float2 e = f2(ex2_approx(-1.4426950408889634f * fabsf(v.x)), ex2_approx(-1.4426950408889634f * fabsf(v.y)));
    const float2 den = __fadd2_rn(e, f2(2.f));
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(den.y));
    const float2 q = __fmul2_rn(e, r);
    const float2 nq = f2(-q.x, -q.y);
    const float2 z = __ffma2_rn(__ffma2_rn(nq, den, e), r, q);
    const float2 z2 = __fmul2_rn(z, z);
    float2 p = __ffma2_rn(z2, f2(1.f / 13.f), f2(1.f / 11.f));
    p = __ffma2_rn(p, z2, f2(1.f / 9.f));
    p = __ffma2_rn(p, z2, f2(1.f / 7.f));
    p = __ffma2_rn(p, z2, f2(1.f / 5.f));
    p = __ffma2_rn(p, z2, f2(1.f / 3.f));
    p = __ffma2_rn(p, z2, f2(1.f));
    return __ffma2_rn(__fadd2_rn(z, z), p, f2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)));
}

// u0 / u1: (bias-added, unscaled) derivative parameters at the two knots of bin r.k
template <bool INVERSE>
__device__ __forceinline__ void rqs16_finish(const RqsLoc& r, float u0, float u1, float lo, float hi, bool want_ld,
                                             float x, float& out, float& ld, int K = kBins) {
    const float span = hi - lo;
    const float km = (float)r.k * STB_RQS_MIN;
    float2 k0 = __ffma2_rn(f2(span), __ffma2_rn(r.c, r.Ek, f2(km)), f2(lo));                   // (x_k, y_k)
    float2 k1 = __ffma2_rn(f2(span), __ffma2_rn(r.c, r.Ek1, f2(km + STB_RQS_MIN)), f2(lo));    // (x_k+1, y_k+1)
    if (r.k == 0) k0 = f2(lo);                                       // knot_0 / knot_K forced to the box
    if (r.k == K - 1) k1 = f2(hi);
    const float2 wh = __fadd2_rn(k1, f2(-k0.x, -k0.y));              // sizes re-derived from the knots (:185)
    const float2 dd = __fadd2_rn(softplus_fast2(f2(u0, u1)), f2(STB_RQS_MIN));
    RqsBin16 b;
    b.xk = k0.x; b.yk = k0.y; b.wk = wh.x; b.hk = wh.y;
    float inv_wk;                                                    // 1 / w_k to ~1 ulp, shared by delta and theta
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_wk) : "f"(b.wk));
    inv_wk = fmaf(fmaf(-b.wk, inv_wk, 1.f), inv_wk, inv_wk);
    b.delta = b.hk * inv_wk;
    b.d0 = dd.x; b.d1 = dd.y;
    const float s = b.d0 + b.d1 - 2.f * b.delta;
    if (!INVERSE) {
        const float theta = (x - b.xk) * inv_wk;
        const float tt = theta * (1.f - theta);
        const float den = b.delta + s * tt;
        out = b.yk + fdiv(b.hk * (b.delta * theta * theta + b.d0 * tt), den);
        ld = rqs16_log_deriv(b, theta, tt, den);
    } else {
        const float dy = x - b.yk;
        const float qa = dy * s + b.hk * (b.delta - b.d0);
        const float qb = b.hk * b.d0 - dy * s;
        const float qc = -b.delta * dy;
        const float disc = qb * qb - 4.f * qa * qc;
        const float root = fdiv(2.f * qc, -qb - fsqrt(disc));
        out = root * b.wk + b.xk;
        // forward log-derivative at the recovered point: theta = (out - x_k) / w_k is `root` up to the
        // rounding of `out` (the reference recomputes it from the rounded output, flow.py:42-47)
        const float tt = root * (1.f - root);
        ld = (want_ld && out >= lo && out <= hi) ? -rqs16_log_deriv(b, root, tt, b.delta + s * tt) : 0.f;
    }
}

// t[0..16): log2(e)-scaled raw (width, height) pairs (destroyed).  cubic_spline.py:104-116,140-151:
// same search; the three bin sizes around bin k that the slopes need are picked from the
// numerators by a select tree on the bits of k and normalised afterwards (the cubic code uses the
// sizes themselves, not knot differences).
template <bool ON_H, bool FULL = true>
__device__ __forceinline__ CubSel cubic16_locate(float2* t, bool shift, float u, int K = kBins) {
    softmax16_num2(t, shift);
    float2 ee, eo;
    const BinSearch16 bs = bin_search16<ON_H, FULL>(t, STB_CUB_MIN, u, ee, eo, K);
    const int k = bs.k;
    // neighbours: k even -> (t[k-1], ee, eo); k odd -> (ee, eo, t[k+1]).  t[k-1] for even k = odd element
    // of the previous pair, t[k+1] for odd k = even element of the next pair: one more 8-way select each
    float2 am[4], bm[2], ap[4], bp[2];
    // prev-odd element of pair j: t[2j - 1] (j = 0 -> 0), next-even: t[2j + 2] (j = 7 -> 0)
#pragma unroll
    for (int i = 0; i < 4; ++i) {                  // select on p3 first: candidates j = i (p3 = 0) or j = i + 4
        am[i] = sel2(bs.p3, t[2 * (i + 4) - 1], (i == 0) ? f2(0.f) : t[2 * i - 1]);
        ap[i] = sel2(bs.p3, (i + 4 == 7) ? f2(0.f) : t[2 * (i + 4) + 2], t[2 * i + 2]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {                  // then p2: j = i or i + 2 within the half
        bm[i] = sel2(bs.p2, am[i + 2], am[i]);
        bp[i] = sel2(bs.p2, ap[i + 2], ap[i]);
    }
    const float2 prev_odd = sel2(bs.p1, bm[1], bm[0]);
    const float2 next_even = sel2(bs.p1, bp[1], bp[0]);
    const float2 ep = sel2(bs.p0, ee, prev_odd), ek = sel2(bs.p0, eo, ee), en = sel2(bs.p0, next_even, eo);
    const float2 mn = f2(STB_CUB_MIN);
    const float2 sp = __ffma2_rn(bs.c, ep, mn), sk = __ffma2_rn(bs.c, ek, mn), sn = __ffma2_rn(bs.c, en, mn);
    const float2 cum = __ffma2_rn(bs.c, bs.Ek, f2((float)k * STB_CUB_MIN));
    CubSel r;
    r.k = k;
    r.wp = sp.x; r.hp = sp.y; r.wk = sk.x; r.hk = sk.y; r.wn = sn.x; r.hn = sn.y;
    r.cw = cum.x; r.ch = cum.y;
    return r;
}

// cubic_spline.py:153-228 (Cardano: three real roots by the trigonometric form, one by cube roots, and the
// almost-quadratic case), same branch structure as stb_math.cuh:cubic_inverse_in_bin with the cheap
// division / square-root forms; the cube root is sign(v) 2^(log2|v| / 3) on MUFU (relative error ~2^-21,
// the reference's own exp(log|v| / 3) in fp32 is no better once |log| > 1).
__device__ __forceinline__ float cbrt_fast(float v) {
    return copysignf(ex2_approx(lg2_approx(fabsf(v)) * 0.33333333333333333f), v) * ((v == 0.f) ? 0.f : 1.f);
}
// atan2(y, x) for y >= 0 (result in [0, pi]) and sin / cos on [0, pi / 3]: minimax polynomials evaluated in fp32
// (absolute error ~1.5e-7 / ~1.1e-7, the class of the library routines) -- atan2f + sincosf with their range reduction
// were 12.5 % of the cubic whole-flow kernel's instructions (profiles/r02_tc_chain_cubic_ncu_summary.txt).
__device__ __forceinline__ float atan2_upper(float y, float x) {
    const float ax = fabsf(x);
    const float mn = fminf(y, ax), mx = fmaxf(y, ax);
    const float a = (mx > 0.f) ? fdiv(mn, mx) : 0.f;
    const float s = a * a;
    float p = fmaf(s, -0.004054553612250179f, 0.02186291228067595f);
    p = fmaf(p, s, -0.055912267216883055f);
    p = fmaf(p, s, 0.09642193534850359f);
    p = fmaf(p, s, -0.13908628358983344f);
    p = fmaf(p, s, 0.19946565495495794f);
    p = fmaf(p, s, -0.33329860782060494f);
    p = fmaf(p, s, 0.9999993355834141f);
    float r = p * a;
    r = (y > ax) ? 1.5707963267948966f - r : r;
    return (x < 0.f) ? 3.141592653589793f - r : r;
}
__device__ __forceinline__ void sincos_third(float t, float& sn, float& cs) {      // t in [0, 1.06]
    const float u = t * t;
    float ps = fmaf(u, 2.679316525301239e-06f, -0.00019832722144834137f);
    ps = fmaf(ps, u, 0.008333291415817233f);
    ps = fmaf(ps, u, -0.16666665826998622f);
    ps = fmaf(ps, u, 0.9999999995289165f);
    sn = ps * t;
    float pc = fmaf(u, 2.4038486070003582e-05f, -0.0013881424341214848f);
    pc = fmaf(pc, u, 0.04166636798661092f);
    pc = fmaf(pc, u, -0.49999995814702086f);
    cs = fmaf(pc, u, 0.9999999990628845f);
}
__device__ __forceinline__ float cubic16_inverse_in_bin(const CubBin& k, float u) {
    float inv_a;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_a) : "f"(k.a));
    inv_a = fmaf(fmaf(-k.a, inv_a, 1.f), inv_a, inv_a);
    const float third = 0.33333333333333333f;
    const float b_ = (k.b * inv_a) * third;
    const float c_ = (k.c * inv_a) * third;
    const float d_ = (k.d - u) * inv_a;
    const float delta_1 = -b_ * b_ + c_;
    const float delta_2 = -c_ * b_ + d_;
    const float delta_3 = b_ * d_ - c_ * c_;
    const float disc = 4.f * delta_1 * delta_3 - delta_2 * delta_2;
    const float dep_1 = -2.f * b_ * delta_1 + delta_2;
    const float dep_2 = delta_1;
    float out;
    if (disc > 0.f) {
        const float theta = atan2_upper(fsqrt(disc), -dep_1) * third;
        float c1, c2;
        sincos_third(theta, c2, c1);
        const float scale = 2.f * fsqrt(-dep_2);
        const float shift = -b_ + k.xl;
        const float r1 = c1 * scale + shift;
        const float r2 = (-0.5f * c1 - 0.5f * 1.7320508075688772f * c2) * scale + shift;
        const float r3 = (-0.5f * c1 + 0.5f * 1.7320508075688772f * c2) * scale + shift;
        const float lo = k.xl - STB_CUB_EPS, hi = k.xr + STB_CUB_EPS;
        const bool ok1 = (lo < r1) && (r1 < hi), ok2 = (lo < r2) && (r2 < hi), ok3 = (lo < r3) && (r3 < hi);
        out = ok1 ? r1 : (ok2 ? r2 : (ok3 ? r3 : r1));
    } else {
        const float sq = fsqrt(-disc);
        out = (cbrt_fast((-dep_1 + sq) * 0.5f) + cbrt_fast((-dep_1 - sq) * 0.5f)) - b_ + k.xl;
    }
    if (fabsf(k.a) < STB_CUB_QUAD) {
        const float qc = k.d - u;
        out = fdiv(-k.c + fsqrt(k.c * k.c - 4.f * k.b * qc), 2.f * k.b) + k.xl;
    }
    return out;
}

template <bool INVERSE>
__device__ __forceinline__ void cubic16_finish(const CubSel& s, float ul, float ur, float lo, float hi, bool want_ld,
                                               float u, float& out, float& ld, int K = kBins) {
    const float span = hi - lo;
    const CubBin b = cubic16_bin(s, ul, ur, K);
    if (!INVERSE) {
        out = cubic_forward_in_bin(b, u, ld) * span + lo;
    } else {
        out = cubic16_inverse_in_bin(b, u) * span + lo;
        ld = 0.f;
        if (want_ld && out >= lo && out <= hi) {
            // forward log-derivative at the recovered point, in the bin the inverse search found
            // (see the note above RqsBin16's users for why the re-search is skipped)
            const float sr = fdiv(out - lo, span) - b.xl;
            ld = -0.69314718055994530942f * lg2_approx(fmaf(fmaf(3.f * b.a, sr, 2.f * b.b), sr, b.c));
        }
    }
}



// -----------------------------------------------------------------------------------------------
// gradient of one RQS element with the softmax numerators in registers (training step)
// -----------------------------------------------------------------------------------------------
// What autograd derives for util/rational_quadratic_spline.py:101-107,180-248 (and flow.py:42-47 for
// the inverse direction), in the register form of rqs16_locate / rqs16_finish:
//   t[0..16)  in : softmax numerators e_i (widths, heights) from softmax16_num2
//             out: gradient wrt the 16 + 16 raw (natural-log domain) width / height logits
//   bs        the bin search result over t (ee / eo: numerators of the pair that contains bin k)
//   u0, u1    raw derivative parameters at the two knots of bin k (box ends: the constant)
// The in-bin map and its log-derivative are differentiated with the 7-variable dual numbers of
// stb_grad.cuh (rqs_bin_dual); knots -> sizes -> logits is back-propagated in closed form:
//   W_i = min + c e_i, c = (1 - 16 min) / sum e  =>  dL/dlogit_i = c e_i (g_i - sum_j s_j g_j),
//   g_i = g_c for i < k (every knot up to x_k+1 moves), g_k for i == k, 0 beyond.
template <bool INVERSE>
__device__ __forceinline__ void rqs16_backward(float2* t, const BinSearch16& bs, float2 ee, float2 eo, float u0,
                                               float u1, float lo, float hi, float x, float g_out, float g_ld,
                                               float& g_x, float& g_u0, float& g_u1) {
    const float span = hi - lo;
    const int k = bs.k;
    const float2 ek = sel2(bs.p0, eo, ee);
    const float2 Ek1 = __fadd2_rn(bs.Ek, ek);
    const float km = (float)k * STB_RQS_MIN;
    float2 k0 = __ffma2_rn(f2(span), __ffma2_rn(bs.c, bs.Ek, f2(km)), f2(lo));
    float2 k1 = __ffma2_rn(f2(span), __ffma2_rn(bs.c, Ek1, f2(km + STB_RQS_MIN)), f2(lo));
    if (k == 0) k0 = f2(lo);
    if (k == kBins - 1) k1 = f2(hi);
    // The MUFU-based forms of the forward kernels (softplus_fast2, fdiv, fsqrt, sigmoid_fast: ~1 ulp) instead of
    // log1pf(expf()), IEEE divisions, sqrtf and expf: ~200 of this element's ~1600 instructions, and the recomputed
    // knots / derivatives now match the forward pass's.
    const float2 dd = __fadd2_rn(softplus_fast2(f2(u0, u1)), f2(STB_RQS_MIN));
    const float d0 = dd.x, d1 = dd.y;
    float xe = x;                                   // point at which S and L = log S' are expanded
    if (INVERSE) {                                  // rational_quadratic_spline.py:212-234, as rqs16_finish<true>
        const float wk = k1.x - k0.x, hk = k1.y - k0.y, delta = fdiv(hk, wk);
        const float dy = x - k0.y, sdd = d0 + d1 - 2.f * delta;
        const float qa = dy * sdd + hk * (delta - d0), qb = hk * d0 - dy * sdd, qc = -delta * dy;
        const float root = fdiv(2.f * qc, -qb - fsqrt(qb * qb - 4.f * qa * qc));
        xe = root * wk + k0.x;
    }
    // Hand-written reverse sweep through the in-bin map S and L = log dS/dx (rational_quadratic_spline.py:
    // 236-248) instead of 7-variable forward duals: ~4x fewer instructions.  With w = x_k+1 - x_k,
    // h = y_k+1 - y_k, delta = h / w, theta = (x - x_k) / w, tt = theta (1 - theta):
    //   N = h (delta theta^2 + d0 tt),  D = delta + (d0 + d1 - 2 delta) tt,  S = y_k + N / D,
    //   P = d1 theta^2 + 2 delta tt + d0 (1 - theta)^2,  M = delta^2 P,  L = log M - 2 log D.
    // F = aS S + aL L is differentiated for the adjoint weights (aS, aL): (g_out, g_ld) in the forward
    // direction; in the inverse direction x = S^-1(y), ld = -L(x), the implicit-function theorem gives
    // g_y = (g_out - g_ld L_x) / S_x and the parameter gradients are those of F with (aS, aL) = (-g_y, -g_ld).
    float gq[RQ_N];
    {
        const float xk = k0.x, xk1 = k1.x, yk = k0.y, yk1 = k1.y;
        const float w = xk1 - xk, h = yk1 - yk;
        const float iw = fdiv(1.f, w);
        const float delta = h * iw;
        const float theta = (xe - xk) * iw, omt = 1.f - theta, tt = theta * omt, om2t = 1.f - 2.f * theta;
        const float sd = d0 + d1 - 2.f * delta;
        const float Q = delta * theta * theta + d0 * tt;              // N = h Q
        const float N = h * Q;
        const float D = delta + sd * tt;
        const float P = d1 * theta * theta + 2.f * delta * tt + d0 * omt * omt;
        const float M = delta * delta * P;
        const float iD = fdiv(1.f, D), iM = fdiv(1.f, M);
        // partials wrt theta (for S_x, L_x and the sweep)
        const float N_t = h * (2.f * delta * theta + d0 * om2t);
        const float D_t = sd * om2t;
        const float M_t = delta * delta * (2.f * d1 * theta + 2.f * delta * om2t - 2.f * d0 * omt);
        const float S_x = M * iD * iD;                               // dS/dx = exp(L)
        const float L_x = (M_t * iM - 2.f * D_t * iD) * iw;
        float aS, aL;
        if (!INVERSE) {
            aS = g_out; aL = g_ld;
            gq[RQ_X] = g_out * S_x + g_ld * L_x;
        } else {
            const float gl = (xe >= lo && xe <= hi) ? g_ld : 0.f;       // recovered point outside the box: ld = 0
            const float gy = fdiv(g_out - gl * L_x, S_x);
            gq[RQ_X] = gy;
            aS = -gy; aL = -gl;
        }
        const float F_N = aS * iD;
        const float F_D = -aS * N * iD * iD - 2.f * aL * iD;
        const float F_M = aL * iM;
        const float F_delta = F_N * h * theta * theta + F_D * (1.f - 2.f * tt) + F_M * (2.f * delta * P + 2.f * delta * delta * tt);
        const float F_theta = F_N * N_t + F_D * D_t + F_M * M_t;
        float F_h = F_N * Q + F_delta * iw;
        const float F_w = -(F_delta * delta + F_theta * theta) * iw;
        gq[RQ_D0] = F_N * h * tt + F_D * tt + F_M * delta * delta * omt * omt;
        gq[RQ_D1] = F_D * tt + F_M * delta * delta * theta * theta;
        gq[RQ_XK1] = F_w;
        gq[RQ_XK] = -F_theta * iw - F_w;
        gq[RQ_YK1] = F_h;
        gq[RQ_YK] = aS - F_h;
    }
    g_x = gq[RQ_X];
    const bool first = (k == 0), last = (k == kBins - 1);
    const float2 gk1 = last ? f2(0.f) : f2(gq[RQ_XK1], gq[RQ_YK1]);
    const float2 gk0 = first ? f2(0.f) : f2(gq[RQ_XK], gq[RQ_YK]);
    const float2 g_c = __fmul2_rn(f2(span), __fadd2_rn(gk0, gk1));  // sizes before bin k
    const float2 g_k = __fmul2_rn(f2(span), gk1);                   // size of bin k
    const float inv_scale = 1.f / (1.f - STB_RQS_MIN * (float)kBins);
    const float2 dot = __fmul2_rn(__fmul2_rn(bs.c, f2(inv_scale)),
                                  __ffma2_rn(bs.Ek, g_c, __fmul2_rn(ek, g_k)));
    const float2 ndot = f2(-dot.x, -dot.y);
    const float2 cA = __fmul2_rn(bs.c, __fadd2_rn(g_c, ndot));
    const float2 cB = __fmul2_rn(bs.c, __fadd2_rn(g_k, ndot));
    const float2 cC = __fmul2_rn(bs.c, ndot);
#pragma unroll
    for (int i = 0; i < kBins; ++i) t[i] = __fmul2_rn(t[i], (i < k) ? cA : ((i == k) ? cB : cC));
    g_u0 = first ? 0.f : gq[RQ_D0] * (u0 > 20.f ? 1.f : sigmoid_fast(u0));        // d softplus / du
    g_u1 = last ? 0.f : gq[RQ_D1] * (u1 > 20.f ? 1.f : sigmoid_fast(u1));
}


// -----------------------------------------------------------------------------------------------
// gradient of one cubic-spline element (util/cubic_spline.py:99-151,229-238 and its inverse) in the same
// register form.  t[0..16): softmax numerators in, gradient wrt the raw width / height logits out; ul / ur:
// raw end-derivative parameters.  The in-bin map is differentiated with the 11-variable duals of
// stb_grad.cuh (cubic_bin_dual: three neighbouring sizes per axis, the two cumulative sums, the end
// derivatives); a size W_i receives  g_cw (i < k)  +  g_wp (i = k - 1)  +  g_wk (i = k)  +  g_wn (i = k + 1).
// u: the element on the unit box; g_out / g_x refer to the un-normalised coordinate (span = hi - lo).
template <bool INVERSE>
__device__ __forceinline__ void cubic16_backward(float2* t, const BinSearch16& bs, float2 ee, float2 eo, float ul,
                                                 float ur, float lo, float hi, float u, float g_out, float g_ld,
                                                 float& g_x, float& g_ul, float& g_ur) {
    const float span = hi - lo;
    const int k = bs.k;
    // numerators of bins k - 1, k, k + 1 (0 beyond the ends): same select trees as cubic16_locate
    float2 am[4], bm[2], ap[4], bp[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        am[i] = sel2(bs.p3, t[2 * (i + 4) - 1], (i == 0) ? f2(0.f) : t[2 * i - 1]);
        ap[i] = sel2(bs.p3, (i + 4 == 7) ? f2(0.f) : t[2 * (i + 4) + 2], t[2 * i + 2]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        bm[i] = sel2(bs.p2, am[i + 2], am[i]);
        bp[i] = sel2(bs.p2, ap[i + 2], ap[i]);
    }
    const float2 prev_odd = sel2(bs.p1, bm[1], bm[0]);
    const float2 next_even = sel2(bs.p1, bp[1], bp[0]);
    const float2 ep = sel2(bs.p0, ee, prev_odd), ek = sel2(bs.p0, eo, ee), en = sel2(bs.p0, next_even, eo);
    const float2 mn = f2(STB_CUB_MIN);
    const float2 sp = __ffma2_rn(bs.c, ep, mn), sk = __ffma2_rn(bs.c, ek, mn), sn = __ffma2_rn(bs.c, en, mn);
    const float2 cum = __ffma2_rn(bs.c, bs.Ek, f2((float)k * STB_CUB_MIN));
    float ue = u;
    if (INVERSE) {
        CubSel s;
        s.k = k; s.wp = sp.x; s.hp = sp.y; s.wk = sk.x; s.hk = sk.y; s.wn = sn.x; s.hn = sn.y; s.cw = cum.x; s.ch = cum.y;
        const CubBin b = cubic16_bin(s, ul, ur);
        float ld_own;
        ue = cubic_inverse_in_bin(b, u, ld_own);
    }
    const bool first = (k == 0), last = (k == kBins - 1);
    Dual<CU_N> S, L;
    cubic_bin_dual(k, kBins, ue, first ? 1.f : sp.x, sk.x, last ? 1.f : sn.x, first ? 1.f : sp.y, sk.y,
                   last ? 1.f : sn.y, cum.x, cum.y, ul, ur, S, L);
    float gq[CU_N];
    if (!INVERSE) {
#pragma unroll
        for (int i = 0; i < CU_N; ++i) gq[i] = g_out * span * S.d[i] + g_ld * L.d[i];
        g_x = gq[CU_U] / span;
    } else {
        const float xo = ue * span + lo;
        const float gl = (xo >= lo && xo <= hi) ? g_ld : 0.f;
        const float gu = (g_out * span - gl * L.d[CU_U]) / S.d[CU_U];
        gq[CU_U] = gu;
#pragma unroll
        for (int i = 1; i < CU_N; ++i) gq[i] = -S.d[i] * gu - gl * L.d[i];
        g_x = gu / span;
    }
    const float2 g_c = f2(gq[CU_CW], gq[CU_CH]);
    const float2 g_p = first ? f2(0.f) : f2(gq[CU_WP], gq[CU_HP]);
    const float2 g_k = f2(gq[CU_WK], gq[CU_HK]);
    const float2 g_n = last ? f2(0.f) : f2(gq[CU_WN], gq[CU_HN]);
    const float2 g_cp = __fadd2_rn(g_c, g_p);                        // bin k - 1: cumulative sum and left neighbour
    const float inv_scale = 1.f / (1.f - STB_CUB_MIN * (float)kBins);
    // sum_j s_j g_j with s_j = e_j / sum:  E_k g_c + e_(k-1) g_p + e_k g_k + e_(k+1) g_n
    float2 acc = __fmul2_rn(bs.Ek, g_c);
    acc = __ffma2_rn(ep, g_p, acc);
    acc = __ffma2_rn(ek, g_k, acc);
    acc = __ffma2_rn(en, g_n, acc);
    const float2 dot = __fmul2_rn(__fmul2_rn(bs.c, f2(inv_scale)), acc);
    const float2 ndot = f2(-dot.x, -dot.y);
    const float2 cA = __fmul2_rn(bs.c, __fadd2_rn(g_c, ndot));       // i < k - 1
    const float2 cB = __fmul2_rn(bs.c, __fadd2_rn(g_cp, ndot));      // i = k - 1
    const float2 cC = __fmul2_rn(bs.c, __fadd2_rn(g_k, ndot));       // i = k
    const float2 cD = __fmul2_rn(bs.c, __fadd2_rn(g_n, ndot));       // i = k + 1
    const float2 cE = __fmul2_rn(bs.c, ndot);                        // beyond
#pragma unroll
    for (int i = 0; i < kBins; ++i) {
        const float2 cf = (i < k - 1) ? cA : ((i == k - 1) ? cB : ((i == k) ? cC : ((i == k + 1) ? cD : cE)));
        t[i] = __fmul2_rn(t[i], cf);
    }
    g_ul = first ? gq[CU_UL] : 0.f;
    g_ur = last ? gq[CU_UR] : 0.f;
}

}  // namespace sp16
}  // namespace stb
