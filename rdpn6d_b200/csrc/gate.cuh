// The mask part of the gate (gdrn_evaluator.py:110-117 on top of engine_utils.py:118-136 get_out_mask),
// shared by the fused solver (pose_solve.cu) and the host-buffer pull kernel (host_api.cu): both must take
// the SAME decision for every pixel, so the arithmetic lives in one place.
#pragma once
#include "common.cuh"

#include <float.h>
#include <math.h>
#include <string.h>

namespace rdpn {

struct RoiConst {
    float fx, fy, cx, cy;
    float ext[3];
    float gthr[3];
    float div;  // depth divisor or 0 (= none)
    float mn, mx;
};

struct RoiGate {
    float hi, lo;     // fast mask filter: a > hi -> in, a < lo -> out, else exact test
    double cut;       // exact: (double)a > / >= (double)b * cut
    float b;          // max - min
    int incl;         // 1: >= (odd mantissa of the threshold), 0: >
};

// host side: fl(a/b) > thr  <=>  a/b > (>=) cut, cut = midpoint between thr and its FP32 successor
// (ties-to-even decides the inclusivity)
inline void host_mask_cut(float thr, double* cut, int* incl) {
    uint32_t bits;
    memcpy(&bits, &thr, sizeof(bits));
    *cut = 0.5 * ((double)thr + (double)nextafterf(thr, INFINITY));
    *incl = (int)(bits & 1u);
}

__device__ __forceinline__ float mask_prob(float m, int mode, float mn, float mx) {
    if (mode == RDPN_MASK_L1) return __fdiv_rn(__fsub_rn(m, mn), __fsub_rn(mx, mn));
    if (mode == RDPN_MASK_BCE) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-m)));
    return m;
}

// L1 mode: the per-ROI constants of the division-free test from the mask's min / max
__device__ __forceinline__ void make_gate(RoiGate& g, float lo, float hi, float mask_thr, double cut, int incl) {
    g.b = __fsub_rn(hi, lo);
    g.cut = cut;
    g.incl = incl;
    const float bt = g.b * mask_thr;
    const bool filt = mask_thr > 1e-30f && mask_thr < 1e30f;
    g.hi = filt ? bt * 1.000002f : INFINITY;
    g.lo = filt ? bt * 0.999998f : -INFINITY;
}

// (mask_prob(m) > mask_thr) without the division for the L1 mode
__device__ __forceinline__ bool mask_pass(float m, int mode, float thr, float mn, const RoiGate& g) {
    if (mode == RDPN_MASK_L1) {
        if (!(g.b > 0.f)) return false;  // flat mask: 0/0 = NaN never passes (engine_utils.py:128 has no eps)
        const float a = __fsub_rn(m, mn);
        if (a > g.hi) return true;
        if (a < g.lo) return false;
        const double l = (double)a, r = __dmul_rn((double)g.b, g.cut);
        return g.incl ? (l >= r) : (l > r);  // NaN -> false
    }
    return mask_prob(m, mode, 0.f, 0.f) > thr;
}

// min / max over 4 quad values folded into running values (fminf / fmaxf skip NaNs, so the result does not
// depend on the reduction order)
__device__ __forceinline__ void minmax4(const float4& m4, float& mn, float& mx) {
    mn = fminf(fminf(fminf(mn, m4.x), fminf(m4.y, m4.z)), m4.w);
    mx = fmaxf(fmaxf(fmaxf(mx, m4.x), fmaxf(m4.y, m4.z)), m4.w);
}

}  // namespace rdpn
