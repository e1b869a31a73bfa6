// Pieces shared by the fused solver (pose_solve.cu) and the three-kernel pipeline (solve_split.cu): the launch
// arguments, the exact S1 pixel arithmetic, the FP32 residual contract, the counter-based sampling stream.
#pragma once
#include "common.cuh"
#include "gate.cuh"
#include "kabsch_math.cuh"

#include <float.h>
#include <math.h>
#include <string.h>

namespace rdpn {

struct SolveArgs {
    rdpn_roi_inputs in;
    const int32_t* hyp_idx;
    const float* t_net;
    rdpn_solve_params prm;
    rdpn_solve_outputs out;
    float sq_cut;  // smallest FP32 x with sqrtf(x) >= thr
    double mask_cut;    // midpoint between mask_thr and its FP32 successor
    int mask_cut_incl;  // ties-to-even: 1 when the quotient may equal the midpoint
};

// the raw planes of one ROI in global memory
struct RoiPlanes {
    const float* depth;
    const float* cx;
    const float* cy;
    const float* cz;
    const float* mask;
    const uint8_t* rid;
};

// one pixel of S1 with the exact oracle arithmetic; cam (and obj in dense mode)
template <bool DENSE>
__device__ __forceinline__ void pixel_s1(const RoiConst& rc, int pix, float d_raw, float cxn, float cyn, float czn,
                                         float (&cam)[3], float (&obj)[3]) {
    const float u = (float)(4 * (pix & 63));
    const float v = (float)(4 * (pix >> 6));
    float d = d_raw;
    if (rc.div != 0.f) d = __fdiv_rn(d, rc.div);                          // data_loader.py:563
    const float X = __fdiv_rn(__fmul_rn(__fsub_rn(u, rc.cx), d), rc.fx);  // :573
    const float Y = __fdiv_rn(__fmul_rn(__fsub_rn(v, rc.cy), d), rc.fy);  // :574
    const float dx = __fmul_rn(__fsub_rn(cxn, 0.5f), rc.ext[0]);          // gdrn_evaluator.py:103-105
    const float dy = __fmul_rn(__fsub_rn(cyn, 0.5f), rc.ext[1]);
    const float dz = __fmul_rn(__fsub_rn(czn, 0.5f), rc.ext[2]);
    if (DENSE) {
        cam[0] = X; cam[1] = Y; cam[2] = d;
        obj[0] = dx; obj[1] = dy; obj[2] = dz;
    } else {
        cam[0] = __fsub_rn(X, dx); cam[1] = __fsub_rn(Y, dy); cam[2] = __fsub_rn(d, dz);
        obj[0] = obj[1] = obj[2] = 0.f;
    }
}

// gather one gated pixel and run S1 on it (cam xyz, w | obj xyz)
template <bool DENSE>
__device__ __forceinline__ void gather_s1(const RoiPlanes& pl, const RoiConst& rc, int p, bool weighted, int mask_mode,
                                          float4& camw, float4& objv) {
    float cam[3], obj[3];
    pixel_s1<DENSE>(rc, p, __ldg(pl.depth + p), __ldg(pl.cx + p), __ldg(pl.cy + p), __ldg(pl.cz + p), cam, obj);
    const float w = weighted ? mask_prob(__ldg(pl.mask + p), mask_mode, rc.mn, rc.mx) : 1.f;
    camw = make_float4(cam[0], cam[1], cam[2], w);
    objv = make_float4(obj[0], obj[1], obj[2], 0.f);
}

__device__ __forceinline__ float resid2_pt(float tx, float ty, float tz, float cx, float cy, float cz) {
    const float dx = __fsub_rn(tx, cx), dy = __fsub_rn(ty, cy), dz = __fsub_rn(tz, cz);
    float d2 = __fmul_rn(dx, dx);
    d2 = __fmaf_rn(dy, dy, d2);
    d2 = __fmaf_rn(dz, dz, d2);
    return d2;
}
// c += (d2 < cut): one FSETP + one predicated IADD (the C++ form compiles to three instructions)
__device__ __forceinline__ void count_if_lt(int& c, float d2, float cut) {
    asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(c) : "f"(d2), "f"(cut));
}
// counter-based stream of the internal hypothesis sampling (include/rdpn6d_b200.h, oracle sample_triplets)
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
// R a + t with the contract's FMA order (oracle/pose_oracle.c:resid2)
__device__ __forceinline__ void xform(const float* P, float ax, float ay, float az, float& x, float& y, float& z) {
    x = __fmaf_rn(P[0], ax, P[3]);
    x = __fmaf_rn(P[1], ay, x);
    x = __fmaf_rn(P[2], az, x);
    y = __fmaf_rn(P[4], ax, P[7]);
    y = __fmaf_rn(P[5], ay, y);
    y = __fmaf_rn(P[6], az, y);
    z = __fmaf_rn(P[8], ax, P[11]);
    z = __fmaf_rn(P[9], ay, z);
    z = __fmaf_rn(P[10], az, z);
}
__device__ __forceinline__ float resid2(const float* P, float ax, float ay, float az, float cx, float cy, float cz) {
    float x, y, z;
    xform(P, ax, ay, az, x, y, z);
    return resid2_pt(x, y, z, cx, cy, cz);
}

// misc.py:134-138: k = log10(1-conf) / log10(1 - w^10), stop once i_ransac > max(k, min_iter).  Kept out of line:
// double pow/log10 are ~1500 instructions that only the (non-default) adaptive mode needs.
static __device__ __noinline__ bool adaptive_stop(int count, int n, int i_ransac, double log_1m_conf, int min_iter) {
    const double wr = (double)count / (double)n;
    const double k = log_1m_conf / log10(1.0 - pow(wr, 10.0));
    return (double)i_ransac > fmax(k, (double)min_iter);
}

// rank-select on the gate bitmap: pixel index of the k-th gated pixel in raster order (internal sampling)
__device__ __forceinline__ int kth_gated_pixel(const uint32_t* selmap, const uint16_t* selpfx, uint32_t k) {
    int lo = 0;
#pragma unroll
    for (int step = RDPN_P / 64; step; step >>= 1)
        if (selpfx[lo + step] <= k) lo += step;
    const unsigned j = k - selpfx[lo];
    return lo * 32 + (int)__fns(selmap[lo], 0, (int)j + 1);
}

}  // namespace rdpn
