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
// The inlier test of the FP32 contract (oracle/pose_oracle.c:inlier_margin): with ncut = -sq_cut(thr),
//   margin = fma(dz, dz, fma(dy, dy, fma(dx, dx, ncut)))      inlier  <=>  margin < 0
// i.e. ||R a + t - c||^2 < cut with the threshold folded into the accumulation: the three squares cost three FFMA and
// the count one LEA.HI on the sign bit -- 7 instructions per (hypothesis, point) instead of 8 with FMUL + FSETP + IADD.
// (The arithmetic never yields -0 or a negative NaN here: x + (-x) is +0 in round-to-nearest and the GPU's NaN is
// 0x7fffffff, so "sign bit set" is exactly "margin < 0".)
__device__ __forceinline__ float margin_pt(float tx, float ty, float tz, float cx, float cy, float cz, float ncut) {
    const float dx = __fsub_rn(tx, cx), dy = __fsub_rn(ty, cy), dz = __fsub_rn(tz, cz);
    float m = __fmaf_rn(dx, dx, ncut);
    m = __fmaf_rn(dy, dy, m);
    m = __fmaf_rn(dz, dz, m);
    return m;
}
__device__ __forceinline__ void count_in(int& c, float margin) { c += (int)(__float_as_uint(margin) >> 31); }
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
__device__ __forceinline__ float margin(const float* P, float ax, float ay, float az, float cx, float cy, float cz, float ncut) {
    float x, y, z;
    xform(P, ax, ay, az, x, y, z);
    return margin_pt(x, y, z, cx, cy, cz, ncut);
}
__device__ __forceinline__ bool is_inlier(const float* P, float ax, float ay, float az, float cx, float cy, float cz, float ncut) {
    return margin(P, ax, ay, az, cx, cy, cz, ncut) < 0.f;
}

// ---------------------------------------------------------------------------------------------
// Inlier scoring of region-sorted correspondences, shared by the fused kernel and the pipeline's K2
// ---------------------------------------------------------------------------------------------
// Where a pass finds the FP32 pose of compacted hypothesis j and where its count goes:
//   PosePlanar  K2: rows of pose j at hyp[r * H + j] (conflict-free LDS.128), counts indexed by j
//   PoseListed  fused kernel: pose of hypothesis h = vlist[j] at hyp[3 h .. 3 h + 2], counts indexed by h
struct PosePlanar {
    const float4* hyp;
    int H;
    __device__ __forceinline__ int slot(int j) const { return j; }
    __device__ __forceinline__ void rows(int j, float4& r0, float4& r1, float4& r2) const {
        r0 = hyp[j]; r1 = hyp[H + j]; r2 = hyp[2 * H + j];
    }
};
struct PoseListed {
    const float4* hyp;
    const uint16_t* vlist;
    __device__ __forceinline__ int slot(int j) const { return (int)vlist[j]; }
    __device__ __forceinline__ void rows(int h, float4& r0, float4& r1, float4& r2) const {
        r0 = hyp[3 * h]; r1 = hyp[3 * h + 1]; r2 = hyp[3 * h + 2];
    }
};

// One pass: K hypotheses per lane (compacted indices j0 + 32 u + lane, u < K) against the staged slots [i0, i1)
// (c0 = slot index of pts[0]).  Every staged point (one LDS.128 broadcast) and every run header is shared by the lane's K
// hypotheses: 4 LDS + 32 K arithmetic instructions per four points, so the more hypotheses a lane carries the fewer
// instructions a pair costs.  Per run the transformed anchor R a + t once per hypothesis (misc.py:108-111 semantics,
// FP32 contract of oracle/pose_oracle.c).
// MEAN (select rule MIN_MEAN_ERR, misc.py:109,113): the sum of the residual norms over ALL points rides along in FP32
// (one MUFU.SQRT + FADD per pair); it only pre-selects the candidates whose mean error the refit kernel recomputes in FP64.
template <int K, class POSE, bool MEAN = false>
__device__ __forceinline__ void score_pass(const float4* __restrict__ pts, const float4* __restrict__ runtab, int nruns,
                                           const POSE pose, int j0, int nvalid, int i0, int i1, int c0, float cut, int* hcnt,
                                           float* herr = nullptr) {
    const int lane = threadIdx.x & 31;
    const float ncut = -cut;
    int sl[K], cnt[K];
    float es[K];
#pragma unroll
    for (int u = 0; u < K; ++u) {
        const int j = j0 + 32 * u + lane;
        sl[u] = pose.slot(j < nvalid ? j : nvalid - 1);  // idle lanes recompute a valid hypothesis
        cnt[u] = 0;
        es[u] = 0.f;
    }
#pragma unroll 1
    for (int k = 0; k < nruns; ++k) {
        const float4 rh = runtab[k];
        const unsigned se = __float_as_uint(rh.w);
        int p = max((int)(se & 0xFFFFu), i0) - c0;  // slot -> index into the staged chunk
        const int e = min((int)(se >> 16), i1) - c0;
        if (p >= e) continue;
        float tx[K], ty[K], tz[K];
#pragma unroll
        for (int u = 0; u < K; ++u) {
            float4 r0, r1, r2;
            pose.rows(sl[u], r0, r1, r2);
            const float P[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
            xform(P, rh.x, rh.y, rh.z, tx[u], ty[u], tz[u]);
        }
        if (!MEAN) {
#pragma unroll 1
            for (; p + 4 <= e; p += 4) {
                const float4 q0 = pts[p], q1 = pts[p + 1], q2 = pts[p + 2], q3 = pts[p + 3];
#pragma unroll
                for (int u = 0; u < K; ++u) {
                    count_in(cnt[u], margin_pt(tx[u], ty[u], tz[u], q0.x, q0.y, q0.z, ncut));
                    count_in(cnt[u], margin_pt(tx[u], ty[u], tz[u], q1.x, q1.y, q1.z, ncut));
                    count_in(cnt[u], margin_pt(tx[u], ty[u], tz[u], q2.x, q2.y, q2.z, ncut));
                    count_in(cnt[u], margin_pt(tx[u], ty[u], tz[u], q3.x, q3.y, q3.z, ncut));
                }
            }
        }
#pragma unroll 1
        for (; p < e; ++p) {
            const float4 q0 = pts[p];
#pragma unroll
            for (int u = 0; u < K; ++u) {
                count_in(cnt[u], margin_pt(tx[u], ty[u], tz[u], q0.x, q0.y, q0.z, ncut));
                if (MEAN) es[u] += sqrtf(resid2_pt(tx[u], ty[u], tz[u], q0.x, q0.y, q0.z));
            }
        }
    }
#pragma unroll
    for (int u = 0; u < K; ++u) {
        const int j = j0 + 32 * u + lane;
        if (j < nvalid && cnt[u]) atomicAdd(&hcnt[sl[u]], cnt[u]);
        if (MEAN && j < nvalid) atomicAdd(&herr[sl[u]], es[u]);
    }
}

// cost of one pass per four points, in issue slots (4 LDS + loop + 32 per hypothesis of a lane)
__device__ __forceinline__ int pass_cost(int k) { return 9 + 32 * k; }

// Warp `wi` of `nw` scores its share of nvalid hypotheses x staged slots [c0, c1): the hypotheses form passes of 4 per
// lane (128 per pass) plus one last pass of 1..4 per lane, the (pass, point) plane is cut into nw slices of equal
// cost, one per warp, whatever the number of valid hypotheses (the cut points are rounded to whole points the same way
// on both sides: nothing is lost or counted twice).  Counts are ADDED to hcnt (shared-memory atomics).
template <class POSE, bool MEAN = false>
__device__ __forceinline__ void score_slices(const float4* __restrict__ pts, const float4* __restrict__ runtab, int nruns,
                                             const POSE pose, int nvalid, int c0, int c1, float cut, int* hcnt, int wi, int nw,
                                             float* herr = nullptr) {
    const int nfull = nvalid >> 7;               // passes with 4 hypotheses per lane
    const int rem = nvalid - (nfull << 7);
    const int klast = (rem + 31) >> 5;           // 0 .. 4 hypotheses per lane in the last pass
    const int ctot = nfull * pass_cost(4) + (klast ? pass_cost(klast) : 0);
    const int m = c1 - c0;
    const long long tot = (long long)ctot * m;
    const int lo = (int)(tot * wi / nw), hi = (int)(tot * (wi + 1) / nw);  // tot < 2^31 for H <= 8192 at 1024 staged slots
    int off = 0;
    for (int ps = 0; ps < nfull + (klast ? 1 : 0); ++ps) {
        const int k = ps < nfull ? 4 : klast;
        const int len = pass_cost(k) * m;
        const int a0 = lo > off ? lo : off, a1 = hi < off + len ? hi : off + len;
        if (a0 < a1) {
            const int i0 = c0 + (a0 - off + pass_cost(k) - 1) / pass_cost(k);
            const int i1 = c0 + (a1 - off + pass_cost(k) - 1) / pass_cost(k);
            const int j0 = ps << 7;
            if (i0 < i1) {
                if (k == 4) score_pass<4, POSE, MEAN>(pts, runtab, nruns, pose, j0, nvalid, i0, i1, c0, cut, hcnt, herr);
                else if (k == 3) score_pass<3, POSE, MEAN>(pts, runtab, nruns, pose, j0, nvalid, i0, i1, c0, cut, hcnt, herr);
                else if (k == 2) score_pass<2, POSE, MEAN>(pts, runtab, nruns, pose, j0, nvalid, i0, i1, c0, cut, hcnt, herr);
                else score_pass<1, POSE, MEAN>(pts, runtab, nruns, pose, j0, nvalid, i0, i1, c0, cut, hcnt, herr);
            }
        }
        off += len;
    }
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
