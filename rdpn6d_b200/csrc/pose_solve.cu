// Dense-correspondence -> pose on sm_100a: the fused per-ROI solver (S1 gate + hypothesis generation +
// H x n inlier scoring + best selection + weighted Kabsch/Umeyama refit) and the batched Kabsch entry.
// (The materialising S1 kernel lives in correspond.cu.)
//
// One CTA owns one ROI.  Its five 16 KB planes (depth, coor_x/y/z, mask) and the 4 KB region-index
// plane are contiguous in HBM, so they are staged with 1-D bulk TMA copies (cp.async.bulk +
// mbarrier, no tensor map) issued by one thread; two CTAs are resident per SM so the copies of one
// ROI overlap the arithmetic of the other.  Everything after the copy stays in shared memory:
// nothing but the 12-float pose (and optional diagnostics) goes back to HBM.
//
// Arithmetic contracts (oracle/pose_oracle.py, oracle/pose_oracle.c):
//   S1       FP32, one IEEE op per step: X = ((u - cx') * d) / fx'   (data_loader.py:563-576),
//            delta = (coor - 0.5) * extent (gdrn_evaluator.py:102-105), gate strict '>'
//            (gdrn_evaluator.py:110-117), L1 mask (m - min) / (max - min) (engine_utils.py:123-128).
//   hyp      FP64 closed-form 3-pair Kabsch, rounded once to FP32 (transform.py:940-951 semantics).
//   scoring  FP32 with explicit FMA order, inlier <=> d2 < sq_cut(thr)  (misc.py:108-111).
//   best     strictly greater count and >= min_inliers, earliest wins (misc.py:121); optional
//            adaptive stop (misc.py:134-138).
//   refit    FP64 accumulation of centroids / cross-covariance over the inliers, Horn rotation,
//            Umeyama scale (transform.py:921-928, 942-948, 971-979).
#include "common.cuh"
#include "kabsch_math.cuh"

#include <float.h>
#include <math.h>
#include <string.h>

namespace rdpn {
extern unsigned long long g_launch_count;

constexpr int ST = 256;              // threads per CTA
constexpr int SW = ST / 32;          // warps
constexpr int QPT = RDPN_P / 4 / ST;  // pixel quads per thread (4)

struct RoiConst {
    float fx, fy, cx, cy;
    float ext[3];
    float gthr[3];
    float div;  // depth divisor or 0 (= none)
    float mn, mx;
};

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

__device__ __forceinline__ float mask_prob(float m, int mode, float mn, float mx) {
    if (mode == RDPN_MASK_L1) return __fdiv_rn(__fsub_rn(m, mn), __fsub_rn(mx, mn));
    if (mode == RDPN_MASK_BCE) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-m)));
    return m;
}

// ---------------------------------------------------------------------------------------------
// fused solver
// ---------------------------------------------------------------------------------------------
// Per ROI (one CTA, 256 threads, two CTAs resident per SM):
//   1  bulk-TMA stage of the raw planes                      (stage_roi)
//   2  mask min/max, then the GATE only (no divisions): 16 pixels per thread -> selection bits
//   3  deterministic counting sort of the gated pixels by region id (warp match_any ranks + per-warp
//      bucket cursors) -> pix[slot]; bucket r occupies slots [bstart[r], bstart[r+1])
//   4  hypothesis generation straight from the raw planes (3 pixels each, FP64 closed form)
//   5  correspondence STAGING: one thread per gated slot computes (cam xyz, w) with the exact S1
//      arithmetic and the list is written in place over the raw planes as float4 AoS
//   6  scoring: one thread per hypothesis; per region bucket the transformed anchor R a + t is
//      computed once (9 FMA) and every point of the bucket costs 1 LDS.128 + 6 FP32 + compare
//   7  best hypothesis, FP64 refit sums, closed-form rotation, outputs
struct RoiGate {
    float hi, lo;     // fast mask filter: a > hi -> in, a < lo -> out, else exact test
    double cut;       // exact: (double)a > / >= (double)b * cut
    float b;          // max - min
    int incl;         // 1: >= (odd mantissa of the threshold), 0: >
};

template <bool DENSE>
struct __align__(128) FusedSmem {
    float tile[5][RDPN_P];                 // raw depth, coor_x, coor_y, coor_z, mask; later tile[0..3] = float4 camw[n]
    float4 objS[DENSE ? RDPN_P : 1];       // dense: object-side AoS (x,y,z,-)
    uint8_t rid[DENSE ? 16 : RDPN_P];
    uint16_t pix[RDPN_P];                  // slot -> pixel
    uint32_t selmap[RDPN_P / 32];          // gate bitmap by pixel
    uint64_t bar;
    RoiConst rc;
    RoiGate gate;
    float red_f[2][SW];
    double red_d[SW][12];
    double bc_d[16];
    int n_sel;
    int n_runs;
    int best_h;
    int n_best;
    int h_eff;
    float pose[12];
    unsigned long long red_k[SW];
    int red_i[SW];
};

// one pixel of S1 with the exact oracle arithmetic; returns cam (and obj in dense mode), d = depth used
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
    }
}

// (mask_prob(m) > mask_thr) without the division for the L1 mode: fl(a/b) > thr  <=>  a/b > (>=) cut
// where cut is the midpoint between thr and its FP32 successor (ties-to-even decides the inclusivity).
__device__ __forceinline__ bool mask_pass(float m, int mode, float thr, const RoiConst& rc, const RoiGate& g) {
    if (mode == RDPN_MASK_L1) {
        if (!(g.b > 0.f)) return false;  // flat mask: 0/0 = NaN never passes (engine_utils.py:128 has no eps)
        const float a = __fsub_rn(m, rc.mn);
        if (a > g.hi) return true;
        if (a < g.lo) return false;
        const double l = (double)a, r = __dmul_rn((double)g.b, g.cut);
        return g.incl ? (l >= r) : (l > r);  // NaN (flat mask: 0/0) -> false
    }
    return mask_prob(m, mode, 0.f, 0.f) > thr;
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

// block-wide sum of NV doubles; result valid in every thread (via s.bc_d[0..NV)).
template <typename SM, int NV>
__device__ __forceinline__ void block_sum(SM& s, double (&v)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) s.red_d[warp][i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double a = 0.0;
        for (int w = 0; w < SW; ++w) a += s.red_d[w][threadIdx.x];
        s.bc_d[threadIdx.x] = a;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = s.bc_d[i];
    __syncthreads();
}

struct FusedLayout {  // byte offsets of the dynamic tail behind FusedSmem
    int anchors;      // float4[R]
    int runtab;       // float4[R]: non-empty buckets (anchor xyz, start | end << 16)
    int bstart;       // int[RB + 1]
    int wrun;         // uint16[SW][RB]     (RB = R + 1: last bucket collects the unselected lanes)
    int hyp;          // float[H][12]
    int hcnt;         // int[H]
    int total;
};

template <bool DENSE>
__global__ void __launch_bounds__(ST, 2) pose_solve_kernel(SolveArgs a, FusedLayout lay) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem<DENSE>& s = *reinterpret_cast<FusedSmem<DENSE>*>(smem_raw);
    float4* anchors = reinterpret_cast<float4*>(smem_raw + lay.anchors);
    float4* runtab = reinterpret_cast<float4*>(smem_raw + lay.runtab);
    int* bstart = reinterpret_cast<int*>(smem_raw + lay.bstart);
    uint16_t* wrun = reinterpret_cast<uint16_t*>(smem_raw + lay.wrun);
    float* hyp = reinterpret_cast<float*>(smem_raw + lay.hyp);  // [H][12]
    int* hcnt = reinterpret_cast<int*>(smem_raw + lay.hcnt);    // [H] counts, -1 = invalid
    const int H = a.prm.num_hyp;
    const int R = DENSE ? 1 : a.in.num_regions;
    const int RB = R + 1;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const rdpn_roi_inputs& in = a.in;

    // first hypothesis triplet of this thread: issue the global loads before anything waits
    int pre_ii[3] = {-1, -1, -1};
    if (t < H) {
        const int32_t* ip = a.hyp_idx + ((size_t)b * H + t) * 3;
        pre_ii[0] = ip[0]; pre_ii[1] = ip[1]; pre_ii[2] = ip[2];
    }
    // ---- 1: stage ----
    if (t == 0) {
        mbar_init(&s.bar, 1);
        mbar_fence_init();
        const size_t o = (size_t)b * RDPN_P;
        const uint32_t plane = RDPN_P * sizeof(float);
        mbar_expect_tx(&s.bar, 5 * plane + (DENSE ? 0 : RDPN_P));
        bulk_g2s(s.tile[4], in.mask + o, plane, &s.bar);
        bulk_g2s(s.tile[0], in.depth + o, plane, &s.bar);
        bulk_g2s(s.tile[1], in.coor_x + o, plane, &s.bar);
        bulk_g2s(s.tile[2], in.coor_y + o, plane, &s.bar);
        bulk_g2s(s.tile[3], in.coor_z + o, plane, &s.bar);
        if (!DENSE) bulk_g2s(s.rid, in.region_idx + o, RDPN_P, &s.bar);
        RoiConst& rc = s.rc;
        rc.fx = in.Kp[4 * b + 0];
        rc.fy = in.Kp[4 * b + 1];
        rc.cx = in.Kp[4 * b + 2];
        rc.cy = in.Kp[4 * b + 3];
        for (int c = 0; c < 3; ++c) {
            const float e = in.extent[3 * b + c];
            rc.ext[c] = e;
            rc.gthr[c] = (float)(0.0001 * (double)e);  // gdrn_evaluator.py:112-114 under numpy-1.23 promotion
        }
        rc.div = in.depth_div ? in.depth_div[b] : 0.f;
        s.best_h = -1;
        s.n_best = 0;
        s.h_eff = H;
    }
    if (!DENSE)
        for (int r = t; r < R; r += ST) {
            const float* ap = in.anchors + ((size_t)b * R + r) * 3;
            anchors[r] = make_float4(ap[0], ap[1], ap[2], 0.f);
        }
    for (int i = t; i < SW * RB; i += ST) wrun[i] = 0;
    if (t < RDPN_P / 32) s.selmap[t] = 0u;
    if (a.out.inlier_mask) {  // zero-fill; inliers are scattered in after the last refit
        uint4* im = reinterpret_cast<uint4*>(a.out.inlier_mask + (size_t)b * RDPN_P);
        im[t] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    mbar_wait(&s.bar, 0);

    // ---- 2: mask min/max + gate.  Thread owns quads q = 32*(SW*k + warp) + lane (k = 0..3): every warp gets
    //         two image rows out of each 16, so the rows the object covers are spread over all warps ----
    if (in.mask_mode == RDPN_MASK_L1) {
        float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            const float4 m4 = reinterpret_cast<const float4*>(s.tile[4])[32 * (SW * k + warp) + lane];
            mn = fminf(fminf(fminf(mn, m4.x), fminf(m4.y, m4.z)), m4.w);
            mx = fmaxf(fmaxf(fmaxf(mx, m4.x), fmaxf(m4.y, m4.z)), m4.w);
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) { s.red_f[0][warp] = mn; s.red_f[1][warp] = mx; }
        __syncthreads();
        if (t == 0) {
            float lo = s.red_f[0][0], hi = s.red_f[1][0];
            for (int w = 1; w < SW; ++w) { lo = fminf(lo, s.red_f[0][w]); hi = fmaxf(hi, s.red_f[1][w]); }
            s.rc.mn = lo;
            s.rc.mx = hi;
            RoiGate& g = s.gate;
            g.b = __fsub_rn(hi, lo);
            g.cut = a.mask_cut;
            g.incl = a.mask_cut_incl;
            const float bt = g.b * in.mask_thr;
            const bool filt = in.mask_thr > 1e-30f && in.mask_thr < 1e30f;
            g.hi = filt ? bt * 1.000002f : INFINITY;
            g.lo = filt ? bt * 0.999998f : -INFINITY;
        }
        __syncthreads();
    }
    const RoiConst rc = s.rc;
    const RoiGate gate = s.gate;
    unsigned selbits = 0u;
#pragma unroll 1
    for (int k = 0; k < QPT; ++k) {
        const int q = 32 * (SW * k + warp) + lane;
        const float4 dq = reinterpret_cast<const float4*>(s.tile[0])[q];
        const float4 xq = reinterpret_cast<const float4*>(s.tile[1])[q];
        const float4 yq = reinterpret_cast<const float4*>(s.tile[2])[q];
        const float4 zq = reinterpret_cast<const float4*>(s.tile[3])[q];
        const float4 m4 = reinterpret_cast<const float4*>(s.tile[4])[q];
        const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
        const float cxn[4] = {xq.x, xq.y, xq.z, xq.w};
        const float cyn[4] = {yq.x, yq.y, yq.z, yq.w};
        const float czn[4] = {zq.x, zq.y, zq.z, zq.w};
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
        unsigned nib = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float d = dd[j];
            if (rc.div != 0.f) {  // zero lanes would drag the warp through div.rn's slow path: 0 / f = +-0
                const float q = __fdiv_rn(d == 0.f ? 1.f : d, rc.div);
                d = d == 0.f ? __fmul_rn(d, copysignf(1.f, rc.div)) : q;
            }
            const float dx = __fmul_rn(__fsub_rn(cxn[j], 0.5f), rc.ext[0]);
            const float dy = __fmul_rn(__fsub_rn(cyn[j], 0.5f), rc.ext[1]);
            const float dz = __fmul_rn(__fsub_rn(czn[j], 0.5f), rc.ext[2]);
            bool sel = (fabsf(dx) > rc.gthr[0]) && (fabsf(dy) > rc.gthr[1]) && (fabsf(dz) > rc.gthr[2]) && (d > 0.f);
            if (sel) sel = mask_pass(mm[j], in.mask_mode, in.mask_thr, rc, gate);  // gdrn_evaluator.py:110-117 (+ depth)
            nib |= (sel ? 1u : 0u) << j;
        }
        selbits |= nib << (4 * k);
        // publish the gate bitmap (4 bits per quad, 8 quads per word)
        unsigned wbits = nib << (4 * (lane & 7));
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 4);
        if ((lane & 7) == 0) s.selmap[q >> 3] = wbits;
    }

    // ---- 3: counting sort by region, deterministic order (warp, k, j, lane) ----
    // pass A: per-warp bucket histogram
    uint16_t* myrun = wrun + warp * RB;
#pragma unroll 1
    for (int k = 0; k < QPT; ++k) {
        const unsigned nib = (selbits >> (4 * k)) & 0xFu;
        if (__ballot_sync(0xffffffffu, nib != 0u) == 0u) continue;
        const uchar4 r4 = DENSE ? make_uchar4(0, 0, 0, 0) : reinterpret_cast<const uchar4*>(s.rid)[32 * (SW * k + warp) + lane];
        const uint8_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool sel = (nib >> j) & 1u;
            if (__ballot_sync(0xffffffffu, sel) == 0u) continue;
            const unsigned key = sel ? (unsigned)rr[j] : (unsigned)R;
            const unsigned m = __match_any_sync(0xffffffffu, key);
            if (sel && lane == __ffs(m) - 1) myrun[key] += (uint16_t)__popc(m);
            __syncwarp();
        }
    }
    __syncthreads();
    // bucket starts and per-warp cursors: warp 0, RPL consecutive buckets per lane, one warp scan
    if (warp == 0) {
        const int RPL = (R + 31) / 32;
        int loc = 0;
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R)
                for (int w = 0; w < SW; ++w) loc += wrun[w * RB + r];
        }
        int x = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        int run = x - loc;
        // run table: one entry per NON-EMPTY bucket = (anchor xyz, start | end << 16), in bucket order
        int ne = 0;
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R) {
                int tot = 0;
                for (int w = 0; w < SW; ++w) tot += wrun[w * RB + r];
                ne += tot > 0 ? 1 : 0;
            }
        }
        int kx = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, kx, o);
            if (lane >= o) kx += y;
        }
        int kk = kx - ne;
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R) {
                bstart[r] = run;
                const int start = run;
                for (int w = 0; w < SW; ++w) {
                    const int c = wrun[w * RB + r];
                    wrun[w * RB + r] = (uint16_t)run;
                    run += c;
                }
                if (run > start) {
                    float4 hd = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : anchors[r];
                    hd.w = __uint_as_float((unsigned)start | ((unsigned)run << 16));
                    runtab[kk++] = hd;
                }
            }
        }
        if (lane == 31) s.n_runs = kx;
        if (lane == 31) { bstart[R] = x; s.n_sel = x; }
    }
    __syncthreads();
    // pass B: assign slots
#pragma unroll 1
    for (int k = 0; k < QPT; ++k) {
        const unsigned nib = (selbits >> (4 * k)) & 0xFu;
        if (__ballot_sync(0xffffffffu, nib != 0u) == 0u) continue;
        const uchar4 r4 = DENSE ? make_uchar4(0, 0, 0, 0) : reinterpret_cast<const uchar4*>(s.rid)[32 * (SW * k + warp) + lane];
        const uint8_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool sel = (nib >> j) & 1u;
            if (__ballot_sync(0xffffffffu, sel) == 0u) continue;
            const unsigned key = sel ? (unsigned)rr[j] : (unsigned)R;
            const unsigned m = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(m) - 1;
            int cur = 0;
            if (sel && lane == leader) {
                cur = myrun[key];
                myrun[key] = (uint16_t)(cur + __popc(m));
            }
            cur = __shfl_sync(0xffffffffu, cur, leader);
            if (sel) s.pix[cur + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(4 * (32 * (SW * k + warp) + lane) + j);
            __syncwarp();
        }
    }

    // ---- 4: hypothesis generation from the raw planes (FP64 closed form), one hypothesis per thread ----
    for (int h = t; h < H; h += ST) {
        int ii[3] = {pre_ii[0], pre_ii[1], pre_ii[2]};
        if (h != t) {
            const int32_t* ip = a.hyp_idx + ((size_t)b * H + h) * 3;
            ii[0] = ip[0]; ii[1] = ip[1]; ii[2] = ip[2];
        }
        bool ok = ((unsigned)ii[0] < RDPN_P) && ((unsigned)ii[1] < RDPN_P) && ((unsigned)ii[2] < RDPN_P);
        float* P = hyp + (size_t)h * 12;
        if (ok) {
#pragma unroll
            for (int v = 0; v < 3; ++v) ok = ok && ((s.selmap[ii[v] >> 5] >> (ii[v] & 31)) & 1u);
        }
        if (ok) {
            double A[3][3], C[3][3];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const int p = ii[v];
                float cam[3], obj[3];
                pixel_s1<DENSE>(rc, p, s.tile[0][p], s.tile[1][p], s.tile[2][p], s.tile[3][p], cam, obj);
                C[v][0] = (double)cam[0]; C[v][1] = (double)cam[1]; C[v][2] = (double)cam[2];
                if (DENSE) {
                    A[v][0] = (double)obj[0]; A[v][1] = (double)obj[1]; A[v][2] = (double)obj[2];
                } else {
                    const float4 an = anchors[s.rid[p]];
                    A[v][0] = (double)an.x; A[v][1] = (double)an.y; A[v][2] = (double)an.z;
                }
            }
            ok = triangle_ok(A[0], A[1], A[2]) && triangle_ok(C[0], C[1], C[2]);
            if (ok) {
                double Rt[12];
                kabsch3(A, C, Rt);
#pragma unroll
                for (int i = 0; i < 12; ++i) P[i] = (float)Rt[i];
            }
        }
        if (!ok) {
#pragma unroll
            for (int i = 0; i < 12; ++i) P[i] = 0.f;
        }
        hcnt[h] = ok ? 0 : -1;
    }
    __syncthreads();  // slots, hypotheses visible; raw planes may be overwritten after the next barrier
    const int n = s.n_sel;

    // ---- 5: staging: thread per slot computes (cam xyz, w) with the exact S1 arithmetic; the AoS list then
    //         replaces raw planes (n <= 1024: the mask plane; larger: the depth/coor planes) and the
    //         region-id plane is rewritten in slot order.  All raw reads precede the barrier.
    const bool small_n = n <= 4 * ST;
    float4* camw_w = reinterpret_cast<float4*>(small_n ? s.tile[4] : s.tile[0]);
    if (small_n) {
        float4 cw[4];
        uint8_t rb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int sl = u * ST + t;
            if (sl < n) {
                const int p = s.pix[sl];
                float cam[3], obj[3];
                pixel_s1<DENSE>(rc, p, s.tile[0][p], s.tile[1][p], s.tile[2][p], s.tile[3][p], cam, obj);
                const float w = a.prm.weighted ? mask_prob(s.tile[4][p], in.mask_mode, rc.mn, rc.mx) : 1.f;
                cw[u] = make_float4(cam[0], cam[1], cam[2], w);
                rb[u] = DENSE ? (uint8_t)0 : s.rid[p];
                if (DENSE) s.objS[sl] = make_float4(obj[0], obj[1], obj[2], 0.f);
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int sl = u * ST + t;
            if (sl < n) {
                camw_w[sl] = cw[u];
                if (!DENSE) s.rid[sl] = rb[u];
            }
        }
    } else {  // rare: more than a quarter of the ROI is gated; same arithmetic, results parked in local memory
        float4 cwl[RDPN_P / ST];
        uint8_t rbl[RDPN_P / ST];
#pragma unroll 1
        for (int u = 0; u < RDPN_P / ST; ++u) {
            const int sl = u * ST + t;
            if (sl < n) {
                const int p = s.pix[sl];
                float cam[3], obj[3];
                pixel_s1<DENSE>(rc, p, s.tile[0][p], s.tile[1][p], s.tile[2][p], s.tile[3][p], cam, obj);
                const float w = a.prm.weighted ? mask_prob(s.tile[4][p], in.mask_mode, rc.mn, rc.mx) : 1.f;
                cwl[u] = make_float4(cam[0], cam[1], cam[2], w);
                rbl[u] = DENSE ? (uint8_t)0 : s.rid[p];
                if (DENSE) s.objS[sl] = make_float4(obj[0], obj[1], obj[2], 0.f);
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int u = 0; u < RDPN_P / ST; ++u) {
            const int sl = u * ST + t;
            if (sl < n) {
                camw_w[sl] = cwl[u];
                if (!DENSE) s.rid[sl] = rbl[u];
            }
        }
    }
    __syncthreads();
    const float4* camw = camw_w;
    const uint8_t* srid = s.rid;  // region id by slot (non-decreasing)

    if (a.out.n_sel && t == 0) a.out.n_sel[b] = n;
    const bool enough = n >= a.prm.min_pts;

    // ---- 6: inlier scoring ----
    if (enough) {
        const int S = (H >= ST) ? 1 : (ST / H);
        const float cut = a.sq_cut;
        for (int item = t; item < H * S; item += ST) {
            const int h = item % H, seg = item / H;
            if (hcnt[h] < 0) continue;
            float P[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) P[i] = hyp[(size_t)h * 12 + i];
            int c = 0;
            if (DENSE) {
                const int i0 = (int)(((long long)n * seg) / S), i1 = (int)(((long long)n * (seg + 1)) / S);
#pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const float4 cp = camw[i];
                    const float4 ap = s.objS[i];
                    count_if_lt(c, resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z), cut);
                }
            } else {
                // slots are sorted by region: per run (non-empty bucket) the transformed anchor R a + t is
                // computed once; every point then costs 1 LDS.128 + 6 FP32 + compare + predicated add.
                const int nruns = s.n_runs;
                for (int k = seg; k < nruns; k += S) {  // runs interleaved over the segments
                    const float4 hd = runtab[k];
                    const unsigned se = __float_as_uint(hd.w);
                    int i = (int)(se & 0xFFFFu);
                    const int e = (int)(se >> 16);
                    float tx, ty, tz;
                    xform(P, hd.x, hd.y, hd.z, tx, ty, tz);
                    for (; i + 4 <= e; i += 4) {
                        const float4 c0 = camw[i], c1 = camw[i + 1], c2 = camw[i + 2], c3 = camw[i + 3];
                        count_if_lt(c, resid2_pt(tx, ty, tz, c0.x, c0.y, c0.z), cut);
                        count_if_lt(c, resid2_pt(tx, ty, tz, c1.x, c1.y, c1.z), cut);
                        count_if_lt(c, resid2_pt(tx, ty, tz, c2.x, c2.y, c2.z), cut);
                        count_if_lt(c, resid2_pt(tx, ty, tz, c3.x, c3.y, c3.z), cut);
                    }
                    for (; i < e; ++i) {
                        const float4 c0 = camw[i];
                        count_if_lt(c, resid2_pt(tx, ty, tz, c0.x, c0.y, c0.z), cut);
                    }
                }
            }
            if (S == 1) hcnt[h] = c;
            else atomicAdd(&hcnt[h], c);
        }
    }
    __syncthreads();

    // ---- 7a: best hypothesis (misc.py:121) with optional adaptive stop (misc.py:134-138) ----
    if (enough) {
        if (a.prm.adaptive) {
            // i_ransac(h) = number of valid hypotheses in [0,h]; stop after the first h with
            // i_ransac > max(k, min_iter), k = log10(1-conf)/log10(1-w^10), w = count/n.
            const double lc = log10(1.0 - (double)a.prm.confidence);
            int running = 0;
            for (int h0 = 0; h0 < H; h0 += ST) {
                const int h = h0 + t;
                const int v = (h < H && hcnt[h] >= 0) ? 1 : 0;
                int x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                if (lane == 31) s.red_i[warp] = x;
                __syncthreads();
                int wbase = 0;
                for (int w = 0; w < warp; ++w) wbase += s.red_i[w];
                int tot = 0;
                for (int w = 0; w < SW; ++w) tot += s.red_i[w];
                const int i_ransac = running + wbase + x;
                if (v) {
                    const double wr = (double)hcnt[h] / (double)n;
                    const double k = lc / log10(1.0 - pow(wr, 10.0));
                    const double lim = fmax(k, (double)a.prm.min_iter);
                    if ((double)i_ransac > lim) atomicMin(&s.h_eff, h + 1);
                }
                running += tot;
                __syncthreads();
            }
        }
        const int heff = s.h_eff;
        unsigned long long key = 0ull;
        for (int h = t; h < heff; h += ST) {
            const int c = hcnt[h];
            if (c >= a.prm.min_inliers && c > 0) {
                const unsigned long long k = ((unsigned long long)(unsigned)c << 32) | (unsigned)(0x7FFFFFFF - h);
                key = k > key ? k : key;
            }
        }
        key = warp_max_u64(key);
        if (lane == 0) s.red_k[warp] = key;
        __syncthreads();
        if (t == 0) {
            unsigned long long k = 0ull;
            for (int w = 0; w < SW; ++w) k = s.red_k[w] > k ? s.red_k[w] : k;
            if (k) {
                s.best_h = 0x7FFFFFFF - (int)(k & 0xFFFFFFFFull);
                s.n_best = (int)(k >> 32);
            }
        }
        __syncthreads();
    }
    const int best = s.best_h;

    // optional diagnostics
    if (a.out.hyp_counts)
        for (int h = t; h < H; h += ST) a.out.hyp_counts[(size_t)b * H + h] = enough ? max(hcnt[h], 0) : 0;
    if (a.out.hyp_poses)
        for (int i = t; i < H * 12; i += ST) a.out.hyp_poses[(size_t)b * H * 12 + i] = enough ? hyp[i] : 0.f;

    if (!enough || best < 0) {
        if (t < 12) a.out.pose[(size_t)b * 12 + t] = -100.f;  // gdrn_evaluator.py:395
        if (t == 0) {
            a.out.n_inliers[b] = 0;
            a.out.status[b] = enough ? RDPN_STATUS_NO_CONSENSUS : RDPN_STATUS_FEW_POINTS;
            if (a.out.best_h) a.out.best_h[b] = -1;
            if (a.out.scale) a.out.scale[b] = 1.f;
        }
        return;
    }

    // ---- 7b: refit on the inliers (misc.py:123-126 -> transform.py:913-980), FP64 accumulation ----
    if (t < 12) s.pose[t] = hyp[(size_t)best * 12 + t];
    __syncthreads();
    const float cut = a.sq_cut;
    float out_scale = 1.f;
    const int iters = a.prm.refit_iters < 1 ? 1 : a.prm.refit_iters;
    for (int it = 0; it < iters; ++it) {
        float P[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) P[i] = s.pose[i];
        // pass 1: weighted centroids; thread handles slots t, t+ST, ... (<= 16 of them)
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        unsigned inl_bits = 0u;
        int slot = 0;
        for (int i = t; i < n; i += ST, ++slot) {
            const float4 cp = camw[i];
            float4 ap;
            if (DENSE) ap = s.objS[i];
            else ap = anchors[srid[i]];
            if (resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z) < cut) {
                inl_bits |= 1u << slot;
                const double w = a.prm.weighted ? (double)cp.w : 1.0;
                acc[0] += w;
                acc[1] += w * cp.x; acc[2] += w * cp.y; acc[3] += w * cp.z;
                acc[4] += w * ap.x; acc[5] += w * ap.y; acc[6] += w * ap.z;
            }
        }
        int tot_inl = warp_sum(__popc(inl_bits));
        if (lane == 0) s.red_i[warp] = tot_inl;
        block_sum<FusedSmem<DENSE>, 7>(s, acc);  // contains the barriers that publish red_i
        tot_inl = 0;
        for (int w = 0; w < SW; ++w) tot_inl += s.red_i[w];
        if (tot_inl < 3) break;  // uniform across the block
        const double isw = 1.0 / acc[0];
        const double mc[3] = {acc[1] * isw, acc[2] * isw, acc[3] * isw};
        const double ma[3] = {acc[4] * isw, acc[5] * isw, acc[6] * isw};
        // pass 2: cross-covariance about the centroids (+ spreads for the Umeyama scale)
        double cov[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        slot = 0;
        for (int i = t; i < n; i += ST, ++slot) {
            if (inl_bits & (1u << slot)) {
                const float4 cp = camw[i];
                float4 ap;
                if (DENSE) ap = s.objS[i];
                else ap = anchors[srid[i]];
                const double w = a.prm.weighted ? (double)cp.w : 1.0;
                const double c0 = cp.x - mc[0], c1 = cp.y - mc[1], c2 = cp.z - mc[2];
                const double a0 = ap.x - ma[0], a1 = ap.y - ma[1], a2 = ap.z - ma[2];
                cov[0] += w * c0 * a0; cov[1] += w * c0 * a1; cov[2] += w * c0 * a2;
                cov[3] += w * c1 * a0; cov[4] += w * c1 * a1; cov[5] += w * c1 * a2;
                cov[6] += w * c2 * a0; cov[7] += w * c2 * a1; cov[8] += w * c2 * a2;
                cov[9] += w * (c0 * c0 + c1 * c1 + c2 * c2);
                cov[10] += w * (a0 * a0 + a1 * a1 + a2 * a2);
            }
        }
        block_sum<FusedSmem<DENSE>, 11>(s, cov);
        if (t == 0) {
            double Rm[9];
            rotation_from_cov(cov, cov[10], cov[9], Rm);
            double sc = 1.0;
            if (a.prm.with_scale) sc = sqrt(cov[9] / cov[10]);  // transform.py:971-975
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double tr = mc[r] - sc * (Rm[3 * r] * ma[0] + Rm[3 * r + 1] * ma[1] + Rm[3 * r + 2] * ma[2]);
                s.pose[4 * r + 0] = (float)(sc * Rm[3 * r + 0]);
                s.pose[4 * r + 1] = (float)(sc * Rm[3 * r + 1]);
                s.pose[4 * r + 2] = (float)(sc * Rm[3 * r + 2]);
                s.pose[4 * r + 3] = (float)tr;
            }
            s.bc_d[15] = sc;
        }
        if (a.out.inlier_mask && it == iters - 1) {  // the inlier set used by the last refit
            slot = 0;
            for (int i = t; i < n; i += ST, ++slot)
                if (inl_bits & (1u << slot)) a.out.inlier_mask[(size_t)b * RDPN_P + s.pix[i]] = 1;
        }
        __syncthreads();
        out_scale = (float)s.bc_d[15];
    }

    // ---- outputs (+ translation sanity, gdrn_evaluator.py:293-296) ----
    if (t == 0) {
        int status = RDPN_STATUS_OK;
        const float tx = s.pose[3], ty = s.pose[7], tz = s.pose[11];
        if (a.t_net) {
            const double d0 = (double)a.t_net[3 * b] - tx, d1 = (double)a.t_net[3 * b + 1] - ty,
                         d2 = (double)a.t_net[3 * b + 2] - tz;
            if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 1.0) {
                status = RDPN_STATUS_T_SANITY;
                s.pose[3] = a.t_net[3 * b];
                s.pose[7] = a.t_net[3 * b + 1];
                s.pose[11] = a.t_net[3 * b + 2];
            }
        }
        a.out.n_inliers[b] = s.n_best;
        a.out.status[b] = status;
        if (a.out.best_h) a.out.best_h[b] = best;
        if (a.out.scale) a.out.scale[b] = out_scale;
    }
    __syncthreads();
    if (t < 12) a.out.pose[(size_t)b * 12 + t] = s.pose[t];
}

// ---------------------------------------------------------------------------------------------
// batched Kabsch / Umeyama (B4): one CTA per problem, points streamed from HBM
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kabsch_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                                                      const float* __restrict__ w, int N, int with_scale,
                                                      float* __restrict__ outM, float* __restrict__ outS) {
    __shared__ double red[8][12];
    __shared__ double bc[12];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* a = src + (size_t)b * N * 3;
    const float* c = dst + (size_t)b * N * 3;
    const float* ww = w ? w + (size_t)b * N : nullptr;
    auto bsum = [&](double* v, int nv) {
        for (int i = 0; i < nv; ++i) {
            v[i] = warp_sum(v[i]);
            if (lane == 0) red[warp][i] = v[i];
        }
        __syncthreads();
        if (t < nv) {
            double x = 0.0;
            for (int k = 0; k < 8; ++k) x += red[k][t];
            bc[t] = x;
        }
        __syncthreads();
        for (int i = 0; i < nv; ++i) v[i] = bc[i];
        __syncthreads();
    };
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = t; i < N; i += 256) {
        const double wi = ww ? (double)ww[i] : 1.0;
        acc[0] += wi;
        acc[1] += wi * c[3 * i]; acc[2] += wi * c[3 * i + 1]; acc[3] += wi * c[3 * i + 2];
        acc[4] += wi * a[3 * i]; acc[5] += wi * a[3 * i + 1]; acc[6] += wi * a[3 * i + 2];
    }
    bsum(acc, 7);
    const double isw = 1.0 / acc[0];
    const double mc[3] = {acc[1] * isw, acc[2] * isw, acc[3] * isw};
    const double ma[3] = {acc[4] * isw, acc[5] * isw, acc[6] * isw};
    double cov[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = t; i < N; i += 256) {
        const double wi = ww ? (double)ww[i] : 1.0;
        const double c0 = c[3 * i] - mc[0], c1 = c[3 * i + 1] - mc[1], c2 = c[3 * i + 2] - mc[2];
        const double a0 = a[3 * i] - ma[0], a1 = a[3 * i + 1] - ma[1], a2 = a[3 * i + 2] - ma[2];
        cov[0] += wi * c0 * a0; cov[1] += wi * c0 * a1; cov[2] += wi * c0 * a2;
        cov[3] += wi * c1 * a0; cov[4] += wi * c1 * a1; cov[5] += wi * c1 * a2;
        cov[6] += wi * c2 * a0; cov[7] += wi * c2 * a1; cov[8] += wi * c2 * a2;
        cov[9] += wi * (c0 * c0 + c1 * c1 + c2 * c2);
        cov[10] += wi * (a0 * a0 + a1 * a1 + a2 * a2);
    }
    bsum(cov, 11);
    if (t == 0) {
        double R[9];
        rotation_from_cov(cov, cov[10], cov[9], R);
        const double sc = with_scale ? sqrt(cov[9] / cov[10]) : 1.0;
        for (int r = 0; r < 3; ++r) {
            outM[(size_t)b * 12 + 4 * r + 0] = (float)(sc * R[3 * r + 0]);
            outM[(size_t)b * 12 + 4 * r + 1] = (float)(sc * R[3 * r + 1]);
            outM[(size_t)b * 12 + 4 * r + 2] = (float)(sc * R[3 * r + 2]);
            outM[(size_t)b * 12 + 4 * r + 3] =
                (float)(mc[r] - sc * (R[3 * r] * ma[0] + R[3 * r + 1] * ma[1] + R[3 * r + 2] * ma[2]));
        }
        if (outS) outS[b] = (float)sc;
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
static float host_sq_cut(float thr) {
    if (!(thr > 0.f)) return 0.f;
    float x = thr * thr;
    while (sqrtf(x) >= thr && x > 0.f) x = nextafterf(x, 0.f);
    while (sqrtf(x) < thr) x = nextafterf(x, INFINITY);
    return x;
}

int check_roi_inputs(const rdpn_roi_inputs* in, bool* dense) {
    if (!in || in->B <= 0) return RDPN_E_BADARG;
    if (!in->depth || !in->Kp || !in->coor_x || !in->coor_y || !in->coor_z || !in->mask || !in->extent)
        return RDPN_E_BADARG;
    if ((in->region_idx == nullptr) != (in->anchors == nullptr)) return RDPN_E_BADARG;
    *dense = in->region_idx == nullptr;
    if (!*dense && (in->num_regions <= 0 || in->num_regions > 255)) return RDPN_E_BADARG;
    if (in->mask_mode < 0 || in->mask_mode > 2) return RDPN_E_BADARG;
    const uintptr_t al = (uintptr_t)in->depth | (uintptr_t)in->coor_x | (uintptr_t)in->coor_y |
                         (uintptr_t)in->coor_z | (uintptr_t)in->mask | (uintptr_t)in->region_idx;
    if (al & 15) return RDPN_E_ALIGN;
    return 0;
}

template <bool DENSE>
static int launch_solve(const SolveArgs& a, cudaStream_t st) {
    const int H = a.prm.num_hyp;
    const int R = DENSE ? 1 : a.in.num_regions;
    const int RB = R + 1;
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    FusedLayout lay;
    size_t off = al(sizeof(FusedSmem<DENSE>));
    lay.anchors = (int)off; off = al(off + (size_t)R * sizeof(float4));
    lay.runtab = (int)off;  off = al(off + (size_t)R * sizeof(float4));
    lay.bstart = (int)off;  off = al(off + (size_t)(RB + 1) * sizeof(int));
    lay.wrun = (int)off;    off = al(off + (size_t)SW * RB * sizeof(uint16_t));
    lay.hyp = (int)off;     off = al(off + (size_t)H * 12 * sizeof(float));
    lay.hcnt = (int)off;    off = al(off + (size_t)H * sizeof(int));
    lay.total = (int)off;
    const size_t smem = off;
    if (smem > 227 * 1024) return RDPN_E_TOOLARGE;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        RDPN_CUDA_TRY(cudaFuncSetAttribute(pose_solve_kernel<DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    pose_solve_kernel<DENSE><<<a.in.B, ST, smem, st>>>(a, lay);
    ++g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

}  // namespace rdpn

extern "C" {

int rdpn_pose_solve(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net,
                    const rdpn_solve_params* prm, const rdpn_solve_outputs* out, void* stream) {
    bool dense = false;
    int rc = rdpn::check_roi_inputs(in, &dense);
    if (rc) return rc;
    if (!d_hyp_idx || !prm || !out || !out->pose || !out->n_inliers || !out->status) return RDPN_E_BADARG;
    if (prm->num_hyp <= 0 || !(prm->inlier_thr > 0.f)) return RDPN_E_BADARG;
    if (out->inlier_mask && ((uintptr_t)out->inlier_mask & 15)) return RDPN_E_ALIGN;
    rdpn::SolveArgs a;
    a.in = *in;
    a.hyp_idx = d_hyp_idx;
    a.t_net = d_t_net;
    a.prm = *prm;
    a.out = *out;
    a.sq_cut = rdpn::host_sq_cut(prm->inlier_thr);
    {
        const float thr = in->mask_thr;
        uint32_t bits;
        memcpy(&bits, &thr, sizeof(bits));
        a.mask_cut = 0.5 * ((double)thr + (double)nextafterf(thr, INFINITY));
        a.mask_cut_incl = (int)(bits & 1u);
    }
    return dense ? rdpn::launch_solve<true>(a, (cudaStream_t)stream) : rdpn::launch_solve<false>(a, (cudaStream_t)stream);
}

int rdpn_kabsch(const float* d_src, const float* d_dst, const float* d_w, int N, int with_scale, float* d_out_M,
                float* d_out_scale, int B, void* stream) {
    if (!d_src || !d_dst || !d_out_M || B <= 0) return RDPN_E_BADARG;
    if (N < 3) return RDPN_E_BADARG;  // transform.py:917-918 raises ValueError
    rdpn::kabsch_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_src, d_dst, d_w, N, with_scale, d_out_M, d_out_scale);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
