// Dense-correspondence -> pose on sm_100a: S1 (fused back-projection + residual + mask gate), the
// fused per-ROI solver (S1 + hypothesis generation + H x n inlier scoring + best selection + weighted
// Kabsch/Umeyama refit) and the batched Kabsch entry.
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

struct RoiArgs {
    rdpn_roi_inputs in;
};

struct SolveArgs {
    rdpn_roi_inputs in;
    const int32_t* hyp_idx;
    const float* t_net;
    rdpn_solve_params prm;
    rdpn_solve_outputs out;
    float sq_cut;  // smallest FP32 x with sqrtf(x) >= thr
};

// ---------------------------------------------------------------------------------------------
// shared-memory layout
// ---------------------------------------------------------------------------------------------
// tiles: [0]=depth [1]=coor_x [2]=coor_y [3]=coor_z [4]=mask, each 4096 floats.
// After S1 (in place):  [0]=w (mask prob) [1..3]=cam xyz planar; mask tile -> sel flags (u8) ;
// after compaction:     tiles[0..3] reinterpret as float4 camw[n]; mask tile -> uint32 meta[n]
//                       (anchor mode: pix | rid << 16) ; dense mode: obj float4 array in tiles2.
template <bool DENSE>
struct __align__(128) SolveSmem {
    float tile[5][RDPN_P];
    float obj[DENSE ? 4 : 1][DENSE ? RDPN_P : 4];  // dense: planar obj xyz, later float4 AoS (x,y,z,pix)
    uint8_t rid[DENSE ? 16 : RDPN_P];
    float4 anchors[DENSE ? 1 : 256];
    uint64_t bar;
    RoiConst rc;
    float red_f[2][SW];
    double red_d[SW][12];
    double bc_d[16];
    int cnt[4][SW];
    int base[4][SW];
    int n_sel;
    int best_h;
    int n_best;
    int h_eff;
    float pose[12];
    unsigned long long red_k[SW];
    int red_i[SW];
};

__device__ __forceinline__ float mask_prob(float m, int mode, float mn, float mx) {
    if (mode == RDPN_MASK_L1) return __fdiv_rn(__fsub_rn(m, mn), __fsub_rn(mx, mn));
    if (mode == RDPN_MASK_BCE) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-m)));
    return m;
}

// Stage the ROI planes with bulk TMA and set up the per-ROI constants.  Returns after the data landed.
template <bool DENSE>
__device__ __forceinline__ void stage_roi(SolveSmem<DENSE>& s, const rdpn_roi_inputs& in, int b) {
    const int t = threadIdx.x;
    if (t == 0) {
        mbar_init(&s.bar, 1);
        mbar_fence_init();
        const size_t o = (size_t)b * RDPN_P;
        const uint32_t plane = RDPN_P * sizeof(float);
        mbar_expect_tx(&s.bar, 5 * plane + (DENSE ? 0 : RDPN_P));
        bulk_g2s(s.tile[0], in.depth + o, plane, &s.bar);
        bulk_g2s(s.tile[1], in.coor_x + o, plane, &s.bar);
        bulk_g2s(s.tile[2], in.coor_y + o, plane, &s.bar);
        bulk_g2s(s.tile[3], in.coor_z + o, plane, &s.bar);
        bulk_g2s(s.tile[4], in.mask + o, plane, &s.bar);
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
    }
    if (!DENSE) {
        const int R = in.num_regions;
        for (int r = t; r < R; r += ST) {
            const float* a = in.anchors + ((size_t)b * R + r) * 3;
            s.anchors[r] = make_float4(a[0], a[1], a[2], 0.f);
        }
    }
    __syncthreads();
    mbar_wait(&s.bar, 0);
}

// S1 for this thread's 4 pixel quads, in place.  Pixel p = 4*(k*ST + t) + j.
//   out: tile[0]=w, tile[1..3]=cam ; sel flags (u8) over the mask tile ; dense: s.obj[0..2] planar.
// Returns the number of gated pixels of this thread per quad-iteration (cnt[k]) and sel bits.
template <bool DENSE>
__device__ __forceinline__ void s1_inplace(SolveSmem<DENSE>& s, const rdpn_roi_inputs& in, int (&cnt)[QPT],
                                           unsigned& selbits) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // ---- mask min / max (engine_utils.py:123-124) ----
    float4 mq[QPT];
#pragma unroll
    for (int k = 0; k < QPT; ++k) mq[k] = reinterpret_cast<const float4*>(s.tile[4])[k * ST + t];
    if (in.mask_mode == RDPN_MASK_L1) {
        float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            mn = fminf(fminf(fminf(mn, mq[k].x), fminf(mq[k].y, mq[k].z)), mq[k].w);
            mx = fmaxf(fmaxf(fmaxf(mx, mq[k].x), fmaxf(mq[k].y, mq[k].z)), mq[k].w);
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) { s.red_f[0][warp] = mn; s.red_f[1][warp] = mx; }
        __syncthreads();
        if (t == 0) {
            float a = s.red_f[0][0], c = s.red_f[1][0];
            for (int w = 1; w < SW; ++w) { a = fminf(a, s.red_f[0][w]); c = fmaxf(c, s.red_f[1][w]); }
            s.rc.mn = a;
            s.rc.mx = c;
        }
    }
    __syncthreads();  // all mask-tile reads done; rc complete
    const RoiConst rc = s.rc;
    selbits = 0u;
    uint8_t* selb = reinterpret_cast<uint8_t*>(s.tile[4]);
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int q = k * ST + t;
        const float4 dq = reinterpret_cast<const float4*>(s.tile[0])[q];
        const float4 xq = reinterpret_cast<const float4*>(s.tile[1])[q];
        const float4 yq = reinterpret_cast<const float4*>(s.tile[2])[q];
        const float4 zq = reinterpret_cast<const float4*>(s.tile[3])[q];
        const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
        const float cxn[4] = {xq.x, xq.y, xq.z, xq.w};
        const float cyn[4] = {yq.x, yq.y, yq.z, yq.w};
        const float czn[4] = {zq.x, zq.y, zq.z, zq.w};
        const float mm[4] = {mq[k].x, mq[k].y, mq[k].z, mq[k].w};
        float ow[4], ox[4], oy[4], oz[4], bx[4], by[4], bz[4];
        const int p0 = 4 * q;
        const float v = (float)(4 * (p0 >> 6));  // row -> crop pixel (stride 4, data_loader.py:625)
        int c4 = 0;
        uchar4 sb;
        uint8_t* sbp = reinterpret_cast<uint8_t*>(&sb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float u = (float)(4 * ((p0 + j) & 63));
            float d = dd[j];
            if (rc.div != 0.f) d = __fdiv_rn(d, rc.div);  // data_loader.py:563
            const float X = __fdiv_rn(__fmul_rn(__fsub_rn(u, rc.cx), d), rc.fx);  // :573
            const float Y = __fdiv_rn(__fmul_rn(__fsub_rn(v, rc.cy), d), rc.fy);  // :574
            const float dx = __fmul_rn(__fsub_rn(cxn[j], 0.5f), rc.ext[0]);  // gdrn_evaluator.py:103-105
            const float dy = __fmul_rn(__fsub_rn(cyn[j], 0.5f), rc.ext[1]);
            const float dz = __fmul_rn(__fsub_rn(czn[j], 0.5f), rc.ext[2]);
            const float w = mask_prob(mm[j], in.mask_mode, rc.mn, rc.mx);
            const bool sel = (w > in.mask_thr) && (fabsf(dx) > rc.gthr[0]) && (fabsf(dy) > rc.gthr[1]) &&
                             (fabsf(dz) > rc.gthr[2]) && (d > 0.f);  // gdrn_evaluator.py:110-117 (+ depth)
            ow[j] = w;
            if (DENSE) {
                ox[j] = X; oy[j] = Y; oz[j] = d;
                bx[j] = dx; by[j] = dy; bz[j] = dz;
            } else {
                ox[j] = __fsub_rn(X, dx); oy[j] = __fsub_rn(Y, dy); oz[j] = __fsub_rn(d, dz);
            }
            sbp[j] = sel ? 1 : 0;
            c4 += sel ? 1 : 0;
            selbits |= (sel ? 1u : 0u) << (4 * k + j);
        }
        cnt[k] = c4;
        reinterpret_cast<float4*>(s.tile[0])[q] = make_float4(ow[0], ow[1], ow[2], ow[3]);
        reinterpret_cast<float4*>(s.tile[1])[q] = make_float4(ox[0], ox[1], ox[2], ox[3]);
        reinterpret_cast<float4*>(s.tile[2])[q] = make_float4(oy[0], oy[1], oy[2], oy[3]);
        reinterpret_cast<float4*>(s.tile[3])[q] = make_float4(oz[0], oz[1], oz[2], oz[3]);
        reinterpret_cast<uchar4*>(selb)[q] = sb;
        if (DENSE) {
            reinterpret_cast<float4*>(s.obj[0])[q] = make_float4(bx[0], bx[1], bx[2], bx[3]);
            reinterpret_cast<float4*>(s.obj[1])[q] = make_float4(by[0], by[1], by[2], by[3]);
            reinterpret_cast<float4*>(s.obj[2])[q] = make_float4(bz[0], bz[1], bz[2], bz[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// S1 standalone (materialising): HBM-bound, 21 B/px in, 29..41 B/px out.
// ---------------------------------------------------------------------------------------------
template <bool DENSE>
__global__ void __launch_bounds__(ST, 2)
    correspond_kernel(RoiArgs a, float* __restrict__ cam, float* __restrict__ obj, float* __restrict__ w,
                      uint8_t* __restrict__ sel, int32_t* __restrict__ nsel) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SolveSmem<DENSE>& s = *reinterpret_cast<SolveSmem<DENSE>*>(smem_raw);
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    stage_roi<DENSE>(s, a.in, b);
    int cnt[QPT];
    unsigned selbits;
    s1_inplace<DENSE>(s, a.in, cnt, selbits);
    __syncthreads();
    const size_t o = (size_t)b * RDPN_P;
    const uint8_t* selb = reinterpret_cast<const uint8_t*>(s.tile[4]);
    int total = 0;
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int q = k * ST + t;
        total += cnt[k];
        reinterpret_cast<float4*>(w + o)[q] = reinterpret_cast<const float4*>(s.tile[0])[q];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            reinterpret_cast<float4*>(cam + (3 * (size_t)b + c) * RDPN_P)[q] = reinterpret_cast<const float4*>(s.tile[1 + c])[q];
        reinterpret_cast<uchar4*>(sel + o)[q] = reinterpret_cast<const uchar4*>(selb)[q];
        if (obj) {
            if (DENSE) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    reinterpret_cast<float4*>(obj + (3 * (size_t)b + c) * RDPN_P)[q] = reinterpret_cast<const float4*>(s.obj[c])[q];
            } else {
                const uchar4 r = reinterpret_cast<const uchar4*>(s.rid)[q];
                const float4 a0 = s.anchors[r.x], a1 = s.anchors[r.y], a2 = s.anchors[r.z], a3 = s.anchors[r.w];
                reinterpret_cast<float4*>(obj + (3 * (size_t)b + 0) * RDPN_P)[q] = make_float4(a0.x, a1.x, a2.x, a3.x);
                reinterpret_cast<float4*>(obj + (3 * (size_t)b + 1) * RDPN_P)[q] = make_float4(a0.y, a1.y, a2.y, a3.y);
                reinterpret_cast<float4*>(obj + (3 * (size_t)b + 2) * RDPN_P)[q] = make_float4(a0.z, a1.z, a2.z, a3.z);
            }
        }
    }
    total = warp_sum(total);
    if (lane == 0) s.red_i[warp] = total;
    __syncthreads();
    if (t == 0) {
        int n = 0;
        for (int i = 0; i < SW; ++i) n += s.red_i[i];
        nsel[b] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// fused solver
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float resid2(const float* P, float ax, float ay, float az, float cx, float cy, float cz) {
    float x = __fmaf_rn(P[0], ax, P[3]);
    x = __fmaf_rn(P[1], ay, x);
    x = __fmaf_rn(P[2], az, x);
    float y = __fmaf_rn(P[4], ax, P[7]);
    y = __fmaf_rn(P[5], ay, y);
    y = __fmaf_rn(P[6], az, y);
    float z = __fmaf_rn(P[8], ax, P[11]);
    z = __fmaf_rn(P[9], ay, z);
    z = __fmaf_rn(P[10], az, z);
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
    float d2 = __fmul_rn(dx, dx);
    d2 = __fmaf_rn(dy, dy, d2);
    d2 = __fmaf_rn(dz, dz, d2);
    return d2;
}

// block-wide sum of NV doubles; result valid in every thread (via s.bc_d[0..NV)).
template <bool DENSE, int NV>
__device__ __forceinline__ void block_sum(SolveSmem<DENSE>& s, double (&v)[NV]) {
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

template <bool DENSE>
__global__ void __launch_bounds__(ST, 2) pose_solve_kernel(SolveArgs a, int smem_hyp_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SolveSmem<DENSE>& s = *reinterpret_cast<SolveSmem<DENSE>*>(smem_raw);
    float* hyp = reinterpret_cast<float*>(smem_raw + smem_hyp_off);  // [H][12]
    const int H = a.prm.num_hyp;
    int* hcnt = reinterpret_cast<int*>(hyp + (size_t)H * 12);        // [H] counts, -1 = invalid
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;

    stage_roi<DENSE>(s, a.in, b);
    int cnt[QPT];
    unsigned selbits;
    s1_inplace<DENSE>(s, a.in, cnt, selbits);

    // ---- ordered compaction bookkeeping: position of every gated pixel in pixel order ----
    int pre[QPT];
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        int x = cnt[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        pre[k] = x - cnt[k];
        if (lane == 31) s.cnt[k][warp] = x;
    }
    if (a.out.inlier_mask) {  // zero-fill; inliers are scattered in after the last refit
        uint4* im = reinterpret_cast<uint4*>(a.out.inlier_mask + (size_t)b * RDPN_P);
        im[t] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();  // S1 results + counts visible
    if (t == 0) {
        int run = 0;
        for (int k = 0; k < QPT; ++k)
            for (int w = 0; w < SW; ++w) { s.base[k][w] = run; run += s.cnt[k][w]; }
        s.n_sel = run;
        s.best_h = -1;
        s.n_best = 0;
        s.h_eff = H;
    }

    // ---- hypothesis generation (FP64 closed form), one hypothesis per thread ----
    {
        const uint8_t* selb = reinterpret_cast<const uint8_t*>(s.tile[4]);
        for (int h = t; h < H; h += ST) {
            const int32_t* ip = a.hyp_idx + ((size_t)b * H + h) * 3;
            const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
            bool ok = ((unsigned)i0 < RDPN_P) && ((unsigned)i1 < RDPN_P) && ((unsigned)i2 < RDPN_P);
            float* P = hyp + (size_t)h * 12;
            if (ok) ok = selb[i0] && selb[i1] && selb[i2];
            if (ok) {
                const int ii[3] = {i0, i1, i2};
                double A[3][3], C[3][3];
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    C[v][0] = (double)s.tile[1][ii[v]];
                    C[v][1] = (double)s.tile[2][ii[v]];
                    C[v][2] = (double)s.tile[3][ii[v]];
                    if (DENSE) {
                        A[v][0] = (double)s.obj[0][ii[v]];
                        A[v][1] = (double)s.obj[1][ii[v]];
                        A[v][2] = (double)s.obj[2][ii[v]];
                    } else {
                        const float4 an = s.anchors[s.rid[ii[v]]];
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
    }
    __syncthreads();  // hypotheses done; planar arrays may now be overwritten
    const int n = s.n_sel;

    // ---- compaction: AoS float4 (cam xyz, w) over tiles[0..3]; meta / obj AoS ----
    {
        float4 cw[QPT][4];
        unsigned meta[QPT][4];
        float4 ob[DENSE ? QPT : 1][4];
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            const int q = k * ST + t;
            const float4 wq = reinterpret_cast<const float4*>(s.tile[0])[q];
            const float4 xq = reinterpret_cast<const float4*>(s.tile[1])[q];
            const float4 yq = reinterpret_cast<const float4*>(s.tile[2])[q];
            const float4 zq = reinterpret_cast<const float4*>(s.tile[3])[q];
            cw[k][0] = make_float4(xq.x, yq.x, zq.x, wq.x);
            cw[k][1] = make_float4(xq.y, yq.y, zq.y, wq.y);
            cw[k][2] = make_float4(xq.z, yq.z, zq.z, wq.z);
            cw[k][3] = make_float4(xq.w, yq.w, zq.w, wq.w);
            if (DENSE) {
                const float4 ax = reinterpret_cast<const float4*>(s.obj[0])[q];
                const float4 ay = reinterpret_cast<const float4*>(s.obj[1])[q];
                const float4 az = reinterpret_cast<const float4*>(s.obj[2])[q];
                ob[k][0] = make_float4(ax.x, ay.x, az.x, __int_as_float(4 * q + 0));
                ob[k][1] = make_float4(ax.y, ay.y, az.y, __int_as_float(4 * q + 1));
                ob[k][2] = make_float4(ax.z, ay.z, az.z, __int_as_float(4 * q + 2));
                ob[k][3] = make_float4(ax.w, ay.w, az.w, __int_as_float(4 * q + 3));
            } else {
                const uchar4 r = reinterpret_cast<const uchar4*>(s.rid)[q];
                meta[k][0] = (unsigned)(4 * q + 0) | ((unsigned)r.x << 16);
                meta[k][1] = (unsigned)(4 * q + 1) | ((unsigned)r.y << 16);
                meta[k][2] = (unsigned)(4 * q + 2) | ((unsigned)r.z << 16);
                meta[k][3] = (unsigned)(4 * q + 3) | ((unsigned)r.w << 16);
            }
        }
        __syncthreads();  // everyone holds its pixels in registers
        float4* camw = reinterpret_cast<float4*>(s.tile[0]);
        unsigned* metaS = reinterpret_cast<unsigned*>(s.tile[4]);
        float4* objS = reinterpret_cast<float4*>(s.obj[0]);
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            int pos = s.base[k][warp] + pre[k];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (selbits & (1u << (4 * k + j))) {
                    camw[pos] = cw[k][j];
                    if (DENSE) objS[pos] = ob[k][j];
                    else metaS[pos] = meta[k][j];
                    ++pos;
                }
        }
    }
    __syncthreads();
    const float4* camw = reinterpret_cast<const float4*>(s.tile[0]);
    const unsigned* metaS = reinterpret_cast<const unsigned*>(s.tile[4]);
    const float4* objS = reinterpret_cast<const float4*>(s.obj[0]);

    if (a.out.n_sel && t == 0) a.out.n_sel[b] = n;
    const bool enough = n >= a.prm.min_pts;

    // ---- S4: inlier scoring, thread per (hypothesis, point segment) ----
    if (enough) {
        const int S = (H >= ST) ? 1 : (ST / H);
        const float cut = a.sq_cut;
        for (int item = t; item < H * S; item += ST) {
            const int h = item % H, seg = item / H;
            if (hcnt[h] < 0 && S == 1) continue;
            const bool valid = hcnt[h] >= 0;
            float P[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) P[i] = hyp[(size_t)h * 12 + i];
            const int i0 = (int)(((long long)n * seg) / S), i1 = (int)(((long long)n * (seg + 1)) / S);
            int c = 0;
            if (valid) {
#pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const float4 cp = camw[i];
                    float4 ap;
                    if (DENSE) ap = objS[i];
                    else ap = s.anchors[metaS[i] >> 16];
                    c += resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z) < cut ? 1 : 0;
                }
                if (S == 1) hcnt[h] = c;
                else atomicAdd(&hcnt[h], c);
            }
        }
    }
    __syncthreads();

    // ---- best hypothesis (misc.py:121) with optional adaptive stop (misc.py:134-138) ----
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

    // ---- S5: refit on the inliers (misc.py:123-126 -> transform.py:913-980), FP64 accumulation ----
    if (t < 12) s.pose[t] = hyp[(size_t)best * 12 + t];
    __syncthreads();
    const float cut = a.sq_cut;
    float out_scale = 1.f;
    const int iters = a.prm.refit_iters < 1 ? 1 : a.prm.refit_iters;
    for (int it = 0; it < iters; ++it) {
        float P[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) P[i] = s.pose[i];
        // pass 1: weighted centroids
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        unsigned inl_bits = 0u;  // thread handles points t, t+ST, ... (<= 16 of them)
        int slot = 0;
        for (int i = t; i < n; i += ST, ++slot) {
            const float4 cp = camw[i];
            float4 ap;
            if (DENSE) ap = objS[i];
            else ap = s.anchors[metaS[i] >> 16];
            if (resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z) < cut) {
                inl_bits |= 1u << slot;
                const double w = a.prm.weighted ? (double)cp.w : 1.0;
                acc[0] += w;
                acc[1] += w * cp.x; acc[2] += w * cp.y; acc[3] += w * cp.z;
                acc[4] += w * ap.x; acc[5] += w * ap.y; acc[6] += w * ap.z;
            }
        }
        const int my_inl = __popc(inl_bits);
        int tot_inl = warp_sum(my_inl);
        if (lane == 0) s.red_i[warp] = tot_inl;
        block_sum<DENSE, 7>(s, acc);  // contains the barriers that publish red_i
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
                if (DENSE) ap = objS[i];
                else ap = s.anchors[metaS[i] >> 16];
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
        block_sum<DENSE, 11>(s, cov);
        if (t == 0) {
            double R[9];
            rotation_from_cov(cov, cov[10], cov[9], R);
            double sc = 1.0;
            if (a.prm.with_scale) sc = sqrt(cov[9] / cov[10]);  // transform.py:971-975
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double tr = mc[r] - sc * (R[3 * r] * ma[0] + R[3 * r + 1] * ma[1] + R[3 * r + 2] * ma[2]);
                s.pose[4 * r + 0] = (float)(sc * R[3 * r + 0]);
                s.pose[4 * r + 1] = (float)(sc * R[3 * r + 1]);
                s.pose[4 * r + 2] = (float)(sc * R[3 * r + 2]);
                s.pose[4 * r + 3] = (float)tr;
            }
            s.bc_d[15] = sc;
        }
        if (it == iters - 1 || true) {
            // remember the inlier set used by this refit; written to HBM only for the last one
            if (a.out.inlier_mask && it == iters - 1) {
                slot = 0;
                for (int i = t; i < n; i += ST, ++slot)
                    if (inl_bits & (1u << slot)) {
                        const unsigned pix = DENSE ? (unsigned)__float_as_int(objS[i].w) : (metaS[i] & 0xFFFFu);
                        a.out.inlier_mask[(size_t)b * RDPN_P + pix] = 1;
                    }
            }
        }
        __syncthreads();
        out_scale = (float)s.bc_d[15];
    }

    // ---- outputs (+ translation sanity, gdrn_evaluator.py:293-296) ----
    if (t == 0) {
        int status = RDPN_STATUS_OK;
        float tx = s.pose[3], ty = s.pose[7], tz = s.pose[11];
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

static int check_inputs(const rdpn_roi_inputs* in, bool* dense) {
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
static int launch_correspond(const rdpn_roi_inputs* in, float* cam, float* obj, float* w, uint8_t* sel, int32_t* nsel,
                             cudaStream_t st) {
    const size_t smem = sizeof(SolveSmem<DENSE>);
    static bool attr_set = false;
    if (!attr_set) {
        RDPN_CUDA_TRY(cudaFuncSetAttribute(correspond_kernel<DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    RoiArgs a;
    a.in = *in;
    correspond_kernel<DENSE><<<in->B, ST, smem, st>>>(a, cam, obj, w, sel, nsel);
    ++g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

template <bool DENSE>
static int launch_solve(const SolveArgs& a, cudaStream_t st) {
    const int H = a.prm.num_hyp;
    const size_t hyp_off = (sizeof(SolveSmem<DENSE>) + 127) & ~(size_t)127;
    const size_t smem = hyp_off + (size_t)H * 12 * sizeof(float) + (size_t)H * sizeof(int);
    if (smem > 227 * 1024) return RDPN_E_TOOLARGE;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        RDPN_CUDA_TRY(cudaFuncSetAttribute(pose_solve_kernel<DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    pose_solve_kernel<DENSE><<<a.in.B, ST, smem, st>>>(a, (int)hyp_off);
    ++g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

}  // namespace rdpn

extern "C" {

int rdpn_correspond(const rdpn_roi_inputs* in, float* d_cam, float* d_obj, float* d_w, uint8_t* d_sel, int32_t* d_nsel,
                    void* stream) {
    bool dense = false;
    int rc = rdpn::check_inputs(in, &dense);
    if (rc) return rc;
    if (!d_cam || !d_w || !d_sel || !d_nsel) return RDPN_E_BADARG;
    if (((uintptr_t)d_cam | (uintptr_t)d_obj | (uintptr_t)d_w | (uintptr_t)d_sel) & 15) return RDPN_E_ALIGN;
    return dense ? rdpn::launch_correspond<true>(in, d_cam, d_obj, d_w, d_sel, d_nsel, (cudaStream_t)stream)
                 : rdpn::launch_correspond<false>(in, d_cam, d_obj, d_w, d_sel, d_nsel, (cudaStream_t)stream);
}

int rdpn_pose_solve(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net,
                    const rdpn_solve_params* prm, const rdpn_solve_outputs* out, void* stream) {
    bool dense = false;
    int rc = rdpn::check_inputs(in, &dense);
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
