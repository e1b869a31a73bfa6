// S1 standalone (materialising): fused back-projection + residual + mask gate for EVERY pixel of every
// ROI, written back to HBM.  Replaces data_loader.py:563-576 + gdrn_evaluator.py:89-126 +
// engine_utils.py:118-136 of the reference for callers that want the dense correspondences themselves.
//
// HBM-bound: 21 B/px in (depth, 3 x coor, mask, region id), 21 B/px out (cam xyz, w, sel) or 33 B/px with
// the object side materialised.  Design for the roofline:
//   * persistent grid, one CTA per SM; ROIs are dealt round-robin (b = blockIdx.x + i * gridDim.x)
//   * a producer warp streams the next ROI's six contiguous planes into the other half of a 2-stage
//     shared-memory ring with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx) while the 8 compute
//     warps work on the current one, so loads, arithmetic and stores of neighbouring ROIs overlap
//   * the per-ROI mask min/max (engine_utils.py:123-124) is a block reduction over the staged plane,
//     which is why the ROI is staged at all (everything else is elementwise)
//   * results go from registers to HBM as 16-byte coalesced stores; shared memory is never written
//     by the compute warps, so the TMA ring needs no cross-proxy fences
// Arithmetic is the exact FP32 contract of oracle/pose_oracle.py (one IEEE op per step).
#include "common.cuh"

#include <float.h>

namespace rdpn {
extern unsigned long long g_launch_count;
int device_sm_count(int* sms);
int ensure_func_smem(const void* func, int slot, size_t bytes);

constexpr int C_CT = 512;           // compute threads (16 warps: the per-pixel IEEE divisions are latency-bound, TLP hides them)
constexpr int C_CW = C_CT / 32;     // compute warps
constexpr int C_NT = C_CT + 32;     // + producer warp
constexpr int C_QPT = RDPN_P / 4 / C_CT;

struct __align__(128) S1Stage {
    float tile[5][RDPN_P];  // depth, coor_x, coor_y, coor_z, mask
    uint8_t rid[RDPN_P];
};

struct __align__(128) S1Smem {
    S1Stage st[2];
    uint64_t full[2];
    uint64_t empty[2];
    float red[2][C_CW];
    int red_i[2][C_CW];
};

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(C_CT) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_cs(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool DENSE>
__global__ void __launch_bounds__(C_NT, 1)
    correspond_kernel(rdpn_roi_inputs in, float* __restrict__ cam, float* __restrict__ obj, float* __restrict__ w,
                      uint8_t* __restrict__ sel, int32_t* __restrict__ nsel) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    S1Smem& s = *reinterpret_cast<S1Smem*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int B = in.B, G = gridDim.x;
    const int my_n = (B - (int)blockIdx.x + G - 1) / G;  // ROIs this CTA owns
    if (t == 0) {
        mbar_init(&s.full[0], 1);
        mbar_init(&s.full[1], 1);
        mbar_init(&s.empty[0], 1);
        mbar_init(&s.empty[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == C_CW) {
        // ---------------- producer warp ----------------
        if (lane == 0) {
            const uint32_t plane = RDPN_P * sizeof(float);
            for (int i = 0; i < my_n; ++i) {
                const int stg = i & 1;
                if (i >= 2) mbar_wait(&s.empty[stg], ((i >> 1) - 1) & 1);
                const size_t o = (size_t)(blockIdx.x + (size_t)i * G) * RDPN_P;
                S1Stage& d = s.st[stg];
                mbar_expect_tx(&s.full[stg], 5 * plane + (DENSE ? 0 : RDPN_P));
                bulk_g2s(d.tile[4], in.mask + o, plane, &s.full[stg]);
                bulk_g2s(d.tile[0], in.depth + o, plane, &s.full[stg]);
                bulk_g2s(d.tile[1], in.coor_x + o, plane, &s.full[stg]);
                bulk_g2s(d.tile[2], in.coor_y + o, plane, &s.full[stg]);
                bulk_g2s(d.tile[3], in.coor_z + o, plane, &s.full[stg]);
                if (!DENSE) bulk_g2s(d.rid, in.region_idx + o, RDPN_P, &s.full[stg]);
            }
        }
        return;
    }

    // ---------------- compute warps ----------------
    for (int i = 0; i < my_n; ++i) {
        const int stg = i & 1;
        const int b = blockIdx.x + i * G;
        // per-ROI constants straight from global (latency hidden behind the TMA wait)
        const float fx = in.Kp[4 * b + 0], fy = in.Kp[4 * b + 1], cxp = in.Kp[4 * b + 2], cyp = in.Kp[4 * b + 3];
        const float e0 = in.extent[3 * b + 0], e1 = in.extent[3 * b + 1], e2 = in.extent[3 * b + 2];
        const float g0 = (float)(0.0001 * (double)e0), g1 = (float)(0.0001 * (double)e1), g2 = (float)(0.0001 * (double)e2);
        const float div = in.depth_div ? in.depth_div[b] : 0.f;
        const float sgx = copysignf(1.f, fx), sgy = copysignf(1.f, fy);
        mbar_wait(&s.full[stg], (i >> 1) & 1);
        const S1Stage& d = s.st[stg];
        float4 mq[C_QPT];
#pragma unroll
        for (int k = 0; k < C_QPT; ++k) mq[k] = reinterpret_cast<const float4*>(d.tile[4])[k * C_CT + t];
        float mn = 0.f, mx = 0.f;
        if (in.mask_mode == RDPN_MASK_L1) {  // engine_utils.py:123-124
            mn = FLT_MAX;
            mx = -FLT_MAX;
#pragma unroll
            for (int k = 0; k < C_QPT; ++k) {
                mn = fminf(fminf(fminf(mn, mq[k].x), fminf(mq[k].y, mq[k].z)), mq[k].w);
                mx = fmaxf(fmaxf(fmaxf(mx, mq[k].x), fmaxf(mq[k].y, mq[k].z)), mq[k].w);
            }
            mn = warp_min(mn);
            mx = warp_max(mx);
            if (lane == 0) { s.red[0][warp] = mn; s.red[1][warp] = mx; }
            bar_compute();
            mn = s.red[0][0];
            mx = s.red[1][0];
#pragma unroll
            for (int ww = 1; ww < C_CW; ++ww) { mn = fminf(mn, s.red[0][ww]); mx = fmaxf(mx, s.red[1][ww]); }
        }
        const float mden = __fsub_rn(mx, mn);
        const size_t o = (size_t)b * RDPN_P;
        int total = 0;
#pragma unroll
        for (int k = 0; k < C_QPT; ++k) {
            const int q = k * C_CT + t;
            const float4 dq = reinterpret_cast<const float4*>(d.tile[0])[q];
            const float4 xq = reinterpret_cast<const float4*>(d.tile[1])[q];
            const float4 yq = reinterpret_cast<const float4*>(d.tile[2])[q];
            const float4 zq = reinterpret_cast<const float4*>(d.tile[3])[q];
            const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
            const float cxn[4] = {xq.x, xq.y, xq.z, xq.w};
            const float cyn[4] = {yq.x, yq.y, yq.z, yq.w};
            const float czn[4] = {zq.x, zq.y, zq.z, zq.w};
            const float mm[4] = {mq[k].x, mq[k].y, mq[k].z, mq[k].w};
            float ow[4], ox[4], oy[4], oz[4], bx[4], by[4], bz[4];
            const int p0 = 4 * q;
            const float v = (float)(4 * (p0 >> 6));  // row -> crop pixel (stride 4, data_loader.py:625)
            uchar4 sb;
            uint8_t* sbp = reinterpret_cast<uint8_t*>(&sb);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = (float)(4 * ((p0 + j) & 63));
                float dz0 = dd[j];
                if (div != 0.f) {                                                        // data_loader.py:563
                    const float q = __fdiv_rn(dz0 == 0.f ? 1.f : dz0, div);
                    dz0 = dz0 == 0.f ? __fmul_rn(dz0, copysignf(1.f, div)) : q;
                }
                // (u - cx') * d / fx' (:573-574).  Background pixels have d = 0, i.e. a zero numerator, which
                // would send the whole warp through div.rn's special-operand slow path; 0 / f is +-0 with the
                // sign of numerator * sign(f), so those lanes divide 1 / f instead and select the signed zero.
                const float nx = __fmul_rn(__fsub_rn(u, cxp), dz0);
                const float ny = __fmul_rn(__fsub_rn(v, cyp), dz0);
                const float qx = __fdiv_rn(nx == 0.f ? 1.f : nx, fx);
                const float qy = __fdiv_rn(ny == 0.f ? 1.f : ny, fy);
                const float X = nx == 0.f ? __fmul_rn(nx, sgx) : qx;
                const float Y = ny == 0.f ? __fmul_rn(ny, sgy) : qy;
                const float dx = __fmul_rn(__fsub_rn(cxn[j], 0.5f), e0);                 // gdrn_evaluator.py:103-105
                const float dy = __fmul_rn(__fsub_rn(cyn[j], 0.5f), e1);
                const float dz = __fmul_rn(__fsub_rn(czn[j], 0.5f), e2);
                float wv = mm[j];
                if (in.mask_mode == RDPN_MASK_L1) wv = __fdiv_rn(__fsub_rn(wv, mn), mden);
                else if (in.mask_mode == RDPN_MASK_BCE) wv = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-wv)));
                const bool ok = (wv > in.mask_thr) && (fabsf(dx) > g0) && (fabsf(dy) > g1) && (fabsf(dz) > g2) &&
                                (dz0 > 0.f);  // gdrn_evaluator.py:110-117 (+ depth validity)
                ow[j] = wv;
                if (DENSE) {
                    ox[j] = X; oy[j] = Y; oz[j] = dz0;
                    bx[j] = dx; by[j] = dy; bz[j] = dz;
                } else {
                    ox[j] = __fsub_rn(X, dx); oy[j] = __fsub_rn(Y, dy); oz[j] = __fsub_rn(dz0, dz);
                }
                sbp[j] = ok ? 1 : 0;
                total += ok ? 1 : 0;
            }
            st_cs(reinterpret_cast<float4*>(w + o) + q, make_float4(ow[0], ow[1], ow[2], ow[3]));
            st_cs(reinterpret_cast<float4*>(cam + (3 * (size_t)b + 0) * RDPN_P) + q, make_float4(ox[0], ox[1], ox[2], ox[3]));
            st_cs(reinterpret_cast<float4*>(cam + (3 * (size_t)b + 1) * RDPN_P) + q, make_float4(oy[0], oy[1], oy[2], oy[3]));
            st_cs(reinterpret_cast<float4*>(cam + (3 * (size_t)b + 2) * RDPN_P) + q, make_float4(oz[0], oz[1], oz[2], oz[3]));
            reinterpret_cast<uchar4*>(sel + o)[q] = sb;
            if (obj) {
                float4 o0, o1, o2;
                if (DENSE) {
                    o0 = make_float4(bx[0], bx[1], bx[2], bx[3]);
                    o1 = make_float4(by[0], by[1], by[2], by[3]);
                    o2 = make_float4(bz[0], bz[1], bz[2], bz[3]);
                } else {
                    const uchar4 r = reinterpret_cast<const uchar4*>(d.rid)[q];
                    const float* ab = in.anchors + (size_t)b * in.num_regions * 3;
                    const float* a0 = ab + 3 * r.x;
                    const float* a1 = ab + 3 * r.y;
                    const float* a2 = ab + 3 * r.z;
                    const float* a3 = ab + 3 * r.w;
                    o0 = make_float4(__ldg(a0), __ldg(a1), __ldg(a2), __ldg(a3));
                    o1 = make_float4(__ldg(a0 + 1), __ldg(a1 + 1), __ldg(a2 + 1), __ldg(a3 + 1));
                    o2 = make_float4(__ldg(a0 + 2), __ldg(a1 + 2), __ldg(a2 + 2), __ldg(a3 + 2));
                }
                st_cs(reinterpret_cast<float4*>(obj + (3 * (size_t)b + 0) * RDPN_P) + q, o0);
                st_cs(reinterpret_cast<float4*>(obj + (3 * (size_t)b + 1) * RDPN_P) + q, o1);
                st_cs(reinterpret_cast<float4*>(obj + (3 * (size_t)b + 2) * RDPN_P) + q, o2);
            }
        }
        total = warp_sum(total);
        if (lane == 0) s.red_i[stg][warp] = total;
        bar_compute();  // every compute thread is done reading this stage
        if (t == 0) {
            int n = 0;
#pragma unroll
            for (int ww = 0; ww < C_CW; ++ww) n += s.red_i[stg][ww];
            nsel[b] = n;
            mbar_arrive(&s.empty[stg]);  // hand the stage back to the producer
        }
    }
}

template <bool DENSE>
static int launch_correspond(const rdpn_roi_inputs* in, float* cam, float* obj, float* w, uint8_t* sel, int32_t* nsel,
                             cudaStream_t st) {
    const size_t smem = sizeof(S1Smem);
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    rc = ensure_func_smem((const void*)correspond_kernel<DENSE>, 4 + (DENSE ? 1 : 0), smem);
    if (rc) return rc;
    const int grid = in->B < sms ? in->B : sms;
    correspond_kernel<DENSE><<<grid, C_NT, smem, st>>>(*in, cam, obj, w, sel, nsel);
    ++g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int check_roi_inputs(const rdpn_roi_inputs* in, bool* dense);

}  // namespace rdpn

extern "C" int rdpn_correspond(const rdpn_roi_inputs* in, float* d_cam, float* d_obj, float* d_w, uint8_t* d_sel,
                               int32_t* d_nsel, void* stream) {
    RDPN_NVTX("rdpn_correspond");
    bool dense = false;
    int rc = rdpn::check_roi_inputs(in, &dense);
    if (rc) return rc;
    if (!d_cam || !d_w || !d_sel || !d_nsel) return RDPN_E_BADARG;
    if (((uintptr_t)d_cam | (uintptr_t)d_obj | (uintptr_t)d_w | (uintptr_t)d_sel) & 15) return RDPN_E_ALIGN;
    return dense ? rdpn::launch_correspond<true>(in, d_cam, d_obj, d_w, d_sel, d_nsel, (cudaStream_t)stream)
                 : rdpn::launch_correspond<false>(in, d_cam, d_obj, d_w, d_sel, d_nsel, (cudaStream_t)stream);
}
