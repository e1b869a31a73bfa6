// f2: the correspondence-feature assembly of GDRN.forward fused into one HBM-bound pass.
//
// Reference (all separate ATen kernels with five intermediate tensors):
//   core/gdrn_modeling/models/GDRN.py:199-222   cat(coor_x,y,z) | cat(roi_coord_2d) | softmax(region[:,1:]) ->
//                                               argmax -> gather fps anchors -> cat | get_mask_prob
//   core/gdrn_modeling/models/conv_pnp_net.py:128-136   cat(region softmax) | x * mask_attention (or concat)
//   core/gdrn_modeling/models/model_utils.py:24-42      mask probability (L1 min-max / sigmoid)
// Output: x [B, C, 64, 64], C = 3 + 5 + 3 (+ R when REGION_ATTENTION) (+ 1 when MASK_ATTENTION == "concat"),
// exactly the tensor ConvPnPNet.features consumes (C = 43 for R = 32, conv_pnp_net.py:73).
//
// The region logits [B, R+1, 64, 64] are the widest tensor on the path (132-260 B/px); they are read
// from HBM exactly once: a thread keeps the R logits of one pixel in registers (lanes = consecutive
// pixels, so every channel access is a coalesced 128-byte line), does max / arg-max, exp, sum and the
// normalisation there, and streams the C output planes out.  Algorithmic bytes per ROI:
// (R + 1 + 3 + 5 + 1) * 16 KB in, C * 16 KB out.
#include "common.cuh"

#include <float.h>

namespace rdpn {
extern unsigned long long g_launch_count;

constexpr int CF_T = 256;

template <int VEC> struct VecT;
template <> struct VecT<4> { typedef float4 type; };
template <> struct VecT<2> { typedef float2 type; };

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float (&v)[VEC]) {
    typename VecT<VEC>::type x = __ldcs(reinterpret_cast<const typename VecT<VEC>::type*>(p));
    const float* xp = reinterpret_cast<const float*>(&x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = xp[j];
}
template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const float (&v)[VEC]) {
    typename VecT<VEC>::type x;
    float* xp = reinterpret_cast<float*>(&x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) xp[j] = v[j];
    __stcs(reinterpret_cast<typename VecT<VEC>::type*>(p), x);
}

// A thread owns VEC consecutive pixels (VEC = 4 for R <= 32, 2 for R <= 64) and keeps their R logits in
// registers: every channel access of a warp is then VEC*128 contiguous bytes and a CTA touches
// VEC*1 KB per channel at a time (DRAM-page friendly; with scalar accesses the same kernel ran at 30 %).
template <int RMAX, int VEC>
__global__ void __launch_bounds__(CF_T)
    coor_feat_kernel(const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ cz,
                     const float* __restrict__ coord2d, const float* __restrict__ region, const float* __restrict__ fps,
                     const float* __restrict__ mask, int R, int mask_mode, int region_attention, int mask_attention,
                     float* __restrict__ out, int C) {
    __shared__ float red[2][CF_T / 32];
    __shared__ float4 anchors[RMAX];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t po = (size_t)b * RDPN_P;
    for (int r = t; r < R; r += CF_T) {
        const float* ap = fps + ((size_t)b * R + r) * 3;
        anchors[r] = make_float4(ap[0], ap[1], ap[2], 0.f);
    }
    // mask probability parameters (model_utils.py:29-34): per-ROI min / max for the L1 mode
    float mn = 0.f, mden = 1.f;
    if (mask_attention != 0 && mask_mode == RDPN_MASK_L1) {
        float lo = FLT_MAX, hi = -FLT_MAX;
        for (int q = t; q < RDPN_P / 4; q += CF_T) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask + po) + q);
            lo = fminf(fminf(fminf(lo, m.x), fminf(m.y, m.z)), m.w);
            hi = fmaxf(fmaxf(fmaxf(hi, m.x), fmaxf(m.y, m.z)), m.w);
        }
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (lane == 0) { red[0][warp] = lo; red[1][warp] = hi; }
        __syncthreads();
        lo = red[0][0];
        hi = red[1][0];
        for (int w = 1; w < CF_T / 32; ++w) { lo = fminf(lo, red[0][w]); hi = fmaxf(hi, red[1][w]); }
        mn = lo;
        mden = __fsub_rn(hi, lo);
    }
    __syncthreads();
    const float* reg = region + ((size_t)b * (R + 1) + 1) * RDPN_P;  // channel 0 is background (GDRN.py:206)
    float* ob = out + (size_t)b * C * RDPN_P;
    for (int p = VEC * t; p < RDPN_P; p += VEC * CF_T) {
        float scale[VEC], mp[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) mp[j] = 1.f;
        if (mask_attention != 0) {
            float m[VEC];
            ld_vec<VEC>(mask + po + p, m);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (mask_mode == RDPN_MASK_L1) mp[j] = __fdiv_rn(__fsub_rn(m[j], mn), mden);
                else if (mask_mode == RDPN_MASK_BCE) mp[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-m[j])));
                else mp[j] = m[j];
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) scale[j] = mask_attention == 1 ? mp[j] : 1.f;  // "mul" (conv_pnp_net.py:134-135)
        float lg[RMAX][VEC];
        float mx[VEC];
        int am[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { mx[j] = -FLT_MAX; am[j] = 0; }
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
            if (r < R) ld_vec<VEC>(reg + (size_t)r * RDPN_P + p, lg[r]);
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
            if (r < R) {
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (lg[r][j] > mx[j]) { mx[j] = lg[r][j]; am[j] = r; }  // first maximum wins (torch.argmax)
            }
        float sum[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) sum[j] = 0.f;
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
            if (r < R) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    lg[r][j] = expf(__fsub_rn(lg[r][j], mx[j]));
                    sum[j] = __fadd_rn(sum[j], lg[r][j]);
                }
            }
        // channels 0-2: coor, 3-7: roi_coord_2d, 8-10: anchor of the arg-max region
        float v[VEC];
        const float* srcs[8] = {cx + po, cy + po, cz + po, coord2d + ((size_t)b * 5 + 0) * RDPN_P, coord2d + ((size_t)b * 5 + 1) * RDPN_P,
                                coord2d + ((size_t)b * 5 + 2) * RDPN_P, coord2d + ((size_t)b * 5 + 3) * RDPN_P,
                                coord2d + ((size_t)b * 5 + 4) * RDPN_P};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            ld_vec<VEC>(srcs[c] + p, v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] *= scale[j];
            st_vec<VEC>(ob + (size_t)c * RDPN_P + p, v);
        }
        float ax[VEC], ay[VEC], az[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float4 an = anchors[am[j]];
            ax[j] = an.x * scale[j]; ay[j] = an.y * scale[j]; az[j] = an.z * scale[j];
        }
        st_vec<VEC>(ob + (size_t)8 * RDPN_P + p, ax);
        st_vec<VEC>(ob + (size_t)9 * RDPN_P + p, ay);
        st_vec<VEC>(ob + (size_t)10 * RDPN_P + p, az);
        int c = 11;
        if (region_attention) {
            // exp(x - max) / sum, then * mask_prob: one IEEE division per pixel (scale / sum) instead of one per
            // channel -- within 1 ulp of the reference's per-element division (tolerance stated in the test)
            float k[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) k[j] = __fdiv_rn(scale[j], sum[j]);
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
                if (r < R) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) lg[r][j] *= k[j];
                    st_vec<VEC>(ob + (size_t)(11 + r) * RDPN_P + p, lg[r]);
                }
            c += R;
        }
        if (mask_attention == 2) st_vec<VEC>(ob + (size_t)c * RDPN_P + p, mp);  // "concat"
    }
}

}  // namespace rdpn

extern "C" int rdpn_coor_feat(const float* d_coor_x, const float* d_coor_y, const float* d_coor_z, const float* d_roi_coord_2d,
                              const float* d_region, const float* d_fps, const float* d_mask, int R, int mask_mode,
                              int region_attention, int mask_attention, float* d_out, int B, void* stream) {
    if (!d_coor_x || !d_coor_y || !d_coor_z || !d_roi_coord_2d || !d_region || !d_fps || !d_out || B <= 0) return RDPN_E_BADARG;
    if (R <= 0 || R > 64) return RDPN_E_TOOLARGE;
    if (mask_attention < 0 || mask_attention > 2 || mask_mode < 0 || mask_mode > 2) return RDPN_E_BADARG;
    if (mask_attention != 0 && !d_mask) return RDPN_E_BADARG;
    if (((uintptr_t)d_coor_x | (uintptr_t)d_coor_y | (uintptr_t)d_coor_z | (uintptr_t)d_roi_coord_2d | (uintptr_t)d_region |
         (uintptr_t)d_mask | (uintptr_t)d_out) & 15)
        return RDPN_E_ALIGN;
    const int C = 11 + (region_attention ? R : 0) + (mask_attention == 2 ? 1 : 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (R <= 32)
        rdpn::coor_feat_kernel<32, 4><<<B, rdpn::CF_T, 0, st>>>(d_coor_x, d_coor_y, d_coor_z, d_roi_coord_2d, d_region, d_fps, d_mask, R,
                                                             mask_mode, region_attention, mask_attention, d_out, C);
    else
        rdpn::coor_feat_kernel<64, 2><<<B, rdpn::CF_T, 0, st>>>(d_coor_x, d_coor_y, d_coor_z, d_roi_coord_2d, d_region, d_fps, d_mask, R,
                                                             mask_mode, region_attention, mask_attention, d_out, C);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}
