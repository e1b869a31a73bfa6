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
// The region logits [B, R+1, 64, 64] are the widest tensor on the path (132-260 B/px); they cross HBM exactly
// once (bulk TMA into shared memory, see the kernel), the soft-max planes leave by bulk-TMA stores.
// Algorithmic bytes per ROI: (R + 1 + 3 + 5 + 1) * 16 KB in, C * 16 KB out.
#include "common.cuh"

#include <float.h>
#include <stdlib.h>

namespace rdpn {
int ensure_func_smem(const void* func, int slot, size_t bytes);
extern unsigned long long g_launch_count;



// One CTA owns a TILE of CF_TP (= threads per CTA) consecutive pixels of one ROI: the R logit rows of the tile (R x 1 KB,
// contiguous in every plane) are staged in shared memory by R bulk-TMA copies behind one mbarrier, a thread owns one
// pixel (conflict-free column access) and sweeps its column three times in shared memory -- max / arg-max,
// exp(l - max) written back in place + its sum in channel order, normalisation in place -- and the finished
// rows leave by bulk-TMA stores.  The logits cross HBM exactly once in each direction and never sit in
// registers (the register-resident version held R x 4 logits in 180 registers: one CTA per SM, load / exp /
// store phases serialised, 54 % of the copy bandwidth at R = 32 and 41 % at R = 64).
// Shared memory: R KB per CTA (3 CTAs per SM at R = 64, 7 at R = 32).

template <int CF_TP>  // pixels per tile = threads per CTA: 512 for R <= 32 (79 % of the copy bandwidth), 256 above (66 %)
__global__ void __launch_bounds__(CF_TP)
    coor_feat_kernel(const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ cz,
                     const float* __restrict__ coord2d, const float* __restrict__ region, const float* __restrict__ fps,
                     const float* __restrict__ mask, int R, int mask_mode, int region_attention, int mask_attention,
                     float* __restrict__ out, int C) {
    extern __shared__ __align__(128) unsigned char cf_smem[];
    float* tile = reinterpret_cast<float*>(cf_smem);  // [R][CF_TP]
    constexpr int CF_T = CF_TP;
    __shared__ float red[2][CF_T / 32];
    __shared__ float4 anchors[64];
    __shared__ __align__(8) uint64_t bar;
    constexpr int TILES = RDPN_P / CF_TP;
    const int b = blockIdx.x / TILES, p0 = (blockIdx.x % TILES) * CF_TP;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t po = (size_t)b * RDPN_P;
    const float* reg = region + ((size_t)b * (R + 1) + 1) * RDPN_P + p0;  // channel 0 is background (GDRN.py:206)
    float* ob = out + (size_t)b * C * RDPN_P + p0;
    if (t == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)R * CF_TP * 4);
        for (int r = 0; r < R; ++r) bulk_g2s(tile + r * CF_TP, reg + (size_t)r * RDPN_P, CF_TP * 4, &bar);
    }
    for (int r = t; r < R; r += CF_T) {
        const float* ap = fps + ((size_t)b * R + r) * 3;
        anchors[r] = make_float4(ap[0], ap[1], ap[2], 0.f);
    }
    // mask probability parameters (model_utils.py:29-34): per-ROI min / max for the L1 mode (the 16 KB plane is
    // re-read by the 16 tiles of the ROI out of L2)
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
    __syncthreads();  // anchors, barrier initialisation
    const int p = p0 + t;
    float mp = 1.f;
    if (mask_attention != 0) {
        const float m = __ldcs(mask + po + p);
        if (mask_mode == RDPN_MASK_L1) mp = __fdiv_rn(__fsub_rn(m, mn), mden);
        else if (mask_mode == RDPN_MASK_BCE) mp = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-m)));
        else mp = m;
    }
    const float scale = mask_attention == 1 ? mp : 1.f;  // "mul" (conv_pnp_net.py:134-135)
    // channels 0-2: coor, 3-7: roi_coord_2d -- while the tile is in flight
    {
        const float* srcs[8] = {cx + po, cy + po, cz + po, coord2d + ((size_t)b * 5 + 0) * RDPN_P,
                                coord2d + ((size_t)b * 5 + 1) * RDPN_P, coord2d + ((size_t)b * 5 + 2) * RDPN_P,
                                coord2d + ((size_t)b * 5 + 3) * RDPN_P, coord2d + ((size_t)b * 5 + 4) * RDPN_P};
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldcs(srcs[c] + p);
#pragma unroll
        for (int c = 0; c < 8; ++c) __stcs(ob + (size_t)c * RDPN_P + t, v[c] * scale);
    }
    mbar_wait(&bar, 0);
    // sweep 1: max / arg-max (first maximum wins, torch.argmax)
    float mx = -FLT_MAX;
    int am = 0;
#pragma unroll 8
    for (int r = 0; r < R; ++r) {
        const float l = tile[r * CF_TP + t];
        if (l > mx) { mx = l; am = r; }
    }
    {   // channels 8-10: anchor of the arg-max region
        const float4 an = anchors[am];
        __stcs(ob + (size_t)8 * RDPN_P + t, an.x * scale);
        __stcs(ob + (size_t)9 * RDPN_P + t, an.y * scale);
        __stcs(ob + (size_t)10 * RDPN_P + t, an.z * scale);
    }
    int c = 11;
    if (region_attention) {
        // sweep 2: e = exp(l - max) in place, sum in channel order
        float sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < R; ++r) {
            const float e = expf(__fsub_rn(tile[r * CF_TP + t], mx));
            tile[r * CF_TP + t] = e;
            sum = __fadd_rn(sum, e);
        }
        // exp(x - max) / sum, then * mask_prob: one IEEE division per pixel (scale / sum) instead of one per
        // channel -- within 1 ulp of the reference's per-element division (tolerance stated in the test)
        const float k = __fdiv_rn(scale, sum);
        // sweep 3: normalise in place, then the rows leave by bulk stores
#pragma unroll 8
        for (int r = 0; r < R; ++r) tile[r * CF_TP + t] *= k;
        fence_proxy_async();
        __syncthreads();
        for (int r = t; r < R; r += CF_T) bulk_s2g(ob + (size_t)(11 + r) * RDPN_P, tile + r * CF_TP, CF_TP * 4);
        bulk_commit();
        c += R;
    }
    if (mask_attention == 2) __stcs(ob + (size_t)c * RDPN_P + t, mp);  // "concat"
    bulk_wait_read_all();  // the tile must stay intact until the bulk stores have read it
}

}  // namespace rdpn

extern "C" int rdpn_coor_feat(const float* d_coor_x, const float* d_coor_y, const float* d_coor_z, const float* d_roi_coord_2d,
                              const float* d_region, const float* d_fps, const float* d_mask, int R, int mask_mode,
                              int region_attention, int mask_attention, float* d_out, int B, void* stream) {
    RDPN_NVTX("rdpn_coor_feat");
    if (!d_coor_x || !d_coor_y || !d_coor_z || !d_roi_coord_2d || !d_region || !d_fps || !d_out || B <= 0) return RDPN_E_BADARG;
    if (R <= 0 || R > 64) return RDPN_E_TOOLARGE;
    if (mask_attention < 0 || mask_attention > 2 || mask_mode < 0 || mask_mode > 2) return RDPN_E_BADARG;
    if (mask_attention != 0 && !d_mask) return RDPN_E_BADARG;
    if (((uintptr_t)d_coor_x | (uintptr_t)d_coor_y | (uintptr_t)d_coor_z | (uintptr_t)d_roi_coord_2d | (uintptr_t)d_region |
         (uintptr_t)d_mask | (uintptr_t)d_out) & 15)
        return RDPN_E_ALIGN;
    const int C = 11 + (region_attention ? R : 0) + (mask_attention == 2 ? 1 : 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int tp = R <= 32 ? 512 : 256;
    const int smem = R * tp * 4;
    {
        const int rc = tp == 512 ? rdpn::ensure_func_smem((const void*)rdpn::coor_feat_kernel<512>, 6, (size_t)smem)
                                 : rdpn::ensure_func_smem((const void*)rdpn::coor_feat_kernel<256>, 7, (size_t)smem);
        if (rc) return rc;
    }
    if (tp == 512)
        rdpn::coor_feat_kernel<512><<<B * (RDPN_P / 512), 512, smem, st>>>(d_coor_x, d_coor_y, d_coor_z, d_roi_coord_2d, d_region, d_fps,
                                                                       d_mask, R, mask_mode, region_attention, mask_attention, d_out, C);
    else
        rdpn::coor_feat_kernel<256><<<B * (RDPN_P / 256), 256, smem, st>>>(d_coor_x, d_coor_y, d_coor_z, d_roi_coord_2d, d_region, d_fps,
                                                                       d_mask, R, mask_mode, region_attention, mask_attention, d_out, C);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}
