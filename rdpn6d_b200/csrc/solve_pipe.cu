// Dense-correspondence -> pose as a pipeline of three kernels (BASELINE.json north_star: one kernel per roofline), chained
// by programmatic dependent launch.  No cooperative launch, no grid barrier, no block-wide barrier on the common path:
// K1 and K3 are grids of INDEPENDENT WARPS (the latency of one ROI's dependent steps is hidden by the SM's other warps,
// each at a different step of a different ROI), K2 is one small CTA per ROI whose warps only meet at one mbarrier wait.
//
//   K1  front_kernel   one WARP per ROI.  HBM / latency bound.  Mask min/max (engine_utils.py:123-124), the gate
//                      (gdrn_evaluator.py:110-117) mask first -- only quads with a passing pixel load depth / coor /
//                      region ids at all (requested by L2 prefetch one pass ahead) --, back-projection + residual (S1,
//                      data_loader.py:530-576) for the gated pixels, written as a compact list in raster order, a stable
//                      counting sort of that list by region id (dense match_any rounds), the hypothesis poses (FP64
//                      closed form from the list, rounded once to FP32; misc.py:91-106) compacted by validity.
//                      Output: one "package" per ROI in the workspace.
//   K2  score_kernel   one 128-thread CTA per ROI.  FP32-issue bound.  Points, run table and hypothesis poses staged in
//                      shared memory by bulk TMA (one mbarrier); up to FOUR hypotheses per lane, the (pass, point) plane
//                      cut into four slices of equal cost, one per warp; per run the transformed anchor R a + t once,
//                      per point 3 FADD + 3 FFMA + LEA.HI per hypothesis (misc.py:108-111; solve_common.cuh).
//                      Output: inlier counts per hypothesis (and FP32 residual sums for select rule MIN_MEAN_ERR).
//   K3  refit_kernel   one WARP per ROI: best hypothesis (misc.py:121, adaptive stop :134-138), inliers of the winner,
//                      18 FP64 moments by warp shuffle, closed-form rotation (transform.py:913-980), every output.
//       refit_minerr_kernel: the same warp walks the hypotheses as the reference loop does and returns the
//                      lowest-mean-error pose (misc.py:113-132).
//
// Hand-over: one flag per ROI and stage (st.release / ld.acquire); K2 / K3 CTAs are scheduled as soon as every CTA of
// the kernel before has started (griddepcontrol.launch_dependents) and wait for exactly the ROI they work on.
// The arithmetic contracts are those of pose_solve.cu (solve_common.cuh): gate, counts, winner and inlier masks are
// bit-identical to the fused kernel, the refit pose equal to FP32 rounding (FP64 sums in a different order).
// Packages live in a caller-provided (or per-stream cached) workspace; batches larger than the workspace holds run
// chunk after chunk.
#include "solve_common.cuh"

#include <stdlib.h>

namespace rdpn {
extern unsigned long long g_launch_count;
int device_sm_count(int* sms);                                           // host_api.cu (per-device cache)
int ensure_func_smem(const void* func, int kernel_slot, size_t bytes);   // host_api.cu (per-device attribute cache)

// ---------------------------------------------------------------------------------------------
// package layout (byte offsets into one ROI's package; every section 128-byte aligned)
// ---------------------------------------------------------------------------------------------
struct PkgLayout {
    unsigned runtab;  // float4[R]      (anchor xyz, start | end << 16) per non-empty region bucket
    unsigned vh;      // uint16[H]      compacted index -> hypothesis index
    unsigned hcnt;    // int32[H]       K2: inlier count per compacted hypothesis
    unsigned herr;    // float[H]       K2, select rule MIN_MEAN_ERR: FP32 sum of the residual norms over all points
    unsigned hyp;     // float4[3][H]   compacted FP32 hypothesis poses, row-planar (row r of pose j at [r * H + j])
    unsigned key;     // uint32[P]      raster list: pixel | region id << 16
    unsigned rast;    // float4[P]      raster list: (cam xyz, w)
    unsigned srid;    // uint8[P]       sorted slot -> region id
    unsigned pix;     // uint16[P]      sorted slot -> pixel
    unsigned slots;   // float4[P]      sorted slot -> (cam xyz, w)
    unsigned long long stride;
};
// header at offset 0: int4 (n gated correspondences, nruns non-empty buckets, nvalid hypotheses, 0)
static PkgLayout make_layout(int H, int R) {
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    PkgLayout l;
    size_t off = 128;
    l.runtab = (unsigned)off; off = al(off + (size_t)R * 16);
    l.vh = (unsigned)off;     off = al(off + (size_t)H * 2);
    l.hcnt = (unsigned)off;   off = al(off + (size_t)H * 4);
    l.herr = (unsigned)off;   off = al(off + (size_t)H * 4);
    l.hyp = (unsigned)off;    off = al(off + (size_t)H * 48);
    l.key = (unsigned)off;    off = al(off + (size_t)RDPN_P * 4);
    l.rast = (unsigned)off;   off = al(off + (size_t)RDPN_P * 16);
    l.srid = (unsigned)off;   off = al(off + RDPN_P);
    l.pix = (unsigned)off;    off = al(off + RDPN_P * 2);
    l.slots = (unsigned)off;  off = al(off + (size_t)RDPN_P * 16);
    l.stride = off;
    return l;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- per-ROI hand-over flags between the kernels (programmatic dependent launch, see launch_split) ----
__device__ __forceinline__ void st_release_i32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// A wait that lasts longer than ~2 s of SM clocks is a protocol bug: trap instead of hanging the device.
__device__ __forceinline__ void wait_flag(const int* p) {
    if (ld_acquire_u32(reinterpret_cast<const unsigned*>(p)) != 0u) return;
    const long long t0 = clock64();
    unsigned ns = 32;
    while (ld_acquire_u32(reinterpret_cast<const unsigned*>(p)) == 0u) {
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        if (clock64() - t0 > (1ll << 32)) __trap();
    }
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// =============================================================================================
// K1: gate + S1 + sort + hypotheses, one warp per ROI
// =============================================================================================
#ifndef RDPN_FRONT_WARPS
#define RDPN_FRONT_WARPS 4          // warps (ROIs) per CTA
#endif
constexpr int FR_W = RDPN_FRONT_WARPS;
constexpr int FR_T = FR_W * 32;
#ifndef RDPN_FRONT_PREFETCH
#define RDPN_FRONT_PREFETCH 1       // L1 prefetch of the next sort trip / the next round's sampled pairs
#endif
#ifndef RDPN_FRONT_CTAS
#define RDPN_FRONT_CTAS 5           // CTAs per SM the register budget is sized for: 96 registers per thread, no spills
                                    // (8 CTAs at 64 registers spill the FP64 hypothesis solve: measured 4 % slower overall)
#endif

struct FrontLayout {  // per-warp shared memory (byte offsets)
    unsigned selmap;   // uint32[128]   gate bitmap by pixel
    unsigned selpfx;   // uint16[132]   gated pixels before word w (raster order), [128] = all
    unsigned cur;      // uint32[R + 1] bucket histogram, then cursors
    unsigned anchors;  // float4[R]
    unsigned scr;      // 2 KB: 8 doubles per lane (3-pair solve) | 16 ints per lane (S > 3 sample indices)
    unsigned per_warp;
    int rpl;           // buckets per lane in the bucket scan
};
static FrontLayout make_front_layout(int R) {
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    FrontLayout l;
    size_t off = 0;
    l.selmap = (unsigned)off;  off = al(off + 128 * 4);
    l.selpfx = (unsigned)off;  off = al(off + 132 * 2);
    l.cur = (unsigned)off;     off = al(off + (size_t)(R + 1) * 4);
    l.anchors = (unsigned)off; off = al(off + (size_t)R * 16);
    l.scr = (unsigned)off;     off = al(off + 2048);
    l.per_warp = (unsigned)((off + 127) & ~(size_t)127);
    l.rpl = (R + 31) / 32;
    return l;
}

// raster rank of a gated pixel (index into the raster list)
__device__ __forceinline__ int raster_rank(const uint32_t* selmap, const uint16_t* selpfx, int px) {
    return (int)selpfx[px >> 5] + __popc(selmap[px >> 5] & ((1u << (px & 31)) - 1u));
}

// Hypothesis from S > 3 pairs (misc.py:72,91 samples random_sample_num = 10): Kabsch of the S pairs,
// transform.py:913-980 semantics; same arithmetic as pose_solve.cu:hyp_from_sample, the pairs read from the raster list.
static __device__ __noinline__ bool hyp_from_sample_list(const uint32_t* selmap, const uint16_t* selpfx, const float4* anchors,
                                                         const float4* rast, const uint32_t* key, int* ii /* [S][32] column */,
                                                         const int32_t* idx_in, uint32_t kroi, int h, int S, float* P) {
    const uint32_t nsel = selpfx[RDPN_P / 32];
    for (int v = 0; v < S; ++v) {
        int px;
        if (idx_in) {
            px = idx_in[v];
        } else {
            // without replacement, as np.random.choice(..., replace=False) at misc.py:91: a draw that repeats an earlier
            // pixel of the sample is re-drawn from the same counter stream (attempt a in bits 20..23 of the counter), up to
            // RDPN_SAMPLE_REDRAWS times; attempt 0 is the plain stream
            px = -1;
            for (int att = 0; att <= RDPN_SAMPLE_REDRAWS && nsel; ++att) {
                const uint32_t kk = fmix32(kroi ^ (uint32_t)(S * h + v) ^ ((uint32_t)att << 20));
                px = kth_gated_pixel(selmap, selpfx, (uint32_t)(((unsigned long long)kk * nsel) >> 32));
                bool dup = false;
                for (int u = 0; u < v; ++u) dup = dup || ii[u * 32] == px;
                if (!dup) break;
            }
        }
        if ((unsigned)px >= RDPN_P) return false;
        if (!((selmap[px >> 5] >> (px & 31)) & 1u)) return false;
        for (int u = 0; u < v; ++u)
            if (ii[u * 32] == px) return false;
        ii[v * 32] = px;
    }
    double m[17];  // sum c (3) | sum a (3) | sum c a^T (9) | sum |c|^2 | sum |a|^2, all about pair 0
#pragma unroll
    for (int i = 0; i < 17; ++i) m[i] = 0.0;
    float c0f[3] = {0.f, 0.f, 0.f}, a0f[3] = {0.f, 0.f, 0.f}, cpf[3] = {0.f, 0.f, 0.f}, apf[3] = {0.f, 0.f, 0.f};
    bool ok_a = false, ok_c = false;
#pragma unroll 1
    for (int v = 0; v < S; ++v) {
        const int rk = raster_rank(selmap, selpfx, ii[v * 32]);
        const float4 cw = rast[rk];
        const float4 ob = anchors[key[rk] >> 16];
        if (v == 0) {
            c0f[0] = cw.x; c0f[1] = cw.y; c0f[2] = cw.z;
            a0f[0] = ob.x; a0f[1] = ob.y; a0f[2] = ob.z;
        } else {
            if (v >= 2 && !(ok_a && ok_c)) {
                const double p0a[3] = {(double)a0f[0], (double)a0f[1], (double)a0f[2]};
                const double p1a[3] = {(double)apf[0], (double)apf[1], (double)apf[2]};
                const double p2a[3] = {(double)ob.x, (double)ob.y, (double)ob.z};
                const double p0c[3] = {(double)c0f[0], (double)c0f[1], (double)c0f[2]};
                const double p1c[3] = {(double)cpf[0], (double)cpf[1], (double)cpf[2]};
                const double p2c[3] = {(double)cw.x, (double)cw.y, (double)cw.z};
                if (!ok_a) ok_a = triangle_ok(p0a, p1a, p2a);
                if (!ok_c) ok_c = triangle_ok(p0c, p1c, p2c);
            }
            const double c[3] = {(double)cw.x - (double)c0f[0], (double)cw.y - (double)c0f[1], (double)cw.z - (double)c0f[2]};
            const double a[3] = {(double)ob.x - (double)a0f[0], (double)ob.y - (double)a0f[1], (double)ob.z - (double)a0f[2]};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                m[i] += c[i];
                m[3 + i] += a[i];
#pragma unroll
                for (int j = 0; j < 3; ++j) m[6 + 3 * i + j] += c[i] * a[j];
            }
            m[15] += c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            m[16] += a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
        }
        cpf[0] = cw.x; cpf[1] = cw.y; cpf[2] = cw.z;
        apf[0] = ob.x; apf[1] = ob.y; apf[2] = ob.z;
    }
    if (!ok_a || !ok_c) return false;
    const double inv = 1.0 / (double)S;
    double Sc[9], R[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Sc[3 * i + j] = m[6 + 3 * i + j] - m[i] * (m[3 + j] * inv);
    const double ga = m[16] - (m[3] * m[3] + m[4] * m[4] + m[5] * m[5]) * inv;
    const double gb = m[15] - (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) * inv;
    rotation_from_cov(Sc, ga, gb, R);
    const double ma[3] = {m[3] * inv + (double)a0f[0], m[4] * inv + (double)a0f[1], m[5] * inv + (double)a0f[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double mc = m[r] * inv + (double)c0f[r];
        P[4 * r + 0] = (float)R[3 * r + 0];
        P[4 * r + 1] = (float)R[3 * r + 1];
        P[4 * r + 2] = (float)R[3 * r + 2];
        P[4 * r + 3] = (float)(mc - (R[3 * r] * ma[0] + R[3 * r + 1] * ma[1] + R[3 * r + 2] * ma[2]));
    }
    return true;
}

// Mask test of one quad -> 4 bits.  L1 mode: the two-sided FP32 filter of gate.cuh decides almost every pixel; the
// exact FP64 comparison runs only when some lane of the warp has a pixel inside the filter's band (warp vote), so the
// common path is 4 FADD + 8 FSETP.  Decisions are those of mask_pass(), bit for bit.
// the exact comparison of the pixels inside the filter's band: rare, kept out of line so that the unrolled mask test
// stays short in the instruction cache
static __device__ __noinline__ unsigned mask_nibble_exact(float a0, float a1, float a2, float a3, unsigned amb, float gb, double cut,
                                                          int incl) {
    const float av[4] = {a0, a1, a2, a3};
    const double r = __dmul_rn((double)gb, cut);
    unsigned nib = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if ((amb >> j) & 1u) {
            const double l = (double)av[j];
            if (incl ? (l >= r) : (l > r)) nib |= 1u << j;  // NaN -> false
        }
    return nib;
}
__device__ __forceinline__ unsigned mask_nibble(const float4& m, int mode, float thr, float mn, const RoiGate& g) {
    const float mm[4] = {m.x, m.y, m.z, m.w};
    unsigned nib = 0u;
    if (mode == RDPN_MASK_L1) {
        if (!(g.b > 0.f)) return 0u;  // flat mask: 0/0 = NaN never passes (engine_utils.py:128 has no eps)
        unsigned amb = 0u;
        float av[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            av[j] = __fsub_rn(mm[j], mn);
            const bool in = av[j] > g.hi, out = av[j] < g.lo;
            nib |= (in ? 1u : 0u) << j;
            amb |= ((!in && !out) ? 1u : 0u) << j;
        }
        if (__any_sync(0xffffffffu, amb != 0u)) nib |= mask_nibble_exact(av[0], av[1], av[2], av[3], amb, g.b, g.cut, g.incl);
        return nib;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) nib |= (mask_pass(mm[j], mode, thr, mn, g) ? 1u : 0u) << j;
    return nib;
}

// L1MASK: the mask mode is RDPN_MASK_L1 (the default of the reference, gdrn_base.py:46) at compile time -- the sigmoid /
// plain-probability code of the other modes stays out of the instruction stream of the unrolled mask test
template <bool MULTI, bool L1MASK>
__global__ void __launch_bounds__(FR_T, RDPN_FRONT_CTAS) front_kernel(SolveArgs a, unsigned char* __restrict__ ws, PkgLayout lay,
                                                                        FrontLayout fl, int* __restrict__ fdone) {
    extern __shared__ __align__(128) unsigned char fsm[];
    pdl_launch_dependents();  // K2's CTAs may take the slots this grid's last wave leaves free; they wait per ROI on fdone
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const rdpn_roi_inputs& in = a.in;
    const int b = blockIdx.x * FR_W + warp;
    if (b >= in.B) return;  // warps are independent: no block-wide barrier anywhere below
    unsigned char* wsm = fsm + (size_t)warp * fl.per_warp;
    uint32_t* selmap = reinterpret_cast<uint32_t*>(wsm + fl.selmap);
    uint16_t* selpfx = reinterpret_cast<uint16_t*>(wsm + fl.selpfx);
    uint32_t* cur = reinterpret_cast<uint32_t*>(wsm + fl.cur);
    float4* anchors = reinterpret_cast<float4*>(wsm + fl.anchors);
    const int H = a.prm.num_hyp, R = in.num_regions;
    unsigned char* pkg = ws + (size_t)b * lay.stride;
    uint32_t* key_g = reinterpret_cast<uint32_t*>(pkg + lay.key);
    float4* rast_g = reinterpret_cast<float4*>(pkg + lay.rast);
    const size_t po = (size_t)b * RDPN_P;
    const float4* mask4 = reinterpret_cast<const float4*>(in.mask + po);
    const float4* dep4 = reinterpret_cast<const float4*>(in.depth + po);
    const float4* cx4 = reinterpret_cast<const float4*>(in.coor_x + po);
    const float4* cy4 = reinterpret_cast<const float4*>(in.coor_y + po);
    const float4* cz4 = reinterpret_cast<const float4*>(in.coor_z + po);
    const uchar4* rid4 = reinterpret_cast<const uchar4*>(in.region_idx + po);

    // everything this warp will certainly read is requested from DRAM now: the mask plane (128 lines) and the ROI's
    // hypothesis samples; the other planes are requested per quad in pass 2
#pragma unroll
    for (int u = 0; u < 4; ++u) asm volatile("prefetch.global.L2 [%0];" ::"l"(in.mask + po + 32 * (32 * u + lane)));
    if (a.hyp_idx) {
        const int hb = a.prm.num_hyp * (MULTI ? a.prm.sample_size : 3) * 4;  // bytes of this ROI's samples
        const char* hp = reinterpret_cast<const char*>(a.hyp_idx) + (size_t)b * hb;
        for (int o = 128 * lane; o < hb; o += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + o));
    }
    // ---- 0: per-ROI constants (every lane: warp-uniform loads), anchors, zeroed histogram ----
    RoiConst rc;
    rc.fx = __ldg(in.Kp + 4 * b + 0);
    rc.fy = __ldg(in.Kp + 4 * b + 1);
    rc.cx = __ldg(in.Kp + 4 * b + 2);
    rc.cy = __ldg(in.Kp + 4 * b + 3);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float e = __ldg(in.extent + 3 * b + c);
        rc.ext[c] = e;
        rc.gthr[c] = (float)(0.0001 * (double)e);  // gdrn_evaluator.py:112-114 under numpy-1.23 promotion
    }
    rc.div = in.depth_div ? __ldg(in.depth_div + b) : 0.f;
    rc.mn = rc.mx = 0.f;
    for (int r = lane; r < R; r += 32) {
        const float* ap = in.anchors + ((size_t)b * R + r) * 3;
        anchors[r] = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), 0.f);
    }
    for (int r = lane; r <= R; r += 32) cur[r] = 0u;
    if (a.out.hyp_counts)
        for (int h = lane; h < H; h += 32) a.out.hyp_counts[(size_t)b * H + h] = 0;

    // ---- 1: mask min / max (engine_utils.py:123-124); lane owns quads q = 32 k + lane ----
    RoiGate gate;
    gate.hi = gate.lo = gate.b = 0.f;
    gate.cut = 0.0;
    gate.incl = 0;
    const int mask_mode = L1MASK ? (int)RDPN_MASK_L1 : in.mask_mode;
    if (mask_mode == RDPN_MASK_L1) {
        float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) minmax4(__ldg(mask4 + 32 * k + lane), mn, mx);
        mn = warp_min(mn);
        mx = warp_max(mx);
        rc.mn = mn;
        rc.mx = mx;
        make_gate(gate, mn, mx, in.mask_thr, a.mask_cut, a.mask_cut_incl);
    }

    // ---- 2: mask test of all 128 pixels of the lane (the gate is a conjunction: mask first).  Quads with a passing
    //         pixel go to a compact list in raster order (quad | mask nibble << 10) and request their depth / coor /
    //         region-id sectors from DRAM now (L2 prefetch), consumed in pass 3 ----
    uint16_t* qlist = reinterpret_cast<uint16_t*>(wsm + fl.scr);  // <= 1024 entries; the scratch is free until pass 6
    int nq = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) selmap[4 * lane + u] = 0u;
#pragma unroll 1
    for (int kk = 0; kk < 4; ++kk) {
        float4 mq[8];
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) mq[k8] = __ldg(mask4 + 32 * (8 * kk + k8) + lane);
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
            const int q = 32 * (8 * kk + k8) + lane;
            const unsigned nib = mask_nibble(mq[k8], mask_mode, in.mask_thr, rc.mn, gate);
            const unsigned bal = __ballot_sync(0xffffffffu, nib != 0u);
            if (nib) {
                qlist[nq + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)((unsigned)q | (nib << 10));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dep4 + q));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cx4 + q));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cy4 + q));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cz4 + q));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rid4 + q));
            }
            nq += __popc(bal);
        }
    }
    __syncwarp();  // quad list, anchors, zeroed histogram / bitmap visible

    // ---- 3: rest of the gate (gdrn_evaluator.py:110-117 + depth validity) on the listed quads, 32 at a time; S1 for
    //         the gated pixels, raster list, bucket histogram, gate bitmap ----
    int running = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < nq; i0 += 32) {
        unsigned nib = 0u;
        int q = 0;
        float dd[4], cxn[4], cyn[4], czn[4], mw[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t rr[4];
        if (i0 + lane < nq) {
            const unsigned ent = qlist[i0 + lane];
            q = (int)(ent & 1023u);
            nib = ent >> 10;
            const float4 dq = __ldg(dep4 + q), xq = __ldg(cx4 + q), yq = __ldg(cy4 + q), zq = __ldg(cz4 + q);
            const uchar4 r4 = __ldg(rid4 + q);
            if (a.prm.weighted) {  // the weights of the quad with the other planes, not pixel by pixel behind the gate
                const float4 m4 = __ldg(mask4 + q);
                mw[0] = m4.x; mw[1] = m4.y; mw[2] = m4.z; mw[3] = m4.w;
            }
            dd[0] = dq.x; dd[1] = dq.y; dd[2] = dq.z; dd[3] = dq.w;
            cxn[0] = xq.x; cxn[1] = xq.y; cxn[2] = xq.z; cxn[3] = xq.w;
            cyn[0] = yq.x; cyn[1] = yq.y; cyn[2] = yq.z; cyn[3] = yq.w;
            czn[0] = zq.x; czn[1] = zq.y; czn[2] = zq.z; czn[3] = zq.w;
            rr[0] = r4.x; rr[1] = r4.y; rr[2] = r4.z; rr[3] = r4.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float dv = dd[j];
                if (rc.div != 0.f) {  // zero lanes would drag the warp through div.rn's slow path: 0 / f = +-0
                    const float qd = __fdiv_rn(dv == 0.f ? 1.f : dv, rc.div);
                    dv = dv == 0.f ? __fmul_rn(dv, copysignf(1.f, rc.div)) : qd;
                }
                const float dx = __fmul_rn(__fsub_rn(cxn[j], 0.5f), rc.ext[0]);
                const float dy = __fmul_rn(__fsub_rn(cyn[j], 0.5f), rc.ext[1]);
                const float dz = __fmul_rn(__fsub_rn(czn[j], 0.5f), rc.ext[2]);
                const bool sel = (fabsf(dx) > rc.gthr[0]) && (fabsf(dy) > rc.gthr[1]) && (fabsf(dz) > rc.gthr[2]) &&
                                 (dv > 0.f) && ((int)rr[j] < R);  // ids outside [0, R) never pass
                if (!sel) nib &= ~(1u << j);
            }
        }
        // raster rank: the list is in quad order and pixel = 4 q + j, so the order within the step is (lane, j)
        const int c = __popc(nib);
        const int incl = warp_incl_scan(c, lane);
        int slot = running + incl - c;
        running += __shfl_sync(0xffffffffu, incl, 31);
        if (nib) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!((nib >> j) & 1u)) continue;
                const int p = 4 * q + j;
                float cam[3], obj[3];
                pixel_s1<false>(rc, p, dd[j], cxn[j], cyn[j], czn[j], cam, obj);
                const float w = a.prm.weighted ? mask_prob(mw[j], mask_mode, rc.mn, rc.mx) : 1.f;
                rast_g[slot] = make_float4(cam[0], cam[1], cam[2], w);
                key_g[slot] = (uint32_t)p | ((uint32_t)rr[j] << 16);
                atomicAdd(&cur[rr[j]], 1u);
                ++slot;
            }
            atomicOr(&selmap[q >> 3], nib << (4 * (q & 7)));  // gate bitmap: 4 bits per quad, 8 quads per word
        }
    }
    const int n = running;
    __syncwarp();

    // ---- 4: bucket starts (ascending region id), run table of the non-empty buckets, cursors ----
    int nruns;
    {
        const int RPL = fl.rpl;
        int loc = 0, ne = 0;
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R) {
                const int c = (int)cur[r];
                loc += c;
                ne += c > 0 ? 1 : 0;
            }
        }
        const int x = warp_incl_scan(loc, lane), kx = warp_incl_scan(ne, lane);
        int run = x - loc, kk = kx - ne;
        float4* runtab_g = reinterpret_cast<float4*>(pkg + lay.runtab);
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R) {
                const int c = (int)cur[r];
                const int start = run;
                cur[r] = (uint32_t)run;
                run += c;
                if (c > 0) {  // one entry per NON-EMPTY bucket = (anchor xyz, start | end << 16)
                    float4 hd = anchors[r];
                    hd.w = __uint_as_float((unsigned)start | ((unsigned)run << 16));
                    runtab_g[kk++] = hd;
                }
            }
        }
        nruns = __shfl_sync(0xffffffffu, kx, 31);
    }
    // gated pixels before each 32-pixel word (raster order): rank queries of the hypothesis stage
    {
        int pc[4], loc = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            pc[u] = __popc(selmap[4 * lane + u]);
            loc += pc[u];
        }
        int run = warp_incl_scan(loc, lane) - loc;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            selpfx[4 * lane + u] = (uint16_t)run;
            run += pc[u];
        }
        if (lane == 31) selpfx[RDPN_P / 32] = (uint16_t)run;
    }
    __syncwarp();  // cursors, selpfx, and this warp's raster list (global) visible to all lanes

    // ---- 5: stable counting sort of the raster list by region id: dense match_any rounds ----
    {
        float4* slots_g = reinterpret_cast<float4*>(pkg + lay.slots);
        uint8_t* srid_g = pkg + lay.srid;
        uint16_t* pix_g = reinterpret_cast<uint16_t*>(pkg + lay.pix);
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += 64) {  // two rounds per trip: both rounds' loads are in flight before the first match
            uint32_t kv[2] = {0u, 0u};
            float4 cw[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = i0 + 32 * u + lane;
                cw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < n) {
                    kv[u] = key_g[i];
                    cw[u] = rast_g[i];
                }
#if RDPN_FRONT_PREFETCH
                if (i + 64 < n) {  // the next trip's lines on their way from L2 while this trip is matched
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(key_g + i + 64));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rast_g + i + 64));
                }
#endif
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (i0 + 32 * u >= n) break;  // warp-uniform
                const bool valid = i0 + 32 * u + lane < n;
                const unsigned rid = valid ? (kv[u] >> 16) : (unsigned)R;
                const unsigned m = __match_any_sync(0xffffffffu, rid);
                const int leader = __ffs(m) - 1;
                int c0 = 0;
                if (valid && lane == leader) {
                    c0 = (int)cur[rid];
                    cur[rid] = (uint32_t)(c0 + __popc(m));
                }
                c0 = __shfl_sync(0xffffffffu, c0, leader);
                if (valid) {
                    const int sl = c0 + __popc(m & ((1u << lane) - 1u));
                    slots_g[sl] = cw[u];
                    srid_g[sl] = (uint8_t)rid;
                    pix_g[sl] = (uint16_t)(kv[u] & 0xFFFFu);
                }
                __syncwarp();
            }
        }
    }

    // ---- 6: hypotheses (FP64 closed form, rounded once to FP32), compacted by validity ----
    const bool enough = n >= a.prm.min_pts;
    int nvalid = 0;
    if (!enough) {
        if (a.out.hyp_poses) {
            float4* hp = reinterpret_cast<float4*>(a.out.hyp_poses + (size_t)b * H * 12);
            for (int i = lane; i < 3 * H; i += 32) hp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        const bool sampling = a.hyp_idx == nullptr;
        const int SS = MULTI ? a.prm.sample_size : 3;
        const uint32_t kroi = fmix32(fmix32(a.prm.seed ^ 0x9e3779b9u) ^ (uint32_t)(a.prm.roi_base + b));
        const uint32_t nsel = (uint32_t)n;
        // explicit samples: the next round's indices are loaded while this round computes
        int nx0 = -1, nx1 = -1, nx2 = -1;
        if (!MULTI && !sampling && lane < H) {
            const int32_t* ip = a.hyp_idx + ((size_t)b * H + lane) * 3;
            nx0 = __ldg(ip); nx1 = __ldg(ip + 1); nx2 = __ldg(ip + 2);
        }
#if RDPN_FRONT_PREFETCH
        int nn0 = -1, nn1 = -1, nn2 = -1;  // the round after
        if (!MULTI && !sampling && lane + 32 < H) {
            const int32_t* ip = a.hyp_idx + ((size_t)b * H + lane + 32) * 3;
            nn0 = __ldg(ip); nn1 = __ldg(ip + 1); nn2 = __ldg(ip + 2);
        }
#endif
#pragma unroll 1
        for (int h0 = 0; h0 < H; h0 += 32) {
            const int h = h0 + lane;
            float P[12];
            bool ok = false;
            const int cu0 = nx0, cu1 = nx1, cu2 = nx2;
#if RDPN_FRONT_PREFETCH
            if (!MULTI && !sampling) {
                // indices run two rounds ahead, so that the pairs of the NEXT round can be requested from L2 now
                nx0 = nn0; nx1 = nn1; nx2 = nn2;
                if (h + 64 < H) {
                    const int32_t* ip = a.hyp_idx + ((size_t)b * H + h + 64) * 3;
                    nn0 = __ldg(ip); nn1 = __ldg(ip + 1); nn2 = __ldg(ip + 2);
                }
                if (h + 32 < H && ((unsigned)nx0 < RDPN_P) && ((unsigned)nx1 < RDPN_P) && ((unsigned)nx2 < RDPN_P)) {
                    const int r0 = min(raster_rank(selmap, selpfx, nx0), n - 1), r1 = min(raster_rank(selmap, selpfx, nx1), n - 1),
                              r2 = min(raster_rank(selmap, selpfx, nx2), n - 1);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rast_g + r0));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rast_g + r1));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rast_g + r2));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(key_g + r0));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(key_g + r1));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(key_g + r2));
                }
            }
#else
            if (!MULTI && !sampling && h + 32 < H) {
                const int32_t* ip = a.hyp_idx + ((size_t)b * H + h + 32) * 3;
                nx0 = __ldg(ip); nx1 = __ldg(ip + 1); nx2 = __ldg(ip + 2);
            }
#endif
            if (h < H) {
                if (MULTI) {
                    ok = hyp_from_sample_list(selmap, selpfx, anchors, rast_g, key_g, reinterpret_cast<int*>(wsm + fl.scr) + lane,
                                              sampling ? nullptr : a.hyp_idx + ((size_t)b * H + h) * SS, kroi, h, SS, P);
                } else {
                    int ii[3];
                    if (!sampling) {
                        ii[0] = cu0;
                        ii[1] = cu1;
                        ii[2] = cu2;
                    } else {
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            const uint32_t kk = fmix32(kroi ^ (uint32_t)(3 * h + v));
                            const uint32_t kth = (uint32_t)(((unsigned long long)kk * nsel) >> 32);
                            ii[v] = nsel ? kth_gated_pixel(selmap, selpfx, kth) : -1;
                        }
                    }
                    ok = ((unsigned)ii[0] < RDPN_P) && ((unsigned)ii[1] < RDPN_P) && ((unsigned)ii[2] < RDPN_P);
                    if (ok) {
#pragma unroll
                        for (int v = 0; v < 3; ++v) ok = ok && ((selmap[ii[v] >> 5] >> (ii[v] & 31)) & 1u);
                    }
                    if (ok) {
                        // register-lean closed form (kabsch_math.cuh): object side first, parked in shared memory
                        double* scr = reinterpret_cast<double*>(wsm + fl.scr) + lane;  // scr[k * 32], k < 8
                        float pf[3][3];
                        float4 cw[3];
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            const int rk = raster_rank(selmap, selpfx, ii[v]);
                            cw[v] = rast_g[rk];
                            const float4 an = anchors[key_g[rk] >> 16];
                            pf[v][0] = an.x; pf[v][1] = an.y; pf[v][2] = an.z;
                        }
                        double ma2, u1a, u2a, v2a;
                        {
                            TriSide sa;
                            tri_side(pf, sa);
                            ok = sa.ok;
                            scr[0 * 32] = sa.e1[0]; scr[1 * 32] = sa.e1[1]; scr[2 * 32] = sa.e1[2];
                            scr[3 * 32] = sa.n[0];  scr[4 * 32] = sa.n[1];  scr[5 * 32] = sa.n[2];
                            scr[6 * 32] = sa.m[0];  scr[7 * 32] = sa.m[1];
                            ma2 = sa.m[2]; u1a = sa.u1; u2a = sa.u2; v2a = sa.v2;
                        }
#pragma unroll
                        for (int v = 0; v < 3; ++v) { pf[v][0] = cw[v].x; pf[v][1] = cw[v].y; pf[v][2] = cw[v].z; }
                        TriSide sc;
                        tri_side(pf, sc);
                        ok = ok && sc.ok;
                        if (ok) kabsch3_sides(scr, 32, ma2, u1a, u2a, v2a, sc, P);
                    }
                }
                if (a.out.hyp_poses) {
                    float4* hp = reinterpret_cast<float4*>(a.out.hyp_poses + ((size_t)b * H + h) * 12);
                    hp[0] = ok ? make_float4(P[0], P[1], P[2], P[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    hp[1] = ok ? make_float4(P[4], P[5], P[6], P[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    hp[2] = ok ? make_float4(P[8], P[9], P[10], P[11]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int j = nvalid + __popc(bal & ((1u << lane) - 1u));
                float4* hp = reinterpret_cast<float4*>(pkg + lay.hyp) + j;  // row-planar: row r of pose j at [r * H + j]
                hp[0] = make_float4(P[0], P[1], P[2], P[3]);
                hp[H] = make_float4(P[4], P[5], P[6], P[7]);
                hp[2 * H] = make_float4(P[8], P[9], P[10], P[11]);
                reinterpret_cast<uint16_t*>(pkg + lay.vh)[j] = (uint16_t)h;
            }
            nvalid += __popc(bal);
        }
    }
    if (lane == 0) {
        *reinterpret_cast<int4*>(pkg) = make_int4(n, nruns, nvalid, 0);
        if (a.out.n_sel) a.out.n_sel[b] = n;
    }
    // every lane's package writes are ordered before the flag: the warp barrier orders them before lane 0's store, whose
    // release at GPU scope is cumulative (the pattern of a block barrier + one releasing thread)
    // (no __threadfence() of all lanes on top: measured -0.6 % of the step)
    __syncwarp();
    if (lane == 0) st_release_i32(fdone + b, 1);
}

// =============================================================================================
// K2: hypotheses x points inlier scoring, one CTA per ROI
// =============================================================================================
constexpr int SC_W = 4;             // warps per CTA
constexpr int SC_T = SC_W * 32;
#ifndef RDPN_SCORE_CHUNK
#define RDPN_SCORE_CHUNK 1024
#endif
constexpr int SC_CHUNK = RDPN_SCORE_CHUNK;  // gated slots staged per pass (16 KB); ROIs with more loop over chunks
#ifndef RDPN_SCORE_REGS_CTAS
#define RDPN_SCORE_REGS_CTAS 8      // register budget: 64 per thread
#endif
struct ScoreLayout {  // dynamic shared memory (byte offsets)
    unsigned pts;     // float4[SC_CHUNK]
    unsigned hyp;     // float4[3][H]
    unsigned runtab;  // float4[R]
    unsigned hcnt;    // int[H]
    unsigned herr;    // float[H] (select rule MIN_MEAN_ERR)
    unsigned bar;     // mbarrier
    unsigned total;
};
static ScoreLayout make_score_layout(int H, int R) {
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    ScoreLayout l;
    size_t off = 0;
    l.pts = (unsigned)off;    off = al(off + (size_t)SC_CHUNK * 16);
    l.hyp = (unsigned)off;    off = al(off + (size_t)H * 48);
    l.runtab = (unsigned)off; off = al(off + (size_t)R * 16);
    l.hcnt = (unsigned)off;   off = al(off + (size_t)H * 4);
    l.herr = (unsigned)off;   off = al(off + (size_t)H * 4);
    l.bar = (unsigned)off;    off = al(off + 8);
    l.total = (unsigned)off;
    return l;
}

// Non-persistent on purpose: a CTA's synchronisation is one mbarrier wait on its bulk-TMA copy (points, run table and
// hypothesis poses into shared memory) and one barrier before the counts are written back; the SM's other resident
// CTAs cover both.  The (pass, point) plane is cut into SC_W slices of equal cost, one per warp, whatever the number
// of valid hypotheses.
template <bool MEAN>
__global__ void __launch_bounds__(SC_T, RDPN_SCORE_REGS_CTAS) score_kernel(int H, int min_pts, float cut, unsigned char* __restrict__ ws,
                                                                            PkgLayout lay, ScoreLayout sl, const int* __restrict__ fdone,
                                                                            int* __restrict__ sdone) {
    extern __shared__ __align__(128) unsigned char sc_raw[];
    const float4* pts = reinterpret_cast<const float4*>(sc_raw + sl.pts);
    const float4* hyp = reinterpret_cast<const float4*>(sc_raw + sl.hyp);
    const float4* runtab = reinterpret_cast<const float4*>(sc_raw + sl.runtab);
    int* hcnt_s = reinterpret_cast<int*>(sc_raw + sl.hcnt);
    float* herr_s = reinterpret_cast<float*>(sc_raw + sl.herr);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sc_raw + sl.bar);
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    unsigned char* pkg = ws + (size_t)b * lay.stride;
    pdl_launch_dependents();  // K3's warps wait per ROI on sdone
    if (threadIdx.x == 0) {
        wait_flag(fdone + b);  // this ROI's package is complete (K1 may still be running elsewhere)
        asm volatile("fence.proxy.async;" ::: "memory");  // written through the generic proxy, read by bulk TMA
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();  // package visible, barrier initialised
    const int4 hd = __ldcg(reinterpret_cast<const int4*>(pkg));
    const int n = hd.x, nruns = hd.y, nvalid = hd.z;
    if (nvalid <= 0 || n <= 0 || n < min_pts) {  // uniform across the CTA: nothing to score
        if (threadIdx.x == 0) st_release_i32(sdone + b, 1);
        return;
    }
    for (int j = threadIdx.x; j < nvalid; j += SC_T) {
        hcnt_s[j] = 0;
        if (MEAN) herr_s[j] = 0.f;
    }
    __syncthreads();  // counts zeroed
    unsigned phase = 0;
    for (int c0 = 0; c0 < n; c0 += SC_CHUNK) {
        const int c1 = min(n, c0 + SC_CHUNK);
        if (c0 > 0) __syncthreads();  // rare: previous chunk fully scored before its buffer is overwritten
        if (threadIdx.x == 0) {
            const uint32_t b_pts = (uint32_t)(c1 - c0) * 16u;
            const uint32_t b_run = c0 == 0 ? (uint32_t)nruns * 16u : 0u;
            const uint32_t b_hyp = c0 == 0 ? (((uint32_t)nvalid * 16u + 15u) & ~15u) : 0u;
            mbar_expect_tx(bar, b_pts + b_run + 3u * b_hyp);
            bulk_g2s(sc_raw + sl.pts, pkg + lay.slots + (size_t)c0 * 16, b_pts, bar);
            if (b_run) {
                bulk_g2s(sc_raw + sl.runtab, pkg + lay.runtab, b_run, bar);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    bulk_g2s(sc_raw + sl.hyp + (size_t)r * H * 16, pkg + lay.hyp + (size_t)r * H * 16, b_hyp, bar);
            }
        }
        mbar_wait_backoff(bar, phase & 1);
        score_slices<PosePlanar, MEAN>(pts, runtab, nruns, PosePlanar{hyp, H}, nvalid, c0, c1, cut, hcnt_s, warp, SC_W, herr_s);
        ++phase;
    }
    __syncthreads();  // every warp's partial counts are in
    int* hcnt = reinterpret_cast<int*>(pkg + lay.hcnt);
    float* herr = reinterpret_cast<float*>(pkg + lay.herr);
    for (int j = threadIdx.x; j < nvalid; j += SC_T) {
        hcnt[j] = hcnt_s[j];
        if (MEAN) herr[j] = herr_s[j];
    }
    __syncthreads();  // + thread 0's release at GPU scope: every thread's counts are visible before the flag
    if (threadIdx.x == 0) st_release_i32(sdone + b, 1);
}

// =============================================================================================
// K3: best hypothesis + refit on its inliers + every output, one warp per ROI
// =============================================================================================
#ifndef RDPN_REFIT_WARPS
#define RDPN_REFIT_WARPS 4
#endif
constexpr int RF_W = RDPN_REFIT_WARPS;  // warps (ROIs) per CTA
#ifndef RDPN_REFIT_CTAS
#define RDPN_REFIT_CTAS 4               // CTAs per SM the register budget is sized for
#endif
constexpr int RF_U = 4;                 // slots per lane in flight

__global__ void __launch_bounds__(RF_W * 32, RDPN_REFIT_CTAS) refit_kernel(SolveArgs a, const unsigned char* __restrict__ ws, PkgLayout lay,
                                                                           const int* __restrict__ sdone) {
    extern __shared__ __align__(16) unsigned char rf_smem[];  // float[RF_W][3 R]: the ROI's anchors, one row per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RF_W + warp;
    if (b >= a.in.B) return;
    const int H = a.prm.num_hyp;
    const unsigned char* pkg = ws + (size_t)b * lay.stride;
    if (lane == 0) wait_flag(sdone + b);  // this ROI's counts are final (K2 may still be running elsewhere)
    __syncwarp();
    // package reads below go to L2 (ld.global.cg): the kernels overlap, so this SM's L1 is not known to be clean
    const int4 h0 = __ldcg(reinterpret_cast<const int4*>(pkg));
    const int n = h0.x, nvalid = h0.z;
    const bool enough = n >= a.prm.min_pts;
    if (a.out.inlier_mask) {  // zero-fill; inliers are scattered in after the last refit
        uint4* im = reinterpret_cast<uint4*>(a.out.inlier_mask + (size_t)b * RDPN_P);
        for (int i = lane; i < RDPN_P / 16; i += 32) im[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // ---- best hypothesis (misc.py:121) with optional adaptive stop (misc.py:134-138) ----
    const int* hcnt = reinterpret_cast<const int*>(pkg + lay.hcnt);
    const uint16_t* vh = reinterpret_cast<const uint16_t*>(pkg + lay.vh);
    int best_j = -1, nbest = 0;
    if (enough && nvalid > 0) {
        int jlim = nvalid;
        if (a.prm.adaptive) {
            // i_ransac of compacted entry j is j + 1; stop after the first j with i_ransac > max(k, min_iter)
            const double lc = log10(1.0 - (double)a.prm.confidence);
            int js = 0x7FFFFFFF;
            for (int j = lane; j < nvalid; j += 32)
                if (adaptive_stop(__ldcg(hcnt + j), n, j + 1, lc, a.prm.min_iter)) js = min(js, j);
            js = warp_min_i(js);
            if (js != 0x7FFFFFFF) jlim = js + 1;
        }
        unsigned long long kbest = 0ull;
        for (int j = lane; j < nvalid; j += 32) {
            const int c = __ldcg(hcnt + j);
            if (a.out.hyp_counts) a.out.hyp_counts[(size_t)b * H + __ldcg(vh + j)] = c;
            if (j < jlim && c >= a.prm.min_inliers && c > 0) {
                const unsigned long long k = ((unsigned long long)(unsigned)c << 32) | (unsigned)(0x7FFFFFFF - j);
                kbest = k > kbest ? k : kbest;
            }
        }
        kbest = warp_max_u64(kbest);
        if (kbest) {
            best_j = 0x7FFFFFFF - (int)(kbest & 0xFFFFFFFFull);
            nbest = (int)(kbest >> 32);
        }
    }
    if (best_j < 0) {
        const float st = enough ? (float)RDPN_STATUS_NO_CONSENSUS : (float)RDPN_STATUS_FEW_POINTS;
        if (lane < 12) a.out.pose[(size_t)b * 12 + lane] = -100.f;  // gdrn_evaluator.py:395
        if (a.out.rows16 && lane < 16)
            a.out.rows16[(size_t)b * 16 + lane] =
                lane < 12 ? -100.f : (lane == 12 ? 0.f : (lane == 13 ? st : (lane == 14 ? (float)n : -1.f)));
        if (lane == 0) {
            a.out.n_inliers[b] = 0;
            a.out.status[b] = enough ? RDPN_STATUS_NO_CONSENSUS : RDPN_STATUS_FEW_POINTS;
            if (a.out.best_h) a.out.best_h[b] = -1;
            if (a.out.scale) a.out.scale[b] = 1.f;
        }
        return;
    }
    const int R3 = 3 * a.in.num_regions;
    float* anc = reinterpret_cast<float*>(rf_smem) + (size_t)warp * R3;
    {
        const float* ag = a.in.anchors + (size_t)b * R3;
        for (int i = lane; i < R3; i += 32) anc[i] = __ldg(ag + i);
        __syncwarp();
    }
    const int best = (int)__ldcg(vh + best_j);
    float P[12];
    {
        const float4* hp = reinterpret_cast<const float4*>(pkg + lay.hyp) + best_j;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float4 v = __ldcg(hp + (size_t)r * H);
            P[4 * r] = v.x; P[4 * r + 1] = v.y; P[4 * r + 2] = v.z; P[4 * r + 3] = v.w;
        }
    }
    const float4* slots = reinterpret_cast<const float4*>(pkg + lay.slots);
    const uint8_t* srid = pkg + lay.srid;
    const uint16_t* pixs = reinterpret_cast<const uint16_t*>(pkg + lay.pix);
    const float ncut = -a.sq_cut;  // the contract's inlier test: margin(..., ncut) < 0 (solve_common.cuh)
    float out_scale = 1.f;
    const int iters = a.prm.refit_iters < 1 ? 1 : a.prm.refit_iters;
    // pivot of the raw moments (exact FP32 differences): slot 0
    const float4 cp0 = __ldcg(slots);
    float4 ap0;
    {
        const int r0 = 3 * (int)__ldcg(srid);
        ap0 = make_float4(anc[r0], anc[r0 + 1], anc[r0 + 2], 0.f);
    }
    for (int it = 0; it < iters; ++it) {
        // FP64 raw moments about the pivot over the inliers of the current pose (misc.py:123-126 -> transform.py:921-928)
        double m[18];  // sum w | w c (3) | w a (3) | w c a^T (9) | w |c|^2 | w |a|^2
#pragma unroll
        for (int i = 0; i < 18; ++i) m[i] = 0.0;
        int ninl = 0;
        const bool mark = a.out.inlier_mask && it == iters - 1;
        unsigned marks = 0u;  // inliers among this lane's first 32 slots (the rest is re-tested when marking)
        for (int i0 = lane; i0 < n; i0 += 32 * RF_U) {
            float4 cpv[RF_U];
            int ridv[RF_U];
#pragma unroll
            for (int u = 0; u < RF_U; ++u) {  // every load of the batch in flight before the first use
                const int i = i0 + 32 * u;
                if (i < n) {
                    cpv[u] = __ldcg(slots + i);
                    ridv[u] = (int)__ldcg(srid + i);
                }
            }
#pragma unroll
            for (int u = 0; u < RF_U; ++u) {
                const int i = i0 + 32 * u;
                if (i >= n) break;
                const float4 cp = cpv[u];
                const float4 ap = make_float4(anc[3 * ridv[u]], anc[3 * ridv[u] + 1], anc[3 * ridv[u] + 2], 0.f);
                if (is_inlier(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z, ncut)) {
                    ++ninl;
                    const int k = (i - lane) >> 5;
                    if (k < 32) marks |= 1u << k;
                    const double w = a.prm.weighted ? (double)cp.w : 1.0;
                    const double c0 = (double)cp.x - (double)cp0.x, c1 = (double)cp.y - (double)cp0.y, c2 = (double)cp.z - (double)cp0.z;
                    const double a0 = (double)ap.x - (double)ap0.x, a1 = (double)ap.y - (double)ap0.y, a2 = (double)ap.z - (double)ap0.z;
                    const double wc0 = w * c0, wc1 = w * c1, wc2 = w * c2;
                    m[0] += w;
                    m[1] += wc0; m[2] += wc1; m[3] += wc2;
                    m[4] += w * a0; m[5] += w * a1; m[6] += w * a2;
                    m[7] += wc0 * a0; m[8] += wc0 * a1; m[9] += wc0 * a2;
                    m[10] += wc1 * a0; m[11] += wc1 * a1; m[12] += wc1 * a2;
                    m[13] += wc2 * a0; m[14] += wc2 * a1; m[15] += wc2 * a2;
                    m[16] += w * (c0 * c0 + c1 * c1 + c2 * c2);
                    m[17] += w * (a0 * a0 + a1 * a1 + a2 * a2);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 18; ++i) m[i] = warp_sum(m[i]);
        ninl = warp_sum(ninl);
        if (ninl < 3) break;  // uniform across the warp; the pose of the previous round stays
        if (mark) {  // the inlier set used by the last refit
            for (int i = lane, k = 0; i < n; i += 32, ++k) {
                bool inl;
                if (k < 32) {
                    inl = (marks >> k) & 1u;
                } else {
                    const float4 cp = __ldcg(slots + i);
                    const int r3 = 3 * (int)__ldcg(srid + i);
                    inl = is_inlier(P, anc[r3], anc[r3 + 1], anc[r3 + 2], cp.x, cp.y, cp.z, ncut);
                }
                if (inl) a.out.inlier_mask[(size_t)b * RDPN_P + __ldcg(pixs + i)] = 1;
            }
        }
        // every lane solves the same 3 x 3 problem (no divergence, no broadcast)
        const double isw = 1.0 / m[0];
        double S[9], Rm[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) S[3 * r + c] = m[7 + 3 * r + c] - m[1 + r] * (m[4 + c] * isw);  // sum w c a^T - (sum w c)(mean a)^T
        const double ga = m[17] - (m[4] * (m[4] * isw) + m[5] * (m[5] * isw) + m[6] * (m[6] * isw));
        const double gb = m[16] - (m[1] * (m[1] * isw) + m[2] * (m[2] * isw) + m[3] * (m[3] * isw));
        rotation_from_cov(S, ga, gb, Rm);
        const double sc = a.prm.with_scale ? sqrt(gb / ga) : 1.0;  // transform.py:971-975
        const double ma0 = m[4] * isw + (double)ap0.x, ma1 = m[5] * isw + (double)ap0.y, ma2 = m[6] * isw + (double)ap0.z;
        const double c0v[3] = {(double)cp0.x, (double)cp0.y, (double)cp0.z};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double mcr = m[1 + r] * isw + c0v[r];
            const double r0 = Rm[3 * r], r1 = Rm[3 * r + 1], r2 = Rm[3 * r + 2];
            P[4 * r + 0] = (float)(sc * r0);
            P[4 * r + 1] = (float)(sc * r1);
            P[4 * r + 2] = (float)(sc * r2);
            P[4 * r + 3] = (float)(mcr - sc * (r0 * ma0 + r1 * ma1 + r2 * ma2));
        }
        out_scale = (float)sc;
    }
    // ---- outputs (+ translation sanity, gdrn_evaluator.py:293-296) ----
    int status = RDPN_STATUS_OK;
    if (a.t_net) {
        const float t0 = a.t_net[3 * b], t1 = a.t_net[3 * b + 1], t2 = a.t_net[3 * b + 2];
        const double d0 = (double)t0 - P[3], d1 = (double)t1 - P[7], d2 = (double)t2 - P[11];
        if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 1.0) {
            status = RDPN_STATUS_T_SANITY;
            P[3] = t0;
            P[7] = t1;
            P[11] = t2;
        }
    }
    float mine = 0.f;  // lane l holds row element l
#pragma unroll
    for (int i = 0; i < 12; ++i)
        if (lane == i) mine = P[i];
    if (lane == 12) mine = (float)nbest;
    if (lane == 13) mine = (float)status;
    if (lane == 14) mine = (float)n;
    if (lane == 15) mine = (float)best;
    if (lane < 12) a.out.pose[(size_t)b * 12 + lane] = mine;
    if (a.out.rows16 && lane < 16) a.out.rows16[(size_t)b * 16 + lane] = mine;
    if (lane == 0) {
        a.out.n_inliers[b] = nbest;
        a.out.status[b] = status;
        if (a.out.best_h) a.out.best_h[b] = best;
        if (a.out.scale) a.out.scale[b] = out_scale;
    }
}

// =============================================================================================
// K3, select rule MIN_MEAN_ERR: the reference loop's return value (lib/pysixd/misc.py:108-142), one warp per ROI
// =============================================================================================
struct WarpRoi {  // what a warp needs to walk its ROI's region-sorted correspondences
    const float4* slots;
    const uint8_t* srid;
    const float* anc;   // shared memory: the ROI's anchors
    int n;
    float4 cp0, ap0;    // pivot of the raw moments (slot 0)
};
// Kabsch / Umeyama refit on the inliers of Pin (misc.py:123-126 -> transform.py:913-980): FP64 raw moments about the pivot by
// warp shuffle, closed-form rotation.  Returns the number of inliers (uniform); Pout is written when there are >= 3.
static __device__ __noinline__ int warp_refit(const WarpRoi& w, const float* Pin, float ncut, int weighted, int with_scale, float* Pout) {
    const int lane = threadIdx.x & 31;
    double m[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) m[i] = 0.0;
    int ninl = 0;
    for (int i = lane; i < w.n; i += 32) {
        const float4 cp = __ldcg(w.slots + i);
        const int r3 = 3 * (int)__ldcg(w.srid + i);
        const float ax = w.anc[r3], ay = w.anc[r3 + 1], az = w.anc[r3 + 2];
        if (is_inlier(Pin, ax, ay, az, cp.x, cp.y, cp.z, ncut)) {
            ++ninl;
            const double wt = weighted ? (double)cp.w : 1.0;
            const double c0 = (double)cp.x - (double)w.cp0.x, c1 = (double)cp.y - (double)w.cp0.y, c2 = (double)cp.z - (double)w.cp0.z;
            const double a0 = (double)ax - (double)w.ap0.x, a1 = (double)ay - (double)w.ap0.y, a2 = (double)az - (double)w.ap0.z;
            const double wc0 = wt * c0, wc1 = wt * c1, wc2 = wt * c2;
            m[0] += wt;
            m[1] += wc0; m[2] += wc1; m[3] += wc2;
            m[4] += wt * a0; m[5] += wt * a1; m[6] += wt * a2;
            m[7] += wc0 * a0; m[8] += wc0 * a1; m[9] += wc0 * a2;
            m[10] += wc1 * a0; m[11] += wc1 * a1; m[12] += wc1 * a2;
            m[13] += wc2 * a0; m[14] += wc2 * a1; m[15] += wc2 * a2;
            m[16] += wt * (c0 * c0 + c1 * c1 + c2 * c2);
            m[17] += wt * (a0 * a0 + a1 * a1 + a2 * a2);
        }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) m[i] = warp_sum(m[i]);
    ninl = warp_sum(ninl);
    if (ninl < 3) return ninl;
    const double isw = 1.0 / m[0];
    double S[9], Rm[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) S[3 * r + c] = m[7 + 3 * r + c] - m[1 + r] * (m[4 + c] * isw);
    const double ga = m[17] - (m[4] * (m[4] * isw) + m[5] * (m[5] * isw) + m[6] * (m[6] * isw));
    const double gb = m[16] - (m[1] * (m[1] * isw) + m[2] * (m[2] * isw) + m[3] * (m[3] * isw));
    rotation_from_cov(S, ga, gb, Rm);
    const double sc = with_scale ? sqrt(gb / ga) : 1.0;
    const double ma0 = m[4] * isw + (double)w.ap0.x, ma1 = m[5] * isw + (double)w.ap0.y, ma2 = m[6] * isw + (double)w.ap0.z;
    const double c0v[3] = {(double)w.cp0.x, (double)w.cp0.y, (double)w.cp0.z};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double mcr = m[1 + r] * isw + c0v[r];
        const double r0 = Rm[3 * r], r1 = Rm[3 * r + 1], r2 = Rm[3 * r + 2];
        Pout[4 * r + 0] = (float)(sc * r0);
        Pout[4 * r + 1] = (float)(sc * r1);
        Pout[4 * r + 2] = (float)(sc * r2);
        Pout[4 * r + 3] = (float)(mcr - sc * (r0 * ma0 + r1 * ma1 + r2 * ma2));
    }
    return ninl;
}
// mean residual norm of a pose over ALL gated points in FP64 (misc.py:109,113: errs.mean()), lane-strided sums combined
// in a fixed order
static __device__ __noinline__ double warp_mean_err(const WarpRoi& w, const float* P) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int i = lane; i < w.n; i += 32) {
        const float4 cp = __ldcg(w.slots + i);
        const int r3 = 3 * (int)__ldcg(w.srid + i);
        const double ax = w.anc[r3], ay = w.anc[r3 + 1], az = w.anc[r3 + 2];
        const double dx = ((double)P[0] * ax + (double)P[1] * ay + (double)P[2] * az + (double)P[3]) - (double)cp.x;
        const double dy = ((double)P[4] * ax + (double)P[5] * ay + (double)P[6] * az + (double)P[7]) - (double)cp.y;
        const double dz = ((double)P[8] * ax + (double)P[9] * ay + (double)P[10] * az + (double)P[11]) - (double)cp.z;
        s += sqrt(dx * dx + dy * dy + dz * dz);
    }
    return warp_sum(s) / (double)w.n;
}

// The loop of misc.py:89-138 on the precomputed counts: hypotheses in order (i_ransac = compacted index + 1),
//   (:113-116) a sample fit whose mean error over all points beats the best so far becomes the best pose;
//   (:118-132) a sample fit that raises the best inlier count (and has >= min_inliers) is refit on its inliers, and the
//              refit becomes the best pose if ITS mean error beats the best so far;
//   (:134-138) the adaptive stop, when enabled, ends the loop.
// K2's FP32 sums pre-select: a sample fit is a candidate for (:113) only if its FP32 mean is within 1e-4 of the best
// FP64 mean so far (the FP32 sum of a few hundred norms is good to ~1e-6), and only candidates are re-evaluated in FP64.
__global__ void __launch_bounds__(RF_W * 32, RDPN_REFIT_CTAS) refit_minerr_kernel(SolveArgs a, const unsigned char* __restrict__ ws,
                                                                                  PkgLayout lay, const int* __restrict__ sdone) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RF_W + warp;
    if (b >= a.in.B) return;
    const int H = a.prm.num_hyp;
    const unsigned char* pkg = ws + (size_t)b * lay.stride;
    if (lane == 0) wait_flag(sdone + b);
    __syncwarp();
    const int4 h0 = __ldcg(reinterpret_cast<const int4*>(pkg));
    const int n = h0.x, nvalid = h0.z;
    const bool enough = n >= a.prm.min_pts;
    if (a.out.inlier_mask) {
        uint4* im = reinterpret_cast<uint4*>(a.out.inlier_mask + (size_t)b * RDPN_P);
        for (int i = lane; i < RDPN_P / 16; i += 32) im[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int* hcnt = reinterpret_cast<const int*>(pkg + lay.hcnt);
    const float* herr = reinterpret_cast<const float*>(pkg + lay.herr);
    const uint16_t* vh = reinterpret_cast<const uint16_t*>(pkg + lay.vh);
    const float4* hp = reinterpret_cast<const float4*>(pkg + lay.hyp);
    float Pbest[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Pbest[i] = -100.f;
    int src = -1, best_inl = 0;
    bool have = false;
    if (enough && nvalid > 0) {
        const int R3 = 3 * a.in.num_regions;
        float* anc = reinterpret_cast<float*>(rf_smem) + (size_t)warp * R3;
        const float* ag = a.in.anchors + (size_t)b * R3;
        for (int i = lane; i < R3; i += 32) anc[i] = __ldg(ag + i);
        __syncwarp();
        WarpRoi w;
        w.slots = reinterpret_cast<const float4*>(pkg + lay.slots);
        w.srid = pkg + lay.srid;
        w.anc = anc;
        w.n = n;
        w.cp0 = __ldcg(w.slots);
        {
            const int r0 = 3 * (int)__ldcg(w.srid);
            w.ap0 = make_float4(anc[r0], anc[r0 + 1], anc[r0 + 2], 0.f);
        }
        int jlim = nvalid;
        if (a.prm.adaptive) {
            const double lc = log10(1.0 - (double)a.prm.confidence);
            int js = 0x7FFFFFFF;
            for (int j = lane; j < nvalid; j += 32)
                if (adaptive_stop(__ldcg(hcnt + j), n, j + 1, lc, a.prm.min_iter)) js = min(js, j);
            js = warp_min_i(js);
            if (js != 0x7FFFFFFF) jlim = js + 1;
        }
        if (a.out.hyp_counts)
            for (int j = lane; j < nvalid; j += 32) a.out.hyp_counts[(size_t)b * H + __ldcg(vh + j)] = __ldcg(hcnt + j);
        double best_err = 1e300;
        const float ncut = -a.sq_cut;
        for (int j = 0; j < jlim; ++j) {  // warp-uniform walk in hypothesis order
            const int c = __ldcg(hcnt + j);
            const double e32 = (double)__ldcg(herr + j) / (double)n;
            const bool cand = !have || e32 < best_err * 1.0001;
            const bool improves = c > best_inl && c >= a.prm.min_inliers;
            if (!cand && !improves) continue;
            float P[12];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 v = __ldcg(hp + (size_t)r * H + j);
                P[4 * r] = v.x; P[4 * r + 1] = v.y; P[4 * r + 2] = v.z; P[4 * r + 3] = v.w;
            }
            if (cand) {  // misc.py:113-116
                const double e = warp_mean_err(w, P);
                if (e < best_err) {
                    best_err = e;
                    have = true;
                    src = (int)__ldcg(vh + j);
#pragma unroll
                    for (int i = 0; i < 12; ++i) Pbest[i] = P[i];
                }
            }
            if (improves) {  // misc.py:118-132
                best_inl = c;
                float Pr[12];
                if (warp_refit(w, P, ncut, a.prm.weighted, a.prm.with_scale, Pr) >= 3) {
                    const double e = warp_mean_err(w, Pr);
                    if (e < best_err) {
                        best_err = e;
                        have = true;
                        src = (int)__ldcg(vh + j);
#pragma unroll
                        for (int i = 0; i < 12; ++i) Pbest[i] = Pr[i];
                    }
                }
            }
        }
        if (have && a.out.inlier_mask) {  // the inliers of the RETURNED pose
            const uint16_t* pixs = reinterpret_cast<const uint16_t*>(pkg + lay.pix);
            for (int i = lane; i < n; i += 32) {
                const float4 cp = __ldcg(w.slots + i);
                const int r3 = 3 * (int)__ldcg(w.srid + i);
                if (is_inlier(Pbest, anc[r3], anc[r3 + 1], anc[r3 + 2], cp.x, cp.y, cp.z, ncut))
                    a.out.inlier_mask[(size_t)b * RDPN_P + __ldcg(pixs + i)] = 1;
            }
        }
    }
    int status = have ? RDPN_STATUS_OK : (enough ? RDPN_STATUS_NO_CONSENSUS : RDPN_STATUS_FEW_POINTS);
    if (have && a.t_net) {
        const float t0 = a.t_net[3 * b], t1 = a.t_net[3 * b + 1], t2 = a.t_net[3 * b + 2];
        const double d0 = (double)t0 - Pbest[3], d1 = (double)t1 - Pbest[7], d2 = (double)t2 - Pbest[11];
        if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 1.0) {
            status = RDPN_STATUS_T_SANITY;
            Pbest[3] = t0;
            Pbest[7] = t1;
            Pbest[11] = t2;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i)
        if (lane == i) mine = Pbest[i];
    if (lane == 12) mine = (float)best_inl;
    if (lane == 13) mine = (float)status;
    if (lane == 14) mine = (float)n;
    if (lane == 15) mine = (float)src;
    if (lane < 12) a.out.pose[(size_t)b * 12 + lane] = mine;
    if (a.out.rows16 && lane < 16) a.out.rows16[(size_t)b * 16 + lane] = mine;
    if (lane == 0) {
        a.out.n_inliers[b] = best_inl;
        a.out.status[b] = status;
        if (a.out.best_h) a.out.best_h[b] = src;
        if (a.out.scale) a.out.scale[b] = 1.f;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
static SolveArgs shifted(const SolveArgs& a, int b0, int nb) {
    SolveArgs c = a;
    const int H = a.prm.num_hyp, S = a.prm.sample_size, R = a.in.num_regions;
    const size_t po = (size_t)b0 * RDPN_P;
    c.in.B = nb;
    c.in.depth += po; c.in.coor_x += po; c.in.coor_y += po; c.in.coor_z += po; c.in.mask += po;
    c.in.Kp += (size_t)b0 * 4;
    c.in.extent += (size_t)b0 * 3;
    if (c.in.depth_div) c.in.depth_div += b0;
    if (c.in.region_idx) c.in.region_idx += po;
    if (c.in.anchors) c.in.anchors += (size_t)b0 * R * 3;
    if (c.hyp_idx) c.hyp_idx += (size_t)b0 * H * S;
    if (c.t_net) c.t_net += (size_t)b0 * 3;
    c.prm.roi_base = a.prm.roi_base + b0;
    c.out.pose += (size_t)b0 * 12;
    c.out.n_inliers += b0;
    c.out.status += b0;
    if (c.out.best_h) c.out.best_h += b0;
    if (c.out.n_sel) c.out.n_sel += b0;
    if (c.out.inlier_mask) c.out.inlier_mask += po;
    if (c.out.hyp_counts) c.out.hyp_counts += (size_t)b0 * H;
    if (c.out.hyp_poses) c.out.hyp_poses += (size_t)b0 * H * 12;
    if (c.out.scale) c.out.scale += b0;
    if (c.out.rows16) c.out.rows16 += (size_t)b0 * 16;
    return c;
}

enum { SLOT_FRONT = 8, SLOT_SCORE = 10, SLOT_REFIT = 12 };  // attribute-cache slots (+ multi; front: + 8 for the generic mask mode)

bool split_supported(const SolveArgs& a, bool dense) {
    if (dense) return false;               // dense mode (no region runs to share the transformed anchor) stays fused
    if (a.prm.num_hyp > 2048) return false;  // poses of all hypotheses are staged in shared memory (48 B each)
    return true;
}

size_t split_pkg_stride(int H, int R, bool dense) {
    (void)dense;
    return (size_t)make_layout(H, R).stride;
}

// Workspace: [fdone[chunk], sdone[chunk] : int, 128-byte aligned] [chunk packages]
static size_t flags_bytes(int chunk) { return (2 * (size_t)chunk * sizeof(int) + 127) & ~(size_t)127; }

size_t split_workspace_bytes(int B, int H, int R, int chunk_rois) {
    int chunk = chunk_rois > 0 && chunk_rois < B ? chunk_rois : B;
    return flags_bytes(chunk) + (size_t)chunk * make_layout(H, R).stride;
}

// K2 and K3 are launched with programmatic stream serialisation (PDL): their CTAs may be scheduled as soon as every CTA
// of the kernel before has STARTED (griddepcontrol.launch_dependents is the first instruction of K1 / K2), i.e. into the
// SM slots the last, partially filled wave of that kernel leaves free, and each of them waits for exactly the ROI it
// works on (acquire-load on the ROI's flag) instead of for the whole grid.  No deadlock: a waiting CTA is only ever
// resident when every CTA it can wait for is resident or done.
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// rdpn_pose_solve_stage_ms: when set, every chunk runs its kernels one after the other (no PDL overlap) with CUDA events
// between them on the launching stream and adds the three durations here
thread_local float* t_stage_ms = nullptr;

template <bool MULTI>
static int launch_chunk(const SolveArgs& a, unsigned char* ws, const PkgLayout& lay, cudaStream_t st) {
    const int B = a.in.B, H = a.prm.num_hyp, R = a.in.num_regions;
    static const bool pdl_env = env_int("RDPN_PIPE_PDL", 1) != 0;
    const bool pdl = pdl_env && !t_stage_ms;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (t_stage_ms) {
        for (auto& e : ev) RDPN_CUDA_TRY(cudaEventCreate(&e));
    }
    const size_t fb = flags_bytes(B);
    int* fdone = reinterpret_cast<int*>(ws);
    int* sdone = fdone + B;
    unsigned char* pk = ws + fb;
    RDPN_CUDA_TRY(cudaMemsetAsync(ws, 0, fb, st));
    if (t_stage_ms) RDPN_CUDA_TRY(cudaEventRecord(ev[0], st));
    {  // K1
        const FrontLayout fl = make_front_layout(R);
        const size_t smem = (size_t)FR_W * fl.per_warp;
        if (smem > 227 * 1024) return RDPN_E_TOOLARGE;
        const bool l1 = a.in.mask_mode == RDPN_MASK_L1;
        void (*kern)(SolveArgs, unsigned char*, PkgLayout, FrontLayout, int*) = l1 ? front_kernel<MULTI, true> : front_kernel<MULTI, false>;
        const int rc = ensure_func_smem((const void*)kern, SLOT_FRONT + (MULTI ? 1 : 0) + (l1 ? 0 : 8), smem);
        if (rc) return rc;
        kern<<<(B + FR_W - 1) / FR_W, FR_T, smem, st>>>(a, pk, lay, fl, fdone);
        ++g_launch_count;
        RDPN_LAUNCH_CHECK();
    }
    if (t_stage_ms) RDPN_CUDA_TRY(cudaEventRecord(ev[1], st));
    {  // K2
        const ScoreLayout sl = make_score_layout(H, R);
        if (sl.total > 227 * 1024) return RDPN_E_TOOLARGE;
        const bool mean = a.prm.select_rule == RDPN_SELECT_MIN_MEAN_ERR;
        void (*kern)(int, int, float, unsigned char*, PkgLayout, ScoreLayout, const int*, int*) = mean ? score_kernel<true> : score_kernel<false>;
        const int rc = ensure_func_smem((const void*)kern, SLOT_SCORE + (mean ? 1 : 0), sl.total);
        if (rc) return rc;
        RDPN_CUDA_TRY(launch_pdl(kern, dim3(B), dim3(SC_T), sl.total, st, pdl, H, a.prm.min_pts, a.sq_cut, pk, lay, sl,
                                 (const int*)fdone, sdone));
        ++g_launch_count;
    }
    if (t_stage_ms) RDPN_CUDA_TRY(cudaEventRecord(ev[2], st));
    {  // K3
        const size_t smem = (size_t)RF_W * 3 * R * sizeof(float);
        const bool mean = a.prm.select_rule == RDPN_SELECT_MIN_MEAN_ERR;
        void (*kern)(SolveArgs, const unsigned char*, PkgLayout, const int*) = mean ? refit_minerr_kernel : refit_kernel;
        const int rc = ensure_func_smem((const void*)kern, SLOT_REFIT + (mean ? 1 : 0), smem);
        if (rc) return rc;
        RDPN_CUDA_TRY(launch_pdl(kern, dim3((B + RF_W - 1) / RF_W), dim3(RF_W * 32), smem, st, pdl, a, (const unsigned char*)pk,
                                 lay, (const int*)sdone));
        ++g_launch_count;
    }
    if (t_stage_ms) {
        RDPN_CUDA_TRY(cudaEventRecord(ev[3], st));
        RDPN_CUDA_TRY(cudaEventSynchronize(ev[3]));
        for (int i = 0; i < 3; ++i) {
            float ms = 0.f;
            RDPN_CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            t_stage_ms[i] += ms;
        }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return 0;
}

void set_stage_timing(float* ms3) { t_stage_ms = ms3; }

// K1 -> K2 -> K3 on the caller's stream, in chunks of as many ROIs as the workspace holds (chunk_rois caps it).
int launch_split(const SolveArgs& a, bool dense, void* ws, size_t ws_bytes, int chunk_rois, cudaStream_t st) {
    if (dense) return RDPN_E_TOOLARGE;
    const int H = a.prm.num_hyp, R = a.in.num_regions, B = a.in.B;
    const PkgLayout lay = make_layout(H, R);
    if (!ws || ((uintptr_t)ws & 127)) return RDPN_E_WORKSPACE;
    int chunk = chunk_rois > 0 && chunk_rois < B ? chunk_rois : B;
    while (chunk > 1 && flags_bytes(chunk) + (size_t)chunk * lay.stride > ws_bytes) chunk = (chunk + 1) / 2;
    if (flags_bytes(chunk) + (size_t)chunk * lay.stride > ws_bytes) return RDPN_E_WORKSPACE;
    const bool multi = a.prm.sample_size > 3;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const SolveArgs c = shifted(a, b0, B - b0 < chunk ? B - b0 : chunk);
        const int rc = multi ? launch_chunk<true>(c, (unsigned char*)ws, lay, st) : launch_chunk<false>(c, (unsigned char*)ws, lay, st);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace rdpn
