// Dense-correspondence -> pose as a pipeline of three kernels, one per roofline (BASELINE.json north_star):
//
//   K1  gate_pack_kernel   HBM-bound.  Persistent, one CTA per SM; a producer warp streams whole ROIs (five FP32
//                          planes + region ids, 86 KB) into a two-stage shared-memory ring by 1-D bulk TMA while two
//                          groups of eight warps each work on one staged ROI: mask min/max, gate, deterministic
//                          counting sort of the gated pixels by region, S1 arithmetic (back-projection + residual) for
//                          the gated pixels only, hypothesis generation (FP64 closed form) from the staged planes.
//                          Output per ROI: a compact "package" (~20 KB touched): sorted correspondences
//                          (cam xyz, w) as float4, region runs with their anchors, the valid hypothesis poses.
//   K2  score_kernel       FP32-issue-bound.  Persistent; a producer warp bulk-TMAs the next ROI's package into the
//                          other half of a two-stage ring while eight warps score hypotheses x points of the current
//                          one (two hypotheses per thread, per region run the transformed anchor once), then a block
//                          arg-max picks the winner (misc.py:121, optional adaptive stop misc.py:134-138).
//   K3  refit_kernel       one WARP per ROI: inlier test of the winner, 18 FP64 moments by warp-shuffle reduction,
//                          closed-form rotation (Horn / QCP), Umeyama scale, translation sanity, every output.
//
// The arithmetic contracts are those of pose_solve.cu (solve_common.cuh); results are bit-identical for counts, masks
// and the winner, and equal to FP32 rounding for the refit pose (FP64 sums in a different order).
// The packages live in a caller-provided (or per-stream cached) workspace; large batches run in chunks of ROIs so that
// a chunk's packages stay in the 126 MB L2 between K1 and K2.
#include "solve_common.cuh"

#include <mutex>

namespace rdpn {
extern unsigned long long g_launch_count;
int device_sm_count(int* sms);                                           // host_api.cu (per-device cache)
int ensure_func_smem(const void* func, int kernel_slot, size_t bytes);   // host_api.cu (per-device attribute cache)

// ---------------------------------------------------------------------------------------------
// package layout
// ---------------------------------------------------------------------------------------------
struct __align__(16) PkgHdr {
    int n;        // gated correspondences
    int nruns;    // non-empty region buckets
    int nvalid;   // valid hypotheses (compacted, ascending h)
    int flags;    // reserved
    int best_j;   // K2: index into the compacted list or -1
    int n_best;   // K2: inlier count of the winner
    int pad[2];
};
struct PkgLayout {
    unsigned runtab;  // float4[R]      (anchor xyz, start | end << 16) per non-empty bucket
    unsigned vh;      // uint16[H]      compacted index -> hypothesis index
    unsigned hyp;     // float[H][12]   compacted FP32 hypothesis poses
    unsigned srid;    // uint8[P]       slot -> region id
    unsigned pix;     // uint16[P]      slot -> pixel
    unsigned slots;   // float4[P]      slot -> (cam xyz, w)   (dense mode: second array of obj xyz behind it)
    unsigned long long stride;
};
static PkgLayout make_layout(int H, int R, bool dense) {
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    PkgLayout l;
    size_t off = al(sizeof(PkgHdr));
    l.runtab = (unsigned)off; off = al(off + (size_t)R * 16);
    l.vh = (unsigned)off;     off = al(off + (size_t)H * 2);
    l.hyp = (unsigned)off;    off = al(off + (size_t)H * 48);
    l.srid = (unsigned)off;   off = al(off + RDPN_P);
    l.pix = (unsigned)off;    off = al(off + RDPN_P * 2);
    l.slots = (unsigned)off;  off = al(off + (size_t)RDPN_P * 16 * (dense ? 2 : 1));
    l.stride = off;
    return l;
}

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// =============================================================================================
// K1: gate + sort + stage + hypotheses
// =============================================================================================
constexpr int PK_GT = 256;                    // threads per group (one ROI at a time per group)
constexpr int PK_GW = PK_GT / 32;
constexpr int PK_NG = 2;                      // groups = ring stages
constexpr int PK_NT = PK_NG * PK_GT + 32;     // + producer warp
constexpr int PK_QPT = RDPN_P / 4 / PK_GT;    // pixel quads per thread (4)
constexpr int PK_MAX_R = 128;                 // larger R: fused kernel

struct __align__(128) PackStage {
    float tile[5][RDPN_P];  // depth, coor_x, coor_y, coor_z, mask
    uint8_t rid[RDPN_P];
};
struct __align__(16) PackGroup {
    uint16_t pix[RDPN_P];
    uint8_t srid[RDPN_P];
    uint32_t selmap[RDPN_P / 32];
    uint16_t selpfx[RDPN_P / 32 + 8];
    RoiConst rc;
    float red_f[2][PK_GW];
    int red_i[PK_GW];
    int red_j[PK_GW];
    int n_sel;
    int n_runs;
};
struct __align__(128) PackSmem {
    PackStage st[PK_NG];
    uint64_t full[PK_NG];
    uint64_t empty[PK_NG];
    PackGroup grp[PK_NG];
};

// one gated pixel from the staged planes -> (cam xyz, w), obj
template <bool DENSE>
__device__ __forceinline__ void stage_s1(const PackStage& d, const RoiConst& rc, int p, bool weighted, int mask_mode,
                                         float4& camw, float4& objv) {
    float cam[3], obj[3];
    pixel_s1<DENSE>(rc, p, d.tile[0][p], d.tile[1][p], d.tile[2][p], d.tile[3][p], cam, obj);
    const float w = weighted ? mask_prob(d.tile[4][p], mask_mode, rc.mn, rc.mx) : 1.f;
    camw = make_float4(cam[0], cam[1], cam[2], w);
    objv = make_float4(obj[0], obj[1], obj[2], 0.f);
}

// S > 3 pairs per hypothesis (misc.py:72,91): same arithmetic as pose_solve.cu:hyp_from_sample, the sample's pixel
// indices packed four to a 64-bit register instead of a scratch array.
struct SampleIdx {
    unsigned long long w[4];
    __device__ __forceinline__ int get(int v) const {
        const unsigned long long x = (v >> 2) == 0 ? w[0] : ((v >> 2) == 1 ? w[1] : ((v >> 2) == 2 ? w[2] : w[3]));
        return (int)((x >> (16 * (v & 3))) & 0xFFFFull);
    }
    __device__ __forceinline__ void set(int v, int px) {
        const unsigned long long m = (unsigned long long)(unsigned)px << (16 * (v & 3));
        if ((v >> 2) == 0) w[0] |= m;
        else if ((v >> 2) == 1) w[1] |= m;
        else if ((v >> 2) == 2) w[2] |= m;
        else w[3] |= m;
    }
};
static_assert(RDPN_MAX_SAMPLE <= 16, "SampleIdx packs 16 indices");

template <bool DENSE>
__device__ __noinline__ bool hyp_from_sample_staged(const PackStage& d, const PackGroup& s, const float4* anchors,
                                                    const int32_t* idx_in, uint32_t kroi, int h, int S, float* P) {
    const RoiConst& rc = s.rc;
    const uint32_t nsel = s.selpfx[RDPN_P / 32];
    SampleIdx ii;
    ii.w[0] = ii.w[1] = ii.w[2] = ii.w[3] = 0ull;
    for (int v = 0; v < S; ++v) {
        int px;
        if (idx_in) {
            px = idx_in[v];
        } else {
            const uint32_t key = fmix32(kroi ^ (uint32_t)(S * h + v));
            px = nsel ? kth_gated_pixel(s.selmap, s.selpfx, (uint32_t)(((unsigned long long)key * nsel) >> 32)) : -1;
        }
        if ((unsigned)px >= RDPN_P) return false;
        if (!((s.selmap[px >> 5] >> (px & 31)) & 1u)) return false;
        for (int u = 0; u < v; ++u)
            if (ii.get(u) == px) return false;
        ii.set(v, px);
    }
    double m[17];  // sum c (3) | sum a (3) | sum c a^T (9) | sum |c|^2 | sum |a|^2, all about pair 0
#pragma unroll
    for (int i = 0; i < 17; ++i) m[i] = 0.0;
    float c0f[3] = {0.f, 0.f, 0.f}, a0f[3] = {0.f, 0.f, 0.f}, cpf[3] = {0.f, 0.f, 0.f}, apf[3] = {0.f, 0.f, 0.f};
    bool ok_a = false, ok_c = false;
#pragma unroll 1
    for (int v = 0; v < S; ++v) {
        float4 cw, ob;
        const int px = ii.get(v);
        stage_s1<DENSE>(d, rc, px, false, RDPN_MASK_RAW, cw, ob);
        if (!DENSE) ob = anchors[d.rid[px]];
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

template <bool DENSE, bool MULTI>
__global__ void __launch_bounds__(PK_NT, 1) gate_pack_kernel(SolveArgs a, unsigned char* __restrict__ ws, PkgLayout lay) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PackSmem& sm = *reinterpret_cast<PackSmem*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const rdpn_roi_inputs& in = a.in;
    const int B = in.B, G = gridDim.x;
    const int my_n = (B - (int)blockIdx.x + G - 1) / G;  // ROIs this CTA owns
    const int H = a.prm.num_hyp;
    const int R = DENSE ? 1 : in.num_regions;
    const int RB = R + 1;
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < PK_NG; ++i) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == PK_NG * PK_GW) {
        // ---------------- producer warp: whole ROIs into the ring ----------------
        if (lane == 0) {
            const uint32_t plane = RDPN_P * sizeof(float);
            for (int i = 0; i < my_n; ++i) {
                const int stg = i % PK_NG;
                if (i >= PK_NG) mbar_wait(&sm.empty[stg], ((i / PK_NG) - 1) & 1);
                const size_t o = (size_t)(blockIdx.x + (size_t)i * G) * RDPN_P;
                PackStage& d = sm.st[stg];
                mbar_expect_tx(&sm.full[stg], 5 * plane + (DENSE ? 0 : RDPN_P));
                bulk_g2s(d.tile[4], in.mask + o, plane, &sm.full[stg]);
                bulk_g2s(d.tile[0], in.depth + o, plane, &sm.full[stg]);
                bulk_g2s(d.tile[1], in.coor_x + o, plane, &sm.full[stg]);
                bulk_g2s(d.tile[2], in.coor_y + o, plane, &sm.full[stg]);
                bulk_g2s(d.tile[3], in.coor_z + o, plane, &sm.full[stg]);
                if (!DENSE) bulk_g2s(d.rid, in.region_idx + o, RDPN_P, &sm.full[stg]);
            }
        }
        return;
    }

    // ---------------- compute groups ----------------
    const int g = warp / PK_GW;            // group = ring stage
    const int gt = t - g * PK_GT;          // thread within the group
    const int gw = gt >> 5;                // warp within the group
    const int bar_id = 1 + g;
    PackGroup& s = sm.grp[g];
    unsigned char* dyn = smem_raw + sizeof(PackSmem);
    const size_t dyn_per_group = ((size_t)R * sizeof(float4) + (size_t)PK_GW * RB * sizeof(uint32_t) + 15) & ~(size_t)15;
    float4* anchors = reinterpret_cast<float4*>(dyn + g * dyn_per_group);
    uint32_t* wrun = reinterpret_cast<uint32_t*>(dyn + g * dyn_per_group + (size_t)R * sizeof(float4));
    const PackStage& d = sm.st[g];
    const int SS = MULTI ? a.prm.sample_size : 3;
    const bool sampling = a.hyp_idx == nullptr;

    for (int it = g; it < my_n; it += PK_NG) {
        const int b = blockIdx.x + it * G;
        unsigned char* pkg = ws + (size_t)b * lay.stride;
        // ---- per-ROI setup that does not need the planes (overlaps the TMA wait) ----
        int pre0 = -1, pre1 = -1, pre2 = -1;
        if (!sampling && gt < H && !MULTI) {
            const int32_t* ip = a.hyp_idx + ((size_t)b * H + gt) * 3;
            pre0 = __ldg(ip);
            pre1 = __ldg(ip + 1);
            pre2 = __ldg(ip + 2);
        }
        if (gt == 0) {
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
            rc.mn = rc.mx = 0.f;
        }
        if (!DENSE)
            for (int r = gt; r < R; r += PK_GT) {
                const float* ap = in.anchors + ((size_t)b * R + r) * 3;
                anchors[r] = make_float4(ap[0], ap[1], ap[2], 0.f);
            }
        for (int i = gt; i < PK_GW * RB; i += PK_GT) wrun[i] = 0;
        if (a.out.hyp_counts)
            for (int h = gt; h < H; h += PK_GT) a.out.hyp_counts[(size_t)b * H + h] = 0;

        mbar_wait(&sm.full[g], (it / PK_NG) & 1);
        // ---- 1: mask min / max (engine_utils.py:123-124) ----
        float4 mq[PK_QPT];
#pragma unroll
        for (int k = 0; k < PK_QPT; ++k) mq[k] = reinterpret_cast<const float4*>(d.tile[4])[32 * (PK_GW * k + gw) + lane];
        if (in.mask_mode == RDPN_MASK_L1) {
            float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
            for (int k = 0; k < PK_QPT; ++k) minmax4(mq[k], mn, mx);
            mn = warp_min(mn);
            mx = warp_max(mx);
            if (lane == 0) { s.red_f[0][gw] = mn; s.red_f[1][gw] = mx; }
        }
        named_bar(bar_id, PK_GT);
        RoiConst rc = s.rc;
        RoiGate gate;
        gate.hi = gate.lo = gate.b = 0.f;
        gate.cut = 0.0;
        gate.incl = 0;
        if (in.mask_mode == RDPN_MASK_L1) {
            float lo = s.red_f[0][0], hi = s.red_f[1][0];
#pragma unroll
            for (int w = 1; w < PK_GW; ++w) { lo = fminf(lo, s.red_f[0][w]); hi = fmaxf(hi, s.red_f[1][w]); }
            rc.mn = lo;
            rc.mx = hi;
            if (gt == 0) { s.rc.mn = lo; s.rc.mx = hi; }  // for the out-of-line S > 3 path (read after later barriers)
            make_gate(gate, lo, hi, in.mask_thr, a.mask_cut, a.mask_cut_incl);
        }

        // ---- 2: gate (gdrn_evaluator.py:110-117 + depth validity), mask first; sort pass A (bucket histogram) ----
        unsigned selbits = 0u;
        uint32_t* hrow = wrun + gw * RB;
#pragma unroll
        for (int k = 0; k < PK_QPT; ++k) {
            const int q = 32 * (PK_GW * k + gw) + lane;
            const float mm[4] = {mq[k].x, mq[k].y, mq[k].z, mq[k].w};
            unsigned nib = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) nib |= (mask_pass(mm[j], in.mask_mode, in.mask_thr, rc.mn, gate) ? 1u : 0u) << j;
            if (nib) {
                const float4 dq = reinterpret_cast<const float4*>(d.tile[0])[q];
                const float4 xq = reinterpret_cast<const float4*>(d.tile[1])[q];
                const float4 yq = reinterpret_cast<const float4*>(d.tile[2])[q];
                const float4 zq = reinterpret_cast<const float4*>(d.tile[3])[q];
                const uchar4 r4 = DENSE ? make_uchar4(0, 0, 0, 0) : reinterpret_cast<const uchar4*>(d.rid)[q];
                const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
                const float cxn[4] = {xq.x, xq.y, xq.z, xq.w};
                const float cyn[4] = {yq.x, yq.y, yq.z, yq.w};
                const float czn[4] = {zq.x, zq.y, zq.z, zq.w};
                const uint8_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
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
                    const bool sel = ((nib >> j) & 1u) && (fabsf(dx) > rc.gthr[0]) && (fabsf(dy) > rc.gthr[1]) &&
                                     (fabsf(dz) > rc.gthr[2]) && (dv > 0.f) && ((int)rr[j] < R);  // ids outside [0, R) never pass
                    if (!sel) nib &= ~(1u << j);
                    else atomicAdd(&hrow[rr[j]], 1u);
                }
            }
            selbits |= nib << (4 * k);
            unsigned wbits = nib << (4 * (lane & 7));  // gate bitmap: 4 bits per quad, 8 quads per word
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 4);
            if ((lane & 7) == 0) s.selmap[q >> 3] = wbits;
        }
        named_bar(bar_id, PK_GT);

        // ---- 3: bucket starts, per-warp cursors, run table (one thread per bucket + block scan) ----
        {
            const bool mine = gt < R;
            int c[PK_GW];
            int tot = 0;
#pragma unroll
            for (int w = 0; w < PK_GW; ++w) {
                c[w] = mine ? (int)wrun[w * RB + gt] : 0;
                tot += c[w];
            }
            const int packed = tot | ((tot > 0 ? 1 : 0) << 16);
            const int pc = (sampling && gt < RDPN_P / 32) ? __popc(s.selmap[gt]) : 0;
            int x = packed, xs = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                const int ys = __shfl_up_sync(0xffffffffu, xs, o);
                if (lane >= o) { x += y; xs += ys; }
            }
            if (lane == 31) { s.red_i[gw] = x; s.red_j[gw] = xs; }
            named_bar(bar_id, PK_GT);
            int base = 0, bases = 0;
#pragma unroll
            for (int w = 0; w < PK_GW; ++w)
                if (w < gw) { base += s.red_i[w]; bases += s.red_j[w]; }
            if (sampling && gt < RDPN_P / 32) {
                s.selpfx[gt] = (uint16_t)(bases + xs - pc);
                if (gt == RDPN_P / 32 - 1) s.selpfx[RDPN_P / 32] = (uint16_t)(bases + xs);
            }
            const int excl = base + x - packed;
            int run = excl & 0xFFFF;
            const int kk = excl >> 16;
            if (mine) {
                const int start = run;
#pragma unroll
                for (int w = 0; w < PK_GW; ++w) {
                    wrun[w * RB + gt] = (uint32_t)run;
                    run += c[w];
                }
                if (tot > 0) {  // one entry per NON-EMPTY bucket = (anchor xyz, start | end << 16)
                    float4 hd = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : anchors[gt];
                    hd.w = __uint_as_float((unsigned)start | ((unsigned)run << 16));
                    reinterpret_cast<float4*>(pkg + lay.runtab)[kk] = hd;
                }
                if (gt == R - 1) { s.n_sel = run; s.n_runs = kk + (tot > 0 ? 1 : 0); }
            }
        }
        named_bar(bar_id, PK_GT);
        // pass B: slots from warp match_any ranks, deterministic order (warp, k, j, lane)
        {
            uint32_t* myrun = wrun + gw * RB;
#pragma unroll 1
            for (int k = 0; k < PK_QPT; ++k) {
                const unsigned nib = (selbits >> (4 * k)) & 0xFu;
                if (__ballot_sync(0xffffffffu, nib != 0u) == 0u) continue;
                const int q = 32 * (PK_GW * k + gw) + lane;
                const uchar4 r4 = DENSE ? make_uchar4(0, 0, 0, 0) : reinterpret_cast<const uchar4*>(d.rid)[q];
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
                        myrun[key] = (uint32_t)(cur + __popc(m));
                    }
                    cur = __shfl_sync(0xffffffffu, cur, leader);
                    if (sel) {
                        const int sl = cur + __popc(m & ((1u << lane) - 1u));
                        s.pix[sl] = (uint16_t)(4 * q + j);
                        s.srid[sl] = (uint8_t)key;
                    }
                    __syncwarp();
                }
            }
        }
        named_bar(bar_id, PK_GT);
        const int n = s.n_sel;
        const bool enough = n >= a.prm.min_pts;

        // ---- 4: S1 for the gated slots, straight to the package (coalesced 16-byte stores) ----
        {
            float4* slots_g = reinterpret_cast<float4*>(pkg + lay.slots);
            float4* objs_g = slots_g + RDPN_P;  // dense only
            uint8_t* srid_g = pkg + lay.srid;
            uint16_t* pix_g = reinterpret_cast<uint16_t*>(pkg + lay.pix);
            for (int sl = gt; sl < n; sl += PK_GT) {
                const int p = s.pix[sl];
                float4 cw, ob;
                stage_s1<DENSE>(d, rc, p, a.prm.weighted != 0, in.mask_mode, cw, ob);
                slots_g[sl] = cw;
                if (DENSE) objs_g[sl] = ob;
                srid_g[sl] = s.srid[sl];
                pix_g[sl] = (uint16_t)p;
            }
        }

        // ---- 5: hypotheses (FP64 closed form, rounded once to FP32), compacted by validity ----
        int nvalid = 0;
        const uint32_t kroi = fmix32(fmix32(a.prm.seed ^ 0x9e3779b9u) ^ (uint32_t)(a.prm.roi_base + b));
        for (int h0 = 0; h0 < H; h0 += PK_GT) {
            const int h = h0 + gt;
            float P[12];
            bool ok = false;
            if (h < H) {
                if (MULTI) {
                    ok = hyp_from_sample_staged<DENSE>(d, s, anchors, sampling ? nullptr : a.hyp_idx + ((size_t)b * H + h) * SS,
                                                       kroi, h, SS, P);
                } else {
                    int ii[3];
                    if (!sampling) {
                        const int32_t* ip = a.hyp_idx + ((size_t)b * H + h) * 3;
                        const bool first = h0 == 0;
                        ii[0] = first ? pre0 : __ldg(ip);
                        ii[1] = first ? pre1 : __ldg(ip + 1);
                        ii[2] = first ? pre2 : __ldg(ip + 2);
                    } else {
                        const uint32_t nsel = s.selpfx[RDPN_P / 32];
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            const uint32_t key = fmix32(kroi ^ (uint32_t)(3 * h + v));
                            const uint32_t k = (uint32_t)(((unsigned long long)key * nsel) >> 32);
                            ii[v] = nsel ? kth_gated_pixel(s.selmap, s.selpfx, k) : -1;
                        }
                    }
                    ok = ((unsigned)ii[0] < RDPN_P) && ((unsigned)ii[1] < RDPN_P) && ((unsigned)ii[2] < RDPN_P);
                    if (ok) {
#pragma unroll
                        for (int v = 0; v < 3; ++v) ok = ok && ((s.selmap[ii[v] >> 5] >> (ii[v] & 31)) & 1u);
                    }
                    if (ok) {
                        float pf[3][3];
                        float4 cw[3], ob[3];
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            stage_s1<DENSE>(d, rc, ii[v], false, in.mask_mode, cw[v], ob[v]);
                            if (!DENSE) ob[v] = anchors[d.rid[ii[v]]];
                            pf[v][0] = ob[v].x; pf[v][1] = ob[v].y; pf[v][2] = ob[v].z;
                        }
                        double scr[8];
                        double ma2, u1a, u2a, v2a;
                        {
                            TriSide sa;
                            tri_side(pf, sa);
                            ok = sa.ok;
                            scr[0] = sa.e1[0]; scr[1] = sa.e1[1]; scr[2] = sa.e1[2];
                            scr[3] = sa.n[0];  scr[4] = sa.n[1];  scr[5] = sa.n[2];
                            scr[6] = sa.m[0];  scr[7] = sa.m[1];
                            ma2 = sa.m[2]; u1a = sa.u1; u2a = sa.u2; v2a = sa.v2;
                        }
#pragma unroll
                        for (int v = 0; v < 3; ++v) { pf[v][0] = cw[v].x; pf[v][1] = cw[v].y; pf[v][2] = cw[v].z; }
                        TriSide sc;
                        tri_side(pf, sc);
                        ok = ok && sc.ok;
                        if (ok) kabsch3_sides(scr, 1, ma2, u1a, u2a, v2a, sc, P);
                    }
                }
                if (a.out.hyp_poses) {
                    float4* hp = reinterpret_cast<float4*>(a.out.hyp_poses + ((size_t)b * H + h) * 12);
                    const bool wr = ok && enough;
                    hp[0] = wr ? make_float4(P[0], P[1], P[2], P[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    hp[1] = wr ? make_float4(P[4], P[5], P[6], P[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    hp[2] = wr ? make_float4(P[8], P[9], P[10], P[11]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s.red_i[gw] = __popc(bal);
            named_bar(bar_id, PK_GT);
            int base = nvalid, tot = 0;
#pragma unroll
            for (int w = 0; w < PK_GW; ++w) {
                const int c = s.red_i[w];
                if (w < gw) base += c;
                tot += c;
            }
            if (ok) {
                const int j = base + __popc(bal & ((1u << lane) - 1u));
                float4* hp = reinterpret_cast<float4*>(pkg + lay.hyp + (size_t)j * 48);
                hp[0] = make_float4(P[0], P[1], P[2], P[3]);
                hp[1] = make_float4(P[4], P[5], P[6], P[7]);
                hp[2] = make_float4(P[8], P[9], P[10], P[11]);
                reinterpret_cast<uint16_t*>(pkg + lay.vh)[j] = (uint16_t)h;
            }
            nvalid += tot;
            named_bar(bar_id, PK_GT);  // red_i is reused by the next round; after the last one: every read of the stage is done
        }
        if (gt == 0) {
            PkgHdr hd;
            hd.n = n;
            hd.nruns = s.n_runs;
            hd.nvalid = enough ? nvalid : 0;
            hd.flags = 0;
            hd.best_j = -1;
            hd.n_best = 0;
            hd.pad[0] = hd.pad[1] = 0;
            *reinterpret_cast<int4*>(pkg) = make_int4(hd.n, hd.nruns, hd.nvalid, hd.flags);
            *reinterpret_cast<int4*>(pkg + 16) = make_int4(-1, 0, 0, 0);
            if (a.out.n_sel) a.out.n_sel[b] = n;
            mbar_arrive1(&sm.empty[g]);  // hand the stage back to the producer
        }
    }
}

// =============================================================================================
// K2: hypotheses x points inlier scoring + best selection
// =============================================================================================
#ifndef RDPN_SCORE_CTAS
#define RDPN_SCORE_CTAS 2
#endif
constexpr int SC_T = 256;            // scoring threads
constexpr int SC_W = SC_T / 32;
constexpr int SC_NT = SC_T + 32;     // + producer warp
constexpr int SC_CHUNK = 1024;       // gated slots resident per stage (dense mode: half, the second half holds obj xyz)

struct ScoreStageHdr { int n, nruns, nvalid, skip; };
struct __align__(128) ScoreSmem {
    uint64_t full[2];
    uint64_t empty[2];
    uint64_t extra;          // multi-chunk ROIs: the consumer-driven reload of the slot buffer
    ScoreStageHdr hdr[2];
    unsigned long long red_k[SC_W];
    int jstop;
};
struct ScoreLayout {  // byte offsets of one stage behind ScoreSmem
    unsigned slots, hyp, runtab, vh, hcnt, total;
};
static ScoreLayout make_score_layout(int H, int R) {
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    ScoreLayout l;
    size_t off = 0;
    l.slots = (unsigned)off;  off = al(off + (size_t)SC_CHUNK * 16);
    l.hyp = (unsigned)off;    off = al(off + (size_t)H * 48);
    l.runtab = (unsigned)off; off = al(off + (size_t)R * 16);
    l.vh = (unsigned)off;     off = al(off + (size_t)H * 2 + 16);
    l.hcnt = (unsigned)off;   off = al(off + (size_t)H * 4);
    l.total = (unsigned)off;
    return l;
}

template <bool DENSE>
__global__ void __launch_bounds__(SC_NT, RDPN_SCORE_CTAS) score_kernel(SolveArgs a, unsigned char* __restrict__ ws, PkgLayout lay,
                                                                         ScoreLayout sl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScoreSmem& sm = *reinterpret_cast<ScoreSmem*>(smem_raw);
    unsigned char* stage_base = smem_raw + ((sizeof(ScoreSmem) + 127) & ~(size_t)127);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int B = a.in.B, G = gridDim.x;
    const int my_n = (B - (int)blockIdx.x + G - 1) / G;
    const int H = a.prm.num_hyp;
    constexpr int CH = DENSE ? SC_CHUNK / 2 : SC_CHUNK;
    if (t == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_init(&sm.empty[0], 1);
        mbar_init(&sm.empty[1], 1);
        mbar_init(&sm.extra, 1);
        mbar_fence_init();
    }
    for (int i = t; i < 2 * H; i += SC_NT) {  // hcnt of both stages starts at zero; every ROI re-zeroes what it used
        const int stg = i / H, j = i - stg * H;
        reinterpret_cast<int*>(stage_base + (size_t)stg * sl.total + sl.hcnt)[j] = 0;
    }
    __syncthreads();

    if (warp == SC_W) {
        // ---------------- producer warp: packages into the ring ----------------
        if (lane == 0) {
            for (int i = 0; i < my_n; ++i) {
                const int stg = i & 1;
                if (i >= 2) mbar_wait(&sm.empty[stg], ((i >> 1) - 1) & 1);
                const unsigned char* pkg = ws + (size_t)(blockIdx.x + (size_t)i * G) * lay.stride;
                const int4 h4 = *reinterpret_cast<const int4*>(pkg);
                const int n = h4.x, nruns = h4.y, nvalid = h4.z;
                const bool skip = n < a.prm.min_pts || n <= 0 || nvalid <= 0;
                sm.hdr[stg].n = n;
                sm.hdr[stg].nruns = nruns;
                sm.hdr[stg].nvalid = nvalid;
                sm.hdr[stg].skip = skip ? 1 : 0;
                if (skip) {
                    mbar_arrive1(&sm.full[stg]);
                    continue;
                }
                unsigned char* st = stage_base + (size_t)stg * sl.total;
                const int n0 = n < CH ? n : CH;
                const uint32_t b_slots = (uint32_t)n0 * 16u, b_hyp = (uint32_t)nvalid * 48u, b_run = (uint32_t)nruns * 16u;
                const uint32_t b_vh = ((uint32_t)nvalid * 2u + 15u) & ~15u;
                mbar_expect_tx(&sm.full[stg], b_slots * (DENSE ? 2u : 1u) + b_hyp + b_run + b_vh);
                bulk_g2s(st + sl.hyp, pkg + lay.hyp, b_hyp, &sm.full[stg]);
                bulk_g2s(st + sl.runtab, pkg + lay.runtab, b_run, &sm.full[stg]);
                bulk_g2s(st + sl.vh, pkg + lay.vh, b_vh, &sm.full[stg]);
                bulk_g2s(st + sl.slots, pkg + lay.slots, b_slots, &sm.full[stg]);
                if (DENSE) bulk_g2s(st + sl.slots + (size_t)CH * 16, pkg + lay.slots + (size_t)RDPN_P * 16, b_slots, &sm.full[stg]);
            }
        }
        return;
    }

    // ---------------- scoring warps ----------------
    const float cut = a.sq_cut;
    unsigned extra_phase = 0;
    for (int i = 0; i < my_n; ++i) {
        const int stg = i & 1;
        const int b = blockIdx.x + i * G;
        unsigned char* pkg = ws + (size_t)b * lay.stride;
        unsigned char* st = stage_base + (size_t)stg * sl.total;
        const float4* camw_s = reinterpret_cast<const float4*>(st + sl.slots);
        const float4* obj_s = camw_s + CH;  // dense only
        const float* hyp = reinterpret_cast<const float*>(st + sl.hyp);
        const float4* runtab = reinterpret_cast<const float4*>(st + sl.runtab);
        const uint16_t* vh = reinterpret_cast<const uint16_t*>(st + sl.vh);
        int* hcnt = reinterpret_cast<int*>(st + sl.hcnt);
        if (t == 0) sm.jstop = 0x7FFFFFFF;
        mbar_wait(&sm.full[stg], (i >> 1) & 1);
        const ScoreStageHdr hd = sm.hdr[stg];
        if (!hd.skip) {
            const int n = hd.n, nruns = hd.nruns, nvalid = hd.nvalid;
            for (int c0 = 0; c0 < n; c0 += CH) {
                const int c1 = min(n, c0 + CH);
                if (c0 > 0) {  // rare: more gated slots than one stage holds -- reload the slot buffer in place
                    named_bar(1, SC_T);
                    if (t == 0) {
                        const uint32_t bytes = (uint32_t)(c1 - c0) * 16u;
                        mbar_expect_tx(&sm.extra, bytes * (DENSE ? 2u : 1u));
                        bulk_g2s(st + sl.slots, pkg + lay.slots + (size_t)c0 * 16, bytes, &sm.extra);
                        if (DENSE)
                            bulk_g2s(st + sl.slots + (size_t)CH * 16, pkg + lay.slots + (size_t)(RDPN_P + c0) * 16, bytes, &sm.extra);
                    }
                    mbar_wait(&sm.extra, extra_phase & 1);
                    ++extra_phase;
                }
                const bool whole = (c0 == 0 && c1 == n);
                if (!DENSE) {
                    // TWO hypotheses per thread (every staged point and every run header is shared by both); W warps
                    // cover all pairs once, the SC_W / W groups of W warps split the runs (warp-aligned segments)
                    const int npairs = (nvalid + 1) >> 1;
                    const int W = max(1, (npairs + 31) >> 5);
                    const bool wide = W >= SC_W;  // H > 512: every thread loops over several pairs, no run split
                    const int S2 = wide ? 1 : SC_W / W;
                    const int seg = wide ? 0 : warp / W;
                    const int j0 = wide ? t : (warp % W) * 32 + lane;
                    for (int j = j0; j < npairs && seg < S2; j += wide ? SC_T : npairs) {
                        const int hA = 2 * j;
                        const bool hasB = 2 * j + 1 < nvalid;
                        const int hB = hasB ? hA + 1 : hA;
                        float PA[12], PB[12];
                        {
                            const float4* pa = reinterpret_cast<const float4*>(hyp + (size_t)hA * 12);
                            const float4* pb = reinterpret_cast<const float4*>(hyp + (size_t)hB * 12);
#pragma unroll
                            for (int r = 0; r < 3; ++r) {
                                const float4 va = pa[r], vb = pb[r];
                                PA[4 * r] = va.x; PA[4 * r + 1] = va.y; PA[4 * r + 2] = va.z; PA[4 * r + 3] = va.w;
                                PB[4 * r] = vb.x; PB[4 * r + 1] = vb.y; PB[4 * r + 2] = vb.z; PB[4 * r + 3] = vb.w;
                            }
                        }
                        int cA = 0, cB = 0;
                        for (int k = seg; k < nruns; k += S2) {
                            const float4 rh = runtab[k];
                            const unsigned se = __float_as_uint(rh.w);
                            int p = (int)(se & 0xFFFFu);
                            int e = (int)(se >> 16);
                            if (!whole) {
                                p = max(p, c0) - c0;
                                e = min(e, c1) - c0;
                                if (p >= e) continue;
                            }
                            float ax, ay, az, bx, by, bz;
                            xform(PA, rh.x, rh.y, rh.z, ax, ay, az);
                            xform(PB, rh.x, rh.y, rh.z, bx, by, bz);
#pragma unroll 1
                            for (; p + 4 <= e; p += 4) {
                                const float4 q0 = camw_s[p], q1 = camw_s[p + 1], q2 = camw_s[p + 2], q3 = camw_s[p + 3];
                                count_if_lt(cA, resid2_pt(ax, ay, az, q0.x, q0.y, q0.z), cut);
                                count_if_lt(cB, resid2_pt(bx, by, bz, q0.x, q0.y, q0.z), cut);
                                count_if_lt(cA, resid2_pt(ax, ay, az, q1.x, q1.y, q1.z), cut);
                                count_if_lt(cB, resid2_pt(bx, by, bz, q1.x, q1.y, q1.z), cut);
                                count_if_lt(cA, resid2_pt(ax, ay, az, q2.x, q2.y, q2.z), cut);
                                count_if_lt(cB, resid2_pt(bx, by, bz, q2.x, q2.y, q2.z), cut);
                                count_if_lt(cA, resid2_pt(ax, ay, az, q3.x, q3.y, q3.z), cut);
                                count_if_lt(cB, resid2_pt(bx, by, bz, q3.x, q3.y, q3.z), cut);
                            }
#pragma unroll 1
                            for (; p < e; ++p) {
                                const float4 q0 = camw_s[p];
                                count_if_lt(cA, resid2_pt(ax, ay, az, q0.x, q0.y, q0.z), cut);
                                count_if_lt(cB, resid2_pt(bx, by, bz, q0.x, q0.y, q0.z), cut);
                            }
                        }
                        if (S2 == 1 && whole) {
                            hcnt[hA] = cA;
                            if (hasB) hcnt[hB] = cB;
                        } else {
                            atomicAdd(&hcnt[hA], cA);
                            if (hasB) atomicAdd(&hcnt[hB], cB);
                        }
                    }
                } else {
                    // dense mode: every pair pays the full transform; S segments of the slot range per hypothesis
                    const int S = (nvalid >= SC_T) ? 1 : (SC_T / nvalid);
                    const int m = c1 - c0;
                    for (int item = t; item < nvalid * S; item += SC_T) {
                        const int h = item % nvalid, seg = item / nvalid;
                        float P[12];
#pragma unroll
                        for (int r = 0; r < 12; ++r) P[r] = hyp[(size_t)h * 12 + r];
                        int c = 0;
                        const int i0 = (int)(((long long)m * seg) / S), i1 = (int)(((long long)m * (seg + 1)) / S);
#pragma unroll 4
                        for (int p = i0; p < i1; ++p) {
                            const float4 cp = camw_s[p];
                            const float4 ap = obj_s[p];
                            count_if_lt(c, resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z), cut);
                        }
                        atomicAdd(&hcnt[h], c);
                    }
                }
            }
            named_bar(1, SC_T);  // counts complete
            // ---- best hypothesis (misc.py:121) with optional adaptive stop (misc.py:134-138) ----
            if (a.prm.adaptive) {
                // i_ransac of compacted entry j is j + 1; stop after the first j with i_ransac > max(k, min_iter)
                const double lc = log10(1.0 - (double)a.prm.confidence);
                for (int j = t; j < nvalid; j += SC_T)
                    if (adaptive_stop(hcnt[j], n, j + 1, lc, a.prm.min_iter)) atomicMin(&sm.jstop, j);
                named_bar(1, SC_T);
            }
            const int jlim = min(nvalid, sm.jstop == 0x7FFFFFFF ? nvalid : sm.jstop + 1);
            unsigned long long key = 0ull;
            for (int j = t; j < jlim; j += SC_T) {
                const int c = hcnt[j];
                if (c >= a.prm.min_inliers && c > 0) {
                    const unsigned long long k = ((unsigned long long)(unsigned)c << 32) | (unsigned)(0x7FFFFFFF - j);
                    key = k > key ? k : key;
                }
            }
            key = warp_max_u64(key);
            if (lane == 0) sm.red_k[warp] = key;
            if (a.out.hyp_counts)
                for (int j = t; j < nvalid; j += SC_T) a.out.hyp_counts[(size_t)b * H + vh[j]] = hcnt[j];
            named_bar(1, SC_T);
            for (int j = t; j < nvalid; j += SC_T) hcnt[j] = 0;  // ready for the stage's next ROI
            if (t == 0) {
                unsigned long long kb = 0ull;
#pragma unroll
                for (int w = 0; w < SC_W; ++w) kb = sm.red_k[w] > kb ? sm.red_k[w] : kb;
                const int bj = kb ? 0x7FFFFFFF - (int)(kb & 0xFFFFFFFFull) : -1;
                *reinterpret_cast<int2*>(pkg + 16) = make_int2(bj, kb ? (int)(kb >> 32) : 0);
            }
        }
        named_bar(1, SC_T);  // every read of the stage (and of red_k / jstop) is done
        if (t == 0) mbar_arrive1(&sm.empty[stg]);
    }
}

// =============================================================================================
// K3: refit on the winner's inliers + every output, one warp per ROI
// =============================================================================================
constexpr int RF_W = 4;       // warps (ROIs) per CTA
#ifndef RDPN_REFIT_CTAS
#define RDPN_REFIT_CTAS 4       // CTAs per SM the register budget is sized for (128 registers per thread)
#endif
constexpr int RF_U = 4;       // slots per lane in flight

template <bool DENSE>
__global__ void __launch_bounds__(RF_W * 32, RDPN_REFIT_CTAS) refit_kernel(SolveArgs a, const unsigned char* __restrict__ ws, PkgLayout lay) {
    extern __shared__ __align__(16) unsigned char rf_smem[];  // float[RF_W][3 R]: the ROI's anchors, one row per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RF_W + warp;
    if (b >= a.in.B) return;
    const unsigned char* pkg = ws + (size_t)b * lay.stride;
    const int4 h0 = *reinterpret_cast<const int4*>(pkg);
    const int2 h1 = *reinterpret_cast<const int2*>(pkg + 16);
    const int n = h0.x, best_j = h1.x, nbest = h1.y;
    const bool enough = n >= a.prm.min_pts;
    if (a.out.inlier_mask) {  // zero-fill; inliers are scattered in after the last refit
        uint4* im = reinterpret_cast<uint4*>(a.out.inlier_mask + (size_t)b * RDPN_P);
        for (int i = lane; i < RDPN_P / 16; i += 32) im[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (!enough || best_j < 0) {
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
    const int R3 = DENSE ? 0 : 3 * a.in.num_regions;
    float* anc = reinterpret_cast<float*>(rf_smem) + (size_t)warp * R3;
    if (!DENSE) {
        const float* ag = a.in.anchors + (size_t)b * R3;
        for (int i = lane; i < R3; i += 32) anc[i] = __ldg(ag + i);
        __syncwarp();
    }
    const int best = (int)reinterpret_cast<const uint16_t*>(pkg + lay.vh)[best_j];
    float P[12];
    {
        const float4* hp = reinterpret_cast<const float4*>(pkg + lay.hyp + (size_t)best_j * 48);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float4 v = hp[r];
            P[4 * r] = v.x; P[4 * r + 1] = v.y; P[4 * r + 2] = v.z; P[4 * r + 3] = v.w;
        }
    }
    const float4* slots = reinterpret_cast<const float4*>(pkg + lay.slots);
    const float4* objs = slots + RDPN_P;  // dense only
    const uint8_t* srid = pkg + lay.srid;
    const uint16_t* pixs = reinterpret_cast<const uint16_t*>(pkg + lay.pix);
    const float cut = a.sq_cut;
    float out_scale = 1.f;
    const int iters = a.prm.refit_iters < 1 ? 1 : a.prm.refit_iters;
    // pivot of the raw moments (exact FP32 differences): slot 0
    float4 cp0 = slots[0], ap0;
    if (DENSE) {
        ap0 = objs[0];
    } else {
        const int r0 = 3 * (int)srid[0];
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
            float4 cpv[RF_U], apv[RF_U];
            int ridv[RF_U];
#pragma unroll
            for (int u = 0; u < RF_U; ++u) {  // every load of the batch in flight before the first use
                const int i = i0 + 32 * u;
                if (i < n) {
                    cpv[u] = slots[i];
                    if (DENSE) apv[u] = objs[i];
                    else ridv[u] = (int)srid[i];
                }
            }
#pragma unroll
            for (int u = 0; u < RF_U; ++u) {
                const int i = i0 + 32 * u;
                if (i >= n) break;
                const float4 cp = cpv[u];
                float4 ap;
                if (DENSE) ap = apv[u];
                else ap = make_float4(anc[3 * ridv[u]], anc[3 * ridv[u] + 1], anc[3 * ridv[u] + 2], 0.f);
                if (resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z) < cut) {
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
                bool in;
                if (k < 32) {
                    in = (marks >> k) & 1u;
                } else {
                    const float4 cp = slots[i];
                    float4 ap;
                    if (DENSE) ap = objs[i];
                    else ap = make_float4(anc[3 * (int)srid[i]], anc[3 * (int)srid[i] + 1], anc[3 * (int)srid[i] + 2], 0.f);
                    in = resid2(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z) < cut;
                }
                if (in) a.out.inlier_mask[(size_t)b * RDPN_P + pixs[i]] = 1;
            }
        }
        // every lane solves the same 3 x 3 problem (no divergence, no broadcast)
        const double isw = 1.0 / m[0];
        double S[9], R[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) S[3 * r + c] = m[7 + 3 * r + c] - m[1 + r] * (m[4 + c] * isw);  // sum w c a^T - (sum w c)(mean a)^T
        const double ga = m[17] - (m[4] * (m[4] * isw) + m[5] * (m[5] * isw) + m[6] * (m[6] * isw));
        const double gb = m[16] - (m[1] * (m[1] * isw) + m[2] * (m[2] * isw) + m[3] * (m[3] * isw));
        rotation_from_cov(S, ga, gb, R);
        const double sc = a.prm.with_scale ? sqrt(gb / ga) : 1.0;  // transform.py:971-975
        const double ma0 = m[4] * isw + (double)ap0.x, ma1 = m[5] * isw + (double)ap0.y, ma2 = m[6] * isw + (double)ap0.z;
        const double c0v[3] = {(double)cp0.x, (double)cp0.y, (double)cp0.z};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double mcr = m[1 + r] * isw + c0v[r];
            const double r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
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

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
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

enum { SLOT_PACK = 8, SLOT_SCORE = 12, SLOT_REFIT = 14 };  // attribute-cache slots (+ dense, + 2 * multi)

template <bool DENSE, bool MULTI>
static int launch_chunk(const SolveArgs& a, unsigned char* ws, const PkgLayout& lay, int sms, cudaStream_t st) {
    const int B = a.in.B, H = a.prm.num_hyp, R = DENSE ? 1 : a.in.num_regions;
    // K1
    {
        const size_t dyn_per_group = ((size_t)R * sizeof(float4) + (size_t)PK_GW * (R + 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
        const size_t smem = sizeof(PackSmem) + PK_NG * dyn_per_group;
        if (smem > 227 * 1024) return RDPN_E_TOOLARGE;
        int rc = ensure_func_smem((const void*)gate_pack_kernel<DENSE, MULTI>, SLOT_PACK + (DENSE ? 1 : 0) + (MULTI ? 2 : 0), smem);
        if (rc) return rc;
        const int grid = B < sms ? B : sms;
        gate_pack_kernel<DENSE, MULTI><<<grid, PK_NT, smem, st>>>(a, ws, lay);
        ++g_launch_count;
        RDPN_LAUNCH_CHECK();
    }
    // K2
    {
        const ScoreLayout sl = make_score_layout(H, R);
        const size_t smem = ((sizeof(ScoreSmem) + 127) & ~(size_t)127) + 2 * (size_t)sl.total;
        if (smem > 227 * 1024) return RDPN_E_TOOLARGE;
        int rc = ensure_func_smem((const void*)score_kernel<DENSE>, SLOT_SCORE + (DENSE ? 1 : 0), smem);
        if (rc) return rc;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > RDPN_SCORE_CTAS) per_sm = RDPN_SCORE_CTAS;
        if (per_sm < 1) per_sm = 1;
        const int cap = sms * per_sm;
        const int grid = B < cap ? B : cap;
        score_kernel<DENSE><<<grid, SC_NT, smem, st>>>(a, ws, lay, sl);
        ++g_launch_count;
        RDPN_LAUNCH_CHECK();
    }
    // K3
    {
        const size_t smem = DENSE ? 0 : (size_t)RF_W * 3 * R * sizeof(float);
        refit_kernel<DENSE><<<(B + RF_W - 1) / RF_W, RF_W * 32, smem, st>>>(a, ws, lay);
        ++g_launch_count;
        RDPN_LAUNCH_CHECK();
    }
    return 0;
}

bool split_supported(const SolveArgs& a, bool dense) {
    if (dense) return false;  // dense mode stays on the fused kernel for now
    if (a.in.num_regions > PK_MAX_R) return false;
    if (a.prm.num_hyp > 2048) return false;  // 2 stages x H x 52 B of shared memory
    return true;
}

size_t split_pkg_stride(int H, int R, bool dense) { return (size_t)make_layout(H, R, dense).stride; }

int launch_split(const SolveArgs& a, bool dense, void* ws, size_t ws_bytes, int chunk_rois, cudaStream_t st) {
    const int H = a.prm.num_hyp, R = dense ? 1 : a.in.num_regions;
    const PkgLayout lay = make_layout(H, R, dense);
    if (!ws || ((uintptr_t)ws & 127)) return RDPN_E_WORKSPACE;
    size_t cap = ws_bytes / lay.stride;
    if (cap < 1) return RDPN_E_WORKSPACE;
    if (chunk_rois > 0 && cap > (size_t)chunk_rois) cap = (size_t)chunk_rois;
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    const bool multi = a.prm.sample_size > 3;
    for (int b0 = 0; b0 < a.in.B; b0 += (int)cap) {
        const int nb = (a.in.B - b0) < (int)cap ? (a.in.B - b0) : (int)cap;
        const SolveArgs c = shifted(a, b0, nb);
        if (dense) rc = multi ? launch_chunk<true, true>(c, (unsigned char*)ws, lay, sms, st) : launch_chunk<true, false>(c, (unsigned char*)ws, lay, sms, st);
        else rc = multi ? launch_chunk<false, true>(c, (unsigned char*)ws, lay, sms, st) : launch_chunk<false, false>(c, (unsigned char*)ws, lay, sms, st);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace rdpn
