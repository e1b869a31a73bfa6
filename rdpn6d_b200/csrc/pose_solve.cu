// Dense-correspondence -> pose on sm_100a: the fused per-ROI solver (S1 gate + hypothesis generation +
// H x n inlier scoring + best selection + weighted Kabsch/Umeyama refit) and the batched Kabsch entry.
// (The materialising S1 kernel lives in correspond.cu.)
//
// One CTA owns one ROI and four CTAs share an SM.  The ROI's planes are read once with coalesced
// 16-byte loads for the gate (after an L2 prefetch); only the gated ~10 % of the pixels are gathered
// again and turned into correspondences, which then live in shared memory.  Nothing but the 12-float
// pose (and optional diagnostics) goes back to HBM.  (The bulk-TMA ring lives in correspond.cu, where
// the whole ROI has to be materialised.)
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
#include "gate.cuh"
#include "kabsch_math.cuh"
#include "solve_common.cuh"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace rdpn {
extern unsigned long long g_launch_count;
int ensure_func_smem(const void* func, int slot, size_t bytes);                                  // host_api.cu
int cached_workspace(cudaStream_t st, size_t need, void** out, size_t* out_bytes);               // host_api.cu
bool split_supported(const SolveArgs& a, bool dense);                                            // solve_split.cu
size_t split_pkg_stride(int H, int R, bool dense);                                               // solve_pipe.cu
size_t split_workspace_bytes(int B, int H, int R, int chunk_rois);                               // solve_pipe.cu
void set_stage_timing(float* ms3);                                                               // solve_pipe.cu
int launch_split(const SolveArgs& a, bool dense, void* ws, size_t ws_bytes, int chunk_rois, cudaStream_t st);

// Optional per-phase cycle stamps (tuning builds only: RDPN_NVCC_EXTRA=-DRDPN_PHASE_CLOCKS, benchmarks/phase_clocks.py)
#ifdef RDPN_PHASE_CLOCKS
__device__ long long* g_phase_clk = nullptr;  // [B][16]
#define PHASE_MARK(i)                                                                    \
    do {                                                                                 \
        if (threadIdx.x == 0 && g_phase_clk) {                                           \
            g_phase_clk[(size_t)blockIdx.x * 16 + (i)] = clock64();                      \
            if ((i) == 0) {                                                              \
                unsigned smid_;                                                          \
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                       \
                g_phase_clk[(size_t)blockIdx.x * 16 + 15] = smid_;                       \
            }                                                                            \
        }                                                                                \
    } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

#ifndef RDPN_SOLVE_THREADS
#define RDPN_SOLVE_THREADS 256
#endif
constexpr int ST = RDPN_SOLVE_THREADS;  // threads per CTA (256; 128 works too: benchmarks show no gain)
constexpr int SW = ST / 32;          // warps
constexpr int QPT = RDPN_P / 4 / ST;  // pixel quads per thread (4)

// ---------------------------------------------------------------------------------------------
// fused solver
// ---------------------------------------------------------------------------------------------
// One CTA (256 threads) per ROI, FOUR CTAs resident per SM (<= 64 registers, ~48 KB shared memory):
// every phase below is short and latency-bound on its own (global gathers, IEEE divisions, FP64
// chains, barriers), so the SM is kept busy by thread-level parallelism across ROIs rather than by
// staging whole ROIs in shared memory (the first version did: 112 KB and 126 registers per CTA capped
// the SM at 16 warps and 57 % issue utilisation).
//   1  L2 prefetch of this ROI's planes; the thread's mask quads into registers; mask min/max (one barrier,
//      every thread folds the eight partials); the thread's first hypothesis triplet is requested now
//   2  GATE, mask first (no divisions): the mask test runs on all 16 pixels of the thread, depth / coor / region
//      ids are loaded only for quads with a passing pixel; sort pass A (per-warp bucket histogram, shared-memory
//      atomics) rides along
//   3  deterministic counting sort of the gated pixels by region id: bucket cursors by one thread per bucket +
//      block scan, slots from warp match_any ranks -> pix[slot], srid[slot]; run table of the non-empty buckets
//   4  hypothesis generation: 3 gathered pixels each, FP64 closed form, rounded once to FP32 (S > 3 pairs per
//      hypothesis: hyp_from_sample, FP64 moments + closed-form rotation)
//   5  per chunk of <= 1024 gated slots: STAGING (one thread per slot gathers the raw pixel and computes
//      (cam xyz, w) with the exact S1 arithmetic into a float4 list in shared memory), then
//   6  SCORING: two hypotheses per thread, warp-aligned groups of warps split the runs; per run the transformed
//      anchor R a + t is computed once (9 FMA per hypothesis) and every point costs 1/2 LDS.128 + 3 FADD + FMUL +
//      2 FFMA + FSETP + predicated IADD per hypothesis
//   7  best hypothesis, FP64 refit moments (two sweeps of nine), closed-form rotation, outputs
#ifndef RDPN_CHUNK_SLOTS
#define RDPN_CHUNK_SLOTS 1024
#endif
constexpr int CHUNK = RDPN_CHUNK_SLOTS;  // gated slots staged in shared memory at a time (dense mode: CHUNK / 2)
static_assert(CHUNK >= 4 * ST, "hypothesis generation parks 8 doubles per thread in the staging buffer");

struct FinishSmem {  // scratch of the select + refit tail
    double red_d[SW][18];
    double bc_d[20];
    int best_h;
    int n_best;
    int h_eff;
    float pose[12];
    unsigned long long red_k[SW];
    int red_i[SW];
    int red_j[SW];
};

struct __align__(128) FusedSmem {
    float4 chunk[CHUNK];            // (cam xyz, w) by slot of the current chunk; dense: second half = obj xyz
    uint16_t pix[RDPN_P];           // slot -> pixel
    uint8_t srid[RDPN_P];           // slot -> region id (non-decreasing)
    uint32_t selmap[RDPN_P / 32];   // gate bitmap by pixel
    uint16_t selpfx[RDPN_P / 32 + 2];  // internal sampling: gated pixels before word w (raster order), [128] = all
    RoiConst rc;
    float red_f[2][SW];
    int n_sel;
    int n_runs;
    FinishSmem fin;
};

// block-wide sum of NV doubles; result valid in every thread (via f.bc_d[0..NV)).
template <int NV>
__device__ __forceinline__ void block_sum(FinishSmem& f, double (&v)[NV]) {
    static_assert(NV <= 18, "red_d / bc_d hold 18 values");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) f.red_d[warp][i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double a = 0.0;
        for (int w = 0; w < SW; ++w) a += f.red_d[w][threadIdx.x];
        f.bc_d[threadIdx.x] = a;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = f.bc_d[i];
    __syncthreads();
}

struct FusedLayout {  // byte offsets of the dynamic tail behind FusedSmem
    int anchors;      // float4[R]
    int runtab;       // float4[R]: non-empty buckets (anchor xyz, start | end << 16)
    int wrun;         // uint32[SW][RB]     per-warp bucket counts, then cursors (RB = R + 1)
    int hyp;          // float[H][12]
    int hcnt;         // int[H]
    int vlist;        // uint16[H] indices of the valid hypotheses
    int planes;       // RoiPlanes copy for the out-of-line S > 3 hypothesis path (MULTI instantiations only)
    int total;
};

// Hypothesis from S > 3 pairs (misc.py:72,91 samples random_sample_num = 10): Kabsch of the S pairs,
// transform.py:913-980 semantics.  FP64 raw moments about the first pair (exact FP32 differences), closed-form
// rotation, rounded once to FP32.  Valid iff the S pixels are gated and pairwise distinct (sampling without
// replacement) and each side has a non-degenerate triangle (p0, p_{v-1}, p_v) (the S = 3 test, along the sample).
// Out of line: the default S = 3 path must not pay registers for it.
template <bool DENSE>
__device__ __noinline__ bool hyp_from_sample(FusedSmem& s, const RoiPlanes& pl, const float4* anchors,
                                             const int32_t* idx_in, uint32_t kroi, int h, int S, float* P) {
    const RoiConst& rc = s.rc;  // pl and rc are the shared-memory copies (no stack traffic for by-reference arguments)
    const uint32_t* selmap = s.selmap;
    const uint16_t* selpfx = s.selpfx;
    // the sample's pixel indices live in the staging buffer (idle until the barrier that ends hypothesis generation):
    // column threadIdx.x of a [RDPN_MAX_SAMPLE][ST] int array, conflict-free.  A per-thread array would sit in local
    // memory, and with four CTAs per SM local memory misses L1 nine times out of ten.
    static_assert(sizeof(float4) * CHUNK >= sizeof(int) * RDPN_MAX_SAMPLE * ST, "staging buffer holds the sample indices");
    int* ii = reinterpret_cast<int*>(s.chunk) + threadIdx.x;
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
                for (int u = 0; u < v; ++u) dup = dup || ii[u * ST] == px;
                if (!dup) break;
            }
        }
        if ((unsigned)px >= RDPN_P) return false;
        if (!((selmap[px >> 5] >> (px & 31)) & 1u)) return false;
        for (int u = 0; u < v; ++u)
            if (ii[u * ST] == px) return false;
        ii[v * ST] = px;
    }
    // One pass over the sample, nothing parked in local memory: pair 0 is the pivot of the FP64 moments (exact FP32
    // differences) and the apex of the degeneracy test -- each side needs one non-degenerate triangle
    // (p0, p_{v-1}, p_v), 2 <= v < S (oracle hypothesis_poses; bit-identical test).
    double m[17];  // sum c (3) | sum a (3) | sum c a^T (9) | sum |c|^2 | sum |a|^2, all about pair 0
#pragma unroll
    for (int i = 0; i < 17; ++i) m[i] = 0.0;
    float c0f[3] = {0.f, 0.f, 0.f}, a0f[3] = {0.f, 0.f, 0.f}, cpf[3] = {0.f, 0.f, 0.f}, apf[3] = {0.f, 0.f, 0.f};
    bool ok_a = false, ok_c = false;
#pragma unroll 1
    for (int v = 0; v < S; ++v) {
        float4 cw, ob;
        const int px = ii[v * ST];
        gather_s1<DENSE>(pl, rc, px, false, RDPN_MASK_RAW, cw, ob);  // unweighted: the mask is not read
        if (!DENSE) ob = anchors[__ldg(pl.rid + px)];
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

#ifndef RDPN_SOLVE_CTAS
#define RDPN_SOLVE_CTAS 4
#endif
#ifndef RDPN_SCORE_PAIRS
#define RDPN_SCORE_PAIRS 1
#endif
// the warp that runs the single-warp sections (bucket cursors, refit solve)
#ifndef RDPN_SERIAL_WARP
#define RDPN_SERIAL_WARP (SW - 1)
#endif
// MULTI: S > 3 pairs per hypothesis (a separate instantiation, so that the default S = 3 kernel carries neither the
// call nor its register pressure)
#ifndef RDPN_MULTI_CTAS
#define RDPN_MULTI_CTAS 3
#endif
template <bool DENSE, bool MULTI = false>
__global__ void __launch_bounds__(ST, MULTI ? RDPN_MULTI_CTAS : RDPN_SOLVE_CTAS) pose_solve_kernel(SolveArgs a, FusedLayout lay) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem& s = *reinterpret_cast<FusedSmem*>(smem_raw);
    FinishSmem& f = s.fin;
    float4* anchors = reinterpret_cast<float4*>(smem_raw + lay.anchors);
    float4* runtab = reinterpret_cast<float4*>(smem_raw + lay.runtab);
    uint32_t* wrun = reinterpret_cast<uint32_t*>(smem_raw + lay.wrun);
    float* hyp = reinterpret_cast<float*>(smem_raw + lay.hyp);  // [H][12]
    int* hcnt = reinterpret_cast<int*>(smem_raw + lay.hcnt);    // [H] counts, -1 = invalid
    uint16_t* vlist = reinterpret_cast<uint16_t*>(smem_raw + lay.vlist);
    const int H = a.prm.num_hyp;
    const int R = DENSE ? 1 : a.in.num_regions;
    const int RB = R + 1;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const rdpn_roi_inputs& in = a.in;
    const size_t po = (size_t)b * RDPN_P;
    RoiPlanes pl;
    pl.depth = in.depth + po;
    pl.cx = in.coor_x + po;
    pl.cy = in.coor_y + po;
    pl.cz = in.coor_z + po;
    pl.mask = in.mask + po;
    pl.rid = DENSE ? nullptr : in.region_idx + po;

    PHASE_MARK(0);
    // ---- 1: prefetch this ROI's planes into L2 (one 64-byte line per thread and plane), constants, zeroing
    {
#pragma unroll
        for (int o = t * 16; o < RDPN_P; o += ST * 16) {  // one 64-byte line per request
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.mask + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.depth + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.cx + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.cy + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.cz + o));
        }
        if (!DENSE && t < 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.rid + t * 64));
    }
    // the thread's first hypothesis triplet is requested now (DRAM) and consumed in phase 4
    const int SS = MULTI ? a.prm.sample_size : 3;  // pairs per hypothesis (3 .. RDPN_MAX_SAMPLE)
    int pre0 = -1, pre1 = -1, pre2 = -1;
    if (a.hyp_idx && t < H && !MULTI) {
        const int32_t* ip = a.hyp_idx + ((size_t)b * H + t) * 3;
        pre0 = __ldg(ip);
        pre1 = __ldg(ip + 1);
        pre2 = __ldg(ip + 2);
    }
    if (t == 0) {
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
        if (MULTI) *reinterpret_cast<RoiPlanes*>(smem_raw + lay.planes) = pl;
        f.best_h = -1;
        f.n_best = 0;
        f.h_eff = H;
    }
    if (!DENSE)
        for (int r = t; r < R; r += ST) {
            const float* ap = in.anchors + ((size_t)b * R + r) * 3;
            anchors[r] = make_float4(ap[0], ap[1], ap[2], 0.f);
        }
    for (int i = t; i < SW * RB; i += ST) wrun[i] = 0;
    if (a.out.inlier_mask) {  // zero-fill; inliers are scattered in after the last refit
        uint4* im = reinterpret_cast<uint4*>(a.out.inlier_mask + (size_t)b * RDPN_P);
        for (int i = t; i < RDPN_P / 16; i += ST) im[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // mask min / max (engine_utils.py:123-124).  Thread owns quads q = 32*(SW*k + warp) + lane (k = 0..3):
    // every warp gets two image rows out of each 16, so the rows the object covers are spread over all warps.
    float4 mq[QPT];  // the thread's mask quads stay in registers for the gate
#pragma unroll
    for (int k = 0; k < QPT; ++k) mq[k] = __ldg(reinterpret_cast<const float4*>(pl.mask) + 32 * (SW * k + warp) + lane);
    if (in.mask_mode == RDPN_MASK_L1) {
        float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
        for (int k = 0; k < QPT; ++k) minmax4(mq[k], mn, mx);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) { s.red_f[0][warp] = mn; s.red_f[1][warp] = mx; }
    }
    __syncthreads();
    RoiConst rc = s.rc;
    RoiGate gate;
    gate.hi = gate.lo = gate.b = 0.f;
    gate.cut = 0.0;
    gate.incl = 0;
    if (in.mask_mode == RDPN_MASK_L1) {  // every thread folds the eight partials itself: no second barrier
        float lo = s.red_f[0][0], hi = s.red_f[1][0];
#pragma unroll
        for (int w = 1; w < SW; ++w) { lo = fminf(lo, s.red_f[0][w]); hi = fmaxf(hi, s.red_f[1][w]); }
        rc.mn = lo;
        rc.mx = hi;
        make_gate(gate, lo, hi, in.mask_thr, a.mask_cut, a.mask_cut_incl);
    }
    PHASE_MARK(1);

    // ---- 2: gate (gdrn_evaluator.py:110-117 + depth validity), no divisions ----
    unsigned selbits = 0u;
    {
        // mask test first (the gate is a conjunction): ~85 % of the quads have no passing pixel and never touch the
        // other four planes -- no L2 -> SM traffic and no arithmetic for them
        unsigned mbits = 0u;
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            const float mm[4] = {mq[k].x, mq[k].y, mq[k].z, mq[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                mbits |= (mask_pass(mm[j], in.mask_mode, in.mask_thr, rc.mn, gate) ? 1u : 0u) << (4 * k + j);
        }
        // the rest of the gate for one quad whose planes have been loaded; also sort pass A: the per-warp bucket
        // histogram (counts do not depend on the order, so shared-memory atomics on the warp's private row do;
        // deterministic ranks are only needed in pass B)
        uint32_t* hrow = wrun + warp * RB;
        auto gate_quad = [&](unsigned nib, const float4& dq, const float4& xq, const float4& yq, const float4& zq,
                             const uchar4& r4) -> unsigned {
            const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
            const float cxn[4] = {xq.x, xq.y, xq.z, xq.w};
            const float cyn[4] = {yq.x, yq.y, yq.z, yq.w};
            const float czn[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float d = dd[j];
                if (rc.div != 0.f) {  // zero lanes would drag the warp through div.rn's slow path: 0 / f = +-0
                    const float qd = __fdiv_rn(d == 0.f ? 1.f : d, rc.div);
                    d = d == 0.f ? __fmul_rn(d, copysignf(1.f, rc.div)) : qd;
                }
                const float dx = __fmul_rn(__fsub_rn(cxn[j], 0.5f), rc.ext[0]);
                const float dy = __fmul_rn(__fsub_rn(cyn[j], 0.5f), rc.ext[1]);
                const float dz = __fmul_rn(__fsub_rn(czn[j], 0.5f), rc.ext[2]);
                const bool sel = (fabsf(dx) > rc.gthr[0]) && (fabsf(dy) > rc.gthr[1]) && (fabsf(dz) > rc.gthr[2]) && (d > 0.f);
                if (!sel) nib &= ~(1u << j);
            }
            if (nib & 1u) atomicAdd(&hrow[r4.x], 1u);
            if (nib & 2u) atomicAdd(&hrow[r4.y], 1u);
            if (nib & 4u) atomicAdd(&hrow[r4.z], 1u);
            if (nib & 8u) atomicAdd(&hrow[r4.w], 1u);
            return nib;
        };
        auto publish = [&](int q, unsigned nib) {  // gate bitmap: 4 bits per quad, 8 quads per word
            unsigned wbits = nib << (4 * (lane & 7));
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 4);
            if ((lane & 7) == 0) s.selmap[q >> 3] = wbits;
        };
#pragma unroll 1
        for (int k = 0; k < QPT; ++k) {
            const int q = 32 * (SW * k + warp) + lane;
            unsigned nib = (mbits >> (4 * k)) & 0xFu;
            if (nib) {
                const float4 dq = __ldg(reinterpret_cast<const float4*>(pl.depth) + q);
                const float4 xq = __ldg(reinterpret_cast<const float4*>(pl.cx) + q);
                const float4 yq = __ldg(reinterpret_cast<const float4*>(pl.cy) + q);
                const float4 zq = __ldg(reinterpret_cast<const float4*>(pl.cz) + q);
                const uchar4 r4 = DENSE ? make_uchar4(0, 0, 0, 0) : __ldg(reinterpret_cast<const uchar4*>(pl.rid) + q);
                nib = gate_quad(nib, dq, xq, yq, zq, r4);
            }
            selbits |= nib << (4 * k);
            publish(q, nib);
        }
    }

    PHASE_MARK(2);
    // ---- 3: counting sort by region, deterministic order (warp, k, j, lane) ----
    uint32_t* myrun = wrun + warp * RB;
    __syncthreads();
    PHASE_MARK(3);
    // bucket starts, per-warp cursors and the run table
    if (R <= ST) {
        // one thread per bucket: its eight per-warp counts in registers, a block-wide exclusive scan of
        // (points | non-empty << 16) over the buckets, then the cursors -- no single-warp stretch
        const bool mine = t < R;
        int c[SW];
        int tot = 0;
#pragma unroll
        for (int w = 0; w < SW; ++w) {
            c[w] = mine ? (int)wrun[w * RB + t] : 0;
            tot += c[w];
        }
        const int packed = tot | ((tot > 0 ? 1 : 0) << 16);
        // internal sampling: the same scan also ranks the gate bitmap (gated pixels before each 32-pixel word)
        const bool sampling = a.hyp_idx == nullptr;
        const int pc = (sampling && t < RDPN_P / 32) ? __popc(s.selmap[t]) : 0;
        int x = packed, xs = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            const int ys = __shfl_up_sync(0xffffffffu, xs, o);
            if (lane >= o) { x += y; xs += ys; }
        }
        if (lane == 31) { f.red_i[warp] = x; f.red_j[warp] = xs; }
        __syncthreads();
        int base = 0, bases = 0;
#pragma unroll
        for (int w = 0; w < SW; ++w)
            if (w < warp) { base += f.red_i[w]; bases += f.red_j[w]; }
        if (sampling && t < RDPN_P / 32) {
            s.selpfx[t] = (uint16_t)(bases + xs - pc);
            if (t == RDPN_P / 32 - 1) s.selpfx[RDPN_P / 32] = (uint16_t)(bases + xs);
        }
        const int excl = base + x - packed;
        int run = excl & 0xFFFF;
        const int kk = excl >> 16;
        if (mine) {
            const int start = run;
#pragma unroll
            for (int w = 0; w < SW; ++w) {
                wrun[w * RB + t] = (uint32_t)run;
                run += c[w];
            }
            if (tot > 0) {  // one entry per NON-EMPTY bucket = (anchor xyz, start | end << 16)
                float4 hd = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : anchors[t];
                hd.w = __uint_as_float((unsigned)start | ((unsigned)run << 16));
                runtab[kk] = hd;
            }
            if (t == R - 1) { s.n_sel = run; s.n_runs = kk + (tot > 0 ? 1 : 0); }
        }
    } else
    if (warp == RDPN_SERIAL_WARP) {
        const int RPL = (R + 31) / 32;
        int loc = 0, ne = 0;
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R) {
                int tot = 0;
                for (int w = 0; w < SW; ++w) tot += wrun[w * RB + r];
                loc += tot;
                ne += tot > 0 ? 1 : 0;
            }
        }
        int x = loc, kx = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            const int ky = __shfl_up_sync(0xffffffffu, kx, o);
            if (lane >= o) { x += y; kx += ky; }
        }
        int run = x - loc, kk = kx - ne;
        for (int u = 0; u < RPL; ++u) {
            const int r = lane * RPL + u;
            if (r < R) {
                const int start = run;
                for (int w = 0; w < SW; ++w) {
                    const int c = wrun[w * RB + r];
                    wrun[w * RB + r] = (uint32_t)run;
                    run += c;
                }
                if (run > start) {  // one entry per NON-EMPTY bucket = (anchor xyz, start | end << 16)
                    float4 hd = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : anchors[r];
                    hd.w = __uint_as_float((unsigned)start | ((unsigned)run << 16));
                    runtab[kk++] = hd;
                }
            }
        }
        if (lane == 31) { s.n_sel = x; s.n_runs = kx; }
    }
    __syncthreads();
    PHASE_MARK(4);
    // pass B: assign slots
#pragma unroll 1
    for (int k = 0; k < QPT; ++k) {
        const unsigned nib = (selbits >> (4 * k)) & 0xFu;
        if (__ballot_sync(0xffffffffu, nib != 0u) == 0u) continue;
        const uchar4 r4 = DENSE ? make_uchar4(0, 0, 0, 0) : __ldg(reinterpret_cast<const uchar4*>(pl.rid) + 32 * (SW * k + warp) + lane);
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
                s.pix[sl] = (uint16_t)(4 * (32 * (SW * k + warp) + lane) + j);
                s.srid[sl] = (uint8_t)key;
            }
            __syncwarp();
        }
    }

    PHASE_MARK(5);
    // ---- 4: hypothesis generation (FP64 closed form), one hypothesis per thread, pixels gathered ----
    for (int h = t; h < H; h += ST) {
        if (MULTI) {  // S-pair hypotheses (misc.py:72,91): out-of-line Kabsch of the sample
            float* P = hyp + (size_t)h * 12;
            const uint32_t kroi = fmix32(fmix32(a.prm.seed ^ 0x9e3779b9u) ^ (uint32_t)(a.prm.roi_base + b));
            const bool ok = hyp_from_sample<DENSE>(s, *reinterpret_cast<const RoiPlanes*>(smem_raw + lay.planes), anchors,
                                                   a.hyp_idx ? a.hyp_idx + ((size_t)b * H + h) * SS : nullptr, kroi, h, SS, P);
            if (!ok) {
#pragma unroll
                for (int i = 0; i < 12; ++i) P[i] = 0.f;
            }
            hcnt[h] = ok ? 0 : -1;
            continue;
        }
        int ii[3];
        if (a.hyp_idx) {
            const int32_t* ip = a.hyp_idx + ((size_t)b * H + h) * 3;
            const bool first = h == t;
            ii[0] = first ? pre0 : ip[0];
            ii[1] = first ? pre1 : ip[1];
            ii[2] = first ? pre2 : ip[2];
        } else {
            // draw the triplet: k-th gated pixel in raster order, k from the counter-based stream (rank-select on
            // the gate bitmap: binary search over the per-word prefix, then the j-th set bit of the word)
            const uint32_t nsel = s.selpfx[RDPN_P / 32];
            const uint32_t kroi = fmix32(fmix32(a.prm.seed ^ 0x9e3779b9u) ^ (uint32_t)(a.prm.roi_base + b));
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const uint32_t key = fmix32(kroi ^ (uint32_t)(3 * h + v));
                const uint32_t k = (uint32_t)(((unsigned long long)key * nsel) >> 32);
                ii[v] = nsel ? kth_gated_pixel(s.selmap, s.selpfx, k) : -1;
            }
        }
        bool ok = ((unsigned)ii[0] < RDPN_P) && ((unsigned)ii[1] < RDPN_P) && ((unsigned)ii[2] < RDPN_P);
        float* P = hyp + (size_t)h * 12;
        if (ok) {
#pragma unroll
            for (int v = 0; v < 3; ++v) ok = ok && ((s.selmap[ii[v] >> 5] >> (ii[v] & 31)) & 1u);
        }
        if (ok) {
            // register-lean closed form (kabsch_math.cuh): the object side is solved first and parked -- eight
            // doubles per thread in the staging buffer, which is idle until the barrier that ends this phase --
            // then the camera side, so the two bases never sit in registers together (the straightforward form
            // spilled ~60 values per thread to local memory that misses the 32 KB L1 half of the time)
            double* scr = reinterpret_cast<double*>(s.chunk) + t;  // scr[k * ST], k < 8: conflict-free columns
            float pf[3][3];
            int rr[3];
            float4 cw[3], ob[3];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                gather_s1<DENSE>(pl, rc, ii[v], false, in.mask_mode, cw[v], ob[v]);
                if (!DENSE) rr[v] = __ldg(pl.rid + ii[v]);
            }
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                if (DENSE) {
                    pf[v][0] = ob[v].x; pf[v][1] = ob[v].y; pf[v][2] = ob[v].z;
                } else {
                    const float4 an = anchors[rr[v]];
                    pf[v][0] = an.x; pf[v][1] = an.y; pf[v][2] = an.z;
                }
            }
            double ma2, u1a, u2a, v2a;
            {
                TriSide sa;
                tri_side(pf, sa);
                ok = sa.ok;
                scr[0 * ST] = sa.e1[0]; scr[1 * ST] = sa.e1[1]; scr[2 * ST] = sa.e1[2];
                scr[3 * ST] = sa.n[0];  scr[4 * ST] = sa.n[1];  scr[5 * ST] = sa.n[2];
                scr[6 * ST] = sa.m[0];  scr[7 * ST] = sa.m[1];
                ma2 = sa.m[2]; u1a = sa.u1; u2a = sa.u2; v2a = sa.v2;
            }
#pragma unroll
            for (int v = 0; v < 3; ++v) { pf[v][0] = cw[v].x; pf[v][1] = cw[v].y; pf[v][2] = cw[v].z; }
            TriSide sc;
            tri_side(pf, sc);
            ok = ok && sc.ok;
            if (ok) {
                // ma[2] rides in a register: patch it into the last scratch use by passing ma2 separately
                kabsch3_sides(scr, ST, ma2, u1a, u2a, v2a, sc, P);
            }
        }
        if (!ok) {
#pragma unroll
            for (int i = 0; i < 12; ++i) P[i] = 0.f;
        }
        hcnt[h] = ok ? 0 : -1;
    }
    // compact the indices of the valid hypotheses (ascending) so that the scoring warps are densely filled.
    // Round 0 (h = t) publishes its per-warp counts BEFORE the barrier that ends hypothesis generation and the
    // list itself is only needed after the staging barrier, so the usual H <= 256 case costs no extra barrier.
    const bool v0 = t < H && hcnt[t] >= 0;
    const unsigned bal0 = __ballot_sync(0xffffffffu, v0);
    if (lane == 0) f.red_i[warp] = __popc(bal0);
    __syncthreads();  // slots, run table, hypotheses, round-0 counts visible
    PHASE_MARK(6);
    int nvalid = 0;
    {
        int base = 0, tot = 0;
        for (int w = 0; w < SW; ++w) {
            const int c = f.red_i[w];
            if (w < warp) base += c;
            tot += c;
        }
        if (v0) vlist[base + __popc(bal0 & ((1u << lane) - 1u))] = (uint16_t)t;
        nvalid = tot;
    }
    for (int h0 = ST; h0 < H; h0 += ST) {  // H > 256: further rounds, two barriers each
        __syncthreads();
        const int h = h0 + t;
        const bool v = h < H && hcnt[h] >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (lane == 0) f.red_i[warp] = __popc(bal);
        __syncthreads();
        int base = nvalid, tot = 0;
        for (int w = 0; w < SW; ++w) {
            const int c = f.red_i[w];
            if (w < warp) base += c;
            tot += c;
        }
        if (v) vlist[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)h;
        nvalid += tot;
    }
    const int n = s.n_sel;
    const int nruns = s.n_runs;
    if (a.out.n_sel && t == 0) a.out.n_sel[b] = n;
    const bool enough = n >= a.prm.min_pts;
    constexpr int CH = DENSE ? CHUNK / 2 : CHUNK;
    float4* camw_s = s.chunk;
    float4* obj_s = s.chunk + CH;  // dense only

    // ---- 5 + 6: per chunk of gated slots: staging, then inlier scoring ----
    if (enough) {
        const float ncut = -a.sq_cut;  // the contract's inlier test: margin_pt(..., ncut) < 0 (solve_common.cuh)
        const int S = (nvalid >= ST || nvalid == 0) ? 1 : (ST / nvalid);
        for (int c0 = 0; c0 < n; c0 += CH) {
            const int c1 = min(n, c0 + CH);
            if (c0 > 0) __syncthreads();  // previous chunk fully scored
            for (int sl = c0 + t; sl < c1; sl += ST) {
                float4 cw, ob;
                gather_s1<DENSE>(pl, rc, s.pix[sl], a.prm.weighted != 0, in.mask_mode, cw, ob);
                camw_s[sl - c0] = cw;
                if (DENSE) obj_s[sl - c0] = ob;
            }
            __syncthreads();
            PHASE_MARK(7);
            if (!DENSE && RDPN_SCORE_PAIRS) {
                // anchor mode, TWO hypotheses per thread: every staged point (one LDS.128) and every run header is
                // shared by both, which removes ~12 % of the scoring instructions (the loop is issue-bound)
                // Segments are WARP-ALIGNED: W warps cover all pairs once, the SW / W groups of W warps split the runs
                // between them (a warp straddling two segments would execute both halves one after the other and
                // hold the whole CTA at the barrier for twice as long).
                const int npairs = (nvalid + 1) >> 1;
                const int W = max(1, (npairs + 31) >> 5);
                const bool wide = W >= SW;  // H > 512: every thread loops over several pairs, no run split
                const int S2 = wide ? 1 : SW / W;
                const int seg = wide ? 0 : warp / W;
                const int j0 = wide ? t : (warp % W) * 32 + lane;
                const bool whole = (c0 == 0 && c1 == n);
                for (int j = j0; j < npairs && seg < S2; j += wide ? ST : npairs) {
                    const int hA = vlist[2 * j];
                    const bool hasB = 2 * j + 1 < nvalid;
                    const int hB = hasB ? vlist[2 * j + 1] : hA;
                    float PA[12], PB[12];
#pragma unroll
                    for (int i = 0; i < 12; ++i) { PA[i] = hyp[(size_t)hA * 12 + i]; PB[i] = hyp[(size_t)hB * 12 + i]; }
                    int cA = 0, cB = 0;
                    for (int k = seg; k < nruns; k += S2) {
                        const float4 hd = runtab[k];
                        const unsigned se = __float_as_uint(hd.w);
                        int i = (int)(se & 0xFFFFu);
                        int e = (int)(se >> 16);
                        if (!whole) {
                            i = max(i, c0) - c0;
                            e = min(e, c1) - c0;
                            if (i >= e) continue;
                        }
                        float ax, ay, az, bx, by, bz;
                        xform(PA, hd.x, hd.y, hd.z, ax, ay, az);
                        xform(PB, hd.x, hd.y, hd.z, bx, by, bz);
#pragma unroll 1
                        for (; i + 4 <= e; i += 4) {
                            const float4 q0 = camw_s[i], q1 = camw_s[i + 1], q2 = camw_s[i + 2], q3 = camw_s[i + 3];
                            count_in(cA, margin_pt(ax, ay, az, q0.x, q0.y, q0.z, ncut));
                            count_in(cB, margin_pt(bx, by, bz, q0.x, q0.y, q0.z, ncut));
                            count_in(cA, margin_pt(ax, ay, az, q1.x, q1.y, q1.z, ncut));
                            count_in(cB, margin_pt(bx, by, bz, q1.x, q1.y, q1.z, ncut));
                            count_in(cA, margin_pt(ax, ay, az, q2.x, q2.y, q2.z, ncut));
                            count_in(cB, margin_pt(bx, by, bz, q2.x, q2.y, q2.z, ncut));
                            count_in(cA, margin_pt(ax, ay, az, q3.x, q3.y, q3.z, ncut));
                            count_in(cB, margin_pt(bx, by, bz, q3.x, q3.y, q3.z, ncut));
                        }
#pragma unroll 1
                        for (; i < e; ++i) {
                            const float4 q0 = camw_s[i];
                            count_in(cA, margin_pt(ax, ay, az, q0.x, q0.y, q0.z, ncut));
                            count_in(cB, margin_pt(bx, by, bz, q0.x, q0.y, q0.z, ncut));
                        }
                    }
                    if (S2 == 1) {
                        hcnt[hA] += cA;
                        if (hasB) hcnt[hB] += cB;
                    } else {
                        atomicAdd(&hcnt[hA], cA);
                        if (hasB) atomicAdd(&hcnt[hB], cB);
                    }
                }
                continue;
            }
            for (int item = t; item < nvalid * S; item += ST) {
                const int h = vlist[item % nvalid], seg = item / nvalid;
                float P[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) P[i] = hyp[(size_t)h * 12 + i];
                int c = 0;
                if (DENSE) {
                    const int m = c1 - c0;
                    const int i0 = (int)(((long long)m * seg) / S), i1 = (int)(((long long)m * (seg + 1)) / S);
#pragma unroll 4
                    for (int i = i0; i < i1; ++i) {
                        const float4 cp = camw_s[i];
                        const float4 ap = obj_s[i];
                        count_in(c, margin(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z, ncut));
                    }
                } else {
                    // slots are sorted by region: per run (non-empty bucket) the transformed anchor R a + t is
                    // computed once; every point then costs 1 LDS.128 + 6 FP32 + compare + predicated add.
                    const bool whole = (c0 == 0 && c1 == n);  // the usual case: everything in one chunk, no clipping
                    for (int k = seg; k < nruns; k += S) {  // runs interleaved over the segments
                        const float4 hd = runtab[k];
                        const unsigned se = __float_as_uint(hd.w);
                        int i = (int)(se & 0xFFFFu);
                        int e = (int)(se >> 16);
                        if (!whole) {
                            i = max(i, c0) - c0;
                            e = min(e, c1) - c0;
                            if (i >= e) continue;
                        }
                        float tx, ty, tz;
                        xform(P, hd.x, hd.y, hd.z, tx, ty, tz);
#pragma unroll 1
                        for (; i + 4 <= e; i += 4) {
                            const float4 q0 = camw_s[i], q1 = camw_s[i + 1], q2 = camw_s[i + 2], q3 = camw_s[i + 3];
                            count_in(c, margin_pt(tx, ty, tz, q0.x, q0.y, q0.z, ncut));
                            count_in(c, margin_pt(tx, ty, tz, q1.x, q1.y, q1.z, ncut));
                            count_in(c, margin_pt(tx, ty, tz, q2.x, q2.y, q2.z, ncut));
                            count_in(c, margin_pt(tx, ty, tz, q3.x, q3.y, q3.z, ncut));
                        }
                        if (i < e) {  // 1..3 left: one masked batch (warp-uniform predicates), indices clamped into the run
                            const float4 q0 = camw_s[i], q1 = camw_s[min(i + 1, e - 1)], q2 = camw_s[min(i + 2, e - 1)];
                            count_in(c, margin_pt(tx, ty, tz, q0.x, q0.y, q0.z, ncut));
                            if (i + 1 < e) count_in(c, margin_pt(tx, ty, tz, q1.x, q1.y, q1.z, ncut));
                            if (i + 2 < e) count_in(c, margin_pt(tx, ty, tz, q2.x, q2.y, q2.z, ncut));
                        }
                    }
                }
                if (S == 1) hcnt[h] += c;
                else atomicAdd(&hcnt[h], c);
            }
        }
    }
    __syncthreads();
    PHASE_MARK(8);
    // slot accessor of the refit: the staged chunk when everything fitted into one, else re-gathered
    const bool one_chunk = n <= CH;
    auto get_slot = [&](int i, float4& cp, float4& ap) {
        if (one_chunk) {
            cp = camw_s[i];
            if (DENSE) ap = obj_s[i];
        } else {
            gather_s1<DENSE>(pl, rc, s.pix[i], a.prm.weighted != 0, in.mask_mode, cp, ap);
        }
        if (!DENSE) ap = anchors[s.srid[i]];
    };

    // ---- 7a: best hypothesis (misc.py:121) with optional adaptive stop (misc.py:134-138) ----
    int best_r = -1, nbest_r = 0;
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
                if (lane == 31) f.red_i[warp] = x;
                __syncthreads();
                int wbase = 0;
                for (int w = 0; w < warp; ++w) wbase += f.red_i[w];
                int tot = 0;
                for (int w = 0; w < SW; ++w) tot += f.red_i[w];
                const int i_ransac = running + wbase + x;
                if (v && adaptive_stop(hcnt[h], n, i_ransac, lc, a.prm.min_iter)) atomicMin(&f.h_eff, h + 1);
                running += tot;
                __syncthreads();
            }
        }
        const int heff = f.h_eff;
        unsigned long long key = 0ull;
        for (int h = t; h < heff; h += ST) {
            const int c = hcnt[h];
            if (c >= a.prm.min_inliers && c > 0) {
                const unsigned long long k = ((unsigned long long)(unsigned)c << 32) | (unsigned)(0x7FFFFFFF - h);
                key = k > key ? k : key;
            }
        }
        key = warp_max_u64(key);
        if (lane == 0) f.red_k[warp] = key;
        __syncthreads();
        unsigned long long kb = 0ull;
#pragma unroll
        for (int w = 0; w < SW; ++w) kb = f.red_k[w] > kb ? f.red_k[w] : kb;  // every thread: no second barrier
        if (kb) {
            best_r = 0x7FFFFFFF - (int)(kb & 0xFFFFFFFFull);
            nbest_r = (int)(kb >> 32);
        }
    }
    const int best = best_r;
    PHASE_MARK(9);

    // optional diagnostics
    if (a.out.hyp_counts)
        for (int h = t; h < H; h += ST) a.out.hyp_counts[(size_t)b * H + h] = enough ? max(hcnt[h], 0) : 0;
    if (a.out.hyp_poses)
        for (int i = t; i < H * 12; i += ST) a.out.hyp_poses[(size_t)b * H * 12 + i] = enough ? hyp[i] : 0.f;

    if (!enough || best < 0) {
        if (t < 12) a.out.pose[(size_t)b * 12 + t] = -100.f;  // gdrn_evaluator.py:395
        if (a.out.rows16 && t < 16) {
            const float st = enough ? (float)RDPN_STATUS_NO_CONSENSUS : (float)RDPN_STATUS_FEW_POINTS;
            a.out.rows16[(size_t)b * 16 + t] = t < 12 ? -100.f : (t == 12 ? 0.f : (t == 13 ? st : (t == 14 ? (float)n : -1.f)));
        }
        if (t == 0) {
            a.out.n_inliers[b] = 0;
            a.out.status[b] = enough ? RDPN_STATUS_NO_CONSENSUS : RDPN_STATUS_FEW_POINTS;
            if (a.out.best_h) a.out.best_h[b] = -1;
            if (a.out.scale) a.out.scale[b] = 1.f;
        }
        return;
    }

    // ---- 7b: refit on the inliers (misc.py:123-126 -> transform.py:913-980), FP64 accumulation ----
    if (t < 12) f.pose[t] = hyp[(size_t)best * 12 + t];  // stays if the refit bails out (< 3 inliers)
    const float ncut = -a.sq_cut;
    float out_scale = 1.f;
    const int iters = a.prm.refit_iters < 1 ? 1 : a.prm.refit_iters;
    for (int it = 0; it < iters; ++it) {
        float P[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) P[i] = it == 0 ? hyp[(size_t)best * 12 + i] : f.pose[i];
        // FP64 raw moments about a pivot (slot 0, exact FP32 differences) over the slots t, t+ST, ... (<= 16 per
        // thread), from which centroids, cross-covariance and spreads follow; the residual cancellation is ~1e2 on
        // 1e-16, far below the FP32 rounding of the result.
        float4 cp0, ap0;
        get_slot(0, cp0, ap0);
        // Two sweeps of nine moments each keep the live FP64 state at 18 registers (one sweep of 18 spills at the
        // 64-register cap of four CTAs per SM); the second sweep re-reads the inliers found by the first.
        unsigned inl_bits = 0u;
        {
            double m[9];  // sum w | w c (3) | w a (3) | w |c|^2 | w |a|^2
#pragma unroll
            for (int i = 0; i < 9; ++i) m[i] = 0.0;
            int slot = 0;
            for (int i = t; i < n; i += ST, ++slot) {
                float4 cp, ap;
                get_slot(i, cp, ap);
                if (is_inlier(P, ap.x, ap.y, ap.z, cp.x, cp.y, cp.z, ncut)) {
                    inl_bits |= 1u << slot;
                    const double w = a.prm.weighted ? (double)cp.w : 1.0;
                    const double c0 = (double)cp.x - (double)cp0.x, c1 = (double)cp.y - (double)cp0.y, c2 = (double)cp.z - (double)cp0.z;
                    const double a0 = (double)ap.x - (double)ap0.x, a1 = (double)ap.y - (double)ap0.y, a2 = (double)ap.z - (double)ap0.z;
                    m[0] += w;
                    m[1] += w * c0; m[2] += w * c1; m[3] += w * c2;
                    m[4] += w * a0; m[5] += w * a1; m[6] += w * a2;
                    m[7] += w * (c0 * c0 + c1 * c1 + c2 * c2);
                    m[8] += w * (a0 * a0 + a1 * a1 + a2 * a2);
                }
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double v = warp_sum(m[i]);
                if (lane == 0) f.red_d[warp][i < 7 ? i : i + 9] = v;  // slots 0..6, 16, 17
            }
        }
        {
            double m[9];  // sum w c a^T
#pragma unroll
            for (int i = 0; i < 9; ++i) m[i] = 0.0;
            int slot = 0;
            for (int i = t; i < n; i += ST, ++slot) {
                if (!(inl_bits & (1u << slot))) continue;
                float4 cp, ap;
                get_slot(i, cp, ap);
                const double w = a.prm.weighted ? (double)cp.w : 1.0;
                const double c0 = w * ((double)cp.x - (double)cp0.x), c1 = w * ((double)cp.y - (double)cp0.y),
                             c2 = w * ((double)cp.z - (double)cp0.z);
                const double a0 = (double)ap.x - (double)ap0.x, a1 = (double)ap.y - (double)ap0.y, a2 = (double)ap.z - (double)ap0.z;
                m[0] += c0 * a0; m[1] += c0 * a1; m[2] += c0 * a2;
                m[3] += c1 * a0; m[4] += c1 * a1; m[5] += c1 * a2;
                m[6] += c2 * a0; m[7] += c2 * a1; m[8] += c2 * a2;
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double v = warp_sum(m[i]);
                if (lane == 0) f.red_d[warp][7 + i] = v;
            }
        }
        const int winl = warp_sum(__popc(inl_bits));
        if (lane == 0) f.red_i[warp] = winl;
        __syncthreads();
        PHASE_MARK(10);
        if (warp == RDPN_SERIAL_WARP) {
            if (lane < 18) {
                double acc = 0.0;
#pragma unroll
                for (int w = 0; w < SW; ++w) acc += f.red_d[w][lane];
                f.bc_d[lane] = acc;
            }
            if (lane == 0) {
                int ti = 0;
#pragma unroll
                for (int w = 0; w < SW; ++w) ti += f.red_i[w];
                f.h_eff = ti;  // h_eff is free after the best selection: reuse as the broadcast slot
            }
            __syncwarp();
            if (f.h_eff >= 3) {
                // The moments stay in shared memory (the per-warp partials are consumed: red_d is scratch now):
                // lanes 0-8 form the cross-covariance, lanes 9 / 10 the spreads, lane 0 solves for the rotation
                // reading S from shared memory, lanes 0-2 assemble one pose row each.  Nothing here holds the 18
                // moments, S and R in registers at once (that version spilled ~140 values of the solving lane).
                double* S = &f.red_d[0][0];  // S[0..8] | ga S[9] | gb S[10] | R S[16..24]
                const double isw = 1.0 / f.bc_d[0];
                if (lane < 9) {
                    const int r = lane / 3, c = lane - 3 * r;
                    S[lane] = f.bc_d[7 + lane] - f.bc_d[1 + r] * (f.bc_d[4 + c] * isw);  // sum w c a^T - (sum w c)(mean a)^T
                } else if (lane == 9) {
                    S[9] = f.bc_d[17] - (f.bc_d[4] * (f.bc_d[4] * isw) + f.bc_d[5] * (f.bc_d[5] * isw) + f.bc_d[6] * (f.bc_d[6] * isw));
                } else if (lane == 10) {
                    S[10] = f.bc_d[16] - (f.bc_d[1] * (f.bc_d[1] * isw) + f.bc_d[2] * (f.bc_d[2] * isw) + f.bc_d[3] * (f.bc_d[3] * isw));
                }
                __syncwarp();
                if (lane == 0) rotation_from_cov(S, S[9], S[10], S + 16);
                __syncwarp();
                const double sc = a.prm.with_scale ? sqrt(S[10] / S[9]) : 1.0;  // transform.py:971-975
                if (lane < 3) {
                    const int r = lane;
                    const double ma0 = f.bc_d[4] * isw + (double)ap0.x, ma1 = f.bc_d[5] * isw + (double)ap0.y,
                                 ma2 = f.bc_d[6] * isw + (double)ap0.z;
                    const double mcr = f.bc_d[1 + r] * isw + (double)(r == 0 ? cp0.x : (r == 1 ? cp0.y : cp0.z));
                    const double r0 = S[16 + 3 * r], r1 = S[17 + 3 * r], r2 = S[18 + 3 * r];
                    f.pose[4 * r + 0] = (float)(sc * r0);
                    f.pose[4 * r + 1] = (float)(sc * r1);
                    f.pose[4 * r + 2] = (float)(sc * r2);
                    f.pose[4 * r + 3] = (float)(mcr - sc * (r0 * ma0 + r1 * ma1 + r2 * ma2));
                }
                if (lane == 0) f.bc_d[19] = sc;
            }
        }
        __syncthreads();
        if (f.h_eff < 3) break;  // uniform across the block
        if (a.out.inlier_mask && it == iters - 1) {  // the inlier set used by the last refit
            int slot = 0;
            for (int i = t; i < n; i += ST, ++slot)
                if (inl_bits & (1u << slot)) a.out.inlier_mask[(size_t)b * RDPN_P + s.pix[i]] = 1;
        }
        out_scale = (float)f.bc_d[19];
        if (it + 1 < iters) __syncthreads();  // next iteration overwrites the reduction scratch
    }

    PHASE_MARK(11);
    // ---- outputs (+ translation sanity, gdrn_evaluator.py:293-296) ----
    if (t == 0) {
        int status = RDPN_STATUS_OK;
        const float tx = f.pose[3], ty = f.pose[7], tz = f.pose[11];
        if (a.t_net) {
            const double d0 = (double)a.t_net[3 * b] - tx, d1 = (double)a.t_net[3 * b + 1] - ty,
                         d2 = (double)a.t_net[3 * b + 2] - tz;
            if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 1.0) {
                status = RDPN_STATUS_T_SANITY;
                f.pose[3] = a.t_net[3 * b];
                f.pose[7] = a.t_net[3 * b + 1];
                f.pose[11] = a.t_net[3 * b + 2];
            }
        }
        a.out.n_inliers[b] = nbest_r;
        a.out.status[b] = status;
        f.red_i[0] = status;
        if (a.out.best_h) a.out.best_h[b] = best;
        if (a.out.scale) a.out.scale[b] = out_scale;
    }
    __syncthreads();
    if (t < 12) a.out.pose[(size_t)b * 12 + t] = f.pose[t];
    if (a.out.rows16 && t < 16)  // gather row: pose(12) | n_inliers | status | n_sel | best_h
        a.out.rows16[(size_t)b * 16 + t] =
            t < 12 ? f.pose[t] : (t == 12 ? (float)nbest_r : (t == 13 ? (float)f.red_i[0] : (t == 14 ? (float)n : (float)best)));
    PHASE_MARK(12);
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

template <bool DENSE, bool MULTI>
static int launch_solve(const SolveArgs& a, cudaStream_t st) {
    const int H = a.prm.num_hyp;
    const int R = DENSE ? 1 : a.in.num_regions;
    const int RB = R + 1;
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    FusedLayout lay;
    size_t off = al(sizeof(FusedSmem));
    lay.anchors = (int)off; off = al(off + (size_t)R * sizeof(float4));
    lay.runtab = (int)off;  off = al(off + (size_t)R * sizeof(float4));
    lay.wrun = (int)off;    off = al(off + (size_t)SW * RB * sizeof(uint32_t));
    lay.hyp = (int)off;     off = al(off + (size_t)H * 12 * sizeof(float));
    lay.hcnt = (int)off;    off = al(off + (size_t)H * sizeof(int));
    lay.vlist = (int)off;   off = al(off + (size_t)H * sizeof(uint16_t));
    lay.planes = (int)off;  if (MULTI) off = al(off + sizeof(RoiPlanes));
    lay.total = (int)off;
    const size_t smem = off;
    if (smem > 227 * 1024) return RDPN_E_TOOLARGE;
    {
        const int rc = ensure_func_smem((const void*)pose_solve_kernel<DENSE, MULTI>, (DENSE ? 1 : 0) + (MULTI ? 2 : 0), smem);
        if (rc) return rc;
    }
    pose_solve_kernel<DENSE, MULTI><<<a.in.B, ST, smem, st>>>(a, lay);
    ++g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

}  // namespace rdpn

extern "C" {

#ifdef RDPN_PHASE_CLOCKS
int rdpn_debug_set_phase_clocks(long long* d_buf) {
    return (int)cudaMemcpyToSymbol(rdpn::g_phase_clk, &d_buf, sizeof(d_buf));
}
#endif

static int env_int_(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
#define RDPN_DEFAULT_CHUNK_ROIS 8192

static int solve_prepare(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net, const rdpn_solve_params* prm,
                         const rdpn_solve_outputs* out, rdpn::SolveArgs* a, bool* dense) {
    int rc = rdpn::check_roi_inputs(in, dense);
    if (rc) return rc;
    if (!prm || !out || !out->pose || !out->n_inliers || !out->status) return RDPN_E_BADARG;  // d_hyp_idx NULL: internal sampling
    if (prm->num_hyp <= 0 || !(prm->inlier_thr > 0.f)) return RDPN_E_BADARG;
    if (prm->sample_size != 0 && (prm->sample_size < 3 || prm->sample_size > RDPN_MAX_SAMPLE)) return RDPN_E_BADARG;
    if (prm->pipeline < RDPN_PIPELINE_AUTO || prm->pipeline > RDPN_PIPELINE_SPLIT || prm->chunk_rois < 0) return RDPN_E_BADARG;
    if (prm->select_rule != RDPN_SELECT_MOST_INLIERS && prm->select_rule != RDPN_SELECT_MIN_MEAN_ERR) return RDPN_E_BADARG;
    if (out->inlier_mask && ((uintptr_t)out->inlier_mask & 15)) return RDPN_E_ALIGN;
    if (out->hyp_poses && ((uintptr_t)out->hyp_poses & 15)) return RDPN_E_ALIGN;
    a->in = *in;
    a->hyp_idx = d_hyp_idx;
    a->t_net = d_t_net;
    a->prm = *prm;
    if (a->prm.sample_size == 0) a->prm.sample_size = 3;
    a->out = *out;
    a->sq_cut = rdpn::host_sq_cut(prm->inlier_thr);
    rdpn::host_mask_cut(in->mask_thr, &a->mask_cut, &a->mask_cut_incl);
    return 0;
}

// which implementation runs.  AUTO: the three-kernel pipeline (solve_pipe.cu) for batches of at least
// RDPN_PIPELINE_MIN_ROIS ROIs (below that the fused kernel's single launch wins: measured on B200, the pipeline is ahead
// from 2048 ROIs on -- 9.9 vs 9.5 M ROI/s there, 13.4 vs 11.5 M at 8192 -- and behind at 1024: profiles/r2/pipeline.jsonl), the fused kernel otherwise or where the pipeline does not apply (dense mode).  The caller's
// prm->pipeline or the environment variable RDPN_SOLVE_PIPELINE ("fused" / "split", tuning) override it.
#define RDPN_PIPELINE_MIN_ROIS 2048
static bool use_split(const rdpn::SolveArgs& a, bool dense, int* err) {
    int mode = a.prm.pipeline;
    if (mode == RDPN_PIPELINE_AUTO) {
        const char* e = getenv("RDPN_SOLVE_PIPELINE");
        if (e && !strcmp(e, "fused")) mode = RDPN_PIPELINE_FUSED;
        if (e && !strcmp(e, "split")) mode = RDPN_PIPELINE_SPLIT;
    }
    const bool ok = rdpn::split_supported(a, dense);
    if (a.prm.select_rule == RDPN_SELECT_MIN_MEAN_ERR) {  // the reference loop's return value: pipeline only
        *err = (!ok || mode == RDPN_PIPELINE_FUSED) ? RDPN_E_TOOLARGE : 0;
        return *err == 0;
    }
    *err = (mode == RDPN_PIPELINE_SPLIT && !ok) ? RDPN_E_TOOLARGE : 0;
    if (mode == RDPN_PIPELINE_AUTO) return ok && a.in.B >= env_int_("RDPN_PIPELINE_MIN_ROIS", RDPN_PIPELINE_MIN_ROIS);
    return mode != RDPN_PIPELINE_FUSED && ok;
}

static int solve_fused(const rdpn::SolveArgs& a, bool dense, cudaStream_t st) {
    if (a.prm.sample_size > 3) return dense ? rdpn::launch_solve<true, true>(a, st) : rdpn::launch_solve<false, true>(a, st);
    return dense ? rdpn::launch_solve<true, false>(a, st) : rdpn::launch_solve<false, false>(a, st);
}

size_t rdpn_pose_solve_workspace_bytes(int B, int num_hyp, int num_regions, int chunk_rois) {
    if (B <= 0 || num_hyp <= 0) return 0;
    const bool dense = num_regions <= 0;
    if (dense) return 128;  // dense mode runs the fused kernel: no workspace
    const int chunk = chunk_rois > 0 ? chunk_rois : env_int_("RDPN_SOLVE_CHUNK", RDPN_DEFAULT_CHUNK_ROIS);
    return rdpn::split_workspace_bytes(B, num_hyp, num_regions, chunk);
}

int rdpn_pose_solve_ws(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net, const rdpn_solve_params* prm,
                       const rdpn_solve_outputs* out, void* d_ws, size_t ws_bytes, void* stream) {
    RDPN_NVTX("rdpn_pose_solve_ws");
    rdpn::SolveArgs a;
    bool dense = false;
    int rc = solve_prepare(in, d_hyp_idx, d_t_net, prm, out, &a, &dense);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int err = 0;
    if (!use_split(a, dense, &err)) return err ? err : solve_fused(a, dense, st);
    const int chunk = a.prm.chunk_rois > 0 ? a.prm.chunk_rois : env_int_("RDPN_SOLVE_CHUNK", RDPN_DEFAULT_CHUNK_ROIS);
    return rdpn::launch_split(a, dense, d_ws, ws_bytes, chunk, st);
}

int rdpn_pose_solve_stage_ms(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net, const rdpn_solve_params* prm,
                             const rdpn_solve_outputs* out, void* d_ws, size_t ws_bytes, void* stream, float* ms3) {
    RDPN_NVTX("rdpn_pose_solve_stage_ms");
    if (!ms3) return RDPN_E_BADARG;
    rdpn::SolveArgs a;
    bool dense = false;
    int rc = solve_prepare(in, d_hyp_idx, d_t_net, prm, out, &a, &dense);
    if (rc) return rc;
    if (!rdpn::split_supported(a, dense)) return RDPN_E_TOOLARGE;
    ms3[0] = ms3[1] = ms3[2] = 0.f;
    rdpn::set_stage_timing(ms3);
    rc = rdpn::launch_split(a, dense, d_ws, ws_bytes, a.prm.chunk_rois, (cudaStream_t)stream);
    rdpn::set_stage_timing(nullptr);
    return rc;
}

int rdpn_pose_solve(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net,
                    const rdpn_solve_params* prm, const rdpn_solve_outputs* out, void* stream) {
    RDPN_NVTX("rdpn_pose_solve");
    rdpn::SolveArgs a;
    bool dense = false;
    int rc = solve_prepare(in, d_hyp_idx, d_t_net, prm, out, &a, &dense);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int err = 0;
    if (!use_split(a, dense, &err)) return err ? err : solve_fused(a, dense, st);
    const int chunk = a.prm.chunk_rois > 0 ? a.prm.chunk_rois : env_int_("RDPN_SOLVE_CHUNK", RDPN_DEFAULT_CHUNK_ROIS);
    const size_t need = rdpn_pose_solve_workspace_bytes(in->B, a.prm.num_hyp, dense ? 0 : in->num_regions, chunk);
    void* ws = nullptr;
    size_t ws_bytes = 0;
    rc = rdpn::cached_workspace(st, need, &ws, &ws_bytes);
    if (rc) return rc;
    return rdpn::launch_split(a, dense, ws, ws_bytes, chunk, st);
}

int rdpn_kabsch(const float* d_src, const float* d_dst, const float* d_w, int N, int with_scale, float* d_out_M,
                float* d_out_scale, int B, void* stream) {
    RDPN_NVTX("rdpn_kabsch");
    if (!d_src || !d_dst || !d_out_M || B <= 0) return RDPN_E_BADARG;
    if (N < 3) return RDPN_E_BADARG;  // transform.py:917-918 raises ValueError
    rdpn::kabsch_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_src, d_dst, d_w, N, with_scale, d_out_M, d_out_scale);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
