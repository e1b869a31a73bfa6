// Farthest point sampling on sm_100a: a persistent cooperative kernel with the cloud resident in
// registers and one hand-rolled grid-wide arg-max per selected point.
//
// Replaces /root/reference/core/csrc/fps/src/farthest_point_sampling.cpp (single-threaded C++,
// O(K*N) with two passes over std::vector per iteration).  Bit-exactness contract (see
// oracle/fps_oracle.c): squared distances ((dx*dx)+(dy*dy))+(dz*dz) in FP32 with no FMA
// (cpp:25), centre (max+min)*(1.f/2.f) (cpp:20,138), strict '>' arg-max from 0 so the lowest index
// wins ties and index 0 is returned when nothing positive is left (cpp:56-73), selected points
// excluded from update and arg-max (cpp:50,66), no update after the last pick (cpp:154).
//
// Design: N points are dealt to G blocks x 512 threads x PPT registers (coalesced at load time,
// never re-read).  Each round a thread updates its PPT running minima against the current pick,
// builds a 64-bit key  (float_bits(min_dist) << 32) | (0xFFFFFFFF - index)  (0 when not a
// candidate), and the key is max-reduced warp -> block -> grid.  The grid step is one
// red.max.u64 + one red.release.add per block on per-round slots, then a spin on ld.acquire: no
// cooperative-groups grid.sync, no slot reset, ~4 L2 round trips per pick.
#include "common.cuh"

#include <cooperative_groups.h>
#include <float.h>
#include <stdlib.h>

namespace rdpn {
extern unsigned long long g_launch_count;

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;

__device__ __forceinline__ float sqdist_nofma(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// order-preserving float <-> uint (for min/max through integer atomics)
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct FpsShared {
    unsigned long long warp_key[FPS_WARPS];
    unsigned long long bcast;
    float fmin[FPS_WARPS][3];
    float fmax[FPS_WARPS][3];
    float ctr[3];
};

// Grid-wide max of a 64-bit key; every thread of every block returns the same value.
__device__ __forceinline__ unsigned long long grid_max_key(unsigned long long key, FpsShared& sh, unsigned long long* keys,
                                                           unsigned* cnt, int slot) {
    key = warp_max_u64(key);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh.warp_key[warp] = key;
    __syncthreads();
    if (warp == 0) {
        unsigned long long k = lane < FPS_WARPS ? sh.warp_key[lane] : 0ull;
        k = warp_max_u64(k);
        if (lane == 0) {
            if (gridDim.x > 1) {
                if (k) red_max_u64(keys + slot, k);
                red_release_add_u32(cnt + slot, 1u);
                while (ld_acquire_u32(cnt + slot) < gridDim.x) {
                }
                k = ld_relaxed_u64(keys + slot);
            }
            sh.bcast = k;
        }
    }
    __syncthreads();
    return sh.bcast;
}

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
    fps_kernel(const float* __restrict__ pts, int* __restrict__ idxs, int pn, int sn, int start,
               unsigned long long* keys, unsigned* cnt, unsigned* bbox) {
    __shared__ FpsShared sh;
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const long long i = ((long long)k * G + b) * FPS_THREADS + t;
        if (i < pn) {
            px[k] = pts[3 * i + 0];
            py[k] = pts[3 * i + 1];
            pz[k] = pts[3 * i + 2];
            md[k] = FLT_MAX;  // cpp:83 / cpp:126
        } else {
            px[k] = py[k] = pz[k] = 0.f;
            md[k] = -1.f;  // never a candidate, never updated (d < -1 is false)
        }
    }
    int slot = 0;
    int cur;
    if (start < 0) {
        // cpp:131-141: bounding-box centre, then min_dist = |p - centre|^2
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (md[k] >= 0.f) {
                mn[0] = fminf(mn[0], px[k]); mx[0] = fmaxf(mx[0], px[k]);
                mn[1] = fminf(mn[1], py[k]); mx[1] = fmaxf(mx[1], py[k]);
                mn[2] = fminf(mn[2], pz[k]); mx[2] = fmaxf(mx[2], pz[k]);
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = warp_min(mn[c]);
            mx[c] = warp_max(mx[c]);
            if (lane == 0) { sh.fmin[warp][c] = mn[c]; sh.fmax[warp][c] = mx[c]; }
        }
        __syncthreads();
        if (t < 3) {
            float lo = FLT_MAX, hi = -FLT_MAX;
            for (int w = 0; w < FPS_WARPS; ++w) { lo = fminf(lo, sh.fmin[w][t]); hi = fmaxf(hi, sh.fmax[w][t]); }
            if (G > 1) {
                atomicMax(bbox + t, ~f2ord(lo));  // max of complement == min
                atomicMax(bbox + 3 + t, f2ord(hi));
            } else {
                sh.ctr[t] = __fmul_rn(__fadd_rn(hi, lo), 0.5f);
            }
        }
        if (G > 1) {
            __syncthreads();
            if (t == 0) {
                __threadfence();
                red_release_add_u32(cnt + slot, 1u);
                while (ld_acquire_u32(cnt + slot) < (unsigned)G) {
                }
            }
            __syncthreads();
            if (t < 3) {
                float lo = ord2f(~__ldcg(bbox + t)), hi = ord2f(__ldcg(bbox + 3 + t));
                sh.ctr[t] = __fmul_rn(__fadd_rn(hi, lo), 0.5f);  // (max+min)*(1.f/2.f), cpp:20,138
            }
            ++slot;
        }
        __syncthreads();
        const float cx = sh.ctr[0], cy = sh.ctr[1], cz = sh.ctr[2];
        unsigned long long key = 0ull;
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (md[k] >= 0.f) {
                const float d = sqdist_nofma(px[k], py[k], pz[k], cx, cy, cz);
                md[k] = d < FLT_MAX ? d : FLT_MAX;  // cpp:141 min(d, FLT_MAX)
                const unsigned i = ((unsigned)k * G + b) * FPS_THREADS + t;
                const unsigned long long kk =
                    md[k] > 0.f ? (((unsigned long long)__float_as_uint(md[k]) << 32) | (0xFFFFFFFFu - i)) : 0ull;
                key = kk > key ? kk : key;
            }
        key = grid_max_key(key, sh, keys, cnt, slot++);  // cpp:149
        cur = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
    } else {
        cur = start;
    }

    for (int it = 0; it < sn; ++it) {
        if (b == 0 && t == 0) idxs[it] = cur;  // cpp:153
        if (it == sn - 1) break;               // cpp:154
        // mark the pick as taken (owner thread only): cpp:152
        {
            const int q = cur / FPS_THREADS;
            if ((cur % FPS_THREADS) == t && (q % G) == b) {
                const int kk = q / G;
#pragma unroll
                for (int k = 0; k < PPT; ++k)
                    if (k == kk) md[k] = -1.f;
            }
        }
        const float cx = __ldg(pts + 3 * (size_t)cur), cy = __ldg(pts + 3 * (size_t)cur + 1),
                    cz = __ldg(pts + 3 * (size_t)cur + 2);
        unsigned long long key = 0ull;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const float d = sqdist_nofma(px[k], py[k], pz[k], cx, cy, cz);
            md[k] = d < md[k] ? d : md[k];  // cpp:52 (taken / padding entries hold -1 and never change)
            const unsigned i = ((unsigned)k * G + b) * FPS_THREADS + t;
            const unsigned long long kk =
                md[k] > 0.f ? (((unsigned long long)__float_as_uint(md[k]) << 32) | (0xFFFFFFFFu - i)) : 0ull;
            key = kk > key ? kk : key;  // cpp:67-71: strict '>' from 0, lowest index on ties
        }
        key = grid_max_key(key, sh, keys, cnt, slot++);
        cur = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;  // cpp:60,72: 0 if none
    }
}

// Clouds beyond the register-resident limit (148 x 512 x 16 = 1.21 M points): same grid-wide arg-max, the cloud and the
// running minima streamed from global memory every pick (20 B per point and pick: HBM-bound, SURVEY 8d's model).
__global__ void __launch_bounds__(FPS_THREADS, 1)
    fps_stream_kernel(const float* __restrict__ pts, float* __restrict__ md, int* __restrict__ idxs, int pn, int sn, int start,
                      unsigned long long* keys, unsigned* cnt, unsigned* bbox) {
    __shared__ FpsShared sh;
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long stride = (long long)G * FPS_THREADS, first = (long long)b * FPS_THREADS + t;
    int slot = 0, cur;
    if (start < 0) {
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (long long i = first; i < pn; i += stride) {
            const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
            mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
            mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
            mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = warp_min(mn[c]);
            mx[c] = warp_max(mx[c]);
            if (lane == 0) { sh.fmin[warp][c] = mn[c]; sh.fmax[warp][c] = mx[c]; }
        }
        __syncthreads();
        if (t < 3) {
            float lo = FLT_MAX, hi = -FLT_MAX;
            for (int w = 0; w < FPS_WARPS; ++w) { lo = fminf(lo, sh.fmin[w][t]); hi = fmaxf(hi, sh.fmax[w][t]); }
            atomicMax(bbox + t, ~f2ord(lo));
            atomicMax(bbox + 3 + t, f2ord(hi));
        }
        __syncthreads();
        if (t == 0) {
            __threadfence();
            red_release_add_u32(cnt + slot, 1u);
            while (ld_acquire_u32(cnt + slot) < (unsigned)G) {
            }
        }
        __syncthreads();
        if (t < 3) sh.ctr[t] = __fmul_rn(__fadd_rn(ord2f(__ldcg(bbox + 3 + t)), ord2f(~__ldcg(bbox + t))), 0.5f);
        ++slot;
        __syncthreads();
        const float cx = sh.ctr[0], cy = sh.ctr[1], cz = sh.ctr[2];
        unsigned long long key = 0ull;
        for (long long i = first; i < pn; i += stride) {
            const float d = sqdist_nofma(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], cx, cy, cz);
            const float m = d < FLT_MAX ? d : FLT_MAX;
            md[i] = m;
            const unsigned long long kk = m > 0.f ? (((unsigned long long)__float_as_uint(m) << 32) | (0xFFFFFFFFu - (unsigned)i)) : 0ull;
            key = kk > key ? kk : key;
        }
        key = grid_max_key(key, sh, keys, cnt, slot++);
        cur = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
    } else {
        for (long long i = first; i < pn; i += stride) md[i] = FLT_MAX;
        cur = start;
    }
    for (int it = 0; it < sn; ++it) {
        if (b == 0 && t == 0) idxs[it] = cur;
        if (it == sn - 1) break;
        const float cx = __ldg(pts + 3 * (size_t)cur), cy = __ldg(pts + 3 * (size_t)cur + 1), cz = __ldg(pts + 3 * (size_t)cur + 2);
        unsigned long long key = 0ull;
        for (long long i = first; i < pn; i += stride) {
            float m = md[i];
            if (i == cur) m = -1.f;  // taken: never updated, never a candidate again (cpp:50,66,152)
            const float d = sqdist_nofma(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], cx, cy, cz);
            m = d < m ? d : m;
            md[i] = m;
            const unsigned long long kk = m > 0.f ? (((unsigned long long)__float_as_uint(m) << 32) | (0xFFFFFFFFu - (unsigned)i)) : 0ull;
            key = kk > key ? kk : key;
        }
        key = grid_max_key(key, sh, keys, cnt, slot++);
        cur = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Small clouds (the sizes the reference's tools run: tools/lm/1_compute_fps.py:26-35, 10^4-10^5 model vertices, <= 256
// picks, one object after the other): ONE THREAD-BLOCK CLUSTER per object, many objects per launch.
// The cloud of an object is dealt to the C <= 8 CTAs of its cluster (registers, as above); the arg-max of a pick is
// block-reduced and every CTA PUSHES its candidate (key + coordinates) into the shared memory of every CTA of the
// cluster, where it is counted by the receiver's own mbarrier -- no cluster barrier, no global atomics, no spinning on
// L2, no remote load.  Same arithmetic, same keys, same tie rule: bit-identical picks.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int FPS_CLUSTER_THREADS_DEFAULT = 128;  // narrowest CTA that holds the cloud at <= 16 points per thread
constexpr int FPS_MAX_CLUSTER = 8;   // the portable cluster size: 8 x 512 x 16 = 65 536 points per object

struct __align__(16) FpsCand {  // a candidate pick travels with its coordinates: the next round needs no global load
    unsigned long long key;
    float x, y, z, pad;
    float pad2[2];
};
struct FpsClusterShared {
    FpsCand wcand[FPS_WARPS];
    FpsCand ccand[2];             // this CTA's candidate, double-buffered by pick parity
    float fmin[FPS_WARPS][3];
    float fmax[FPS_WARPS][3];
    float bb[6];                  // this CTA's bounding box (min xyz, max xyz)
    // push exchange: every CTA of the cluster stores its candidate into EVERY CTA (st.async, 16 + 4 bytes) and the
    // stores complete the destination's mbarrier; double-buffered by pick parity
    uint4 pk[2][FPS_MAX_CLUSTER];  // {distance bits, ~index, x, y}, slot = sending rank
    float pz[2][FPS_MAX_CLUSTER];
    uint64_t bar[2];
    uint4 wk4[2][FPS_WARPS];       // the warps' candidates, reduced by warp 0 after the block barrier; by pick parity, so
    float wz[2][FPS_WARPS];        // that the block barrier of pick i + 1 separates warp 0's reads of pick i from the writes of i + 2
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// remote (or own) shared-memory store that completes `bytes` on the destination CTA's mbarrier when it lands
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(a), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// warp-wide max of a 64-bit key by two 32-bit REDUX instead of five shuffle rounds
__device__ __forceinline__ unsigned long long warp_max_key_redux(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

template <int PPT, bool PUSH, int CT /* threads per CTA */>
__global__ void __launch_bounds__(CT, 1)
    fps_cluster_kernel(const float* __restrict__ pts_all, const int* __restrict__ offs, int* __restrict__ idxs_all, int sn,
                       const int* __restrict__ starts, int one_pn, int one_start) {
    __shared__ FpsClusterShared sh;
    constexpr int CW = CT / 32;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int obj = blockIdx.x / C, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int o0 = offs ? offs[obj] : 0;
    const int pn = offs ? offs[obj + 1] - o0 : one_pn;
    const float* pts = pts_all + 3 * (size_t)o0;
    int* idxs = idxs_all + (size_t)obj * sn;
    const int start = offs ? (starts ? starts[obj] : -1) : one_start;
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int i = (k * C + rank) * CT + t;
        if (i < pn) {
            px[k] = pts[3 * (size_t)i + 0];
            py[k] = pts[3 * (size_t)i + 1];
            pz[k] = pts[3 * (size_t)i + 2];
            md[k] = FLT_MAX;
        } else {
            px[k] = py[k] = pz[k] = 0.f;
            md[k] = -1.f;
        }
    }
    // Cluster-wide max of a key whose owner thread also knows the candidate's coordinates (bx, by, bz).  Returns the
    // winning key and leaves its coordinates in (cx, cy, cz).
    // PUSH (default): two REDUX in the warp, the warps' candidates meet in local shared memory at one block barrier,
    // warp 0 reduces them and stores the CTA's candidate into slot [rank] of EVERY CTA of the cluster (st.async: 16 + 4
    // bytes, each store completing the DESTINATION's mbarrier); a CTA waits on its OWN mbarrier for the 20 C bytes and
    // every thread picks the best of the <= 8 candidates from local shared memory.  No cluster barrier, no remote load
    // on the path of a pick.  Slots and barriers are double-buffered by pick parity: a CTA can receive pick i + 1 while
    // it still reads pick i, and nobody can send pick i + 2 before every CTA of the cluster has sent pick i + 1, i.e.
    // has finished reading pick i.
    // !PUSH (RDPN_FPS_EXCHANGE=barrier): block reduce, ONE cluster barrier, C remote reads of 8 bytes and one of 12.
    int par = 0;
    unsigned xchg = 0;  // exchanges done (PUSH): barrier xchg & 1, phase parity (xchg >> 1) & 1
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (PUSH) {
        if (t == 0) {
            mbar_init(&sh.bar[0], 1);
            mbar_init(&sh.bar[1], 1);
            mbar_fence_init();
            mbar_expect_tx(&sh.bar[0], (uint32_t)(C * (1) * 20));
            mbar_expect_tx(&sh.bar[1], (uint32_t)(C * (1) * 20));
        }
        if (t < 2 * FPS_MAX_CLUSTER) (&sh.pk[0][0])[t] = make_uint4(0u, 0u, 0u, 0u);  // slots of absent ranks: key 0
        cluster.sync();  // every barrier of the cluster is armed (and every slot zeroed) before the first store can arrive
    }
#ifdef RDPN_FPS_TIMING  // cycles of thread 0 between the marks of the exchange, summed over the picks, printed at the end
    long long tk_last = clock64(), tk_sum[5] = {0, 0, 0, 0, 0};
#define RDPN_FPS_TICK(i) { const long long now_ = clock64(); tk_sum[i] += now_ - tk_last; tk_last = now_; }
#else
#define RDPN_FPS_TICK(i)
#endif
    auto cluster_max_key = [&](unsigned long long key, float bx, float by, float bz) -> unsigned long long {
        RDPN_FPS_TICK(0);
        const unsigned long long wk = warp_max_key_redux(key);
        if (PUSH) {
            // stage 1 inside the CTA: the warps' candidates meet in local shared memory at ONE block barrier; warp 0
            // reduces the sixteen and pushes the CTA's candidate to every CTA of the cluster (2 C mbarrier transactions
            // per CTA and pick instead of 32 C: the transactions of one barrier serialise, ~4 cycles each)
            const int b = (int)(xchg & 1u);
            if (wk ? key == wk : lane == 0) {
                sh.wk4[b][warp] = make_uint4((unsigned)(wk >> 32), (unsigned)wk, __float_as_uint(bx), __float_as_uint(by));
                sh.wz[b][warp] = bz;
            }
            RDPN_FPS_TICK(1);
            __syncthreads();
            RDPN_FPS_TICK(2);
            if (warp == 0) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                float z = 0.f;
                if (lane < CW) { v = sh.wk4[b][lane]; z = sh.wz[b][lane]; }
                const unsigned mh = __reduce_max_sync(0xffffffffu, v.x);
                const unsigned ml = __reduce_max_sync(0xffffffffu, v.x == mh ? v.y : 0u);
                if ((mh | ml) ? (lane < CW && v.x == mh && v.y == ml) : lane == 0) {
                    const uint32_t a4 = smem_u32(&sh.pk[b][rank]), az = smem_u32(&sh.pz[b][rank]), ab = smem_u32(&sh.bar[b]);
                    for (int r = 0; r < C; ++r) {
                        const uint32_t rb = mapa_u32(ab, (uint32_t)r);
                        st_async_v4(mapa_u32(a4, (uint32_t)r), v.x, v.y, v.z, v.w, rb);
                        st_async_b32(mapa_u32(az, (uint32_t)r), __float_as_uint(z), rb);
                    }
                }
            }
            const uint32_t ph = (xchg >> 1) & 1u;
            while (!mbar_try_cluster(&sh.bar[b], ph)) {
            }
            RDPN_FPS_TICK(3);
            if (t == 0) mbar_expect_tx(&sh.bar[b], (uint32_t)(C * 20));  // armed for the exchange after the next one
            // every thread reads the (at most eight) CTA candidates by broadcast loads and picks the best in a tree;
            // the slots of ranks >= C stay zero (key 0 never wins)
            unsigned long long kk[FPS_MAX_CLUSTER];
            int wi[FPS_MAX_CLUSTER];
#pragma unroll
            for (int r = 0; r < FPS_MAX_CLUSTER; ++r) {
                const uint2 v = *reinterpret_cast<const uint2*>(&sh.pk[b][r]);
                kk[r] = ((unsigned long long)v.x << 32) | v.y;
                wi[r] = r;
            }
#pragma unroll
            for (int w = 1; w < FPS_MAX_CLUSTER; w *= 2)
#pragma unroll
                for (int r = 0; r < FPS_MAX_CLUSTER; r += 2 * w)
                    if (kk[r + w] > kk[r]) { kk[r] = kk[r + w]; wi[r] = wi[r + w]; }
            const unsigned long long best = kk[0];
            if (best) {
                const uint4 v = sh.pk[b][wi[0]];
                cx = __uint_as_float(v.z); cy = __uint_as_float(v.w); cz = sh.pz[b][wi[0]];
            } else {  // nothing positive left: the reference returns index 0 (cpp:60,72)
                cx = __ldg(pts); cy = __ldg(pts + 1); cz = __ldg(pts + 2);
            }
            RDPN_FPS_TICK(4);
            ++xchg;
            return best;
        }
        if (wk ? key == wk : lane == 0) {  // keys embed the point index: exactly one owner (lane 0 when nothing is left)
            FpsCand c;
            c.key = wk; c.x = bx; c.y = by; c.z = bz; c.pad = 0.f; c.pad2[0] = c.pad2[1] = 0.f;
            sh.wcand[warp] = c;
        }
        __syncthreads();
        if (warp == 0) {
            const unsigned long long k = lane < CW ? sh.wcand[lane].key : 0ull;
            const unsigned long long bk = warp_max_key_redux(k);
            if (bk ? (lane < CW && k == bk) : lane == 0) sh.ccand[par] = sh.wcand[lane];
        }
        cluster.sync();
        unsigned long long best = 0ull;
        int br = 0;
        for (int r = 0; r < C; ++r) {
            const unsigned long long k = cluster.map_shared_rank(&sh.ccand[par], r)->key;
            if (k > best) { best = k; br = r; }
        }
        if (best) {
            const FpsCand* w = cluster.map_shared_rank(&sh.ccand[par], br);
            cx = w->x; cy = w->y; cz = w->z;
        } else {  // nothing positive left: the reference returns index 0 (cpp:60,72)
            cx = __ldg(pts); cy = __ldg(pts + 1); cz = __ldg(pts + 2);
        }
        par ^= 1;  // the next pick writes the other slot: its barrier orders that write after every read of this one
        return best;
    };
    int cur;
    if (start < 0) {
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (md[k] >= 0.f) {
                mn[0] = fminf(mn[0], px[k]); mx[0] = fmaxf(mx[0], px[k]);
                mn[1] = fminf(mn[1], py[k]); mx[1] = fmaxf(mx[1], py[k]);
                mn[2] = fminf(mn[2], pz[k]); mx[2] = fmaxf(mx[2], pz[k]);
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = warp_min(mn[c]);
            mx[c] = warp_max(mx[c]);
            if (lane == 0) { sh.fmin[warp][c] = mn[c]; sh.fmax[warp][c] = mx[c]; }
        }
        __syncthreads();
        if (t < 3) {
            float lo = FLT_MAX, hi = -FLT_MAX;
            for (int w = 0; w < CW; ++w) { lo = fminf(lo, sh.fmin[w][t]); hi = fmaxf(hi, sh.fmax[w][t]); }
            sh.bb[t] = lo;
            sh.bb[3 + t] = hi;
        }
        cluster.sync();
        float ctr[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float lo = FLT_MAX, hi = -FLT_MAX;
            for (int r = 0; r < C; ++r) {
                const float* rb = cluster.map_shared_rank(sh.bb, r);
                lo = fminf(lo, rb[c]);
                hi = fmaxf(hi, rb[3 + c]);
            }
            ctr[c] = __fmul_rn(__fadd_rn(hi, lo), 0.5f);  // (max+min)*(1.f/2.f), cpp:20,138
        }
        // the thread's best candidate by FP32 compare (strict '>' in ascending k = ascending index: the lowest index
        // wins ties, cpp:67-71); the 64-bit key is built once per thread, not once per point
        float bd = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (md[k] >= 0.f) {
                const float d = sqdist_nofma(px[k], py[k], pz[k], ctr[0], ctr[1], ctr[2]);
                md[k] = d < FLT_MAX ? d : FLT_MAX;  // cpp:141
                if (md[k] > bd) { bd = md[k]; bk = k; bx = px[k]; by = py[k]; bz = pz[k]; }
            }
        unsigned long long key =
            bd > 0.f ? (((unsigned long long)__float_as_uint(bd) << 32) | (0xFFFFFFFFu - (((unsigned)bk * C + rank) * CT + t))) : 0ull;
        key = cluster_max_key(key, bx, by, bz);  // cpp:149
        cur = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
    } else {
        cur = start;
        cx = __ldg(pts + 3 * (size_t)cur); cy = __ldg(pts + 3 * (size_t)cur + 1); cz = __ldg(pts + 3 * (size_t)cur + 2);
    }
    for (int it = 0; it < sn; ++it) {
        if (rank == 0 && t == 0) idxs[it] = cur;  // cpp:153
        if (it == sn - 1) break;                  // cpp:154
        if (it == 0 || !PUSH) {  // with the push exchange the owner recognises its own key below (no divisions per pick)
            const int q = cur / CT;
            if ((cur % CT) == t && (q % C) == rank) {
                const int kk = q / C;
#pragma unroll
                for (int k = 0; k < PPT; ++k)
                    if (k == kk) md[k] = -1.f;  // cpp:152
            }
        }
        float bd = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const float d = sqdist_nofma(px[k], py[k], pz[k], cx, cy, cz);
            md[k] = d < md[k] ? d : md[k];  // cpp:52 (taken / padding entries hold -1 and never change)
            if (md[k] > bd) { bd = md[k]; bk = k; bx = px[k]; by = py[k]; bz = pz[k]; }  // cpp:67-71
        }
        unsigned long long key =
            bd > 0.f ? (((unsigned long long)__float_as_uint(bd) << 32) | (0xFFFFFFFFu - (((unsigned)bk * C + rank) * CT + t))) : 0ull;
        const unsigned long long mine = key;
        key = cluster_max_key(key, bx, by, bz);
        cur = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;  // cpp:60,72
        if (PUSH && key && mine == key) {  // cpp:152: the pick is taken (keys embed the index: exactly one thread matches;
#pragma unroll                             // when nothing is left the pick is index 0, whose distance is already <= 0)
            for (int k = 0; k < PPT; ++k)
                if (k == bk) md[k] = -1.f;
        }
    }
#ifdef RDPN_FPS_TIMING
    if (PUSH && t == 0 && rank == 0 && obj == 0)
        printf("fps timing C=%d PPT=%d picks=%d: dist %lld redux %lld block-barrier %lld wait %lld select %lld cycles/pick\n", C, PPT, sn,
               tk_sum[0] / sn, tk_sum[1] / sn, tk_sum[2] / sn, tk_sum[3] / sn, tk_sum[4] / sn);
#endif
    cluster.sync();  // no CTA leaves while a sibling may still read its shared memory
}

// nobj clusters of C CTAs; offs == nullptr: one object of one_pn points.  Picks C and the points per thread so that the
// rate per pick is best and returns RDPN_E_TOOLARGE beyond 8 CTAs x 512 threads x 16 points.
static int fps_cluster_launch(const float* d_pts, const int* d_offs, int32_t* d_idxs, int nobj, int max_pn, int sn, const int* d_starts,
                              int one_start, cudaStream_t st) {
    // threads per CTA x points per thread (benchmarks/fps_ppt_probe.py on B200): the fixed cost of a pick is paid per
    // warp (reductions, barrier, selection), the distance updates per point and SM
    int ct = FPS_CLUSTER_THREADS_DEFAULT;
    if (const char* e = getenv("RDPN_FPS_CLUSTER_THREADS")) {
        const int v = atoi(e);
        ct = v <= 128 ? 128 : v <= 256 ? 256 : 512;
    }
    while (ct < 512 && (long long)FPS_MAX_CLUSTER * ct * 16 < max_pn) ct *= 2;
    int ppt = 2;
    while (ppt < 16 && (long long)FPS_MAX_CLUSTER * ct * ppt < max_pn) ppt *= 2;
    if (const char* e = getenv("RDPN_FPS_CLUSTER_PPT")) {  // tuning: more points per thread = narrower cluster
        const int v = atoi(e);
        ppt = 1;
        while (ppt < 16 && (ppt < v || (long long)FPS_MAX_CLUSTER * ct * ppt < max_pn)) ppt *= 2;
    }
    if ((long long)FPS_MAX_CLUSTER * ct * ppt < max_pn) return RDPN_E_TOOLARGE;
    int C = (int)(((long long)max_pn + (long long)ct * ppt - 1) / ((long long)ct * ppt));
    if (C < 1) C = 1;
    // RDPN_FPS_EXCHANGE = barrier: the cluster-barrier exchange, kept for A/B timing (benchmarks/fps_small.py)
    const char* xe = getenv("RDPN_FPS_EXCHANGE");
    const bool push = !(xe && xe[0] == 'b');
#define RDPN_FPS_PICK2(P, T) (push ? (const void*)fps_cluster_kernel<P, true, T> : (const void*)fps_cluster_kernel<P, false, T>)
#define RDPN_FPS_PICK(P) (ct == 128 ? RDPN_FPS_PICK2(P, 128) : ct == 256 ? RDPN_FPS_PICK2(P, 256) : RDPN_FPS_PICK2(P, 512))
    const void* fn = nullptr;
    switch (ppt) {
        case 1: fn = RDPN_FPS_PICK(1); break;
        case 2: fn = RDPN_FPS_PICK(2); break;
        case 4: fn = RDPN_FPS_PICK(4); break;
        case 8: fn = RDPN_FPS_PICK(8); break;
        default: fn = RDPN_FPS_PICK(16); break;
    }
#undef RDPN_FPS_PICK
#undef RDPN_FPS_PICK2
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nobj * C));
    cfg.blockDim = dim3((unsigned)ct);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void* args[] = {(void*)&d_pts, (void*)&d_offs, (void*)&d_idxs, (void*)&sn, (void*)&d_starts, (void*)&max_pn, (void*)&one_start};
    RDPN_CUDA_TRY(cudaLaunchKernelExC(&cfg, fn, args));
    ++g_launch_count;
    return 0;
}

// Per-axis mean in FP64 with a FIXED summation order (run-to-run identical): every block sums a fixed slice, block
// partials are combined in block order by the block that finishes last.
__global__ void fps_center_kernel(const float* __restrict__ pts, int pn, double* __restrict__ partial, unsigned* __restrict__ ticket,
                                  double* __restrict__ center) {
    __shared__ double red[8][3];
    __shared__ bool last;
    double s[3] = {0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pn; i += (long long)gridDim.x * blockDim.x) {
        s[0] += pts[3 * i];
        s[1] += pts[3 * i + 1];
        s[2] += pts[3 * i + 2];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < 3; ++c) {
        const double v = warp_sum(s[c]);
        if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w][threadIdx.x];
        partial[3 * blockIdx.x + threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < 3) {
        __threadfence();
        double a = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) a += __ldcg(partial + 3 * b + threadIdx.x);
        center[threadIdx.x] = a / (double)pn;
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

__global__ void fps_gather_kernel(const float* __restrict__ pts, const int* __restrict__ idxs, int sn, float* __restrict__ out) {
    // out[i] = pts[idxs[i]] (fps_utils.py:21)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sn * 3; i += gridDim.x * blockDim.x)
        out[i] = pts[3 * (size_t)idxs[i / 3] + (i % 3)];
}

static int fps_launch(const float* d_pts, int32_t* d_idxs, int pn, int sn, int start, void* d_ws, size_t ws_bytes,
                      cudaStream_t st) {
    if (!d_pts || !d_idxs || pn <= 0 || sn <= 0 || start >= pn) return RDPN_E_BADARG;
    if (!d_ws || ws_bytes < rdpn_fps_workspace_bytes(sn)) return RDPN_E_WORKSPACE;
    if (pn <= FPS_MAX_CLUSTER * FPS_THREADS * 16 && !getenv("RDPN_FPS_NO_CLUSTER")) {
        // small clouds: one thread-block cluster, arg-max through distributed shared memory (above ~32 k points the
        // cooperative grid's wider spread wins: benchmarks/fps_small.py)
        return fps_cluster_launch(d_pts, nullptr, d_idxs, 1, pn, sn, nullptr, start, st);
    }
    int dev = 0, sms = 0, coop = 0;
    RDPN_CUDA_TRY(cudaGetDevice(&dev));
    RDPN_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RDPN_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return RDPN_E_NOCOOP;
    int gmax = sms;  // one 512-thread block per SM keeps up to 16 points per thread in registers
    if (const char* e = getenv("RDPN_FPS_BLOCKS")) {
        int v = atoi(e);
        if (v > 0 && v < gmax) gmax = v;
    }
    int ppt = 1;
    while (ppt < 16 && (long long)gmax * FPS_THREADS * ppt < pn) ppt *= 2;
    const bool streaming = (long long)gmax * FPS_THREADS * ppt < pn;
    // beyond the register-resident limit the running minima live behind the workspace header: pn more floats
    if (streaming && ws_bytes < rdpn_fps_workspace_bytes(sn) + (size_t)pn * sizeof(float)) return RDPN_E_TOOLARGE;
    int G = streaming ? gmax : (int)(((long long)pn + (long long)FPS_THREADS * ppt - 1) / ((long long)FPS_THREADS * ppt));
    // workspace: keys[sn+2] u64 | cnt[sn+2] u32 | bbox[6] u32   (all zero-initialised per call)
    unsigned long long* keys = (unsigned long long*)d_ws;
    unsigned* cnt = (unsigned*)(keys + sn + 2);
    unsigned* bbox = cnt + sn + 2;
    RDPN_CUDA_TRY(cudaMemsetAsync(d_ws, 0, rdpn_fps_workspace_bytes(sn), st));
    if (streaming) {
        float* md = reinterpret_cast<float*>((char*)d_ws + rdpn_fps_workspace_bytes(sn));
        void* sargs[] = {(void*)&d_pts, (void*)&md, (void*)&d_idxs, (void*)&pn, (void*)&sn, (void*)&start, (void*)&keys, (void*)&cnt,
                         (void*)&bbox};
        RDPN_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)fps_stream_kernel, dim3(G), dim3(FPS_THREADS), sargs, 0, st));
        ++g_launch_count;
        return 0;
    }
    void* args[] = {(void*)&d_pts, (void*)&d_idxs, (void*)&pn, (void*)&sn, (void*)&start, (void*)&keys, (void*)&cnt, (void*)&bbox};
    const void* fn = nullptr;
    switch (ppt) {
        case 1: fn = (const void*)fps_kernel<1>; break;
        case 2: fn = (const void*)fps_kernel<2>; break;
        case 4: fn = (const void*)fps_kernel<4>; break;
        case 8: fn = (const void*)fps_kernel<8>; break;
        default: fn = (const void*)fps_kernel<16>; break;
    }
    RDPN_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(FPS_THREADS), args, 0, st));
    ++g_launch_count;
    return 0;
}

}  // namespace rdpn

extern "C" {

size_t rdpn_fps_workspace_bytes(int sn) {
    if (sn < 0) sn = 0;
    size_t b = (size_t)(sn + 2) * 8 + (size_t)(sn + 2) * 4 + 6 * 4;
    return (b + 255) & ~(size_t)255;
}

int rdpn_fps_init_center(const float* d_pts, int32_t* d_idxs, int pn, int sn, void* d_ws, size_t ws_bytes,
                         void* stream) {
    RDPN_NVTX("rdpn_fps_init_center");
    return rdpn::fps_launch(d_pts, d_idxs, pn, sn, -1, d_ws, ws_bytes, (cudaStream_t)stream);
}

int rdpn_fps_from_index(const float* d_pts, int32_t* d_idxs, int pn, int sn, int start, void* d_ws, size_t ws_bytes,
                        void* stream) {
    RDPN_NVTX("rdpn_fps_from_index");
    if (start < 0) return RDPN_E_BADARG;
    return rdpn::fps_launch(d_pts, d_idxs, pn, sn, start, d_ws, ws_bytes, (cudaStream_t)stream);
}

int rdpn_fps_gather(const float* d_pts, const int32_t* d_idxs, int pn, int sn, float* d_out, double* d_center,
                    void* stream) {
    RDPN_NVTX("rdpn_fps_gather");
    if (!d_pts || !d_idxs || !d_out || pn <= 0 || sn <= 0) return RDPN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    rdpn::fps_gather_kernel<<<(sn * 3 + 255) / 256, 256, 0, st>>>(d_pts, d_idxs, sn, d_out);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    if (d_center) {
        int blocks = (pn + 256 * 8 - 1) / (256 * 8);
        if (blocks > 592) blocks = 592;
        if (blocks < 1) blocks = 1;
        void* scratch = nullptr;  // block partials + ticket, stream-ordered
        const size_t bytes = (size_t)blocks * 3 * sizeof(double) + 16;
        RDPN_CUDA_TRY(cudaMallocAsync(&scratch, bytes, st));
        RDPN_CUDA_TRY(cudaMemsetAsync((char*)scratch + (size_t)blocks * 3 * sizeof(double), 0, 16, st));
        rdpn::fps_center_kernel<<<blocks, 256, 0, st>>>(d_pts, pn, (double*)scratch,
                                                        (unsigned*)((char*)scratch + (size_t)blocks * 3 * sizeof(double)), d_center);
        ++rdpn::g_launch_count;
        const cudaError_t e = cudaGetLastError();
        cudaFreeAsync(scratch, st);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int rdpn_fps_batch(const float* d_pts, const int32_t* d_offsets, int nobj, int max_pn, int sn, const int32_t* d_starts,
                   int32_t* d_idxs, void* stream) {
    RDPN_NVTX("rdpn_fps_batch");
    if (!d_pts || !d_offsets || !d_idxs || nobj <= 0 || max_pn <= 0 || sn <= 0) return RDPN_E_BADARG;
    return rdpn::fps_cluster_launch(d_pts, d_offsets, d_idxs, nobj, max_pn, sn, d_starts, -1, (cudaStream_t)stream);
}

}  // extern "C"
