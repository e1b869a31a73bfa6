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

__global__ void fps_gather_kernel(const float* __restrict__ pts, const int* __restrict__ idxs, int pn, int sn,
                                  float* __restrict__ out, double* __restrict__ center) {
    // block 0: out[i] = pts[idxs[i]] (fps_utils.py:21); all blocks: FP64 per-axis sums for the mean row
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < sn * 3; i += blockDim.x) out[i] = pts[3 * (size_t)idxs[i / 3] + (i % 3)];
    if (center == nullptr) return;
    double s[3] = {0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pn; i += (long long)gridDim.x * blockDim.x) {
        s[0] += pts[3 * i];
        s[1] += pts[3 * i + 1];
        s[2] += pts[3 * i + 2];
    }
    for (int c = 0; c < 3; ++c) {
        double v = warp_sum(s[c]);
        if ((threadIdx.x & 31) == 0) atomicAdd(center + c, v / (double)pn);
    }
}

static int fps_launch(const float* d_pts, int32_t* d_idxs, int pn, int sn, int start, void* d_ws, size_t ws_bytes,
                      cudaStream_t st) {
    if (!d_pts || !d_idxs || pn <= 0 || sn <= 0 || start >= pn) return RDPN_E_BADARG;
    if (!d_ws || ws_bytes < rdpn_fps_workspace_bytes(sn)) return RDPN_E_WORKSPACE;
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
    if ((long long)gmax * FPS_THREADS * ppt < pn) return RDPN_E_TOOLARGE;
    int G = (int)(((long long)pn + (long long)FPS_THREADS * ppt - 1) / ((long long)FPS_THREADS * ppt));
    // workspace: keys[sn+2] u64 | cnt[sn+2] u32 | bbox[6] u32   (all zero-initialised per call)
    unsigned long long* keys = (unsigned long long*)d_ws;
    unsigned* cnt = (unsigned*)(keys + sn + 2);
    unsigned* bbox = cnt + sn + 2;
    RDPN_CUDA_TRY(cudaMemsetAsync(d_ws, 0, rdpn_fps_workspace_bytes(sn), st));
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
    return rdpn::fps_launch(d_pts, d_idxs, pn, sn, -1, d_ws, ws_bytes, (cudaStream_t)stream);
}

int rdpn_fps_from_index(const float* d_pts, int32_t* d_idxs, int pn, int sn, int start, void* d_ws, size_t ws_bytes,
                        void* stream) {
    if (start < 0) return RDPN_E_BADARG;
    return rdpn::fps_launch(d_pts, d_idxs, pn, sn, start, d_ws, ws_bytes, (cudaStream_t)stream);
}

int rdpn_fps_gather(const float* d_pts, const int32_t* d_idxs, int pn, int sn, float* d_out, double* d_center,
                    void* stream) {
    if (!d_pts || !d_idxs || !d_out || pn <= 0 || sn <= 0) return RDPN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = 1;
    if (d_center) {
        RDPN_CUDA_TRY(cudaMemsetAsync(d_center, 0, 3 * sizeof(double), st));
        blocks = (pn + 256 * 8 - 1) / (256 * 8);
        if (blocks > 592) blocks = 592;
        if (blocks < 1) blocks = 1;
    }
    rdpn::fps_gather_kernel<<<blocks, 256, 0, st>>>(d_pts, d_idxs, pn, sn, d_out, d_center);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
