// Microbenchmark: does packed FP32 (FADD2/FMUL2/FFMA2, sm_100a) lift the issue limit of the inlier-scoring loop?
//   ffma / ffma2        : dependent FMA chains, 8 per thread
//   score_scalar        : the scoring inner loop of pose_solve.cu (2 hypotheses per thread, LDS.128 per point)
//   score_packed        : same arithmetic per element, two POINTS per packed instruction (SoA quads in smem)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return (u64)__float_as_uint(a) | ((u64)__float_as_uint(b) << 32); }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void count_if_ltu(int& c, float d2, float cut) {  // d2 >= +0 or NaN: unsigned bit order == float order
    asm("{\n.reg .pred p;\nsetp.lt.u32 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(c) : "r"(__float_as_uint(d2)), "r"(__float_as_uint(cut)));
}
__device__ __forceinline__ void count_if_lt(int& c, float d2, float cut) {
    asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(c) : "f"(d2), "f"(cut));
}

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
    u64 x[8];
    const u64 a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma2(x[i], a2, b2);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += lo(x[i]) + hi(x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

constexpr int NP = 512;
__global__ void __launch_bounds__(256, 4) k_score_scalar(int* out, int iters, float cut) {
    __shared__ float4 pts[NP];
    for (int i = threadIdx.x; i < NP; i += 256) pts[i] = make_float4(i * 0.001f, i * 0.002f, 1.f + i * 0.0005f, 1.f);
    __syncthreads();
    const float ax = threadIdx.x * 0.001f, ay = 0.3f, az = 1.1f, bx = 0.2f, by = threadIdx.x * 0.002f, bz = 1.2f;
    int cA = 0, cB = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int i = 0; i < NP; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 q = pts[i + j];
                float dx = __fsub_rn(ax, q.x), dy = __fsub_rn(ay, q.y), dz = __fsub_rn(az, q.z);
                float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                count_if_lt(cA, d2, cut);
                dx = __fsub_rn(bx, q.x); dy = __fsub_rn(by, q.y); dz = __fsub_rn(bz, q.z);
                d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                count_if_lt(cB, d2, cut);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = cA + (cB << 16);
}
// SoA quads: [x0..x3][y0..y3][z0..z3] per group of 4 points
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_score_packed(int* out, int iters, float cut) {
    __shared__ float4 pts[NP / 4 * 3];
    for (int i = threadIdx.x; i < NP; i += 256) {
        float* p = reinterpret_cast<float*>(pts) + (i / 4) * 12 + (i & 3);
        p[0] = i * 0.001f; p[4] = i * 0.002f; p[8] = 1.f + i * 0.0005f;
    }
    __syncthreads();
    const float ax = threadIdx.x * 0.001f, ay = 0.3f, az = 1.1f, bx = 0.2f, by = threadIdx.x * 0.002f, bz = 1.2f;
    const u64 Ax = pk(ax, ax), Ay = pk(ay, ay), Az = pk(az, az), Bx = pk(bx, bx), By = pk(by, by), Bz = pk(bz, bz);
    int cA = 0, cB = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int g = 0; g < NP / 4; ++g) {
            const ulonglong2 X = reinterpret_cast<const ulonglong2*>(pts)[3 * g];
            const ulonglong2 Y = reinterpret_cast<const ulonglong2*>(pts)[3 * g + 1];
            const ulonglong2 Z = reinterpret_cast<const ulonglong2*>(pts)[3 * g + 2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const u64 xx = h ? X.y : X.x, yy = h ? Y.y : Y.x, zz = h ? Z.y : Z.x;
                u64 d = sub2(Ax, xx), e = sub2(Ay, yy), f = sub2(Az, zz);
                u64 s = fma2(f, f, fma2(e, e, mul2(d, d)));
                if (MODE == 2) { count_if_ltu(cA, lo(s), cut); count_if_ltu(cA, hi(s), cut); }
                else if (MODE == 0) { count_if_lt(cA, lo(s), cut); count_if_lt(cA, hi(s), cut); }
                else { cA -= (int)(lo(s) < cut ? -1 : 0) ; int m0, m1; asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m0) : "f"(lo(s)), "f"(cut)); asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m1) : "f"(hi(s)), "f"(cut)); cA = cA - m0 - m1; }
                d = sub2(Bx, xx); e = sub2(By, yy); f = sub2(Bz, zz);
                s = fma2(f, f, fma2(e, e, mul2(d, d)));
                if (MODE == 2) { count_if_ltu(cB, lo(s), cut); count_if_ltu(cB, hi(s), cut); }
                else if (MODE == 0) { count_if_lt(cB, lo(s), cut); count_if_lt(cB, hi(s), cut); }
                else { int m0, m1; asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m0) : "f"(lo(s)), "f"(cut)); asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(m1) : "f"(hi(s)), "f"(cut)); cB = cB - m0 - m1; }
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = cA + (cB << 16);
}

template <class F>
static float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, grid = sms * 8;
    float* out; cudaMalloc(&out, grid * 256 * 4);
    int iters = 20000;
    float ms = timeit([&] { k_ffma<<<grid, 256>>>(out, iters, 1.0001f, 0.5f); });
    printf("{\"probe\":\"ffma\",\"tflops\":%.2f}\n", 2.0 * 8 * iters * grid * 256 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_ffma2<<<grid, 256>>>(out, iters, 1.0001f, 0.5f); });
    printf("{\"probe\":\"ffma2\",\"tflops\":%.2f}\n", 4.0 * 8 * iters * grid * 256 / (ms * 1e-3) / 1e12);
    const int grid2 = sms * 4, it2 = 200;
    const double pairs = 2.0 * NP * it2 * grid2 * 256;
    ms = timeit([&] { k_score_scalar<<<grid2, 256>>>((int*)out, it2, 0.01f); });
    printf("{\"probe\":\"score_scalar\",\"Gpairs_s\":%.1f,\"clk_per_warp_pair_per_smsp\":%.2f}\n", pairs / (ms * 1e-3) / 1e9,
           (ms * 1e-3) * p.clockRate * 1e3 / (pairs / 32 / (sms * 4)));
    ms = timeit([&] { k_score_packed<0><<<grid2, 256>>>((int*)out, it2, 0.01f); });
    printf("{\"probe\":\"score_packed_setp\",\"Gpairs_s\":%.1f,\"clk_per_warp_pair_per_smsp\":%.2f}\n", pairs / (ms * 1e-3) / 1e9,
           (ms * 1e-3) * p.clockRate * 1e3 / (pairs / 32 / (sms * 4)));
    ms = timeit([&] { k_score_packed<1><<<grid2, 256>>>((int*)out, it2, 0.01f); });
    printf("{\"probe\":\"score_packed_set_iadd3\",\"Gpairs_s\":%.1f,\"clk_per_warp_pair_per_smsp\":%.2f}\n", pairs / (ms * 1e-3) / 1e9,
           (ms * 1e-3) * p.clockRate * 1e3 / (pairs / 32 / (sms * 4)));
    ms = timeit([&] { k_score_packed<2><<<grid2, 256>>>((int*)out, it2, 0.01f); });
    printf("{\"probe\":\"score_packed_isetp\",\"Gpairs_s\":%.1f,\"clk_per_warp_pair_per_smsp\":%.2f}\n", pairs / (ms * 1e-3) / 1e9,
           (ms * 1e-3) * p.clockRate * 1e3 / (pairs / 32 / (sms * 4)));
    return 0;
}
