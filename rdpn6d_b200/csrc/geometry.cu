// Small geometry kernels around the pose path: ROI crop intrinsics (a1), generic back-projection
// (B4), pose assembly from (rot | rot6d, centroid, z) with allocentric->egocentric on the GPU (B3),
// region arg-max (f2) and the FP32 FMA throughput probe used as the scoring-stage roofline peak.
#include "common.cuh"

#include <math.h>

namespace rdpn {
extern unsigned long long g_launch_count;

// a1: K' = [[A],[0,0,1]] . K with A the rot=0 crop affine of core/utils/data_utils.py:111-152 and
// data_loader.py:553-568.  The reference stores its three source points in float32 (:136-142) before
// cv2.getAffineTransform, so the scale factors are (crop/2) / float32-rounded differences; this kernel
// repeats those roundings in FP64 (same closed form and operation order as oracle roi_affine()).
__global__ void roi_intrinsics_kernel(const float* __restrict__ K, const float* __restrict__ center,
                                      const float* __restrict__ scale, int crop_res, float* __restrict__ Kp, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double x0 = (double)center[2 * b], y0 = (double)center[2 * b + 1];
    const double h = __dmul_rn(0.5, (double)scale[b]);
    const double y1 = (double)__double2float_rn(__dsub_rn(y0, h));
    const double e1 = __dsub_rn(y0, y1);
    const double e = (double)__double2float_rn(e1);
    const double x2 = (double)__double2float_rn(__dsub_rn(x0, e));
    const double e2 = __dsub_rn(x0, x2);
    const double half = 0.5 * (double)crop_res;
    const double a00 = __ddiv_rn(half, e2), a11 = __ddiv_rn(half, e1);
    const double tx = __dsub_rn(half, __dmul_rn(a00, x0));
    const double ty = __dsub_rn(half, __dmul_rn(a11, y0));
    const float* k = K + 9 * (size_t)b;
    Kp[4 * b + 0] = (float)__dmul_rn(a00, (double)k[0]);
    Kp[4 * b + 1] = (float)__dmul_rn(a11, (double)k[4]);
    Kp[4 * b + 2] = (float)__dadd_rn(__dmul_rn(a00, (double)k[2]), tx);
    Kp[4 * b + 3] = (float)__dadd_rn(__dmul_rn(a11, (double)k[5]), ty);
}

// lib/pysixd/misc.py:319-349: out[h,w,:] = ((w - cx) * d / fx, (h - cy) * d / fy, d)
__global__ void backproject_kernel(const float* __restrict__ depth, const float* __restrict__ K, int k_stride,
                                   float* __restrict__ out, int B, int H, int W) {
    const size_t total = (size_t)B * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const int b = (int)(i / ((size_t)W * H));
        const float* k = K + (size_t)k_stride * b;
        const float d = depth[i];
        // zero numerators (no depth) bypass div.rn's special-operand slow path: 0 / f = +-0
        const float nx = __fmul_rn(__fsub_rn((float)w, k[2]), d), ny = __fmul_rn(__fsub_rn((float)h, k[5]), d);
        const float qx = __fdiv_rn(nx == 0.f ? 1.f : nx, k[0]), qy = __fdiv_rn(ny == 0.f ? 1.f : ny, k[4]);
        const float X = nx == 0.f ? __fmul_rn(nx, copysignf(1.f, k[0])) : qx;
        const float Y = ny == 0.f ? __fmul_rn(ny, copysignf(1.f, k[4])) : qy;
        out[3 * i + 0] = X;
        out[3 * i + 1] = Y;
        out[3 * i + 2] = d;
    }
}

// pose_from_pred_centroid_z.py:52-141 (test branch) + utils.py:39-94 + rot_reps.py:34-49
// trans_mode: 0 = centroid relative to the box + z (pose_from_pred_centroid_z.py), 1 = absolute 2-D centre + absolute z
// (pose_from_pred_centroid_z_abs.py:27-49), 2 = translation given in `centroid` [B,3] (pose_from_pred.py:21-23).
// rot_is_6d: 0 = 3x3 matrix, 1 = rot6d, 2 = quaternion (w, x, y, z), normalised internally (RT_transform.py:177-183).
__global__ void centroid_z_kernel(const float* __restrict__ rot_in, int rot_is_6d, const float* __restrict__ centroid,
                                  const float* __restrict__ zval, const float* __restrict__ K,
                                  const float* __restrict__ center, const float* __restrict__ rr,
                                  const float* __restrict__ wh, int is_allo, int z_rel, float* __restrict__ rot_out,
                                  float* __restrict__ trans_out, int B, int trans_mode) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float tx, ty, z;
    if (trans_mode == 2) {
        tx = centroid[3 * b];
        ty = centroid[3 * b + 1];
        z = centroid[3 * b + 2];
    } else {
        const float* k = K + 9 * (size_t)b;
        float cx, cy;
        if (trans_mode == 0) {  // :68-74  c = centroid * wh + center
            cx = __fadd_rn(__fmul_rn(centroid[2 * b], wh[2 * b]), center[2 * b]);
            cy = __fadd_rn(__fmul_rn(centroid[2 * b + 1], wh[2 * b + 1]), center[2 * b + 1]);
        } else {
            cx = centroid[2 * b];
            cy = centroid[2 * b + 1];
        }
        z = zval[b];
        if (z_rel && trans_mode == 0) z = __fmul_rn(z, rr[b]);  // :84
        tx = __fdiv_rn(__fmul_rn(z, __fsub_rn(cx, k[2])), k[0]);  // :102
        ty = __fdiv_rn(__fmul_rn(z, __fsub_rn(cy, k[5])), k[4]);
    }
    trans_out[3 * b + 0] = tx;
    trans_out[3 * b + 1] = ty;
    trans_out[3 * b + 2] = z;
    float Rm[9];
    double Rq[9];
    bool from_quat = false;
    if (rot_is_6d == 2) {  // transforms3d quat2mat in float64, as the reference's numpy path (pose_from_pred.py:29-43)
        const float* q = rot_in + 4 * (size_t)b;
        const double w = q[0], x = q[1], y = q[2], zq = q[3];
        const double Nq = w * w + x * x + y * y + zq * zq;
        from_quat = true;
        if (Nq < 2.220446049250313e-16) {
            for (int i = 0; i < 9; ++i) Rq[i] = (i % 4 == 0) ? 1.0 : 0.0;
        } else {
            const double s2 = 2.0 / Nq, X = x * s2, Y = y * s2, Z = zq * s2;
            const double wX = w * X, wY = w * Y, wZ = w * Z, xX = x * X, xY = x * Y, xZ = x * Z, yY = y * Y, yZ = y * Z, zZ = zq * Z;
            Rq[0] = 1.0 - (yY + zZ); Rq[1] = xY - wZ;         Rq[2] = xZ + wY;
            Rq[3] = xY + wZ;         Rq[4] = 1.0 - (xX + zZ); Rq[5] = yZ - wX;
            Rq[6] = xZ - wY;         Rq[7] = yZ + wX;         Rq[8] = 1.0 - (xX + yY);
        }
        for (int i = 0; i < 9; ++i) Rm[i] = (float)Rq[i];
    } else if (rot_is_6d) {  // rot_reps.py:34-49: columns x, y, z
        const float* p = rot_in + 6 * (size_t)b;
        float x0 = p[0], x1 = p[1], x2 = p[2];
        float n = fmaxf(sqrtf(x0 * x0 + x1 * x1 + x2 * x2), 1e-8f);
        x0 /= n; x1 /= n; x2 /= n;
        float z0 = x1 * p[5] - x2 * p[4], z1 = x2 * p[3] - x0 * p[5], z2 = x0 * p[4] - x1 * p[3];
        n = fmaxf(sqrtf(z0 * z0 + z1 * z1 + z2 * z2), 1e-8f);
        z0 /= n; z1 /= n; z2 /= n;
        const float y0 = z1 * x2 - z2 * x1, y1 = z2 * x0 - z0 * x2, y2 = z0 * x1 - z1 * x0;
        Rm[0] = x0; Rm[1] = y0; Rm[2] = z0;
        Rm[3] = x1; Rm[4] = y1; Rm[5] = z1;
        Rm[6] = x2; Rm[7] = y2; Rm[8] = z2;
    } else {
        for (int i = 0; i < 9; ++i) Rm[i] = rot_in[9 * (size_t)b + i];
    }
    if (is_allo) {
        // utils.py:57-66: rotate by the angle between the optical axis and the ray to the object
        const double t0 = tx, t1 = ty, t2 = z;
        const double nrm = sqrt(t0 * t0 + t1 * t1 + t2 * t2);
        const double o0 = t0 / nrm, o1 = t1 / nrm, o2 = t2 / nrm;
        const double cosang = fmin(1.0, fmax(-1.0, o2));
        const double angle = acos(cosang);
        if (angle > 0.0) {
            // axis = cam_ray x obj_ray = (-o1, o0, 0), normalised (transforms3d axangle2mat)
            double ax = -o1, ay = o0;
            const double an = sqrt(ax * ax + ay * ay);
            ax /= an; ay /= an;
            const double c = cos(angle), s = sin(angle), C = 1.0 - c;
            const double M[9] = {ax * ax * C + c, ax * ay * C, ay * s,
                                 ax * ay * C, ay * ay * C + c, -ax * s,
                                 -ay * s, ax * s, c};
            float out[9];
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc) {
                    const double r0 = from_quat ? Rq[cc] : (double)Rm[cc], r1 = from_quat ? Rq[3 + cc] : (double)Rm[3 + cc],
                                 r2 = from_quat ? Rq[6 + cc] : (double)Rm[6 + cc];
                    out[3 * r + cc] = (float)(M[3 * r] * r0 + M[3 * r + 1] * r1 + M[3 * r + 2] * r2);
                }
            for (int i = 0; i < 9; ++i) Rm[i] = out[i];
        }
    }
    for (int i = 0; i < 9; ++i) rot_out[9 * (size_t)b + i] = Rm[i];
}

// GDRN.py:206-209: argmax over channels 1..R of region[B,R+1,P]; softmax is monotone so the raw
// logits give the same index; first maximum wins (torch.argmax).  One thread per pixel quad.
__global__ void region_argmax_kernel(const float* __restrict__ region, int R, uint8_t* __restrict__ out, int B) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // quad index over B*P/4
    const size_t nq = (size_t)B * (RDPN_P / 4);
    if (q >= nq) return;
    const int b = (int)(q / (RDPN_P / 4));
    const int qq = (int)(q % (RDPN_P / 4));
    const float4* base = reinterpret_cast<const float4*>(region + ((size_t)b * (R + 1) + 1) * RDPN_P) + qq;
    float4 best = __ldcs(base);
    uchar4 bi = make_uchar4(0, 0, 0, 0);
    for (int r = 1; r < R; ++r) {
        const float4 v = __ldcs(base + (size_t)r * (RDPN_P / 4));
        if (v.x > best.x) { best.x = v.x; bi.x = (uint8_t)r; }
        if (v.y > best.y) { best.y = v.y; bi.y = (uint8_t)r; }
        if (v.z > best.z) { best.z = v.z; bi.z = (uint8_t)r; }
        if (v.w > best.w) { best.w = v.w; bi.w = (uint8_t)r; }
    }
    reinterpret_cast<uchar4*>(out)[q] = bi;
}

// lib/pysixd/misc.py:288-316 (calc_emb_bp_fast) and :352-371 (backproject_v2): out[v,u,:] = m * R^T (d * Kinv (u,v,1)^T - T) in
// float64 (the reference's numpy einsum promotes to float64), m = (d != 0) for calc_emb_bp_fast, R = I, T = 0 for
// backproject_v2.  mats: Kinv (9) | R (9) | T (3), doubles.
__global__ void backproject_kinv_kernel(const float* __restrict__ depth, const double* __restrict__ mats, int H, int W,
                                        double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const double u = (double)(i % W), v = (double)(i / W), d = (double)depth[i];
    const double* Ki = mats;
    const double* R = mats + 9;
    const double* T = mats + 18;
    double p[3], q[3];
    for (int r = 0; r < 3; ++r) p[r] = d * ((Ki[3 * r] * u + Ki[3 * r + 1] * v) + Ki[3 * r + 2]) - T[r];
    for (int c = 0; c < 3; ++c) q[c] = (R[c] * p[0] + R[3 + c] * p[1]) + R[6 + c] * p[2];  // R^T p
    const double m = depth[i] != 0.f ? 1.0 : 0.0;
    for (int c = 0; c < 3; ++c) out[3 * (size_t)i + c] = q[c] * m;
}

// lib/pysixd/pose_error.py:315-337 (adi): mean over the ground-truth-posed model points of the distance to the nearest
// estimate-posed model point.  Brute force: one thread per ground-truth point, the estimate-posed cloud staged through
// shared memory in tiles; FP64 distances, block partial sums combined in block order by the last block.
// NN = false is pose_error.py:297-312 (add): the distance to the SAME point under the estimated pose.
template <bool NN>
__global__ void adi_kernel(const float* __restrict__ pts, int n, const double* __restrict__ poses /* R_est t_est R_gt t_gt: 24 */,
                           double* __restrict__ partial, unsigned* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double tile[256][3];
    __shared__ double red[8];
    __shared__ bool last;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double* Re = poses;
    const double* te = poses + 9;
    const double* Rg = poses + 12;
    const double* tg = poses + 21;
    double g[3] = {0, 0, 0};
    if (i < n) {
        const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        for (int r = 0; r < 3; ++r) g[r] = Rg[3 * r] * x + Rg[3 * r + 1] * y + Rg[3 * r + 2] * z + tg[r];
    }
    double best = 1e300;
    if (!NN && i < n) {
        const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        best = 0.0;
        for (int r = 0; r < 3; ++r) {
            const double d = g[r] - (Re[3 * r] * x + Re[3 * r + 1] * y + Re[3 * r + 2] * z + te[r]);
            best += d * d;
        }
    }
    for (int j0 = 0; NN && j0 < n; j0 += 256) {
        const int j = j0 + threadIdx.x;
        if (j < n) {
            const double x = pts[3 * j], y = pts[3 * j + 1], z = pts[3 * j + 2];
            for (int r = 0; r < 3; ++r) tile[threadIdx.x][r] = Re[3 * r] * x + Re[3 * r + 1] * y + Re[3 * r + 2] * z + te[r];
        }
        __syncthreads();
        const int m = min(256, n - j0);
        for (int k = 0; k < m; ++k) {
            const double dx = g[0] - tile[k][0], dy = g[1] - tile[k][1], dz = g[2] - tile[k][2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            best = d2 < best ? d2 : best;
        }
        __syncthreads();
    }
    double v = i < n ? sqrt(best) : 0.0;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += red[w];
        partial[blockIdx.x] = a;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            double tot = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) tot += __ldcg(partial + b);
            *out = tot / (double)n;
            *ticket = 0u;
        }
    }
}

__global__ void fp32_probe_kernel(float* out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace rdpn

extern "C" {

int rdpn_roi_intrinsics(const float* d_K, const float* d_center, const float* d_scale, int crop_res, float* d_Kp, int B,
                        void* stream) {
    RDPN_NVTX("rdpn_roi_intrinsics");
    if (!d_K || !d_center || !d_scale || !d_Kp || B <= 0 || crop_res <= 0) return RDPN_E_BADARG;
    rdpn::roi_intrinsics_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_K, d_center, d_scale, crop_res, d_Kp, B);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_backproject(const float* d_depth, const float* d_K, int k_stride, float* d_out, int B, int H, int W,
                     void* stream) {
    RDPN_NVTX("rdpn_backproject");
    if (!d_depth || !d_K || !d_out || B <= 0 || H <= 0 || W <= 0 || (k_stride != 0 && k_stride != 9)) return RDPN_E_BADARG;
    const size_t total = (size_t)B * H * W;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    rdpn::backproject_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_depth, d_K, k_stride, d_out, B, H, W);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_centroid_z_to_pose(const float* d_rot_in, int rot_is_6d, const float* d_centroid, const float* d_z,
                            const float* d_K, const float* d_center, const float* d_resize_ratio, const float* d_wh,
                            int is_allo, int z_type_rel, float* d_rot_out, float* d_trans_out, int B, void* stream) {
    RDPN_NVTX("rdpn_centroid_z_to_pose");
    if (!d_rot_in || !d_centroid || !d_z || !d_K || !d_center || !d_wh || !d_rot_out || !d_trans_out || B <= 0)
        return RDPN_E_BADARG;
    if (z_type_rel && !d_resize_ratio) return RDPN_E_BADARG;
    rdpn::centroid_z_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_rot_in, rot_is_6d, d_centroid, d_z, d_K, d_center, d_resize_ratio, d_wh, is_allo, z_type_rel, d_rot_out,
        d_trans_out, B, 0);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_backproject_kinv(const float* d_depth, const double* d_mats, int H, int W, double* d_out, void* stream) {
    RDPN_NVTX("rdpn_backproject_kinv");
    if (!d_depth || !d_mats || !d_out || H <= 0 || W <= 0) return RDPN_E_BADARG;
    rdpn::backproject_kinv_kernel<<<(H * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_depth, d_mats, H, W, d_out);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_adi(const float* d_pts, int n, const double* d_poses, double* d_scratch, double* d_out, void* stream) {
    RDPN_NVTX("rdpn_adi");
    if (!d_pts || !d_poses || !d_scratch || !d_out || n <= 0) return RDPN_E_BADARG;
    const int blocks = (n + 255) / 256;
    // scratch: blocks partial sums + one ticket word (zero before the first use; the kernel leaves it zero)
    rdpn::adi_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_pts, n, d_poses, d_scratch + 1, (unsigned*)d_scratch, d_out);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_add(const float* d_pts, int n, const double* d_poses, double* d_scratch, double* d_out, void* stream) {
    RDPN_NVTX("rdpn_add");
    if (!d_pts || !d_poses || !d_scratch || !d_out || n <= 0) return RDPN_E_BADARG;
    const int blocks = (n + 255) / 256;
    rdpn::adi_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_pts, n, d_poses, d_scratch + 1, (unsigned*)d_scratch, d_out);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_assemble_pose(const float* d_rot_in, int rot_kind, const float* d_trans_or_centroid, const float* d_z, const float* d_K,
                       int trans_mode, int is_allo, float* d_rot_out, float* d_trans_out, int B, void* stream) {
    RDPN_NVTX("rdpn_assemble_pose");
    if (!d_rot_in || !d_trans_or_centroid || !d_rot_out || !d_trans_out || B <= 0) return RDPN_E_BADARG;
    if (rot_kind < 0 || rot_kind > 2 || (trans_mode != 1 && trans_mode != 2)) return RDPN_E_BADARG;
    if (trans_mode == 1 && (!d_z || !d_K)) return RDPN_E_BADARG;
    rdpn::centroid_z_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_rot_in, rot_kind, d_trans_or_centroid, d_z, d_K, nullptr, nullptr, nullptr, is_allo, 0, d_rot_out, d_trans_out, B, trans_mode);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_region_argmax(const float* d_region, int R, uint8_t* d_region_idx, int B, void* stream) {
    RDPN_NVTX("rdpn_region_argmax");
    if (!d_region || !d_region_idx || B <= 0 || R <= 0 || R > 255) return RDPN_E_BADARG;
    if (((uintptr_t)d_region | (uintptr_t)d_region_idx) & 15) return RDPN_E_ALIGN;
    const size_t nq = (size_t)B * (RDPN_P / 4);
    rdpn::region_argmax_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_region, R, d_region_idx, B);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

int rdpn_fp32_peak_probe(int iters, double* out_flops) {
    RDPN_NVTX("rdpn_fp32_peak_probe");
    if (iters <= 0 || !out_flops) return RDPN_E_BADARG;
    int dev = 0, sms = 0;
    RDPN_CUDA_TRY(cudaGetDevice(&dev));
    RDPN_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    float* buf = nullptr;
    RDPN_CUDA_TRY(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rdpn::fp32_probe_kernel<<<blocks, threads>>>(buf, iters / 4 + 1, 1.f);  // warm-up
    cudaEventRecord(e0);
    rdpn::fp32_probe_kernel<<<blocks, threads>>>(buf, iters, 1.f);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    rdpn::g_launch_count += 2;
    if (err != cudaSuccess) return (int)err;
    *out_flops = (double)blocks * threads * (double)iters * 64.0 * 2.0 / ((double)ms * 1e-3);
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// f4: region / residual targets -- core/utils/data_utils.py:229-244 (xyz_to_region).
//   region id = 1 + argmin_r ||xyz - fps_r||  (scipy cdist: float64, sqrt(sum of squares), first minimum),
//   0 where xyz == (0,0,0) (background); delta = xyz - fps[region-1] for every pixel (also background).
// xyz [B,P,3] (HWC as the loader holds it), fps [B,R,3]; out region [B,P] uint8, delta [B,P,3].
// ------------------------------------------------------------------------------------------------
namespace rdpn {
__global__ void xyz_to_region_kernel(const float* __restrict__ xyz, const float* __restrict__ fps, int R, int P,
                                     uint8_t* __restrict__ region, float* __restrict__ delta, int B) {
    extern __shared__ float sfps[];  // [R*3]
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < R * 3; i += blockDim.x) sfps[i] = fps[(size_t)b * R * 3 + i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float* v = xyz + ((size_t)b * P + p) * 3;
    const float x = v[0], y = v[1], z = v[2];
    int best = 0;
    double bd = 1e300;
    for (int r = 0; r < R; ++r) {
        const double dx = __dsub_rn((double)x, (double)sfps[3 * r]), dy = __dsub_rn((double)y, (double)sfps[3 * r + 1]),
                     dz = __dsub_rn((double)z, (double)sfps[3 * r + 2]);
        const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        if (d < bd) { bd = d; best = r; }
    }
    const bool fg = (x != 0.f) || (y != 0.f) || (z != 0.f);  // data_utils.py:232-233
    region[(size_t)b * P + p] = fg ? (uint8_t)(best + 1) : (uint8_t)0;
    float* o = delta + ((size_t)b * P + p) * 3;
    o[0] = __fsub_rn(x, sfps[3 * best]);
    o[1] = __fsub_rn(y, sfps[3 * best + 1]);
    o[2] = __fsub_rn(z, sfps[3 * best + 2]);
}
}  // namespace rdpn

extern "C" int rdpn_xyz_to_region(const float* d_xyz, const float* d_fps, int R, int P, uint8_t* d_region, float* d_delta, int B,
                                  void* stream) {
    RDPN_NVTX("rdpn_xyz_to_region");
    if (!d_xyz || !d_fps || !d_region || !d_delta || B <= 0 || P <= 0 || R <= 0 || R > 254) return RDPN_E_BADARG;
    dim3 grid((P + 255) / 256, B);
    rdpn::xyz_to_region_kernel<<<grid, 256, (size_t)R * 3 * sizeof(float), (cudaStream_t)stream>>>(d_xyz, d_fps, R, P, d_region, d_delta, B);
    ++rdpn::g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}
