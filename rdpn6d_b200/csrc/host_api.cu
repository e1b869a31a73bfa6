// Host-facing side of the C ABI: version / error strings, the exact drop-in FPS symbols of the
// reference's cffi extension (HOST pointers, core/csrc/fps/src/ext.h:1-14), and the host-buffer
// pose-solve plugin call with a context that owns device scratch and pipelines chunks of ROIs over
// two streams so that host->device copies overlap the kernels.
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

namespace rdpn {
unsigned long long g_launch_count = 0;
}

extern "C" {

int rdpn_version(void) { return RDPN_VERSION; }

unsigned long long rdpn_launch_count(void) { return rdpn::g_launch_count; }

const char* rdpn_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case RDPN_E_BADARG: return "rdpn: bad argument";
        case RDPN_E_ALIGN: return "rdpn: ROI-map pointer not 16-byte aligned";
        case RDPN_E_WORKSPACE: return "rdpn: workspace too small";
        case RDPN_E_TOOLARGE: return "rdpn: problem too large for the kernel";
        case RDPN_E_NOCOOP: return "rdpn: device lacks cooperative launch";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "rdpn: unknown error";
    }
}

// ------------------------------------------------------------------------------------------------
// drop-in FPS symbols (host pointers, synchronous, no return code -- as the reference: ext.h:1-14).
// Failures cannot be reported through the reference's signature; they abort loudly instead of
// silently falling back to a CPU path.
// ------------------------------------------------------------------------------------------------
static void fps_host(float* pts, int* idxs, int pn, int sn, int start) {
    if (pn <= 0 || sn <= 0) return;
    float* d_pts = nullptr;
    int* d_idx = nullptr;
    void* d_ws = nullptr;
    const size_t ws = rdpn_fps_workspace_bytes(sn);
    cudaError_t e = cudaMalloc(&d_pts, (size_t)pn * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_idx, (size_t)sn * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_ws, ws);
    if (e == cudaSuccess) e = cudaMemcpy(d_pts, pts, (size_t)pn * 3 * sizeof(float), cudaMemcpyHostToDevice);
    int rc = (int)e;
    if (rc == 0)
        rc = start < 0 ? rdpn_fps_init_center(d_pts, d_idx, pn, sn, d_ws, ws, nullptr)
                       : rdpn_fps_from_index(d_pts, d_idx, pn, sn, start, d_ws, ws, nullptr);
    if (rc == 0) rc = (int)cudaMemcpy(idxs, d_idx, (size_t)sn * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_pts);
    cudaFree(d_idx);
    cudaFree(d_ws);
    if (rc != 0) {
        fprintf(stderr, "rdpn6d_b200: farthest_point_sampling failed on the GPU: %s (no CPU fallback)\n",
                rdpn_error_string(rc));
        abort();
    }
}

void farthest_point_sampling_init_center(float* pts, int* idxs, int pn, int sn) { fps_host(pts, idxs, pn, sn, -1); }

void farthest_point_sampling(float* pts, int* idxs, int pn, int sn) {
    if (pn <= 0) return;
    srand((unsigned)time(0));  // farthest_point_sampling.cpp:93-94
    fps_host(pts, idxs, pn, sn, rand() % pn);
}

// ------------------------------------------------------------------------------------------------
// host-buffer pose solve
// ------------------------------------------------------------------------------------------------
#define RDPN_CHUNK 256  // ROIs per pipeline stage (22 MB of maps)

struct rdpn_ctx {
    int device;
    cudaStream_t st[2];
    unsigned char* buf[2];
    size_t buf_bytes;
};

int rdpn_ctx_create(int device, rdpn_ctx** out_ctx) {
    if (!out_ctx) return RDPN_E_BADARG;
    RDPN_CUDA_TRY(cudaSetDevice(device));
    rdpn_ctx* c = (rdpn_ctx*)calloc(1, sizeof(rdpn_ctx));
    c->device = device;
    for (int i = 0; i < 2; ++i) RDPN_CUDA_TRY(cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking));
    *out_ctx = c;
    return 0;
}

void rdpn_ctx_destroy(rdpn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 2; ++i) {
        if (c->buf[i]) cudaFree(c->buf[i]);
        if (c->st[i]) cudaStreamDestroy(c->st[i]);
    }
    free(c);
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int rdpn_pose_solve_host(rdpn_ctx* c, const rdpn_roi_inputs* h, const int32_t* h_hyp, const float* h_tnet,
                         const rdpn_solve_params* prm, const rdpn_solve_outputs* ho) {
    if (!c || !h || !h_hyp || !prm || !ho || h->B <= 0 || prm->num_hyp <= 0) return RDPN_E_BADARG;
    if (!ho->pose || !ho->n_inliers || !ho->status) return RDPN_E_BADARG;
    if ((h->region_idx == nullptr) != (h->anchors == nullptr)) return RDPN_E_BADARG;
    RDPN_CUDA_TRY(cudaSetDevice(c->device));
    const bool dense = h->region_idx == nullptr;
    const int H = prm->num_hyp, R = dense ? 0 : h->num_regions;
    const size_t P = RDPN_P, CH = RDPN_CHUNK;
    // device layout of one stage
    size_t off = 0;
    const size_t o_depth = off; off += al256(CH * P * 4);
    const size_t o_cx = off; off += al256(CH * P * 4);
    const size_t o_cy = off; off += al256(CH * P * 4);
    const size_t o_cz = off; off += al256(CH * P * 4);
    const size_t o_mask = off; off += al256(CH * P * 4);
    const size_t o_rid = off; off += al256(CH * P);
    const size_t o_anc = off; off += al256(CH * (size_t)(R > 0 ? R : 1) * 12);
    const size_t o_kp = off; off += al256(CH * 16);
    const size_t o_ext = off; off += al256(CH * 12);
    const size_t o_div = off; off += al256(CH * 4);
    const size_t o_hyp = off; off += al256(CH * (size_t)H * 12);
    const size_t o_tnet = off; off += al256(CH * 12);
    const size_t o_pose = off; off += al256(CH * 48);
    const size_t o_ninl = off; off += al256(CH * 4);
    const size_t o_stat = off; off += al256(CH * 4);
    const size_t o_best = off; off += al256(CH * 4);
    const size_t o_nsel = off; off += al256(CH * 4);
    const size_t o_scale = off; off += al256(CH * 4);
    const size_t o_imask = off; off += al256(CH * P);
    const size_t o_hcnt = off; off += al256(CH * (size_t)H * 4);
    const size_t o_hpose = off; off += al256(ho->hyp_poses ? CH * (size_t)H * 48 : 16);
    if (off > c->buf_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (c->buf[i]) cudaFree(c->buf[i]);
            c->buf[i] = nullptr;
            RDPN_CUDA_TRY(cudaMalloc(&c->buf[i], off));
        }
        c->buf_bytes = off;
    }
    const int B = h->B;
    int rc = 0;
    for (int b0 = 0, stage = 0; b0 < B && rc == 0; b0 += RDPN_CHUNK, stage ^= 1) {
        const size_t nb = (size_t)((B - b0) < RDPN_CHUNK ? (B - b0) : RDPN_CHUNK);
        cudaStream_t st = c->st[stage];
        unsigned char* d = c->buf[stage];
        const size_t po = (size_t)b0 * P;
#define H2D(dst_off, src, bytes) RDPN_CUDA_TRY(cudaMemcpyAsync(d + (dst_off), (src), (bytes), cudaMemcpyHostToDevice, st))
        H2D(o_depth, h->depth + po, nb * P * 4);
        H2D(o_cx, h->coor_x + po, nb * P * 4);
        H2D(o_cy, h->coor_y + po, nb * P * 4);
        H2D(o_cz, h->coor_z + po, nb * P * 4);
        H2D(o_mask, h->mask + po, nb * P * 4);
        if (!dense) {
            H2D(o_rid, h->region_idx + po, nb * P);
            H2D(o_anc, h->anchors + (size_t)b0 * R * 3, nb * R * 12);
        }
        H2D(o_kp, h->Kp + (size_t)b0 * 4, nb * 16);
        H2D(o_ext, h->extent + (size_t)b0 * 3, nb * 12);
        if (h->depth_div) H2D(o_div, h->depth_div + b0, nb * 4);
        H2D(o_hyp, h_hyp + (size_t)b0 * H * 3, nb * H * 12);
        if (h_tnet) H2D(o_tnet, h_tnet + (size_t)b0 * 3, nb * 12);
#undef H2D
        rdpn_roi_inputs di = *h;
        di.B = (int)nb;
        di.depth = (const float*)(d + o_depth);
        di.coor_x = (const float*)(d + o_cx);
        di.coor_y = (const float*)(d + o_cy);
        di.coor_z = (const float*)(d + o_cz);
        di.mask = (const float*)(d + o_mask);
        di.region_idx = dense ? nullptr : (const uint8_t*)(d + o_rid);
        di.anchors = dense ? nullptr : (const float*)(d + o_anc);
        di.Kp = (const float*)(d + o_kp);
        di.extent = (const float*)(d + o_ext);
        di.depth_div = h->depth_div ? (const float*)(d + o_div) : nullptr;
        rdpn_solve_outputs dout;
        memset(&dout, 0, sizeof(dout));
        dout.pose = (float*)(d + o_pose);
        dout.n_inliers = (int32_t*)(d + o_ninl);
        dout.status = (int32_t*)(d + o_stat);
        dout.best_h = ho->best_h ? (int32_t*)(d + o_best) : nullptr;
        dout.n_sel = ho->n_sel ? (int32_t*)(d + o_nsel) : nullptr;
        dout.scale = ho->scale ? (float*)(d + o_scale) : nullptr;
        dout.inlier_mask = ho->inlier_mask ? (uint8_t*)(d + o_imask) : nullptr;
        dout.hyp_counts = ho->hyp_counts ? (int32_t*)(d + o_hcnt) : nullptr;
        dout.hyp_poses = ho->hyp_poses ? (float*)(d + o_hpose) : nullptr;
        rc = rdpn_pose_solve(&di, (const int32_t*)(d + o_hyp), h_tnet ? (const float*)(d + o_tnet) : nullptr, prm, &dout, st);
        if (rc) break;
#define D2H(dst, src_off, bytes) RDPN_CUDA_TRY(cudaMemcpyAsync((dst), d + (src_off), (bytes), cudaMemcpyDeviceToHost, st))
        D2H(ho->pose + (size_t)b0 * 12, o_pose, nb * 48);
        D2H(ho->n_inliers + b0, o_ninl, nb * 4);
        D2H(ho->status + b0, o_stat, nb * 4);
        if (ho->best_h) D2H(ho->best_h + b0, o_best, nb * 4);
        if (ho->n_sel) D2H(ho->n_sel + b0, o_nsel, nb * 4);
        if (ho->scale) D2H(ho->scale + b0, o_scale, nb * 4);
        if (ho->inlier_mask) D2H(ho->inlier_mask + po, o_imask, nb * P);
        if (ho->hyp_counts) D2H(ho->hyp_counts + (size_t)b0 * H, o_hcnt, nb * H * 4);
        if (ho->hyp_poses) D2H(ho->hyp_poses + (size_t)b0 * H * 12, o_hpose, nb * H * 48);
#undef D2H
    }
    cudaError_t e0 = cudaStreamSynchronize(c->st[0]);
    cudaError_t e1 = cudaStreamSynchronize(c->st[1]);
    if (rc) return rc;
    if (e0 != cudaSuccess) return (int)e0;
    if (e1 != cudaSuccess) return (int)e1;
    return 0;
}

}  // extern "C"
