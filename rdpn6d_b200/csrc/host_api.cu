// Host-facing side of the C ABI: version / error strings, the exact drop-in FPS symbols of the
// reference's cffi extension (HOST pointers, core/csrc/fps/src/ext.h:1-14), and the host-buffer
// pose-solve plugin call with a context that owns device scratch and pipelines chunks of ROIs over
// two streams so that host->device copies overlap the kernels.
#include "common.cuh"
#include "gate.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

namespace rdpn {
unsigned long long g_launch_count = 0;

// ------------------------------------------------------------------------------------------------
// Gated pull: the solver only ever reads depth / coor / region-id values of pixels whose MASK test
// passes (the gate is a conjunction, gdrn_evaluator.py:110-117), i.e. of ~10-15 % of a ROI.  When the
// caller's planes live in pinned (device-mapped) host memory, only the mask plane is copied by the copy
// engine; this kernel evaluates the mask test with the solver's own arithmetic (gate.cuh) and fetches the
// other planes over PCIe only for the G-quad groups (G x 16 bytes per plane) that contain a passing pixel,
// writing them to the device planes the solver reads.  Everything else in those planes is never looked at.
// One CTA per ROI, thread -> quads exactly as in pose_solve_kernel.
// ------------------------------------------------------------------------------------------------
struct PullArgs {
    const float* d_mask;       // device [nb,P] (copied)
    const float* h_depth;      // host-mapped planes of this chunk
    const float* h_cx;
    const float* h_cy;
    const float* h_cz;
    const uint8_t* h_rid;      // or nullptr (dense mode)
    float* d_depth;
    float* d_cx;
    float* d_cy;
    float* d_cz;
    uint8_t* d_rid;
    unsigned long long* pulled_quads;  // device counter (quads fetched per plane)
    // small per-ROI arrays fetched by the same kernel (all nullptr: they were copied instead)
    const int32_t* h_hyp;      // [nb,H,3]
    const float* h_anchors;    // [nb,R,3] or nullptr
    const float* h_kp;         // [nb,4]
    const float* h_ext;        // [nb,3]
    const float* h_div;        // [nb] or nullptr
    const float* h_tnet;       // [nb,3] or nullptr
    int32_t* d_hyp;
    float* d_anchors;
    float* d_kp;
    float* d_ext;
    float* d_div;
    float* d_tnet;
    int H3;                    // H * 3
    int R3;                    // R * 3
    int mask_mode;
    float mask_thr;
    double mask_cut;
    int mask_cut_incl;
};

template <int G>
__global__ void __launch_bounds__(256) pull_gated_kernel(PullArgs a) {
    constexpr int SW = 8, QPT = 4;
    __shared__ float red_f[2][SW];
    __shared__ RoiGate s_gate;
    __shared__ float s_mn;
    __shared__ int s_cnt[SW];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t po = (size_t)b * RDPN_P;
    const float4* mask4 = reinterpret_cast<const float4*>(a.d_mask + po);
    float4 m4[QPT];
#pragma unroll
    for (int k = 0; k < QPT; ++k) m4[k] = __ldg(mask4 + 32 * (SW * k + warp) + lane);
    if (a.mask_mode == RDPN_MASK_L1) {
        float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
        for (int k = 0; k < QPT; ++k) minmax4(m4[k], mn, mx);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) { red_f[0][warp] = mn; red_f[1][warp] = mx; }
        __syncthreads();
        if (t == 0) {
            float lo = red_f[0][0], hi = red_f[1][0];
            for (int w = 1; w < SW; ++w) { lo = fminf(lo, red_f[0][w]); hi = fmaxf(hi, red_f[1][w]); }
            s_mn = lo;
            make_gate(s_gate, lo, hi, a.mask_thr, a.mask_cut, a.mask_cut_incl);
        }
        __syncthreads();
    }
    const RoiGate gate = s_gate;
    const float mn = a.mask_mode == RDPN_MASK_L1 ? s_mn : 0.f;
    bool need[QPT];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const float mm[4] = {m4[k].x, m4[k].y, m4[k].z, m4[k].w};
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) any = any || mask_pass(mm[j], a.mask_mode, a.mask_thr, mn, gate);
        unsigned bal = __ballot_sync(0xffffffffu, any);
        // fetch granularity: groups of G consecutive quads (G * 16 bytes per plane)
        const unsigned gm = (G >= 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(G - 1));
        need[k] = (bal & gm) != 0u;
        cnt += need[k] ? 1 : 0;
    }
    // all loads first (PCIe latency is microseconds: keep every request of the thread in flight), then the stores
    float4 dq[QPT], xq[QPT], yq[QPT], zq[QPT];
    uchar4 rq[QPT];
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int q = 32 * (SW * k + warp) + lane;
        if (need[k]) {
            dq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_depth + po) + q);
            xq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_cx + po) + q);
            yq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_cy + po) + q);
            zq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_cz + po) + q);
            if (a.h_rid) rq[k] = __ldcs(reinterpret_cast<const uchar4*>(a.h_rid + po) + q);
        }
    }
    if (a.h_hyp) {  // hypothesis triplets, anchors and scalars ride along while the plane requests are in flight
        const int32_t* hs = a.h_hyp + (size_t)b * a.H3;
        int32_t* hd = a.d_hyp + (size_t)b * a.H3;
        for (int i = t; i < a.H3; i += 256) hd[i] = __ldcs(hs + i);
        if (a.h_anchors)
            for (int i = t; i < a.R3; i += 256) a.d_anchors[(size_t)b * a.R3 + i] = __ldcs(a.h_anchors + (size_t)b * a.R3 + i);
        if (t < 4) a.d_kp[4 * b + t] = __ldcs(a.h_kp + 4 * b + t);
        else if (t < 7) a.d_ext[3 * b + t - 4] = __ldcs(a.h_ext + 3 * b + t - 4);
        else if (t < 10) { if (a.h_tnet) a.d_tnet[3 * b + t - 7] = __ldcs(a.h_tnet + 3 * b + t - 7); }
        else if (t == 10) { if (a.h_div) a.d_div[b] = __ldcs(a.h_div + b); }
    }
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int q = 32 * (SW * k + warp) + lane;
        if (need[k]) {
            reinterpret_cast<float4*>(a.d_depth + po)[q] = dq[k];
            reinterpret_cast<float4*>(a.d_cx + po)[q] = xq[k];
            reinterpret_cast<float4*>(a.d_cy + po)[q] = yq[k];
            reinterpret_cast<float4*>(a.d_cz + po)[q] = zq[k];
            if (a.h_rid) reinterpret_cast<uchar4*>(a.d_rid + po)[q] = rq[k];
        }
    }
    if (a.pulled_quads) {
        cnt = warp_sum(cnt);
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        if (t == 0) {
            int tot = 0;
            for (int w = 0; w < SW; ++w) tot += s_cnt[w];
            atomicAdd(a.pulled_quads, (unsigned long long)tot);
        }
    }
}

static int launch_pull(const PullArgs& a, int nb, int gran, cudaStream_t st) {
    switch (gran) {
        case 1: pull_gated_kernel<1><<<nb, 256, 0, st>>>(a); break;
        case 2: pull_gated_kernel<2><<<nb, 256, 0, st>>>(a); break;
        case 4: pull_gated_kernel<4><<<nb, 256, 0, st>>>(a); break;
        case 8: pull_gated_kernel<8><<<nb, 256, 0, st>>>(a); break;
        case 16: pull_gated_kernel<16><<<nb, 256, 0, st>>>(a); break;
        default: return RDPN_E_BADARG;
    }
    ++g_launch_count;
    RDPN_LAUNCH_CHECK();
    return 0;
}

// pinned / registered host memory is addressable from the device under UVA
static bool device_can_read_host(const void* p) {
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost && at.devicePointer != nullptr;
}
static const void* device_view(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    cudaPointerGetAttributes(&at, p);
    return at.devicePointer;
}
}  // namespace rdpn

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
#define RDPN_CHUNK_MAX 1024
#define RDPN_STAGES 4  // pipeline depth: stage buffers / streams in flight

struct rdpn_ctx {
    int device;
    cudaStream_t st[RDPN_STAGES];
    unsigned char* buf[RDPN_STAGES];
    size_t buf_bytes;
    int transfer;   // RDPN_TRANSFER_*
    int gran;       // pull granularity in quads (16-byte units per plane)
    int chunk;      // ROIs per pipeline stage
    int count;      // count the bytes that crossed the bus (one extra 8-byte read-back per call)
    unsigned long long* d_pulled;
    unsigned long long pulled_seen;
    unsigned long long last_h2d_bytes;
    int last_transfer;
};

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int rdpn_ctx_create(int device, rdpn_ctx** out_ctx) {
    if (!out_ctx) return RDPN_E_BADARG;
    RDPN_CUDA_TRY(cudaSetDevice(device));
    rdpn_ctx* c = (rdpn_ctx*)calloc(1, sizeof(rdpn_ctx));
    c->device = device;
    for (int i = 0; i < RDPN_STAGES; ++i) RDPN_CUDA_TRY(cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking));
    RDPN_CUDA_TRY(cudaMalloc(&c->d_pulled, sizeof(unsigned long long)));
    RDPN_CUDA_TRY(cudaMemset(c->d_pulled, 0, sizeof(unsigned long long)));
    c->transfer = RDPN_TRANSFER_AUTO;
    c->gran = env_int("RDPN_PULL_GRAN", 2);
    c->chunk = env_int("RDPN_HOST_CHUNK", 256);
    *out_ctx = c;
    return 0;
}

void rdpn_ctx_destroy(rdpn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < RDPN_STAGES; ++i) {
        if (c->buf[i]) cudaFree(c->buf[i]);
        if (c->st[i]) cudaStreamDestroy(c->st[i]);
    }
    if (c->d_pulled) cudaFree(c->d_pulled);
    free(c);
}

int rdpn_ctx_set_option(rdpn_ctx* c, int key, int value) {
    if (!c) return RDPN_E_BADARG;
    switch (key) {
        case RDPN_OPT_TRANSFER:
            if (value < RDPN_TRANSFER_AUTO || value > RDPN_TRANSFER_PULL) return RDPN_E_BADARG;
            c->transfer = value;
            return 0;
        case RDPN_OPT_PULL_GRANULARITY:
            if (value != 1 && value != 2 && value != 4 && value != 8 && value != 16) return RDPN_E_BADARG;
            c->gran = value;
            return 0;
        case RDPN_OPT_CHUNK_ROIS:
            if (value < 1 || value > RDPN_CHUNK_MAX) return RDPN_E_BADARG;
            c->chunk = value;
            return 0;
        case RDPN_OPT_COUNT_BYTES:
            c->count = value != 0;
            return 0;
        default:
            return RDPN_E_BADARG;
    }
}

unsigned long long rdpn_ctx_last_h2d_bytes(const rdpn_ctx* c) { return c ? c->last_h2d_bytes : 0ull; }
int rdpn_ctx_last_transfer(const rdpn_ctx* c) { return c ? c->last_transfer : 0; }

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int rdpn_pose_solve_host(rdpn_ctx* c, const rdpn_roi_inputs* h, const int32_t* h_hyp, const float* h_tnet,
                         const rdpn_solve_params* prm, const rdpn_solve_outputs* ho) {
    if (!c || !h || !h_hyp || !prm || !ho || h->B <= 0 || prm->num_hyp <= 0) return RDPN_E_BADARG;
    if (!ho->pose || !ho->n_inliers || !ho->status) return RDPN_E_BADARG;
    if (!h->depth || !h->coor_x || !h->coor_y || !h->coor_z || !h->mask || !h->Kp || !h->extent) return RDPN_E_BADARG;
    if ((h->region_idx == nullptr) != (h->anchors == nullptr)) return RDPN_E_BADARG;
    RDPN_CUDA_TRY(cudaSetDevice(c->device));
    const bool dense = h->region_idx == nullptr;
    const int H = prm->num_hyp, R = dense ? 0 : h->num_regions;
    const size_t P = RDPN_P, CH = (size_t)c->chunk;
    // transfer strategy: gated pull needs the big planes in device-mapped (pinned / registered) host memory
    const bool mapped = rdpn::device_can_read_host(h->depth) && rdpn::device_can_read_host(h->coor_x) &&
                        rdpn::device_can_read_host(h->coor_y) && rdpn::device_can_read_host(h->coor_z) &&
                        rdpn::device_can_read_host(h->region_idx);
    if (c->transfer == RDPN_TRANSFER_PULL && !mapped) return RDPN_E_BADARG;
    const bool pull = c->transfer == RDPN_TRANSFER_PULL || (c->transfer == RDPN_TRANSFER_AUTO && mapped);
    const bool aligned = !(((uintptr_t)h->depth | (uintptr_t)h->coor_x | (uintptr_t)h->coor_y | (uintptr_t)h->coor_z |
                            (uintptr_t)h->region_idx) & 15);
    if (pull && !aligned) return RDPN_E_ALIGN;
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
        for (int i = 0; i < RDPN_STAGES; ++i) {
            if (c->buf[i]) cudaFree(c->buf[i]);
            c->buf[i] = nullptr;
            RDPN_CUDA_TRY(cudaMalloc(&c->buf[i], off));
            RDPN_CUDA_TRY(cudaMemset(c->buf[i], 0, off));  // planes the gated pull never writes stay defined
        }
        c->buf_bytes = off;
    }
    const int B = h->B;
    int rc = 0;
    unsigned long long copied = 0;
    const float* hv_depth = pull ? (const float*)rdpn::device_view(h->depth) : nullptr;
    const float* hv_cx = pull ? (const float*)rdpn::device_view(h->coor_x) : nullptr;
    const float* hv_cy = pull ? (const float*)rdpn::device_view(h->coor_y) : nullptr;
    const float* hv_cz = pull ? (const float*)rdpn::device_view(h->coor_z) : nullptr;
    const uint8_t* hv_rid = pull && !dense ? (const uint8_t*)rdpn::device_view(h->region_idx) : nullptr;
    // the small per-ROI arrays ride along in the pull kernel when they are mapped too (saves 6 copies per chunk)
    const bool pull_small = pull && rdpn::device_can_read_host(h_hyp) && rdpn::device_can_read_host(h->anchors) &&
                            rdpn::device_can_read_host(h->Kp) && rdpn::device_can_read_host(h->extent) &&
                            rdpn::device_can_read_host(h->depth_div) && rdpn::device_can_read_host(h_tnet);
    const int32_t* hv_hyp = pull_small ? (const int32_t*)rdpn::device_view(h_hyp) : nullptr;
    const float* hv_anc = pull_small ? (const float*)rdpn::device_view(h->anchors) : nullptr;
    const float* hv_kp = pull_small ? (const float*)rdpn::device_view(h->Kp) : nullptr;
    const float* hv_ext = pull_small ? (const float*)rdpn::device_view(h->extent) : nullptr;
    const float* hv_div = pull_small ? (const float*)rdpn::device_view(h->depth_div) : nullptr;
    const float* hv_tnet = pull_small ? (const float*)rdpn::device_view(h_tnet) : nullptr;
    // results are written straight into the caller's buffers when those are mapped (no device -> host copies)
    const bool direct_out = pull && rdpn::device_can_read_host(ho->pose) && rdpn::device_can_read_host(ho->n_inliers) &&
                            rdpn::device_can_read_host(ho->status) && rdpn::device_can_read_host(ho->best_h) &&
                            rdpn::device_can_read_host(ho->n_sel) && rdpn::device_can_read_host(ho->scale) &&
                            !ho->inlier_mask && !ho->hyp_counts && !ho->hyp_poses;
    for (int b0 = 0, stage = 0; b0 < B && rc == 0; b0 += c->chunk, stage = (stage + 1) % RDPN_STAGES) {
        const size_t nb = (size_t)((B - b0) < c->chunk ? (B - b0) : c->chunk);
        cudaStream_t st = c->st[stage];
        unsigned char* d = c->buf[stage];
        const size_t po = (size_t)b0 * P;
#define H2D(dst_off, src, bytes)                                                                           \
    do {                                                                                                   \
        RDPN_CUDA_TRY(cudaMemcpyAsync(d + (dst_off), (src), (bytes), cudaMemcpyHostToDevice, st));         \
        copied += (bytes);                                                                                 \
    } while (0)
        H2D(o_mask, h->mask + po, nb * P * 4);
        if (!pull) {
            H2D(o_depth, h->depth + po, nb * P * 4);
            H2D(o_cx, h->coor_x + po, nb * P * 4);
            H2D(o_cy, h->coor_y + po, nb * P * 4);
            H2D(o_cz, h->coor_z + po, nb * P * 4);
            if (!dense) H2D(o_rid, h->region_idx + po, nb * P);
        }
        if (!pull_small) {
            if (!dense) H2D(o_anc, h->anchors + (size_t)b0 * R * 3, nb * R * 12);
            H2D(o_kp, h->Kp + (size_t)b0 * 4, nb * 16);
            H2D(o_ext, h->extent + (size_t)b0 * 3, nb * 12);
            if (h->depth_div) H2D(o_div, h->depth_div + b0, nb * 4);
            H2D(o_hyp, h_hyp + (size_t)b0 * H * 3, nb * H * 12);
            if (h_tnet) H2D(o_tnet, h_tnet + (size_t)b0 * 3, nb * 12);
        } else {
            copied += nb * ((size_t)R * 12 + 16 + 12 + (h->depth_div ? 4 : 0) + (size_t)H * 12 + (h_tnet ? 12 : 0));
        }
#undef H2D
        if (pull) {
            rdpn::PullArgs pa;
            pa.d_mask = (const float*)(d + o_mask);
            pa.h_depth = hv_depth + po;
            pa.h_cx = hv_cx + po;
            pa.h_cy = hv_cy + po;
            pa.h_cz = hv_cz + po;
            pa.h_rid = dense ? nullptr : hv_rid + po;
            pa.d_depth = (float*)(d + o_depth);
            pa.d_cx = (float*)(d + o_cx);
            pa.d_cy = (float*)(d + o_cy);
            pa.d_cz = (float*)(d + o_cz);
            pa.d_rid = (uint8_t*)(d + o_rid);
            pa.pulled_quads = c->count ? c->d_pulled : nullptr;
            memset(&pa.h_hyp, 0, (char*)&pa.H3 - (char*)&pa.h_hyp);
            if (pull_small) {
                pa.h_hyp = hv_hyp + (size_t)b0 * H * 3;
                pa.h_anchors = dense ? nullptr : hv_anc + (size_t)b0 * R * 3;
                pa.h_kp = hv_kp + (size_t)b0 * 4;
                pa.h_ext = hv_ext + (size_t)b0 * 3;
                pa.h_div = h->depth_div ? hv_div + b0 : nullptr;
                pa.h_tnet = h_tnet ? hv_tnet + (size_t)b0 * 3 : nullptr;
                pa.d_hyp = (int32_t*)(d + o_hyp);
                pa.d_anchors = (float*)(d + o_anc);
                pa.d_kp = (float*)(d + o_kp);
                pa.d_ext = (float*)(d + o_ext);
                pa.d_div = (float*)(d + o_div);
                pa.d_tnet = (float*)(d + o_tnet);
            }
            pa.H3 = H * 3;
            pa.R3 = R * 3;
            pa.mask_mode = h->mask_mode;
            pa.mask_thr = h->mask_thr;
            rdpn::host_mask_cut(h->mask_thr, &pa.mask_cut, &pa.mask_cut_incl);
            rc = rdpn::launch_pull(pa, (int)nb, c->gran, st);
            if (rc) break;
        }
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
        if (direct_out) {
            dout.pose = (float*)rdpn::device_view(ho->pose) + (size_t)b0 * 12;
            dout.n_inliers = (int32_t*)rdpn::device_view(ho->n_inliers) + b0;
            dout.status = (int32_t*)rdpn::device_view(ho->status) + b0;
            dout.best_h = ho->best_h ? (int32_t*)rdpn::device_view(ho->best_h) + b0 : nullptr;
            dout.n_sel = ho->n_sel ? (int32_t*)rdpn::device_view(ho->n_sel) + b0 : nullptr;
            dout.scale = ho->scale ? (float*)rdpn::device_view(ho->scale) + b0 : nullptr;
        }
        rc = rdpn_pose_solve(&di, (const int32_t*)(d + o_hyp), h_tnet ? (const float*)(d + o_tnet) : nullptr, prm, &dout, st);
        if (rc) break;
        if (direct_out) continue;
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
    cudaError_t es = cudaSuccess;
    for (int i = 0; i < RDPN_STAGES; ++i) {
        const cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (es == cudaSuccess) es = e;
    }
    if (rc) return rc;
    if (es != cudaSuccess) return (int)es;
    c->last_transfer = pull ? RDPN_TRANSFER_PULL : RDPN_TRANSFER_COPY;
    c->last_h2d_bytes = copied;
    if (pull && c->count) {
        unsigned long long tot = 0;
        RDPN_CUDA_TRY(cudaMemcpy(&tot, c->d_pulled, sizeof(tot), cudaMemcpyDeviceToHost));
        const unsigned long long quads = tot - c->pulled_seen;
        c->pulled_seen = tot;
        c->last_h2d_bytes += quads * (unsigned long long)(4 * 16 + (dense ? 0 : 4));
    }
    return 0;
}

}  // extern "C"
