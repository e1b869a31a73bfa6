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

#include <mutex>
#include <vector>

namespace rdpn {
unsigned long long g_launch_count = 0;

// ------------------------------------------------------------------------------------------------
// Per-device launch state.  Function attributes (opt-in dynamic shared memory) and the SM count belong to a device,
// not to the process: every launcher asks here, keyed by the CURRENT device, under a mutex.
// ------------------------------------------------------------------------------------------------
#define RDPN_MAX_DEVICES 64
#define RDPN_ATTR_SLOTS 32
struct DevState {
    int sms;
    size_t attr[RDPN_ATTR_SLOTS];
};
static std::mutex g_dev_mu;
static DevState g_dev[RDPN_MAX_DEVICES];

int device_sm_count(int* sms) {
    int dev = 0;
    RDPN_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= RDPN_MAX_DEVICES) return RDPN_E_BADARG;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev[dev].sms) RDPN_CUDA_TRY(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
    *sms = g_dev[dev].sms;
    return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize >= bytes for `func` on the current device (slot: one per kernel instantiation)
int ensure_func_smem(const void* func, int slot, size_t bytes) {
    int dev = 0;
    RDPN_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= RDPN_MAX_DEVICES || slot < 0 || slot >= RDPN_ATTR_SLOTS) return RDPN_E_BADARG;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (bytes > g_dev[dev].attr[slot]) {
        RDPN_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        g_dev[dev].attr[slot] = bytes;
    }
    return 0;
}

// Workspace of the plain rdpn_pose_solve entry (no workspace argument): one buffer per (device, stream), grown on
// demand.  Growing synchronises that stream once (the old buffer may be in use); steady state allocates nothing.
struct WsEntry {
    int dev;
    cudaStream_t st;
    void* p;
    size_t bytes;
};
static std::mutex g_ws_mu;
static std::vector<WsEntry> g_ws;

int cached_workspace(cudaStream_t st, size_t need, void** out, size_t* out_bytes) {
    int dev = 0;
    RDPN_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ws_mu);
    WsEntry* e = nullptr;
    for (auto& w : g_ws)
        if (w.dev == dev && w.st == st) e = &w;
    if (!e) {
        g_ws.push_back(WsEntry{dev, st, nullptr, 0});
        e = &g_ws.back();
    }
    if (e->bytes < need) {
        if (e->p) {
            RDPN_CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(e->p);
            e->p = nullptr;
            e->bytes = 0;
        }
        RDPN_CUDA_TRY(cudaMalloc(&e->p, need));
        e->bytes = need;
    }
    *out = e->p;
    *out_bytes = e->bytes;
    return 0;
}

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
    const int32_t* h_hyp;      // [nb,H,S]
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
    int H3;                    // H * S
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
        if (need[k]) {  // each plane is optional: planes that already live on the device are used in place
            if (a.h_depth) dq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_depth + po) + q);
            if (a.h_cx) xq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_cx + po) + q);
            if (a.h_cy) yq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_cy + po) + q);
            if (a.h_cz) zq[k] = __ldcs(reinterpret_cast<const float4*>(a.h_cz + po) + q);
            if (a.h_rid) rq[k] = __ldcs(reinterpret_cast<const uchar4*>(a.h_rid + po) + q);
        }
    }
    // hypothesis triplets, anchors and scalars ride along while the plane requests are in flight (each optional)
    if (a.h_hyp) {
        const int32_t* hs = a.h_hyp + (size_t)b * a.H3;
        int32_t* hd = a.d_hyp + (size_t)b * a.H3;
        if (((a.H3 & 3) | (int)(((uintptr_t)hs | (uintptr_t)hd) & 15)) == 0) {  // 16 bytes per request
            for (int i = t; i < a.H3 / 4; i += 256) reinterpret_cast<int4*>(hd)[i] = __ldcs(reinterpret_cast<const int4*>(hs) + i);
        } else {
            for (int i = t; i < a.H3; i += 256) hd[i] = __ldcs(hs + i);
        }
    }
    if (a.h_anchors) {
        const float* as = a.h_anchors + (size_t)b * a.R3;
        float* ad = a.d_anchors + (size_t)b * a.R3;
        if (((a.R3 & 3) | (int)(((uintptr_t)as | (uintptr_t)ad) & 15)) == 0) {
            for (int i = t; i < a.R3 / 4; i += 256) reinterpret_cast<float4*>(ad)[i] = __ldcs(reinterpret_cast<const float4*>(as) + i);
        } else {
            for (int i = t; i < a.R3; i += 256) ad[i] = __ldcs(as + i);
        }
    }
    if (t < 4) { if (a.h_kp) a.d_kp[4 * b + t] = __ldcs(a.h_kp + 4 * b + t); }
    else if (t < 7) { if (a.h_ext) a.d_ext[3 * b + t - 4] = __ldcs(a.h_ext + 3 * b + t - 4); }
    else if (t < 10) { if (a.h_tnet) a.d_tnet[3 * b + t - 7] = __ldcs(a.h_tnet + 3 * b + t - 7); }
    else if (t == 10) { if (a.h_div) a.d_div[b] = __ldcs(a.h_div + b); }
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int q = 32 * (SW * k + warp) + lane;
        if (need[k]) {
            if (a.h_depth) reinterpret_cast<float4*>(a.d_depth + po)[q] = dq[k];
            if (a.h_cx) reinterpret_cast<float4*>(a.d_cx + po)[q] = xq[k];
            if (a.h_cy) reinterpret_cast<float4*>(a.d_cy + po)[q] = yq[k];
            if (a.h_cz) reinterpret_cast<float4*>(a.d_cz + po)[q] = zq[k];
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

// Where a caller's buffer lives decides how it reaches the solver.
enum Loc { LOC_NULL = 0, LOC_DEVICE, LOC_MAPPED, LOC_PAGEABLE };
enum Move { MOVE_NONE = 0, MOVE_IN_PLACE, MOVE_PULL, MOVE_COPY };
struct Buf {
    const void* host;  // the caller's pointer
    const void* dev;   // device-usable address (device memory, or pinned / registered host memory under UVA)
    Loc loc;
    Move move;
};
static Buf classify(const void* p) {
    Buf b = {p, nullptr, LOC_NULL, MOVE_NONE};
    if (!p) return b;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        b.loc = LOC_PAGEABLE;
    } else if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
        b.loc = LOC_DEVICE;
        b.dev = at.devicePointer ? at.devicePointer : p;
    } else if (at.type == cudaMemoryTypeHost && at.devicePointer) {
        b.loc = LOC_MAPPED;
        b.dev = at.devicePointer;
    } else {
        b.loc = LOC_PAGEABLE;
    }
    return b;
}
static void plan(Buf& b, bool may_pull) {
    b.move = b.loc == LOC_NULL ? MOVE_NONE
             : b.loc == LOC_DEVICE ? MOVE_IN_PLACE
             : (b.loc == LOC_MAPPED && may_pull) ? MOVE_PULL : MOVE_COPY;
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
    // clouds beyond the register-resident limit (1.21 M points) stream their running minima from the workspace: pn
    // more floats behind the header (rdpn_fps_init_center's contract)
    const size_t ws = rdpn_fps_workspace_bytes(sn) + (pn > 1000000 ? (size_t)pn * sizeof(float) : 0);
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
    RDPN_NVTX("farthest_point_sampling");
    if (pn <= 0) return;
    srand((unsigned)time(0));  // farthest_point_sampling.cpp:93-94
    fps_host(pts, idxs, pn, sn, rand() % pn);
}

// ------------------------------------------------------------------------------------------------
// host-buffer pose solve
// ------------------------------------------------------------------------------------------------
#define RDPN_CHUNK_MAX 1024
#define RDPN_STAGES 4  // pipeline depth: stage buffers / streams in flight
#define RDPN_MAX_INFLIGHT 8  // submitted calls that may be outstanding (ring of completion events)

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
    cudaEvent_t ev[RDPN_MAX_INFLIGHT][RDPN_STAGES];  // completion of submitted calls (created on first use)
    unsigned next_ticket;
    unsigned counted_ticket;  // submitted calls whose pulled quads are already in pulled_seen
};

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// The context's entry points run on the context's device and leave the CALLER's current device as they found it
// (PyTorch and other users of the runtime API keep a per-thread current device).
struct DeviceGuard {
    int prev;
    cudaError_t err;
    explicit DeviceGuard(int device) : prev(-1), err(cudaSuccess) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) err = cudaSetDevice(device);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int rdpn_ctx_create(int device, rdpn_ctx** out_ctx) {
    RDPN_NVTX("rdpn_ctx_create");
    if (!out_ctx) return RDPN_E_BADARG;
    DeviceGuard guard(device);
    RDPN_CUDA_TRY(guard.err);
    rdpn_ctx* c = (rdpn_ctx*)calloc(1, sizeof(rdpn_ctx));
    c->device = device;
    for (int i = 0; i < RDPN_STAGES; ++i) RDPN_CUDA_TRY(cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking));
    RDPN_CUDA_TRY(cudaMalloc(&c->d_pulled, sizeof(unsigned long long)));
    RDPN_CUDA_TRY(cudaMemset(c->d_pulled, 0, sizeof(unsigned long long)));
    c->transfer = RDPN_TRANSFER_AUTO;
    c->gran = env_int("RDPN_PULL_GRAN", 2);
    c->chunk = env_int("RDPN_HOST_CHUNK", 0);  // 0 = by call size (host_queue)
    *out_ctx = c;
    return 0;
}

void rdpn_ctx_destroy(rdpn_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    for (int i = 0; i < RDPN_STAGES; ++i) {
        if (c->buf[i]) cudaFree(c->buf[i]);
        if (c->st[i]) cudaStreamDestroy(c->st[i]);
    }
    if (c->d_pulled) cudaFree(c->d_pulled);
    for (int t = 0; t < RDPN_MAX_INFLIGHT; ++t)
        for (int i = 0; i < RDPN_STAGES; ++i)
            if (c->ev[t][i]) cudaEventDestroy(c->ev[t][i]);
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

struct HostCallInfo {  // what the synchronous form reports after the call
    unsigned long long moved;     // bytes of copied tensors and whole small arrays
    unsigned long long per_quad;  // bytes per fetched 16-byte quad group of the gated pull
    bool any_pull, pull_planes;
};

// Queues the whole call (transfers, gated pull, solver, result copies) on the context's streams; does not wait.
static int host_queue(rdpn_ctx* c, const rdpn_roi_inputs* h, const int32_t* h_hyp, const float* h_tnet,
                      const rdpn_solve_params* prm, const rdpn_solve_outputs* ho, HostCallInfo* info) {
    using namespace rdpn;
    if (!c || !h || !prm || !ho || h->B <= 0 || prm->num_hyp <= 0) return RDPN_E_BADARG;  // h_hyp NULL: internal sampling
    if (!ho->pose || !ho->n_inliers || !ho->status) return RDPN_E_BADARG;
    if (!h->depth || !h->coor_x || !h->coor_y || !h->coor_z || !h->mask || !h->Kp || !h->extent) return RDPN_E_BADARG;
    if ((h->region_idx == nullptr) != (h->anchors == nullptr)) return RDPN_E_BADARG;
    DeviceGuard guard(c->device);
    RDPN_CUDA_TRY(guard.err);
    const bool dense = h->region_idx == nullptr;
    const int H = prm->num_hyp, R = dense ? 0 : h->num_regions;
    if (prm->sample_size != 0 && (prm->sample_size < 3 || prm->sample_size > RDPN_MAX_SAMPLE)) return RDPN_E_BADARG;
    const int SS = prm->sample_size ? prm->sample_size : 3;  // pairs per hypothesis
    // ROIs per pipeline chunk: measured on B200 (benchmarks/host_path.py) 256 is best for calls of ~1024 ROIs (four chunks
    // keep the four stages busy), 512 for calls of 4096 and more (+2 %)
    const int chunk = c->chunk > 0 ? c->chunk : (h->B >= 4096 ? 512 : 256);
    const size_t P = RDPN_P, CH = (size_t)chunk;
    const bool may_pull = c->transfer != RDPN_TRANSFER_COPY;

    // every buffer is classified on its own: device memory is used in place, pinned memory is pulled (planes: only
    // where the mask passes) or written directly (results), pageable memory goes through the stage buffers
    enum { I_DEPTH, I_CX, I_CY, I_CZ, I_RID, I_MASK, I_ANC, I_KP, I_EXT, I_DIV, I_HYP, I_TNET, NIN };
    Buf in[NIN] = {classify(h->depth), classify(h->coor_x), classify(h->coor_y), classify(h->coor_z),
                   classify(h->region_idx), classify(h->mask), classify(h->anchors), classify(h->Kp),
                   classify(h->extent), classify(h->depth_div), classify(h_hyp), classify(h_tnet)};
    for (int i = 0; i < NIN; ++i) plan(in[i], may_pull);
    if (in[I_MASK].move == MOVE_PULL) in[I_MASK].move = MOVE_COPY;  // the whole mask plane is needed (min / max)
    bool any_pull = false, pull_planes = false;
    for (int i = 0; i < NIN; ++i) any_pull = any_pull || in[i].move == MOVE_PULL;
    for (int i = I_DEPTH; i <= I_RID; ++i) {
        pull_planes = pull_planes || in[i].move == MOVE_PULL;
        if (c->transfer == RDPN_TRANSFER_PULL && in[i].loc == LOC_PAGEABLE) return RDPN_E_BADARG;
        if (in[i].move != MOVE_COPY && ((uintptr_t)in[i].dev & 15)) return RDPN_E_ALIGN;
    }
    enum { O_POSE, O_NINL, O_STAT, O_BEST, O_NSEL, O_SCALE, O_IMASK, O_HCNT, O_HPOSE, NOUT };
    Buf out[NOUT] = {classify(ho->pose), classify(ho->n_inliers), classify(ho->status), classify(ho->best_h),
                     classify(ho->n_sel), classify(ho->scale), classify(ho->inlier_mask), classify(ho->hyp_counts),
                     classify(ho->hyp_poses)};
    for (int i = 0; i < NOUT; ++i) {
        plan(out[i], may_pull && i <= O_SCALE);  // MOVE_PULL on an output = the kernel writes it directly
        if (i == O_IMASK && out[i].move == MOVE_IN_PLACE && ((uintptr_t)out[i].dev & 15)) return RDPN_E_ALIGN;
    }

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
    const size_t o_hyp = off; off += al256(CH * (size_t)H * SS * 4);
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
        for (int i = 0; i < RDPN_STAGES; ++i) RDPN_CUDA_TRY(cudaStreamSynchronize(c->st[i]));  // submitted calls may use them
        for (int i = 0; i < RDPN_STAGES; ++i) {
            if (c->buf[i]) cudaFree(c->buf[i]);
            c->buf[i] = nullptr;
            RDPN_CUDA_TRY(cudaMalloc(&c->buf[i], off));
            RDPN_CUDA_TRY(cudaMemset(c->buf[i], 0, off));  // planes the gated pull never writes stay defined
        }
        c->buf_bytes = off;
    }
    const size_t in_off[NIN] = {o_depth, o_cx, o_cy, o_cz, o_rid, o_mask, o_anc, o_kp, o_ext, o_div, o_hyp, o_tnet};
    const size_t in_roi_bytes[NIN] = {P * 4, P * 4, P * 4, P * 4, P, P * 4, (size_t)R * 12, 16, 12, 4, (size_t)H * SS * 4, 12};
    const size_t out_off[NOUT] = {o_pose, o_ninl, o_stat, o_best, o_nsel, o_scale, o_imask, o_hcnt, o_hpose};
    const size_t out_roi_bytes[NOUT] = {48, 4, 4, 4, 4, 4, P, (size_t)H * 4, (size_t)H * 48};

    const int B = h->B;
    int rc = 0;
    unsigned long long moved = 0;
    // bulk copies of many tensors queue best over two streams (more only interleaves them on the copy engines); the
    // pull pipeline (one bulk copy + two kernels per chunk) wants all four stages
    const int nstages = any_pull ? RDPN_STAGES : 2;
    for (int b0 = 0, stage = 0; b0 < B && rc == 0; b0 += chunk, stage = (stage + 1) % nstages) {
        const size_t nb = (size_t)((B - b0) < chunk ? (B - b0) : chunk);
        cudaStream_t st = c->st[stage];
        unsigned char* d = c->buf[stage];
        // where the solver finds input i of this chunk, and the copies
        const void* src[NIN];
        for (int i = 0; i < NIN; ++i) {
            const size_t skip = (size_t)b0 * in_roi_bytes[i];
            if (in[i].move == MOVE_IN_PLACE) {
                src[i] = (const unsigned char*)in[i].dev + skip;
            } else if (in[i].move == MOVE_NONE) {
                src[i] = nullptr;
            } else {
                src[i] = d + in_off[i];
                if (in[i].move == MOVE_COPY) {
                    RDPN_CUDA_TRY(cudaMemcpyAsync(d + in_off[i], (const unsigned char*)in[i].host + skip, nb * in_roi_bytes[i],
                                                  cudaMemcpyHostToDevice, st));
                    moved += nb * in_roi_bytes[i];
                } else if (i > I_RID) {
                    moved += nb * in_roi_bytes[i];  // small array fetched whole by the pull kernel
                }
            }
        }
        if (any_pull) {
            PullArgs pa;
            memset(&pa, 0, sizeof(pa));
            auto hsrc = [&](int i) -> const void* {
                return in[i].move == MOVE_PULL ? (const unsigned char*)in[i].dev + (size_t)b0 * in_roi_bytes[i] : nullptr;
            };
            pa.d_mask = (const float*)src[I_MASK];
            pa.h_depth = (const float*)hsrc(I_DEPTH);
            pa.h_cx = (const float*)hsrc(I_CX);
            pa.h_cy = (const float*)hsrc(I_CY);
            pa.h_cz = (const float*)hsrc(I_CZ);
            pa.h_rid = (const uint8_t*)hsrc(I_RID);
            pa.d_depth = (float*)(d + o_depth);
            pa.d_cx = (float*)(d + o_cx);
            pa.d_cy = (float*)(d + o_cy);
            pa.d_cz = (float*)(d + o_cz);
            pa.d_rid = (uint8_t*)(d + o_rid);
            pa.pulled_quads = (c->count && pull_planes) ? c->d_pulled : nullptr;
            pa.h_hyp = (const int32_t*)hsrc(I_HYP);
            pa.h_anchors = (const float*)hsrc(I_ANC);
            pa.h_kp = (const float*)hsrc(I_KP);
            pa.h_ext = (const float*)hsrc(I_EXT);
            pa.h_div = (const float*)hsrc(I_DIV);
            pa.h_tnet = (const float*)hsrc(I_TNET);
            pa.d_hyp = (int32_t*)(d + o_hyp);
            pa.d_anchors = (float*)(d + o_anc);
            pa.d_kp = (float*)(d + o_kp);
            pa.d_ext = (float*)(d + o_ext);
            pa.d_div = (float*)(d + o_div);
            pa.d_tnet = (float*)(d + o_tnet);
            pa.H3 = H * SS;
            pa.R3 = R * 3;
            pa.mask_mode = h->mask_mode;
            pa.mask_thr = h->mask_thr;
            host_mask_cut(h->mask_thr, &pa.mask_cut, &pa.mask_cut_incl);
            rc = launch_pull(pa, (int)nb, c->gran, st);
            if (rc) break;
        }
        rdpn_roi_inputs di = *h;
        di.B = (int)nb;
        di.depth = (const float*)src[I_DEPTH];
        di.coor_x = (const float*)src[I_CX];
        di.coor_y = (const float*)src[I_CY];
        di.coor_z = (const float*)src[I_CZ];
        di.mask = (const float*)src[I_MASK];
        di.region_idx = (const uint8_t*)src[I_RID];
        di.anchors = (const float*)src[I_ANC];
        di.Kp = (const float*)src[I_KP];
        di.extent = (const float*)src[I_EXT];
        di.depth_div = (const float*)src[I_DIV];
        // outputs: in place (device), direct (pinned) or staged + copied back
        void* dst[NOUT];
        for (int i = 0; i < NOUT; ++i) {
            const size_t skip = (size_t)b0 * out_roi_bytes[i];
            dst[i] = out[i].move == MOVE_NONE ? nullptr
                     : (out[i].move == MOVE_IN_PLACE || out[i].move == MOVE_PULL) ? (void*)((unsigned char*)out[i].dev + skip)
                     : (void*)(d + out_off[i]);
        }
        rdpn_solve_outputs dout;
        memset(&dout, 0, sizeof(dout));
        dout.pose = (float*)dst[O_POSE];
        dout.n_inliers = (int32_t*)dst[O_NINL];
        dout.status = (int32_t*)dst[O_STAT];
        dout.best_h = (int32_t*)dst[O_BEST];
        dout.n_sel = (int32_t*)dst[O_NSEL];
        dout.scale = (float*)dst[O_SCALE];
        dout.inlier_mask = (uint8_t*)dst[O_IMASK];
        dout.hyp_counts = (int32_t*)dst[O_HCNT];
        dout.hyp_poses = (float*)dst[O_HPOSE];
        rdpn_solve_params pc = *prm;
        pc.roi_base = prm->roi_base + b0;  // internal sampling must not depend on the chunking
        rc = rdpn_pose_solve(&di, (const int32_t*)src[I_HYP], (const float*)src[I_TNET], &pc, &dout, st);
        if (rc) break;
        for (int i = 0; i < NOUT; ++i)
            if (out[i].move == MOVE_COPY)
                RDPN_CUDA_TRY(cudaMemcpyAsync((unsigned char*)out[i].host + (size_t)b0 * out_roi_bytes[i], d + out_off[i],
                                              nb * out_roi_bytes[i], cudaMemcpyDeviceToHost, st));
    }
    info->moved = moved;
    info->any_pull = any_pull;
    info->pull_planes = pull_planes;
    info->per_quad = 0;
    for (int i = I_DEPTH; i <= I_CZ; ++i) info->per_quad += in[i].move == MOVE_PULL ? 16 : 0;
    info->per_quad += in[I_RID].move == MOVE_PULL ? 4 : 0;
    return rc;
}

static int host_drain(rdpn_ctx* c) {
    cudaError_t es = cudaSuccess;
    for (int i = 0; i < RDPN_STAGES; ++i) {
        const cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (es == cudaSuccess) es = e;
    }
    return (int)es;
}

int rdpn_pose_solve_host(rdpn_ctx* c, const rdpn_roi_inputs* h, const int32_t* h_hyp, const float* h_tnet,
                         const rdpn_solve_params* prm, const rdpn_solve_outputs* ho) {
    RDPN_NVTX("rdpn_pose_solve_host");
    HostCallInfo info;
    memset(&info, 0, sizeof(info));
    if (c && c->count && c->counted_ticket != c->next_ticket) {  // submitted calls moved the counter: step over them
        const int e0 = host_drain(c);
        if (e0) return e0;
        RDPN_CUDA_TRY(cudaMemcpy(&c->pulled_seen, c->d_pulled, sizeof(c->pulled_seen), cudaMemcpyDeviceToHost));
        c->counted_ticket = c->next_ticket;
    }
    const int rc = host_queue(c, h, h_hyp, h_tnet, prm, ho, &info);
    if (rc == RDPN_E_BADARG || rc == RDPN_E_ALIGN) return rc;  // rejected before anything was queued
    const int es = host_drain(c);
    if (rc) return rc;
    if (es) return es;
    c->last_transfer = info.any_pull ? RDPN_TRANSFER_PULL : RDPN_TRANSFER_COPY;
    c->last_h2d_bytes = info.moved;
    if (info.pull_planes && c->count) {
        unsigned long long tot = 0;
        RDPN_CUDA_TRY(cudaMemcpy(&tot, c->d_pulled, sizeof(tot), cudaMemcpyDeviceToHost));
        const unsigned long long quads = tot - c->pulled_seen;
        c->pulled_seen = tot;
        c->last_h2d_bytes += quads * info.per_quad;
    }
    return 0;
}

int rdpn_pose_solve_host_submit(rdpn_ctx* c, const rdpn_roi_inputs* h, const int32_t* h_hyp, const float* h_tnet,
                                const rdpn_solve_params* prm, const rdpn_solve_outputs* ho, int* out_ticket) {
    RDPN_NVTX("rdpn_pose_solve_host_submit");
    if (!c || !out_ticket) return RDPN_E_BADARG;
    const int t = (int)(c->next_ticket % RDPN_MAX_INFLIGHT);
    DeviceGuard guard(c->device);
    RDPN_CUDA_TRY(guard.err);
    for (int i = 0; i < RDPN_STAGES; ++i) {
        if (!c->ev[t][i]) RDPN_CUDA_TRY(cudaEventCreateWithFlags(&c->ev[t][i], cudaEventDisableTiming));
        else RDPN_CUDA_TRY(cudaEventSynchronize(c->ev[t][i]));  // the ring is full: wait for its oldest call
    }
    HostCallInfo info;
    memset(&info, 0, sizeof(info));
    const int rc = host_queue(c, h, h_hyp, h_tnet, prm, ho, &info);
    if (rc) {
        if (rc != RDPN_E_BADARG && rc != RDPN_E_ALIGN) host_drain(c);
        return rc;
    }
    for (int i = 0; i < RDPN_STAGES; ++i) RDPN_CUDA_TRY(cudaEventRecord(c->ev[t][i], c->st[i]));
    c->last_transfer = info.any_pull ? RDPN_TRANSFER_PULL : RDPN_TRANSFER_COPY;
    c->next_ticket++;
    *out_ticket = t;
    return 0;
}

int rdpn_ctx_wait(rdpn_ctx* c, int ticket) {
    RDPN_NVTX("rdpn_ctx_wait");
    if (!c || ticket < 0 || ticket >= RDPN_MAX_INFLIGHT || !c->ev[ticket][0]) return RDPN_E_BADARG;
    for (int i = 0; i < RDPN_STAGES; ++i) RDPN_CUDA_TRY(cudaEventSynchronize(c->ev[ticket][i]));
    return 0;
}

}  // extern "C"
