/*
 * rdpn6d_b200.h -- C ABI of librdpn6d_b200.so (hand-written sm_100a CUDA kernels).
 *
 * Drop-in boundary for RDPN6D's test-time dense-correspondence -> pose path and its FPS extension.
 * Plain pointers and sizes only; no torch / C++ types.  Each entry cites the reference interface it
 * replaces (paths relative to the RDPN6D repository root).
 *
 * Conventions
 *   - "d_" arguments are DEVICE pointers; entry points that take them are stream-ordered, allocate
 *     nothing, never synchronise, and return 0 on success, a positive cudaError_t, or a negative
 *     RDPN_E_* code.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - "h_" arguments are HOST pointers; those entry points copy in, launch, copy out and
 *     synchronise before returning (they are what a CPU caller of the reference binds to).
 *   - All float tensors are FP32, C-contiguous.  ROI maps are 64 x 64 (P = 4096 pixels), the
 *     reference's BACKBONE.OUTPUT_RES (configs/_base_/gdrn_base.py:26).
 *   - There is NO CPU fallback anywhere in this library.
 */
#ifndef RDPN6D_B200_H
#define RDPN6D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDPN_VERSION 200 /* 0.2.0: three-kernel pipeline (rdpn_pose_solve_ws), rdpn_solve_params gained pipeline / chunk_rois */

/* negative error codes (positive values are cudaError_t) */
#define RDPN_E_BADARG (-1)    /* NULL / non-positive size / unsupported option */
#define RDPN_E_ALIGN (-2)     /* a ROI-map pointer is not 16-byte aligned (bulk-TMA requirement) */
#define RDPN_E_WORKSPACE (-3) /* workspace too small */
#define RDPN_E_TOOLARGE (-4)  /* problem exceeds what the kernel supports */
#define RDPN_E_NOCOOP (-5)    /* device lacks cooperative launch */

/* status codes written per ROI by rdpn_pose_solve (cf. gdrn_evaluator.py:293-301, 393-395) */
#define RDPN_STATUS_OK 0
#define RDPN_STATUS_FEW_POINTS 1   /* fewer than min_pts gated correspondences: pose = -100 fill  */
#define RDPN_STATUS_T_SANITY 2     /* te(t_est, t_net) > 1 m: translation replaced by t_net       */
#define RDPN_STATUS_NO_CONSENSUS 3 /* no valid hypothesis reached min_inliers: pose = -100 fill   */

/* mask post-processing modes (engine_utils.py:118-136 get_out_mask) */
#define RDPN_MAX_SAMPLE 16 /* largest sample_size (pairs per hypothesis) */
#define RDPN_SAMPLE_REDRAWS 7 /* kernel-drawn samples of S > 3 pairs: re-draws of a vertex that repeats an earlier pixel */

#define RDPN_MASK_RAW 0 /* mask already is a probability                     */
#define RDPN_MASK_L1 1  /* per-ROI (m - min) / (max - min), no epsilon       */
#define RDPN_MASK_BCE 2 /* sigmoid                                           */

int rdpn_version(void);
const char* rdpn_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * B1  Farthest point sampling -- replaces core/csrc/fps/src/ext.h:1-14 and
 *     core/csrc/fps/src/farthest_point_sampling.cpp:166-204.
 * ---------------------------------------------------------------------------------------------- */

/* Exact drop-in symbols (same names, same signatures, HOST pointers, synchronous):
 *   pts [pn,3] float, idxs [sn] int (written).  ext.h:1-6 and ext.h:9-14.
 * farthest_point_sampling() keeps the reference's "random start" contract (cpp:93-94) by drawing the
 * start index with srand(time(0)); rand() % pn on the host; everything after that is the GPU kernel. */
void farthest_point_sampling(float* pts, int* idxs, int pn, int sn);
void farthest_point_sampling_init_center(float* pts, int* idxs, int pn, int sn);

/* Device-pointer entries.  d_ws: scratch of at least rdpn_fps_workspace_bytes(sn) bytes.
 * Indices are bit-exact with the reference C++ (squared FP32 distances without FMA, lowest index
 * wins ties, index 0 when nothing is left: cpp:40-73).
 * Three implementations behind the same entry, by cloud size: <= 65 536 points one thread-block cluster (the CTAs'
 * candidates pushed into each other's shared memory, one mbarrier wait per pick); up to 148 x 512 x 16 = 1.21 M points a persistent cooperative grid with the
 * cloud in registers; above that the same grid streaming the cloud and the running minima from global memory, which
 * needs pn more floats behind the workspace header (ws_bytes >= rdpn_fps_workspace_bytes(sn) + 4 pn; RDPN_E_TOOLARGE
 * otherwise). */
size_t rdpn_fps_workspace_bytes(int sn);
int rdpn_fps_init_center(const float* d_pts, int32_t* d_idxs, int pn, int sn, void* d_ws, size_t ws_bytes,
                         void* stream);
int rdpn_fps_from_index(const float* d_pts, int32_t* d_idxs, int pn, int sn, int start, void* d_ws,
                        size_t ws_bytes, void* stream);
/* Gather pts[idxs] (fps_utils.py:21) and, when d_center != NULL, the per-axis mean row appended by
 * get_fps_and_center (core/utils/data_utils.py:217-226) computed in FP64. out: [sn,3] (+[1,3]). */
int rdpn_fps_gather(const float* d_pts, const int32_t* d_idxs, int pn, int sn, float* d_out, double* d_center,
                    void* stream);
/* Many objects in ONE launch (the loop of tools/lm/1_compute_fps.py:26-35 runs get_fps_and_center object by object):
 * d_pts holds the clouds back to back, object o = points [d_offsets[o], d_offsets[o + 1]) (nobj + 1 offsets), every
 * object gets sn picks into d_idxs[o * sn ..] (indices relative to the object's first point).  d_starts == NULL:
 * farthest_point_sampling_init_center for every object; else d_starts[o] is the first pick of object o
 * (farthest_point_sampling).  One thread-block cluster per object: max_pn (the largest cloud) <= 65 536 points,
 * RDPN_E_TOOLARGE above -- use the per-object entries there. */
int rdpn_fps_batch(const float* d_pts, const int32_t* d_offsets, int nobj, int max_pn, int sn, const int32_t* d_starts,
                   int32_t* d_idxs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a1  ROI crop intrinsics -- core/utils/data_utils.py:111-152 (rot = 0) + data_loader.py:553-568.
 *     K [B,9] row-major, center [B,2], scale [B] -> Kp [B,4] = (fx', fy', cx', cy') of the
 *     crop_res x crop_res crop (256 in the reference).  Computed in FP64, rounded once to FP32.
 * ---------------------------------------------------------------------------------------------- */
int rdpn_roi_intrinsics(const float* d_K, const float* d_center, const float* d_scale, int crop_res, float* d_Kp,
                        int B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B4  Generic back-projection -- lib/pysixd/misc.py:319-349 (backproject / backproject_th).
 *     depth [B,H,W], K [B,9] (or one K when k_stride == 0) -> out [B,H,W,3]; (X*depth)/fx order.
 * ---------------------------------------------------------------------------------------------- */
int rdpn_backproject(const float* d_depth, const float* d_K, int k_stride, float* d_out, int B, int H, int W,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * ROI inputs shared by S1 and the fused solver.  Per ROI b (P = 4096, planar, row-major 64x64):
 *   depth  [B,P]   ROI depth at crop pixels (4i,4j), metres, 0 = no measurement
 *                  (data_loader.py:532-535, 563, 625)
 *   Kp     [B,4]   (fx',fy',cx',cy') of the 256 crop (rdpn_roi_intrinsics)
 *   depth_div [B] or NULL: divide depth by this first (resize_ratio, data_loader.py:563)
 *   coor_x/y/z [B,P]  head outputs in [0,1] (cdpn_rot_head_region.py:190-198); delta=(c-0.5)*extent
 *   mask   [B,P]   raw head mask; mask_mode selects get_out_mask's post-processing
 *   extent [B,3]   object extents (roi_extent)
 *   region_idx [B,P] uint8 and anchors [B,R,3]: anchor mode (GDRN.py:206-218); both NULL: dense mode
 *                  (obj = delta, cam = back-projected point)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rdpn_roi_inputs {
    const float* depth;
    const float* Kp;
    const float* depth_div;
    const float* coor_x;
    const float* coor_y;
    const float* coor_z;
    const float* mask;
    const float* extent;
    const uint8_t* region_idx;
    const float* anchors;
    int32_t num_regions; /* R (<= 255); ignored in dense mode */
    int32_t mask_mode;   /* RDPN_MASK_*                      */
    float mask_thr;      /* MASK_THR_TEST, 0.5 (gdrn_base.py) */
    int32_t B;
} rdpn_roi_inputs;

/* S1 (materialising): fused back-projection + residual + mask gate for every pixel.
 *   d_cam  [B,3,P]  camera-side point of the correspondence (q - delta | q)
 *   d_obj  [B,3,P]  object-side point (anchor | delta); may be NULL in anchor mode (region_idx says it)
 *   d_w    [B,P]    mask probability
 *   d_sel  [B,P]    uint8 gate: mask_prob > mask_thr && |delta_c| > 1e-4*extent_c (all c) && depth > 0
 *                   (gdrn_evaluator.py:110-117 + depth validity)
 *   d_nsel [B]      int32 number of gated pixels
 * Replaces data_loader.py:563-576 + gdrn_evaluator.py:89-126 + engine_utils.py:118-136. */
int rdpn_correspond(const rdpn_roi_inputs* in, float* d_cam, float* d_obj, float* d_w, uint8_t* d_sel,
                    int32_t* d_nsel, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B2  Fused pose solve: S1 + hypothesis generation + inlier scoring + best selection + weighted
 *     Kabsch/Umeyama refit, one launch for the whole batch.  Replaces the per-ROI CPU loop of
 *     gdrn_evaluator.py:316-435 (process_pnp_ransac) with the 3D-3D composite of
 *     lib/pysixd/misc.py:58-142 (RANSAC semantics) and lib/pysixd/transform.py:913-980 (Kabsch).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rdpn_solve_params {
    float inlier_thr;    /* metres; inlier <=> ||R a + t - c|| < thr (strict, misc.py:111)          */
    int32_t num_hyp;     /* H                                                                       */
    int32_t min_pts;     /* 4  (gdrn_evaluator.py:380)                                              */
    int32_t min_inliers; /* 4  (misc.py:121)                                                        */
    int32_t weighted;    /* 0: unweighted refit (transform.py), 1: mask-probability weights         */
    int32_t refit_iters; /* >= 1: refit on inliers, re-score, refit ...                             */
    int32_t with_scale;  /* Umeyama scale (transform.py:971-975)                                    */
    int32_t adaptive;    /* misc.py:134-138 early stop emulated on the hypothesis order             */
    float confidence;    /* 0.995 (misc.py:73)                                                      */
    int32_t min_iter;    /* 10 (misc.py:63)                                                         */
    uint32_t seed;       /* internal sampling (hyp_idx == NULL): stream seed                        */
    int32_t roi_base;    /* internal sampling: global index of ROI 0 of this call (shards / chunks  */
                         /* of one job pass their offset so that results do not depend on batching) */
    int32_t sample_size; /* S: correspondences per hypothesis, 3 .. RDPN_MAX_SAMPLE; 0 means 3.     */
                         /* misc.py:72,91 samples random_sample_num = 10 pairs per iteration        */
    int32_t pipeline;    /* RDPN_PIPELINE_*: which implementation runs (results are the same)       */
    int32_t chunk_rois;  /* pipeline: ROIs per pass through the three kernels (0 = default, 8192)   */
    int32_t select_rule; /* RDPN_SELECT_*: which pose the RANSAC stage returns                      */
} rdpn_solve_params;

/* Which pose wins (lib/pysixd/misc.py:113-132):
 *   MOST_INLIERS   the earliest hypothesis with the largest inlier count, refit on its inliers (misc.py:121-126
 *                  without the mean-error bookkeeping; SURVEY section 7's parity definition)
 *   MIN_MEAN_ERR   the reference loop's return value: every hypothesis that raises the best inlier count is refit on
 *                  its inliers, and of all sample fits and refits seen on the way the pose with the lowest mean
 *                  residual over ALL gated points is returned (misc.py:113-120, 127-132) */
#define RDPN_SELECT_MOST_INLIERS 0
#define RDPN_SELECT_MIN_MEAN_ERR 1

/* Implementations of the solve (identical counts / masks / winner; refit pose equal to FP32 rounding):
 *   AUTO   the three-kernel pipeline for batches of >= 2048 ROIs in anchor mode, else the fused kernel
 *   FUSED  one kernel, one CTA per ROI (csrc/pose_solve.cu): lowest latency for small batches
 *   SPLIT  csrc/solve_pipe.cu: front (warp per ROI: gate, back-projection + residual, region sort, hypothesis
 *          poses) -> score (CTA per ROI, bulk-TMA staged, FP32 cores) -> best + refit (warp per ROI), the three
 *          launches chained by programmatic dependent launch; RDPN_E_TOOLARGE where it does not apply (dense mode) */
#define RDPN_PIPELINE_AUTO 0
#define RDPN_PIPELINE_FUSED 1
#define RDPN_PIPELINE_SPLIT 2

typedef struct rdpn_solve_outputs {
    float* pose;           /* [B,12] row-major 3x4 (R|t), FP32                                       */
    int32_t* n_inliers;    /* [B] inliers of the winning hypothesis                                  */
    int32_t* status;       /* [B] RDPN_STATUS_*                                                      */
    int32_t* best_h;       /* [B] winning hypothesis index or -1            (may be NULL)            */
    int32_t* n_sel;        /* [B] gated correspondences                     (may be NULL)            */
    uint8_t* inlier_mask;  /* [B,P] pixels used by the last refit           (may be NULL)            */
    int32_t* hyp_counts;   /* [B,H] inlier count per hypothesis (0 = invalid) (may be NULL)          */
    float* hyp_poses;      /* [B,H,12] FP32 hypothesis poses                (may be NULL)            */
    float* scale;          /* [B] Umeyama scale                             (may be NULL)            */
    float* rows16;         /* [B,16] pose(12) | n_inliers | status | n_sel | best_h as FP32: the dense row
                              block that is all-gathered across GPUs          (may be NULL)            */
} rdpn_solve_outputs;

/* hyp_idx [B,H,S] int32 absolute pixel indices (0..4095), S = prm->sample_size (3 by default); t_net [B,3] or
 * NULL (translation sanity).
 *
 * A hypothesis is valid iff its S pixels passed the gate, are pairwise distinct (the reference samples without
 * replacement, misc.py:91) and, on the object side and on the camera side, some triangle (p0, p_{v-1}, p_v),
 * 2 <= v < S, of the sample is non-degenerate (sin^2 of the angle at p0 > 1e-6; for S = 3 that is the one triangle there is).  S = 3: the
 * closed-form 3-pair Kabsch; S > 3: Kabsch of the S pairs (FP64 moments, closed-form rotation), both rounded once
 * to FP32 (transform.py:913-980 semantics).
 *
 * hyp_idx == NULL: the solver draws the triplets itself, as the reference's loop does with np.random.choice
 * (misc.py:91), from a counter-based stream so that runs are reproducible and independent of batching:
 *     g[0..n)  = the ROI's gated pixels in raster order
 *     fmix32(x): x ^= x >> 16; x *= 0x85ebca6b; x ^= x >> 13; x *= 0xc2b2ae35; x ^= x >> 16      (uint32)
 *     key      = fmix32(fmix32(fmix32(seed ^ 0x9e3779b9) ^ (roi_base + b)) ^ (S * h + v))
 *     pixel of vertex v of hypothesis h = g[(uint64(key) * n) >> 32]
 * S = 3: the three draws are independent, so a sample may repeat a pixel: such a hypothesis is invalid like any other
 * and does not consume an iteration.  S > 3: WITHOUT replacement, as misc.py:91 samples -- a vertex that repeats an earlier
 * pixel of its sample is re-drawn with key' = fmix32(kroi ^ (S * h + v) ^ (attempt << 20)), attempt = 1 .. RDPN_SAMPLE_REDRAWS
 * (attempt 0 is the key above), so small ROIs keep their hypotheses (with independent draws ~45 / n of the 10-pair samples
 * would be lost); a sample still repeating a pixel after the last attempt is invalid.
 * oracle/pose_oracle.py:sample_triplets is the same arithmetic; tests feed its output back as explicit hyp_idx and demand
 * bit-identical results). */
int rdpn_pose_solve(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net,
                    const rdpn_solve_params* prm, const rdpn_solve_outputs* out, void* stream);
/* Same call with a caller-owned scratch buffer for the pipeline's per-ROI packages (raster list and region-sorted
 * list of correspondences, region runs, hypothesis poses and counts: ~175 KB of address space, ~32 KB touched per ROI)
 * and hand-over flags: allocates nothing and never synchronises, so it can be captured in a CUDA graph.  d_ws must be
 * 128-byte aligned and hold at least one package (rdpn_pose_solve_workspace_bytes(1, ..));
 * rdpn_pose_solve_workspace_bytes(B, H, R, chunk_rois) is the size at which a chunk of min(B, chunk_rois) ROIs goes
 * through each kernel in one launch (R = 0: dense mode, fused kernel, no workspace needed).  rdpn_pose_solve itself keeps one such buffer per (device,
 * stream) and grows it on demand (growing synchronises that stream once). */
size_t rdpn_pose_solve_workspace_bytes(int B, int num_hyp, int num_regions, int chunk_rois);
int rdpn_pose_solve_ws(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net,
                       const rdpn_solve_params* prm, const rdpn_solve_outputs* out, void* d_ws, size_t ws_bytes,
                       void* stream);
/* Measurement aid (bench.py's per-kernel roofline lines): the same solve through the pipeline with its three kernels
 * run strictly one after the other and CUDA events between them on `stream`; synchronises, then returns the
 * durations in milliseconds: ms3[0] front (gate + back-projection + residual + sort + hypotheses), ms3[1] scoring,
 * ms3[2] best + refit.  RDPN_E_TOOLARGE where the pipeline does not apply. */
int rdpn_pose_solve_stage_ms(const rdpn_roi_inputs* in, const int32_t* d_hyp_idx, const float* d_t_net,
                             const rdpn_solve_params* prm, const rdpn_solve_outputs* out, void* d_ws, size_t ws_bytes,
                             void* stream, float* ms3);

/* ------------------------------------------------------------------------------------------------
 * B4  Batched weighted Kabsch / Umeyama -- lib/pysixd/transform.py:913-1029
 *     (affine_matrix_from_points(shear=False, usesvd=True) / superimposition_matrix).
 *     src, dst [B,N,3]; w [B,N] or NULL; out_M [B,12] (3x4, maps src -> dst); out_scale [B] or NULL.
 *     Warp-shuffle segmented reduction (FP64 accumulate) + closed-form rotation per ROI.
 * ---------------------------------------------------------------------------------------------- */
int rdpn_kabsch(const float* d_src, const float* d_dst, const float* d_w, int N, int with_scale, float* d_out_M,
                float* d_out_scale, int B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B3  Pose assembly -- core/gdrn_modeling/models/pose_from_pred_centroid_z.py:52-141 (test branch)
 *     with allocentric_to_egocentric (core/utils/utils.py:39-94) done on the GPU for all ROIs
 *     (the reference loops on the CPU with one device->host sync per ROI).
 *     rot_in: [B,9] rotation matrices, or [B,6] rot6d when rot_is_6d (core/utils/rot_reps.py:34-49).
 *     z_type_rel: 1 = "REL" (z * resize_ratio), 0 = "ABS".
 * ---------------------------------------------------------------------------------------------- */
int rdpn_centroid_z_to_pose(const float* d_rot_in, int rot_is_6d, const float* d_centroid, const float* d_z,
                            const float* d_K, const float* d_center, const float* d_resize_ratio,
                            const float* d_wh, int is_allo, int z_type_rel, float* d_rot_out, float* d_trans_out,
                            int B, void* stream);
/* The two sibling heads of the same family, test branches:
 *   trans_mode 2  core/gdrn_modeling/models/pose_from_pred.py:21-58: translation given (d_trans_or_centroid = [B,3])
 *   trans_mode 1  .../pose_from_pred_centroid_z_abs.py:21-92: absolute 2-D centre [B,2] + absolute z [B], K [B,9]
 * rot_kind: 0 = [B,9] matrices, 1 = [B,6] rot6d, 2 = [B,4] quaternions (w,x,y,z; normalised internally as
 * RT_transform.quat_trans_to_pose_m does, lib/pysixd/RT_transform.py:177-183).  is_allo: allocentric -> egocentric. */
int rdpn_assemble_pose(const float* d_rot_in, int rot_kind, const float* d_trans_or_centroid, const float* d_z, const float* d_K,
                       int trans_mode, int is_allo, float* d_rot_out, float* d_trans_out, int B, void* stream);

/* lib/pysixd/misc.py:288-316 (calc_emb_bp_fast / calc_xyz_bp_fast) and :352-371 (backproject_v2): the organised cloud of a
 * depth map through the INVERSE intrinsics, float64 like the reference's numpy einsum:
 *   out[v,u,:] = (d != 0) * R^T (d * Kinv (u,v,1)^T - T).   d_mats: Kinv (9) | R (9) | T (3) doubles (R = I, T = 0 gives
 * backproject_v2).  depth [H,W] FP32 -> out [H,W,3] FP64. */
int rdpn_backproject_kinv(const float* d_depth, const double* d_mats, int H, int W, double* d_out, void* stream);
/* lib/pysixd/pose_error.py:315-337 (adi): mean nearest-neighbour distance between the model points under the
 * ground-truth and under the estimated pose (brute force on the GPU, FP64).  d_pts [n,3] FP32; d_poses: R_est (9) t_est
 * (3) R_gt (9) t_gt (3) doubles; d_scratch: 1 + ceil(n / 256) doubles, zeroed before the first use; d_out: one double. */
int rdpn_adi(const float* d_pts, int n, const double* d_poses, double* d_scratch, double* d_out, void* stream);
/* lib/pysixd/pose_error.py:297-312 (add): mean distance between the SAME model point under the two poses (objects
 * without indistinguishable views).  Arguments as rdpn_adi. */
int rdpn_add(const float* d_pts, int n, const double* d_poses, double* d_scratch, double* d_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f2  Region arg-max -- GDRN.py:206-209: argmax over channels 1..R of region [B,R+1,P] -> uint8.
 * ---------------------------------------------------------------------------------------------- */
int rdpn_region_argmax(const float* d_region, int R, uint8_t* d_region_idx, int B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f2  Correspondence-feature assembly for the unchanged ConvPnPNet -- GDRN.py:199-222 +
 *     conv_pnp_net.py:128-136 + model_utils.py:24-42 in one pass over the region logits:
 *     out [B,C,P] = [coor_x,y,z | roi_coord_2d (5) | fps[argmax softmax(region[:,1:])] (3) | softmax (R, when
 *     region_attention) ] * mask_prob (mask_attention 1 = "mul") [| mask_prob (2 = "concat")];  0 = "none".
 *     C = 11 + (region_attention ? R : 0) + (mask_attention == 2).  R <= 64.
 * ---------------------------------------------------------------------------------------------- */
int rdpn_coor_feat(const float* d_coor_x, const float* d_coor_y, const float* d_coor_z, const float* d_roi_coord_2d,
                   const float* d_region, const float* d_fps, const float* d_mask, int R, int mask_mode,
                   int region_attention, int mask_attention, float* d_out, int B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f4  Region / residual targets -- core/utils/data_utils.py:229-244 (xyz_to_region): nearest FPS anchor per
 *     pixel (float64 distances as scipy cdist, first minimum), region ids 1..R (0 = background where
 *     xyz == 0) and delta = xyz - anchor.  xyz [B,P,3] (HWC), fps [B,R,3] -> region [B,P] u8, delta [B,P,3].
 * ---------------------------------------------------------------------------------------------- */
int rdpn_xyz_to_region(const float* d_xyz, const float* d_fps, int R, int P, uint8_t* d_region, float* d_delta, int B,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * f1  ROI depth crop from the full frame -- cv2.warpAffine(depth, A, (256,256), INTER_LINEAR)[::4, ::4]
 *     of core/gdrn_modeling/data_loader.py:532-535, 625 (core/utils/data_utils.py:81-96), sampled
 *     directly at the 64 x 64 kept positions.  depth_imgs [N,H,W] metres; img_idx [B] (NULL: image 0);
 *     center [B,2], scale [B] as in rdpn_roi_intrinsics; out [B,out_res,out_res].  The result feeds
 *     rdpn_correspond / rdpn_pose_solve as `depth` (depth_div = resize_ratio reproduces :563).
 * ---------------------------------------------------------------------------------------------- */
int rdpn_roi_crop_depth(const float* d_depth_imgs, int H, int W, const int32_t* d_img_idx, const float* d_center,
                        const float* d_scale, int crop_res, int out_res, float* d_out, int B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer plugin call (what a CPU caller of the reference's evaluator binds): the pointers in `in`,
 * hyp_idx, t_net and `out` may be HOST pointers -- or, buffer by buffer, device pointers, which are then used
 * in place (the deployment split of the reference: head outputs already on the GPU, models/GDRN.py:291-297,
 * loader tensors on the host, data_loader.py:417-421).  Synchronous.  The context owns device scratch and
 * streams; ROIs are pipelined in chunks over four stages so that transfers overlap the kernels.
 *
 * Transfer strategies (results are bit-identical):
 *   RDPN_TRANSFER_COPY  every input tensor is copied host -> device by the copy engine.
 *   RDPN_TRANSFER_PULL  "gated pull": only the mask plane (+ the small per-ROI arrays and the hypothesis
 *                       triplets) is copied; a kernel evaluates the mask test of the gate
 *                       (gdrn_evaluator.py:110-117, engine_utils.py:118-136) and reads depth / coor_x/y/z /
 *                       region_idx straight from the caller's buffers over PCIe, only where a pixel can
 *                       pass (typically 10-20 % of a ROI).  Needs depth, coor_*, region_idx in pinned or
 *                       cudaHostRegister'ed memory, 16-byte aligned.
 *   RDPN_TRANSFER_AUTO  (default) decided per buffer: device memory in place, pinned memory pulled (results:
 *                       written by the kernel directly), pageable memory copied.
 * ---------------------------------------------------------------------------------------------- */
#define RDPN_TRANSFER_AUTO 0
#define RDPN_TRANSFER_COPY 1
#define RDPN_TRANSFER_PULL 2

#define RDPN_OPT_TRANSFER 1          /* RDPN_TRANSFER_*                                                 */
#define RDPN_OPT_PULL_GRANULARITY 2  /* 16-byte quads fetched together per plane: 1, 2, 4, 8 or 16      */
#define RDPN_OPT_CHUNK_ROIS 3        /* ROIs per pipeline stage (1..1024, default 256)                  */
#define RDPN_OPT_COUNT_BYTES 4       /* 1: measure the bytes that cross the bus (one 8-byte read-back)  */

typedef struct rdpn_ctx rdpn_ctx;
int rdpn_ctx_create(int device, rdpn_ctx** out_ctx);
void rdpn_ctx_destroy(rdpn_ctx* ctx);
int rdpn_ctx_set_option(rdpn_ctx* ctx, int key, int value);
int rdpn_pose_solve_host(rdpn_ctx* ctx, const rdpn_roi_inputs* h_in, const int32_t* h_hyp_idx, const float* h_t_net,
                         const rdpn_solve_params* prm, const rdpn_solve_outputs* h_out);
/* Asynchronous form of rdpn_pose_solve_host for a loop that processes batch after batch, as the reference's
 * gdrn_inference_on_dataset does (core/gdrn_modeling/gdrn_evaluator.py:649: model forward, then evaluator.process,
 * per batch), and must keep the bus busy across steps: _submit queues
 * the whole call (transfers, gated pull, solver, result copies) on the context's streams and returns a ticket; the
 * outputs are complete once rdpn_ctx_wait(ctx, ticket) returns.  Calls complete in submission order per pipeline
 * stage; up to 8 may be outstanding (a ninth submit first waits for the oldest).  Inputs and outputs of an outstanding
 * call must not be touched, and concurrent calls need distinct output buffers.  Pinned or device buffers keep the
 * call asynchronous; pageable ones are copied synchronously by the CUDA runtime.  rdpn_ctx_last_h2d_bytes describes
 * synchronous calls only. */
int rdpn_pose_solve_host_submit(rdpn_ctx* ctx, const rdpn_roi_inputs* h_in, const int32_t* h_hyp_idx,
                                const float* h_t_net, const rdpn_solve_params* prm, const rdpn_solve_outputs* h_out,
                                int* out_ticket);
int rdpn_ctx_wait(rdpn_ctx* ctx, int ticket);
/* Host -> device bytes of the last rdpn_pose_solve_host call: copied tensors, plus the fetched sectors of
 * the gated pull when RDPN_OPT_COUNT_BYTES is on.  rdpn_ctx_last_transfer: the strategy that call used. */
unsigned long long rdpn_ctx_last_h2d_bytes(const rdpn_ctx* ctx);
int rdpn_ctx_last_transfer(const rdpn_ctx* ctx);
/* Number of kernel launches issued by this library in this process (bench.py's gpu_launches). */
unsigned long long rdpn_launch_count(void);

/* FP32 FMA throughput probe (roofline denominator for the scoring stage): runs `iters` dependent FMA
 * chains on every SM and returns achieved FLOP/s in *out_flops (synchronous). */
int rdpn_fp32_peak_probe(int iters, double* out_flops);

#ifdef __cplusplus
}
#endif
#endif /* RDPN6D_B200_H */
