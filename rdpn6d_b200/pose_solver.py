"""Python host side of the dense-correspondence -> pose path (PyTorch tensors in and out).

`correspond` = stage S1 materialised (back-projection + residual + mask gate);
`pose_solve` = the fused solver: one launch for S1 + hypothesis generation + H x n inlier scoring +
best selection + weighted Kabsch/Umeyama refit (csrc/pose_solve.cu through the C ABI).

Inputs follow the reference's tensors at the evaluator boundary
(/root/reference/core/gdrn_modeling/gdrn_evaluator.py:187-207, models/GDRN.py:291-297):
  depth        [B,64,64]   ROI depth (channel 2 of roi_coord_2d before the resize_ratio division, or
                           pass roi_coord_2d[:,2] with depth_div=None -- it is already divided)
  Kp           [B,4]       crop intrinsics (fx',fy',cx',cy'); see geometry.roi_intrinsics
  coor_x/y/z   [B,1,64,64] or [B,64,64] head outputs
  mask         [B,1,64,64] or [B,64,64] raw head mask
  extent       [B,3]       roi_extent
  region_idx   [B,64,64] uint8 (geometry.region_argmax of out_dict["region"]) + anchors [B,R,3] (fps)
  hyp_idx      [B,H,S] int32 absolute pixel indices of each hypothesis' S correspondences (S = 3 by default;
               the reference's loop samples random_sample_num = 10, misc.py:72,91)
PoseSolver / correspond take CUDA tensors; HostPoseSolver takes CPU tensors and moves them itself (host-buffer
plugin call).  Either way the arithmetic runs in the CUDA kernels: there is no CPU compute path.
"""
import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

MASK_RAW, MASK_L1, MASK_BCE = 0, 1, 2
_MASK_MODES = {"raw": MASK_RAW, "none": MASK_RAW, "l1": MASK_L1, "bce": MASK_BCE}
STATUS_OK, STATUS_FEW_POINTS, STATUS_T_SANITY, STATUS_NO_CONSENSUS = 0, 1, 2, 3
_PIPELINES = {"auto": _lib.PIPELINE_AUTO, "fused": _lib.PIPELINE_FUSED, "split": _lib.PIPELINE_SPLIT}


def _workspace(B, H, R, chunk_rois, dev):
    """Scratch for the pipeline's per-ROI packages (rdpn_pose_solve_workspace_bytes), a uint8 CUDA tensor."""
    n = int(_lib.lib().rdpn_pose_solve_workspace_bytes(int(B), int(H), int(R), int(chunk_rois)))
    return torch.empty(max(n, 128), dtype=torch.uint8, device=dev)
P = 64 * 64


def _map(x, name, B):
    """[B,1,64,64] | [B,64,64] float32 contiguous CUDA, 16-byte aligned."""
    if x.dim() == 4:
        assert x.shape[1] == 1, (name, x.shape)
        x = x[:, 0]
    assert x.shape == (B, 64, 64), (name, tuple(x.shape))
    if not x.is_cuda:
        raise RuntimeError("rdpn6d_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    x = x.detach().to(torch.float32).contiguous()
    if x.data_ptr() % 16:  # bulk-TMA staging needs 16-byte aligned planes; a fresh allocation is aligned
        x = x.clone()
    return x


def _vec(x, shape, name, dtype=torch.float32):
    assert tuple(x.shape) == tuple(shape), (name, tuple(x.shape), shape)
    if not x.is_cuda:
        raise RuntimeError("rdpn6d_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    return x.detach().to(dtype).contiguous()


def _mask_mode(m):
    if isinstance(m, str):
        if m.lower() == "ce":
            raise ValueError("mask_mode 'ce' is resolved by the tensor entries (argmax over the two channels, then 'raw'); "
                             "this entry takes the arg-maxed plane with mask_mode='raw'")
        return _MASK_MODES[m.lower()]
    return int(m)


def out_mask_ce(pred_mask):
    """engine_utils.py:131-132, MASK_LOSS_TYPE == "CE": the mask is the arg-max over the class channels, [B,C,64,64] ->
    [B,64,64] float32 of 0 / 1 (the gate's `> mask_thr` then reads it as a probability, mask_mode 'raw')."""
    return torch.argmax(pred_mask, dim=1).to(torch.float32)


def _resolve_mask(mask, mode):
    """(plane, integer mode) for the C struct: 'ce' becomes the arg-maxed plane in 'raw' mode."""
    if isinstance(mode, str) and mode.lower() == "ce":
        return out_mask_ce(mask), MASK_RAW
    return mask, _mask_mode(mode)


class _Inputs:
    """Keeps the prepared tensors alive next to the C struct that points at them."""

    def __init__(self, depth, Kp, coor_x, coor_y, coor_z, mask, extent, region_idx=None, anchors=None,
                 depth_div=None, mask_mode=MASK_L1, mask_thr=0.5):
        B = depth.shape[0]
        self.B = B
        self.dev = depth.device
        mask, mask_mode = _resolve_mask(mask, mask_mode)
        self.t = dict(
            depth=_map(depth, "depth", B), coor_x=_map(coor_x, "coor_x", B), coor_y=_map(coor_y, "coor_y", B),
            coor_z=_map(coor_z, "coor_z", B), mask=_map(mask, "mask", B),
            Kp=_vec(Kp, (B, 4), "Kp"), extent=_vec(extent, (B, 3), "extent"))
        if depth_div is not None:
            self.t["depth_div"] = _vec(depth_div.reshape(B), (B,), "depth_div")
        R = 0
        if (region_idx is None) != (anchors is None):
            raise ValueError("region_idx and anchors must be given together (anchor mode) or both None (dense mode)")
        if region_idx is not None:
            R = anchors.shape[1]
            if R > 255:
                raise ValueError("num_regions must be <= 255")
            rid = region_idx.detach()
            assert rid.shape == (B, 64, 64), tuple(rid.shape)
            # Region ids are 0-based anchor indices (argmax over the R foreground channels, GDRN.py:206-218); the
            # loader's roi_region uses 1..R with 0 = background (data_utils.py:229-244) and must be shifted by the
            # caller.  The kernels drop pixels whose id is outside [0, R) from the gate; a wider integer dtype is
            # checked here before the cast to uint8 would wrap it (RDPN_CHECK_INPUTS=1 also checks uint8 inputs: it
            # costs a device synchronisation).
            if rid.dtype != torch.uint8 or os.environ.get("RDPN_CHECK_INPUTS") == "1":
                lo, hi = int(rid.min()), int(rid.max())
                if lo < 0 or hi >= max(R, 1):
                    raise ValueError("region_idx must hold anchor indices in [0, %d): found [%d, %d]" % (R, lo, hi))
            self.t["region_idx"] = rid.to(torch.uint8).contiguous()
            self.t["anchors"] = _vec(anchors, (B, R, 3), "anchors")
        s = _lib.RoiInputs()
        for k in ("depth", "Kp", "depth_div", "coor_x", "coor_y", "coor_z", "mask", "extent", "region_idx", "anchors"):
            setattr(s, k, self.t[k].data_ptr() if k in self.t else None)
        s.num_regions = R
        s.mask_mode = _mask_mode(mask_mode)
        s.mask_thr = float(mask_thr)
        s.B = B
        self.struct = s


def correspond(depth, Kp, coor_x, coor_y, coor_z, mask, extent, region_idx=None, anchors=None, depth_div=None,
               mask_mode=MASK_L1, mask_thr=0.5, want_obj=True, stream=None):
    """Stage S1.  Returns dict(cam[B,3,P], obj[B,3,P] | None, w[B,P], sel[B,P] uint8, n_sel[B] int32)."""
    L = _lib.lib()
    inp = _Inputs(depth, Kp, coor_x, coor_y, coor_z, mask, extent, region_idx, anchors, depth_div, mask_mode, mask_thr)
    B, dev = inp.B, inp.dev
    cam = torch.empty(B, 3, P, dtype=torch.float32, device=dev)
    obj = torch.empty(B, 3, P, dtype=torch.float32, device=dev) if want_obj else None
    w = torch.empty(B, P, dtype=torch.float32, device=dev)
    sel = torch.empty(B, P, dtype=torch.uint8, device=dev)
    nsel = torch.empty(B, dtype=torch.int32, device=dev)
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    with torch.cuda.device(dev):
        rc = L.rdpn_correspond(ctypes.byref(inp.struct), cam.data_ptr(), obj.data_ptr() if want_obj else None,
                               w.data_ptr(), sel.data_ptr(), nsel.data_ptr(), st)
    _lib.check(rc, "correspond")
    return dict(cam=cam, obj=obj, w=w, sel=sel, n_sel=nsel)


@dataclass
class PoseSolveResult:
    pose: torch.Tensor  # [B,3,4] float32 (R|t), -100 where status is FEW_POINTS / NO_CONSENSUS
    n_inliers: torch.Tensor  # [B] int32
    status: torch.Tensor  # [B] int32
    best_h: torch.Tensor  # [B] int32
    n_sel: torch.Tensor  # [B] int32
    inlier_mask: Optional[torch.Tensor] = None  # [B,64,64] uint8
    hyp_counts: Optional[torch.Tensor] = None  # [B,H] int32
    hyp_poses: Optional[torch.Tensor] = None  # [B,H,3,4] float32
    scale: Optional[torch.Tensor] = None  # [B] float32
    rows: Optional[torch.Tensor] = None  # [B,16] float32 written by the kernel (see rows16)

    def rows16(self):
        """[B,16] float32 rows for the multi-GPU gather: pose(12) | n_inliers | status | n_sel | best_h."""
        if self.rows is not None:
            return self.rows
        B = self.pose.shape[0]
        return torch.cat([self.pose.reshape(B, 12), self.n_inliers.float()[:, None], self.status.float()[:, None],
                          self.n_sel.float()[:, None], self.best_h.float()[:, None]], dim=1)


def _hyp_arg(hyp_idx, B, num_hyp, sample_size, conv):
    """(H, S, tensor | None) of a call: explicit hyp_idx [B,H,S] or the solver's internal sampling."""
    if hyp_idx is None:
        return num_hyp, sample_size, None
    if hyp_idx.dim() != 3 or hyp_idx.shape[0] != B or not 3 <= hyp_idx.shape[2] <= _lib.MAX_SAMPLE:
        raise ValueError("hyp_idx must be [B,H,S] with 3 <= S <= %d, got %s" % (_lib.MAX_SAMPLE, tuple(hyp_idx.shape)))
    H, S = hyp_idx.shape[1], hyp_idx.shape[2]
    return H, S, conv(hyp_idx, (B, H, S))


class PoseSolver:
    """Reusable launcher: output buffers are allocated once per (B, H) and reused (graph friendly)."""

    def __init__(self, inlier_thr=0.005, min_pts=4, min_inliers=4, weighted=False, refit_iters=1, with_scale=False,
                 adaptive=False, confidence=0.995, min_iter=10, mask_mode=MASK_L1, mask_thr=0.5,
                 want_inlier_mask=False, want_hyp=False, num_hyp=256, seed=0, sample_size=3, pipeline="auto", chunk_rois=0,
                 select_rule="most_inliers"):
        """pipeline: "auto" (three-kernel pipeline gate_pack -> score -> refit where it applies, else the fused kernel),
        "fused", "split"; chunk_rois: ROIs per pass through the three kernels (0 = library default).
        select_rule: "most_inliers" (misc.py:121-126: earliest hypothesis with the largest count, refit on its inliers) or
        "min_mean_err" (the reference loop's return value, misc.py:113-132: of all sample fits and of the refits of the
        hypotheses that raised the best count, the pose with the lowest mean residual over ALL gated points; anchor mode,
        runs the three-kernel pipeline whatever the batch size).
        num_hyp / seed / sample_size: used when a call passes hyp_idx=None -- the kernel then draws num_hyp samples of
        sample_size pixels itself from a counter-based stream (include/rdpn6d_b200.h), the stand-in for np.random.choice
        at misc.py:91.  With explicit hyp_idx [B,H,S] both H and S come from the tensor."""
        if not 3 <= int(sample_size) <= _lib.MAX_SAMPLE:
            raise ValueError("sample_size must be in 3..%d" % _lib.MAX_SAMPLE)
        self.sample_size = int(sample_size)
        self.prm = dict(inlier_thr=float(inlier_thr), min_pts=int(min_pts), min_inliers=int(min_inliers),
                        weighted=int(bool(weighted)), refit_iters=int(refit_iters), with_scale=int(bool(with_scale)),
                        adaptive=int(bool(adaptive)), confidence=float(confidence), min_iter=int(min_iter),
                        seed=int(seed) & 0xFFFFFFFF, pipeline=_PIPELINES[pipeline] if isinstance(pipeline, str) else int(pipeline),
                        chunk_rois=int(chunk_rois),
                        select_rule={"most_inliers": _lib.SELECT_MOST_INLIERS, "min_mean_err": _lib.SELECT_MIN_MEAN_ERR}[select_rule])
        self.num_hyp = int(num_hyp)
        self._ws = {}
        self.mask_mode = mask_mode
        self.mask_thr = mask_thr
        self.want_inlier_mask = want_inlier_mask
        self.want_hyp = want_hyp
        self._out = {}

    def _buffers(self, B, H, dev):
        key = (B, H, str(dev))
        if key not in self._out:
            o = dict(pose=torch.empty(B, 12, dtype=torch.float32, device=dev),
                     n_inliers=torch.empty(B, dtype=torch.int32, device=dev),
                     status=torch.empty(B, dtype=torch.int32, device=dev),
                     best_h=torch.empty(B, dtype=torch.int32, device=dev),
                     n_sel=torch.empty(B, dtype=torch.int32, device=dev),
                     scale=torch.empty(B, dtype=torch.float32, device=dev),
                     rows16=torch.empty(B, 16, dtype=torch.float32, device=dev))
            if self.want_inlier_mask:
                o["inlier_mask"] = torch.empty(B, 64, 64, dtype=torch.uint8, device=dev)
            if self.want_hyp:
                o["hyp_counts"] = torch.empty(B, H, dtype=torch.int32, device=dev)
                o["hyp_poses"] = torch.empty(B, H, 3, 4, dtype=torch.float32, device=dev)
            self._out[key] = o
        return self._out[key]

    def __call__(self, depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx=None, region_idx=None, anchors=None,
                 depth_div=None, t_net=None, stream=None, roi_base=0):
        """hyp_idx=None: the kernel draws self.num_hyp triplets per ROI itself (seeded; roi_base = global index of
        ROI 0 of this call, so that shards of one job reproduce the single-GPU result)."""
        L = _lib.lib()
        inp = _Inputs(depth, Kp, coor_x, coor_y, coor_z, mask, extent, region_idx, anchors, depth_div,
                      self.mask_mode, self.mask_thr)
        B, dev = inp.B, inp.dev
        H, S, hyp = _hyp_arg(hyp_idx, B, self.num_hyp, self.sample_size, lambda x, shp: _vec(x, shp, "hyp_idx", torch.int32))
        tn = _vec(t_net, (B, 3), "t_net") if t_net is not None else None
        prm = _lib.SolveParams(num_hyp=H, roi_base=int(roi_base), sample_size=S, **self.prm)
        o = self._buffers(B, H, dev)
        outs = _lib.SolveOutputs()
        for k in ("pose", "n_inliers", "status", "best_h", "n_sel", "inlier_mask", "hyp_counts", "hyp_poses", "scale", "rows16"):
            setattr(outs, k, o[k].data_ptr() if k in o else None)
        st = (stream or torch.cuda.current_stream(dev)).cuda_stream
        wkey = (B, H, inp.struct.num_regions, str(dev), st)
        if wkey not in self._ws:
            self._ws[wkey] = _workspace(B, H, inp.struct.num_regions, self.prm["chunk_rois"], dev)
        ws = self._ws[wkey]
        with torch.cuda.device(dev):
            rc = L.rdpn_pose_solve_ws(ctypes.byref(inp.struct), hyp.data_ptr() if hyp is not None else None,
                                      tn.data_ptr() if tn is not None else None, ctypes.byref(prm), ctypes.byref(outs),
                                      ws.data_ptr(), ws.numel(), st)
        _lib.check(rc, "pose_solve")
        self._keep = (inp, hyp, tn)  # keep inputs alive until the next call (stream-ordered use)
        return PoseSolveResult(pose=o["pose"].view(B, 3, 4), n_inliers=o["n_inliers"], status=o["status"],
                               best_h=o["best_h"], n_sel=o["n_sel"], inlier_mask=o.get("inlier_mask"),
                               hyp_counts=o.get("hyp_counts"), hyp_poses=o.get("hyp_poses"), scale=o["scale"], rows=o["rows16"])


class SolvePlan:
    """A fully prepared launch (C structs built once): `launch()` is a single ctypes call, so a
    benchmark or serving loop pays no per-step Python tensor bookkeeping."""

    def __init__(self, solver, inp, hyp, tn, prm, outs, result, ws):
        self._solver, self._inp, self._hyp, self._tn, self._prm, self._outs = solver, inp, hyp, tn, prm, outs
        self.result = result
        self._ws = ws  # the plan's own package scratch: plans may run concurrently on different streams
        self._fn = _lib.lib().rdpn_pose_solve_ws
        self._args = (ctypes.byref(inp.struct), hyp.data_ptr() if hyp is not None else None,
                      tn.data_ptr() if tn is not None else None, ctypes.byref(prm), ctypes.byref(outs),
                      ws.data_ptr(), ws.numel())
        self._dev = inp.dev

    def launch(self, stream=None):
        st = (stream or torch.cuda.current_stream(self._dev)).cuda_stream
        rc = self._fn(*self._args, st)
        if rc:
            _lib.check(rc, "pose_solve")
        return self.result

    def stage_ms(self, stream=None):
        """Durations (ms) of the pipeline's three kernels on this plan's inputs, run one after the other with CUDA
        events between them (rdpn_pose_solve_stage_ms): (front, score, refit).  Synchronises."""
        st = (stream or torch.cuda.current_stream(self._dev)).cuda_stream
        ms = (ctypes.c_float * 3)()
        _lib.check(_lib.lib().rdpn_pose_solve_stage_ms(*self._args, st, ms), "pose_solve_stage_ms")
        return float(ms[0]), float(ms[1]), float(ms[2])


def make_plan(solver, depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx=None, region_idx=None, anchors=None,
              depth_div=None, t_net=None, roi_base=0):
    """Prepare a reusable launch of `solver` on fixed device buffers (private output buffers)."""
    inp = _Inputs(depth, Kp, coor_x, coor_y, coor_z, mask, extent, region_idx, anchors, depth_div,
                  solver.mask_mode, solver.mask_thr)
    B, dev = inp.B, inp.dev
    H, S, hyp = _hyp_arg(hyp_idx, B, solver.num_hyp, solver.sample_size, lambda x, shp: _vec(x, shp, "hyp_idx", torch.int32))
    tn = _vec(t_net, (B, 3), "t_net") if t_net is not None else None
    prm = _lib.SolveParams(num_hyp=H, roi_base=int(roi_base), sample_size=S, **solver.prm)
    solver._out.pop((B, H, str(dev)), None)
    o = solver._buffers(B, H, dev)
    solver._out.pop((B, H, str(dev)), None)  # the plan owns these buffers
    outs = _lib.SolveOutputs()
    for k in ("pose", "n_inliers", "status", "best_h", "n_sel", "inlier_mask", "hyp_counts", "hyp_poses", "scale", "rows16"):
        setattr(outs, k, o[k].data_ptr() if k in o else None)
    res = PoseSolveResult(pose=o["pose"].view(B, 3, 4), n_inliers=o["n_inliers"], status=o["status"],
                          best_h=o["best_h"], n_sel=o["n_sel"], inlier_mask=o.get("inlier_mask"),
                          hyp_counts=o.get("hyp_counts"), hyp_poses=o.get("hyp_poses"), scale=o["scale"], rows=o["rows16"])
    ws = _workspace(B, H, inp.struct.num_regions, solver.prm["chunk_rois"], dev)
    return SolvePlan(solver, inp, hyp, tn, prm, outs, res, ws)


def pose_solve(depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx=None, region_idx=None, anchors=None,
               depth_div=None, t_net=None, stream=None, roi_base=0, **kw):
    """One-shot functional form of PoseSolver (see its constructor for the keyword arguments)."""
    return PoseSolver(**kw)(depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx, region_idx, anchors,
                            depth_div, t_net, stream, roi_base)


def sample_hypotheses(sel, H, generator=None, sample_size=3):
    """Draw hyp_idx [B,H,S] int32 uniformly (with replacement) from each ROI's gated pixels.

    The reference draws its RANSAC samples with np.random.choice inside the loop
    (lib/pysixd/misc.py:91); here randomness is lifted out into an explicit tensor so that runs are
    reproducible.  sel: [B,P] uint8/bool CUDA tensor (correspond()["sel"]).  ROIs with no gated pixel get
    index 0 (their hypotheses are invalid and the solver reports FEW_POINTS).
    """
    B = sel.shape[0]
    w = sel.reshape(B, -1).float()
    empty = w.sum(dim=1) == 0
    w[empty, 0] = 1.0
    idx = torch.multinomial(w, H * sample_size, replacement=True, generator=generator)
    return idx.view(B, H, sample_size).to(torch.int32)


class HostPoseSolver:
    """Plugin call for buffers that are not (all) on the GPU yet (rdpn_pose_solve_host): results come back in CPU
    tensors.  Every input may be a CUDA tensor (used in place -- the head outputs of models/GDRN.py:291-297), a
    PINNED CPU tensor (the loader's roi_coord_2d / cam / roi_extent, data_loader.py:417-421) or a pageable one.
    The library moves what has to move and launches the fused solver in a 4-stage pipeline; pinned planes are not
    copied but fetched over PCIe only where the mask test passes ("gated pull", include/rdpn6d_b200.h), pinned
    outputs are written by the kernel directly.  There is still no CPU compute path.

    transfer: "auto" (pull when the buffers are pinned), "copy", "pull".
    """

    _TRANSFER = {"auto": _lib.TRANSFER_AUTO, "copy": _lib.TRANSFER_COPY, "pull": _lib.TRANSFER_PULL}

    def __init__(self, device=0, transfer="auto", chunk_rois=None, pull_granularity=None, count_bytes=False, pin_outputs=True,
                 **solver_kw):
        self._L = _lib.lib()
        self._ctx = ctypes.c_void_p()
        self.device = int(device)
        _lib.check(self._L.rdpn_ctx_create(int(device), ctypes.byref(self._ctx)), "ctx_create")
        self.set_option(_lib.OPT_TRANSFER, self._TRANSFER[transfer])
        if chunk_rois is not None:
            self.set_option(_lib.OPT_CHUNK_ROIS, int(chunk_rois))
        if pull_granularity is not None:
            self.set_option(_lib.OPT_PULL_GRANULARITY, int(pull_granularity))
        self.set_option(_lib.OPT_COUNT_BYTES, int(bool(count_bytes)))
        self.pin_outputs = pin_outputs
        s = PoseSolver(**solver_kw)
        self.prm, self.mask_mode, self.mask_thr, self.num_hyp = s.prm, s.mask_mode, s.mask_thr, s.num_hyp
        self.sample_size = s.sample_size
        self.want_inlier_mask, self.want_hyp = s.want_inlier_mask, s.want_hyp
        self._out = {}

    def set_option(self, key, value):
        _lib.check(self._L.rdpn_ctx_set_option(self._ctx, key, value), "ctx_set_option")

    @property
    def last_h2d_bytes(self):
        return int(self._L.rdpn_ctx_last_h2d_bytes(self._ctx))

    @property
    def last_transfer(self):
        return {_lib.TRANSFER_COPY: "copy", _lib.TRANSFER_PULL: "pull"}.get(int(self._L.rdpn_ctx_last_transfer(self._ctx)))

    def close(self):
        if self._ctx:
            self._L.rdpn_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _cpu(self, x, shape, dtype, name):
        """Any mix of CPU and CUDA tensors is accepted: the library classifies every buffer on its own (device memory is
        used in place, pinned memory is pulled / written directly, pageable memory is copied)."""
        if x.is_cuda and x.device.index != self.device:
            raise RuntimeError("rdpn6d_b200: %s is on %s, the context is on cuda:%d" % (name, x.device, self.device))
        x = x.detach()
        if x.dtype != dtype or not x.is_contiguous():
            x = x.to(dtype).contiguous()  # note: loses pinning; pass float32 / uint8 / int32 contiguous tensors to pull
        return x.reshape(shape)

    def _buffers(self, B, H):
        key = (B, H)
        if key not in self._out:
            mk = (lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()) if self.pin_outputs else \
                 (lambda *s, dtype: torch.empty(*s, dtype=dtype))
            o = dict(pose=mk(B, 12, dtype=torch.float32), n_inliers=mk(B, dtype=torch.int32), status=mk(B, dtype=torch.int32),
                     best_h=mk(B, dtype=torch.int32), n_sel=mk(B, dtype=torch.int32), scale=mk(B, dtype=torch.float32))
            if self.want_inlier_mask:
                o["inlier_mask"] = mk(B, 64, 64, dtype=torch.uint8)
            if self.want_hyp:
                o["hyp_counts"] = mk(B, H, dtype=torch.int32)
                o["hyp_poses"] = mk(B, H, 3, 4, dtype=torch.float32)
            self._out[key] = o
        return self._out[key]

    def plan(self, depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx=None, region_idx=None, anchors=None,
             depth_div=None, t_net=None, roi_base=0, private_outputs=False):
        """Prepare a call on fixed buffers: returns a zero-argument callable that is one ctypes call (a serving loop
        that refills the same pinned buffers pays no per-step Python bookkeeping) and yields the PoseSolveResult.

        The callable also has the asynchronous pair `ticket = run.submit()` / `result = run.wait(ticket)`
        (rdpn_pose_solve_host_submit / rdpn_ctx_wait): a loop that submits step i + 1 before it waits for step i keeps
        the bus busy across steps.  Plans that are in flight together need private_outputs=True (own result buffers)."""
        B = depth.shape[0]
        H, S, hyp = _hyp_arg(hyp_idx, B, self.num_hyp, self.sample_size, lambda x, shp: self._cpu(x, shp, torch.int32, "hyp_idx"))
        f32, t = torch.float32, {}
        for name, x in (("depth", depth), ("coor_x", coor_x), ("coor_y", coor_y), ("coor_z", coor_z), ("mask", mask)):
            t[name] = self._cpu(x, (B, P), f32, name)
        t["Kp"] = self._cpu(Kp, (B, 4), f32, "Kp")
        t["extent"] = self._cpu(extent, (B, 3), f32, "extent")
        if (region_idx is None) != (anchors is None):
            raise ValueError("region_idx and anchors must be given together (anchor mode) or both None (dense mode)")
        R = 0
        if region_idx is not None:
            R = anchors.shape[1]
            t["region_idx"] = self._cpu(region_idx, (B, P), torch.uint8, "region_idx")
            t["anchors"] = self._cpu(anchors, (B, R, 3), f32, "anchors")
        if depth_div is not None:
            t["depth_div"] = self._cpu(depth_div, (B,), f32, "depth_div")
        tn = self._cpu(t_net, (B, 3), f32, "t_net") if t_net is not None else None
        s = _lib.RoiInputs()
        for k in ("depth", "Kp", "depth_div", "coor_x", "coor_y", "coor_z", "mask", "extent", "region_idx", "anchors"):
            setattr(s, k, t[k].data_ptr() if k in t else None)
        s.num_regions, s.mask_mode, s.mask_thr, s.B = R, _mask_mode(self.mask_mode), float(self.mask_thr), B
        prm = _lib.SolveParams(num_hyp=H, roi_base=int(roi_base), sample_size=S, **self.prm)
        if private_outputs:
            self._out.pop((B, H), None)
        o = self._buffers(B, H)
        if private_outputs:
            self._out.pop((B, H), None)  # the plan owns these buffers
        outs = _lib.SolveOutputs()
        for k in ("pose", "n_inliers", "status", "best_h", "n_sel", "inlier_mask", "hyp_counts", "hyp_poses", "scale"):
            setattr(outs, k, o[k].data_ptr() if k in o else None)
        res = PoseSolveResult(pose=o["pose"].view(B, 3, 4), n_inliers=o["n_inliers"], status=o["status"], best_h=o["best_h"],
                              n_sel=o["n_sel"], inlier_mask=o.get("inlier_mask"), hyp_counts=o.get("hyp_counts"),
                              hyp_poses=o.get("hyp_poses"), scale=o["scale"], rows=None)
        fn, ctx = self._L.rdpn_pose_solve_host, self._ctx
        cargs = (ctx, ctypes.byref(s), hyp.data_ptr() if hyp is not None else None,
                 tn.data_ptr() if tn is not None else None, ctypes.byref(prm), ctypes.byref(outs))
        keep = (t, hyp, tn, s, prm, outs)  # the C structs point into these

        def run(_keep=keep):
            rc = fn(*cargs)
            if rc:
                _lib.check(rc, "pose_solve_host")
            return res

        fn_submit, fn_wait, ticket = self._L.rdpn_pose_solve_host_submit, self._L.rdpn_ctx_wait, ctypes.c_int(-1)

        def submit():
            rc = fn_submit(*cargs, ctypes.byref(ticket))
            if rc:
                _lib.check(rc, "pose_solve_host_submit")
            return ticket.value

        def wait(tk):
            rc = fn_wait(ctx, tk)
            if rc:
                _lib.check(rc, "ctx_wait")
            return res

        run.submit, run.wait = submit, wait
        return run

    def __call__(self, depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx=None, region_idx=None, anchors=None,
                 depth_div=None, t_net=None, roi_base=0):
        return self.plan(depth, Kp, coor_x, coor_y, coor_z, mask, extent, hyp_idx, region_idx, anchors, depth_div, t_net,
                         roi_base)()
