"""Evaluator hook (surface B2 of SURVEY.md 8b): a batched GPU replacement for the per-ROI CPU loop of
GDRN_Evaluator.process_pnp_ransac (/root/reference/core/gdrn_modeling/gdrn_evaluator.py:316-435).

Same call shape as DatasetEvaluator.process(inputs, outputs, out_dict) (:128): `inputs` is the list of
per-image dicts produced by the reference loader, `outputs` the list of {"time": ...} dicts and
`out_dict` the model's test-time dict (models/GDRN.py:291-297).  One fused kernel launch handles every
ROI of the step; the only device->host transfer is the [n,16] result block.  Behaviour kept from the
reference: warn-and-fallback instead of raising (:294-301, :393-395), BOP rows with R row-major and t in
millimetres (:483-513), per-image time accounting (:420-423).

Wiring in the reference:  PNP_TYPE "gpu_ransac_kabsch" in process() :137-145.
"""
import logging
import time

import numpy as np
import torch

from . import geometry
from .pose_solver import MASK_L1, PoseSolver, STATUS_OK, STATUS_T_SANITY

logger = logging.getLogger(__name__)
PNP_TYPE = "gpu_ransac_kabsch"


def pose_prediction_to_json(pose_est, scene_id, im_id, obj_id, score=None, pose_time=-1):
    """gdrn_evaluator.py:483-513."""
    if score is None:
        score = 1.0
    rot = np.asarray(pose_est)[:3, :3]
    trans = np.asarray(pose_est)[:3, 3]
    return [{"scene_id": scene_id, "im_id": im_id, "obj_id": obj_id, "score": score,
             "R": rot.flatten().tolist(), "t": (1000 * trans).flatten().tolist(), "time": pose_time}]


class GpuRansacKabsch:
    """Stateful processor holding the solver configuration and the accumulated predictions."""

    def __init__(self, num_hyp=256, inlier_thr=0.005, mask_thr=0.5, mask_mode=MASK_L1, weighted=False, refit_iters=1,
                 label_to_obj_id=None, depth_is_scale_normalised=True, seed=0, sample_size=3, adaptive=False, roi_offset=0):
        """depth_is_scale_normalised: roi_coord_2d[:, 2] holds depth / resize_ratio (data_loader.py:563); the
        solver multiplies it back (depth_div = 1 / resize_ratio) so that the 3D-3D solve is metric.
        sample_size: pairs per RANSAC sample (3 = minimal; misc.py:72 uses random_sample_num = 10); adaptive: the
        reference loop's early stop (misc.py:134-138).  roi_offset: global index of the first ROI this evaluator will
        see -- in a multi-rank evaluation the `begin` of this rank's InferenceSampler shard
        (my_distributed_sampler.py:189-192, e.g. distributed.shard_range(total, rank, world)[0]) -- so that every rank
        draws from its own part of the counter-based sampling stream and results do not depend on the world size."""
        self.num_hyp = num_hyp
        # the kernel draws the RANSAC triplets itself (seeded counter-based stream): no separate S1 pass, no
        # multinomial -- the step is one launch
        self.solver = PoseSolver(inlier_thr=inlier_thr, mask_thr=mask_thr, mask_mode=mask_mode, weighted=weighted,
                                 refit_iters=refit_iters, num_hyp=num_hyp, seed=seed, sample_size=sample_size,
                                 adaptive=adaptive)
        self.mask_thr, self.mask_mode = mask_thr, mask_mode
        self.label_to_obj_id = label_to_obj_id or (lambda label: int(label) + 1)
        self.depth_is_scale_normalised = depth_is_scale_normalised
        self.roi_offset = int(roi_offset)
        self._roi_base = self.roi_offset  # global index of the next ROI: the sampling stream does not depend on batching
        self._predictions = []

    def reset(self):
        self._predictions = []
        self._roi_base = self.roi_offset

    def process(self, inputs, outputs, out_dict):
        """Fills self._predictions like process_pnp_ransac; returns the rows appended by this call."""
        dev = out_dict["mask"].device
        t0 = time.perf_counter()
        cat = lambda key: torch.cat([torch.as_tensor(d[key]).to(dev) for d in inputs], dim=0)
        coord2d = cat("roi_coord_2d").float()  # [n,5,64,64]; ch 0-2 = depth xyz (data_loader.py:624-625)
        K = cat("cam").float()
        extent = cat("roi_extent").float()
        center = cat("bbox_center").float()
        scale = cat("scale").float().reshape(-1)
        rr = cat("resize_ratio").float().reshape(-1)
        n = coord2d.shape[0]
        Kp = geometry.roi_intrinsics(K, center, scale, 256)
        depth = coord2d[:, 2].contiguous()
        depth_div = (1.0 / rr) if self.depth_is_scale_normalised else None
        region_idx = geometry.region_argmax(out_dict["region"])
        fps = out_dict.get("fps")
        if fps is None:
            fps = cat("fps").float()
        if fps.dim() == 2:
            fps = fps[None].expand(n, -1, -1)
        args = (depth, Kp, out_dict["coor_x"], out_dict["coor_y"], out_dict["coor_z"], out_dict["mask"], extent)
        t_net = out_dict["trans"].detach().float() if "trans" in out_dict else None
        res = self.solver(*args, None, region_idx=region_idx, anchors=fps, depth_div=depth_div, t_net=t_net,
                          roi_base=self._roi_base)
        self._roi_base += n
        rows = res.rows16().cpu().numpy()  # the single device->host copy of the step
        net_rot = out_dict["rot"].detach().cpu().numpy() if "rot" in out_dict else None
        net_t = t_net.cpu().numpy() if t_net is not None else None
        gpu_time = time.perf_counter() - t0
        out_i = -1
        new_rows = []
        for i, (_input, output) in enumerate(zip(inputs, outputs)):
            start = time.perf_counter()
            json_results = []
            for inst_i in range(len(_input["roi_img"])):
                out_i += 1
                scene_id, im_id = str(_input["scene_im_id"][inst_i]).split("/")
                obj_id = self.label_to_obj_id(_input["roi_cls"][inst_i])
                if obj_id is None:
                    continue
                status = int(rows[out_i, 13])
                pose = rows[out_i, :12].reshape(3, 4).astype(np.float32)
                if status not in (STATUS_OK, STATUS_T_SANITY):
                    logger.warning("num points: %d (status %d)", int(rows[out_i, 14]), status)  # :393-395
                    if net_rot is not None and net_t is not None:  # :298-301 keep the network pose
                        pose = np.hstack([net_rot[out_i], net_t[out_i].reshape(3, 1)]).astype(np.float32)
                score = _input["score"][inst_i] if "score" in _input else 1.0
                json_results.extend(pose_prediction_to_json(pose, scene_id, int(im_id), obj_id, score=float(score),
                                                            pose_time=output.get("time", 0.0)))
            output["time"] = output.get("time", 0.0) + gpu_time / max(len(inputs), 1) + time.perf_counter() - start
            for item in json_results:
                item["time"] = output["time"]
            new_rows.extend(json_results)
        self._predictions.extend(new_rows)
        return new_rows
