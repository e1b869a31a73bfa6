"""rdpn6d_b200 -- B200-native (sm_100a) dense-correspondence -> pose path of RDPN6D.

Only what the path needs: csrc/ (hand-written CUDA kernels + the C ABI of include/rdpn6d_b200.h) and
the Python host side mirroring the reference's call surfaces:

  fps_utils        farthest_point_sampling / get_fps_and_center   (core/csrc/fps/fps_utils.py)
  pose_solver      correspond / PoseSolver / pose_solve            (gdrn_evaluator.py process_pnp_ransac)
  geometry         backproject_th / kabsch / superimposition_matrix / roi_intrinsics / region_argmax
  pose_from_pred   pose_from_pred_centroid_z                       (models/pose_from_pred_centroid_z.py)
  evaluator        GpuRansacKabsch.process                         (DatasetEvaluator.process hook)
  distributed      shard_range / gather_rows / solve_sharded       (InferenceSampler + all_gather)
  synth            deterministic synthetic ROI batches (numpy only)

Importing the package does not load CUDA; the first op call loads librdpn6d_b200.so and raises if it
is missing.  There is no CPU fallback.
"""
__version__ = "0.1.0"
