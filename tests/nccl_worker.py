"""torchrun worker for test_nccl_sharded_solve_two_gpus: shard 50 ROIs over the ranks, gather, compare
with the unsharded solve on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdpn6d_b200 import distributed as D  # noqa: E402
from rdpn6d_b200 import pose_solver, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    total = 51
    b = synth.make_batch(total, H=64, seed=5)
    full = {k: torch.from_numpy(v).cuda() for k, v in b.items() if v is not None}
    solver = pose_solver.PoseSolver(inlier_thr=0.005)
    local = D.shard_batch(full, rank, world)
    rows = D.solve_sharded(solver, local, total)
    ref = D.solve_sharded(pose_solver.PoseSolver(inlier_thr=0.005), full, total, group=None) if False else None
    r = pose_solver.PoseSolver(inlier_thr=0.005)(full["depth"], full["Kp"], full["coor"][:, 0], full["coor"][:, 1], full["coor"][:, 2],
                                                  full["mask"], full["extent"], full["hyp_idx"], full["region_idx"], full["anchors"])
    ok = torch.equal(rows, r.rows16())
    # kernel-drawn samples: every rank draws from its own part of the counter-based stream (roi_base = shard begin), so the
    # gathered rows equal the single-GPU solve whatever the world size
    samp = pose_solver.PoseSolver(inlier_thr=0.005, num_hyp=64, seed=3)
    local2 = {k: v for k, v in local.items() if k != "hyp_idx"}
    rows2 = D.solve_sharded(samp, local2, total)
    r2 = pose_solver.PoseSolver(inlier_thr=0.005, num_hyp=64, seed=3)(
        full["depth"], full["Kp"], full["coor"][:, 0], full["coor"][:, 1], full["coor"][:, 2], full["mask"], full["extent"], None,
        full["region_idx"], full["anchors"])
    ok = ok and torch.equal(rows2, r2.rows16())
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag) == 1:
        print("NCCL_GATHER_OK")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
