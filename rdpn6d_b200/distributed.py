"""Multi-GPU layer: ROI batches shard across ranks, one dense all-gather of the poses at the end.

ROIs are independent, so the path needs no data-path collective.  Partitioning is the reference's
InferenceSampler rule (/root/reference/core/utils/my_distributed_sampler.py:189-192): contiguous blocks
of ceil(B/W) so that concatenating the ranks' results restores the original order.  The only collective
replaces the reference's pickle-based `all_gather(self._predictions)` + `synchronize()`
(core/gdrn_modeling/gdrn_evaluator.py:439-442) with one `all_gather_into_tensor` of a dense
[shard,16] float32 block per rank (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world_size):
    """my_distributed_sampler.py:189-192 -> (begin, end)."""
    shard = (total - 1) // world_size + 1 if total > 0 else 0
    begin = min(shard * rank, total)
    end = min(shard * (rank + 1), total)
    return begin, end


def shard_size(total, world_size):
    return (total - 1) // world_size + 1 if total > 0 else 0


def shard_batch(batch, rank, world_size):
    """Slice every array/tensor of a batch dict along dim 0 with the InferenceSampler rule."""
    total = next(v for v in batch.values() if v is not None).shape[0]
    b, e = shard_range(total, rank, world_size)
    return {k: (None if v is None else v[b:e]) for k, v in batch.items()}


def gather_rows(rows_local, total, group=None):
    """All-gather per-ROI result rows [n_local,C] into [total,C] in the original ROI order.

    The last ranks' shards may be short (or empty); every rank pads to ceil(total/W) rows so a single
    fixed-size all_gather_into_tensor suffices, and the padding is dropped afterwards.
    """
    if not (dist.is_available() and dist.is_initialized()):
        assert rows_local.shape[0] == total
        return rows_local
    W = dist.get_world_size(group)
    shard = shard_size(total, W)
    C = rows_local.shape[1]
    send = rows_local
    if rows_local.shape[0] < shard:
        pad = torch.zeros(shard - rows_local.shape[0], C, dtype=rows_local.dtype, device=rows_local.device)
        send = torch.cat([rows_local, pad], dim=0)
    send = send.contiguous()
    out = torch.empty(W * shard, C, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:total]


def solve_sharded(solver, batch_local, total, group=None, roi_base=None):
    """Run the solver on this rank's shard and gather [total,16] rows on every rank.

    batch_local: dict of CUDA tensors with keys depth, Kp, coor (or coor_x/y/z), mask, extent, and optionally hyp_idx
    (absent / None: the kernel draws the samples itself), region_idx, anchors, depth_div, t_net.
    roi_base: global index of this shard's first ROI, default = the InferenceSampler rule's `begin` for this rank
    (my_distributed_sampler.py:189-192) -- it keys the kernel-drawn sampling stream, so results do not depend on the
    world size.
    """
    n = batch_local["depth"].shape[0]
    if roi_base is None:
        W, r = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_available() and dist.is_initialized() else (1, 0)
        roi_base = shard_range(total, r, W)[0]
    if n > 0:
        if "coor" in batch_local:
            cx, cy, cz = batch_local["coor"][:, 0], batch_local["coor"][:, 1], batch_local["coor"][:, 2]
        else:
            cx, cy, cz = batch_local["coor_x"], batch_local["coor_y"], batch_local["coor_z"]
        res = solver(batch_local["depth"], batch_local["Kp"], cx, cy, cz, batch_local["mask"], batch_local["extent"],
                     batch_local.get("hyp_idx"), region_idx=batch_local.get("region_idx"),
                     anchors=batch_local.get("anchors"), depth_div=batch_local.get("depth_div"),
                     t_net=batch_local.get("t_net"), roi_base=roi_base)
        rows = res.rows16()
    else:
        rows = torch.zeros(0, 16, dtype=torch.float32, device=batch_local["depth"].device)
    return gather_rows(rows, total, group)


def _parse_cpulist(txt):
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host
    buffers: first-touch then places them on that node and the host-buffer plugin call (copies and the gated pull
    alike) does not cross the socket interconnect.  Returns the node, or None when the topology cannot be read
    (the process is then left alone).  The reference launches one process per GPU the same way
    (core/gdrn_modeling/main_gdrn.py via launch) but leaves placement to the OS."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:], rest)
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
