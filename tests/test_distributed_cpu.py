"""CPU: the N>1 host logic (shard rule + dense all-gather of result rows) over gloo, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rdpn6d_b200 import distributed as D


def test_shard_rule_matches_inference_sampler():
    # /root/reference/core/utils/my_distributed_sampler.py:189-192
    for total in (1, 2, 7, 8, 1024, 65536, 65537):
        for W in (1, 2, 3, 4, 8):
            shard = (total - 1) // W + 1
            covered = []
            for r in range(W):
                b, e = D.shard_range(total, r, W)
                assert b == min(shard * r, total) and e == min(shard * (r + 1), total)
                covered += list(range(b, e))
            assert covered == list(range(total))
    assert D.shard_range(0, 0, 2) == (0, 0)


def test_shard_rule_matches_reference_sampler_run_from_source(golden_dir):
    """class InferenceSampler (my_distributed_sampler.py:170-199) executed from source (oracle/gen_golden.py:gen_sampler):
    shard_range gives every rank the index range the reference's sampler gives it (None = an empty shard)."""
    import json

    cases = json.load(open(os.path.join(golden_dir, "sampler_golden.json")))
    assert len(cases) >= 40
    for size, world, ranges in cases:
        for rank, want in enumerate(ranges):
            b, e = D.shard_range(size, rank, world)
            if want is None:
                assert b == e
            else:
                assert [b, e] == want, (size, world, rank)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 16, dtype=torch.float32).reshape(total, 16)
        b, e = D.shard_range(total, rank, world)
        out = D.gather_rows(full[b:e].clone(), total)
        batch = {"x": full, "none": None}
        sh = D.shard_batch(batch, rank, world)
        ok = torch.equal(out, full) and sh["none"] is None and torch.equal(sh["x"], full[b:e])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 10), (2, 7), (3, 4), (2, 1)])
def test_gather_rows_gloo(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_gather_rows_single_process_is_identity():
    x = torch.randn(5, 16)
    assert D.gather_rows(x, 5) is x


def test_numa_helper_parses_cpulists_and_never_raises():
    from rdpn6d_b200.distributed import _parse_cpulist, bind_to_gpu_numa_node

    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    assert bind_to_gpu_numa_node(0) in (None, 0, 1, 2, 3, 4, 5, 6, 7)  # no GPU / no NUMA info here: leaves the process alone
