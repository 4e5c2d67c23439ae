"""World-size-2 gloo tests (CPU) of the data-parallel host logic: shard bounds, the single all-gather,
and the DP invariant `gathered == single-rank result` with a stand-in per-sample function."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from strajnet_b200 import parallel as P


def test_shard_bounds():
    assert [P.shard_bounds(128, r, 8) for r in (0, 3, 7)] == [(0, 16), (48, 64), (112, 128)]
    assert P.shard_bounds(16, 0, 1) == (0, 16)
    with pytest.raises(ValueError):
        P.shard_bounds(10, 0, 4)
    with pytest.raises(ValueError):
        P.shard_bounds(8, 4, 4)
    covered = sorted(i for r in range(4) for i in range(*P.shard_bounds(12, r, 4)))
    assert covered == list(range(12))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeModel:
    """Per-sample function standing in for the CUDA forward (samples are independent, as in STrajNet)."""

    def __call__(self, ogm, map_img, training=True, obs=None, occ=None, flow=None):
        s = ogm.sum(dim=(1, 2, 3)) + 2 * obs.sum(dim=(1, 2)) + flow.mean(dim=(1, 2))
        return s[:, None, None, None].expand(-1, 4, 4, 32).contiguous()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        B = 6
        inp = dict(ogm=torch.randn(B, 3, 3, 2, generator=g), map_img=torch.randn(B, 2, generator=g),
                   obs=torch.randn(B, 4, 5, generator=g), occ=torch.randn(B, 2, generator=g),
                   flow=torch.randn(B, 3, 3, generator=g))
        dp = P.DataParallelSTrajNet(_FakeModel())
        y = dp(inp["ogm"], inp["map_img"], training=False, obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
        ref = _FakeModel()(inp["ogm"], inp["map_img"], obs=inp["obs"], occ=inp["occ"], flow=inp["flow"])
        ok = tuple(y.shape) == (B, 4, 4, 32) and torch.equal(y, ref)
        sh = P.shard_inputs(inp, rank, world)
        ok = ok and sh["ogm"].shape[0] == B // world
        # the slot-based gatherer bench.py / serving use: on CPU (gloo) make_gatherer must give the plain all-gather path
        ag = P.make_gatherer((3, 4, 4, 32), torch.float32, "cpu", slots=2)
        ok = ok and isinstance(ag, P.NcclAllGather)
        for s in range(2):
            ag.shard(s).fill_(float(10 * s + rank + 1))
            ag.gather(s)
            want = torch.cat([torch.full((3, 4, 4, 32), float(10 * s + r + 1)) for r in range(world)])
            ok = ok and torch.equal(ag.full[s], want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_dp_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_data_parallel_wrapper_defaults_and_validation():
    """The wrapper is inference-only: training defaults to False (the model raises on True), and missing inputs are
    reported before any sharding."""
    import inspect
    import pytest
    from strajnet_b200.parallel import DataParallelSTrajNet
    assert inspect.signature(DataParallelSTrajNet.__call__).parameters["training"].default is False
    dp = DataParallelSTrajNet(model=lambda *a, **k: None)
    with pytest.raises(ValueError):
        dp(None, None, obs=None, occ=None, flow=None)
