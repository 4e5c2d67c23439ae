"""Data-parallel inference over the GPUs of one box (BASELINE config 4): samples are independent
(no cross-sample op anywhere in STrajNet.call), so the batch is split contiguously by
rank, weights are replicated, and the ONLY collective is one all-gather of the output grids.

One process per GPU (torchrun); `torch.distributed` (NCCL over NVLink/NVSwitch, gloo in CPU tests)
is used purely as plumbing.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def bind_to_local_numa(device_index: int) -> dict:
    """Pin this process (and with it the first-touch placement of its pinned host buffers) to the CPUs of the NUMA node
    the GPU hangs off.  With one process per GPU on a two-socket box, pinned staging buffers otherwise land wherever the
    launcher started the process, and every host<->device copy of half the ranks crosses the socket interconnect.
    Call before allocating pinned memory.  Best effort: returns {} when sysfs does not say."""
    import os
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "pci": bdf}
    except Exception:  # noqa: BLE001 -- no GPU, no sysfs, restricted affinity: leave the process where it is
        return {}


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range [lo, hi) of `rank`; the global batch must divide evenly (all-gather needs equal shards)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_inputs(inputs: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    B = next(iter(inputs.values())).shape[0]
    lo, hi = shard_bounds(B, rank, world)
    return {k: v[lo:hi] for k, v in inputs.items()}


def gather_outputs(local_out: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of the path: all-gather [B/world,256,256,32] -> [B,256,256,32] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_out
    local_out = local_out.contiguous()
    full = torch.empty((world * local_out.shape[0],) + tuple(local_out.shape[1:]), dtype=local_out.dtype,
                       device=local_out.device)
    dist.all_gather_into_tensor(full, local_out, group=group)
    return full


class DataParallelSTrajNet:
    """`model(global_inputs)` on every rank -> the full [B,256,256,32] output on every rank."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    def __call__(self, ogm, map_img, training=False, obs=None, occ=None, mapt=None, flow=None):
        # inference-only wrapper: `training` defaults to False here (the model raises on True)
        if obs is None or occ is None or flow is None:
            raise ValueError("DataParallelSTrajNet: obs, occ and flow are required")
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        sh = shard_inputs(dict(ogm=ogm, map_img=map_img, obs=obs, occ=occ, flow=flow), rank, world)
        y = self.model(sh["ogm"], sh["map_img"], training=training, obs=sh["obs"], occ=sh["occ"], flow=sh["flow"])
        return gather_outputs(y, self.group)


class PeerAllGather:
    """The same single all-gather of the output grids, moved by the COPY ENGINES over NVLink peer memory.

    Why: the NCCL all-gather is a kernel; at 8 ranks each rank receives 7 x 134 MB per step and that kernel shares the SMs
    with the next step's forward for ~1/3 of the step (measured 3.02 -> 3.57 ms / step at N = 8).  Here every rank's
    forward writes its shard straight into a symmetric-memory buffer (`torch.distributed._symmetric_memory`, CUDA VMM
    handles exchanged once at construction); per step and slot the side stream then runs
        barrier  (every rank's shard of this slot is complete)
        world - 1 peer-to-peer pulls `full[r] <- shard@rank r` as plain device-to-device memcpys (copy engines; rank
                  r - 1, r - 2, ... so that no two ranks pull from the same peer at the same time)
        barrier  (every rank has finished reading: the shards of this slot may be overwritten)
    and no SM is taken from the forward.  `slots` shard buffers let the gather of step i overlap the forward of step i+1.

    Construction is collective and may fail (no peer access, symmetric memory unavailable, gloo): use `make_gatherer`,
    which agrees across ranks and falls back to the NCCL path.
    """

    def __init__(self, shard_shape, dtype, device, group=None, slots: int = 2):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.shape, self.dtype, self.device = tuple(shard_shape), dtype, torch.device(device)
        self.shards = [symm.empty(*self.shape, dtype=dtype, device=self.device) for _ in range(slots)]
        self.handles = [symm.rendezvous(t, self.group) for t in self.shards]
        self.full = [torch.empty((self.world * self.shape[0],) + self.shape[1:], dtype=dtype, device=self.device)
                     for _ in range(slots)]
        self.stream = torch.cuda.Stream(self.device)
        self.kind = "copy-engine peer pulls over symmetric memory"

    def shard(self, slot: int) -> torch.Tensor:
        """Where this rank's forward writes its output of `slot`."""
        return self.shards[slot]

    def gather(self, slot: int, after: torch.cuda.Event) -> torch.cuda.Event:
        """Enqueue the gather of `slot` behind `after` (recorded once the shard is complete); returns the event after
        which `full[slot]` holds every rank's shard AND `shard(slot)` may be overwritten."""
        h = self.handles[slot]
        chunks = self.full[slot].chunk(self.world)
        done = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(after)
            h.barrier(channel=slot)
            for step in range(self.world):
                r = (self.rank - step) % self.world
                src = self.shards[slot] if r == self.rank else h.get_buffer(r, self.shape, self.dtype)
                chunks[r].copy_(src, non_blocking=True)
            h.barrier(channel=slot)
            done.record(self.stream)
        return done


class NcclAllGather:
    """Same interface over `dist.all_gather_into_tensor` (the fallback, and the path of the gloo CPU tests)."""

    def __init__(self, shard_shape, dtype, device, group=None, slots: int = 2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        shape = tuple(shard_shape)
        self.shards = [torch.empty(shape, dtype=dtype, device=device) for _ in range(slots)]
        self.full = [torch.empty((self.world * shape[0],) + shape[1:], dtype=dtype, device=device) for _ in range(slots)]
        self.stream = torch.cuda.Stream(device) if torch.device(device).type == "cuda" else None
        self.kind = "NCCL all_gather_into_tensor" if self.stream is not None else "all_gather_into_tensor"

    def shard(self, slot: int) -> torch.Tensor:
        return self.shards[slot]

    def gather(self, slot: int, after=None):
        if self.stream is None:  # CPU / gloo
            if self.world > 1:
                dist.all_gather_into_tensor(self.full[slot], self.shards[slot], group=self.group)
            else:
                self.full[slot].copy_(self.shards[slot])
            return None
        done = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            if after is not None:
                self.stream.wait_event(after)
            if self.world > 1:
                dist.all_gather_into_tensor(self.full[slot], self.shards[slot], group=self.group)
            else:
                self.full[slot].copy_(self.shards[slot])
            done.record(self.stream)
        return done


def make_gatherer(shard_shape, dtype, device, group=None, slots: int = 2, prefer_copy_engine: bool = True):
    """Collective constructor: the copy-engine gatherer when EVERY rank can build it and it reproduces the NCCL result
    on a rank- and position-dependent pattern, else the NCCL one.  SJ_GATHER=nccl forces the fallback."""
    import os
    dev = torch.device(device)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    want = prefer_copy_engine and dev.type == "cuda" and world > 1 and os.environ.get("SJ_GATHER", "") != "nccl"
    if not want:
        return NcclAllGather(shard_shape, dtype, device, group, slots)
    # symm.rendezvous is collective: agree that every rank can get that far BEFORE any rank enters it, so that a rank
    # which cannot (module missing, no peer access) sends everybody to the NCCL path instead of leaving them hanging
    ag, ok = None, 1
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401
        if not all(torch.cuda.can_device_access_peer(dev.index, d) for d in range(torch.cuda.device_count()) if d != dev.index):
            ok = 0
    except Exception:  # noqa: BLE001
        ok = 0
    flag = torch.tensor([ok], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        return NcclAllGather(shard_shape, dtype, device, group, slots)
    try:
        ag = PeerAllGather(shard_shape, dtype, dev, group, slots)
    except Exception as e:  # noqa: BLE001 -- any failure means "use NCCL"
        import warnings
        warnings.warn(f"copy-engine all-gather unavailable on rank {dist.get_rank(group)}: {e}")
        ok = 0
    flag.fill_(ok)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        # self-check on every slot: shard = rank stamp, gathered result must equal the NCCL gather of the same shards
        rank = dist.get_rank(group)
        try:
            for s in range(slots):
                # position-dependent pattern (a constant stamp cannot see an offset / stride error inside a shard)
                pat = torch.arange(ag.shard(s).numel(), device=dev, dtype=torch.float32).remainder_(8191.0)
                ag.shard(s).copy_((pat * (rank + 1) + 10 * s).reshape(ag.shard(s).shape).to(ag.shard(s).dtype))
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                ag.gather(s, ev).synchronize()
                ref = torch.empty_like(ag.full[s])
                dist.all_gather_into_tensor(ref, ag.shard(s), group=group)
                if not torch.equal(ref, ag.full[s]):
                    ok = 0
        except Exception as e:  # noqa: BLE001
            import warnings
            warnings.warn(f"copy-engine all-gather self-check failed on rank {rank}: {e}")
            ok = 0
        flag.fill_(ok)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return ag
    return NcclAllGather(shard_shape, dtype, device, group, slots)
