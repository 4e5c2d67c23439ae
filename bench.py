#!/usr/bin/env python
"""Benchmark of the occupancy-flow forward path (BASELINE.json metric: frames/sec @ 256x256x8, batch 16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3|5|2|1]

One "step" = one forward of STrajNet over one batch of synthetic inputs per GPU.  --config selects the BASELINE.json
configuration: 3 (default, headline: cfg256 + FG-MSA, batch 16, bf16), 5 (cfg512 / large_ogm, batch 4, bf16), 2 (cfg256,
batch 1, fp32: the parity configuration), 1 (a single SwinTransformerBlock on a 64x64x32 grid, CPU time beside it).
N > 1 is launched by torchrun, one rank per GPU; batches shard data-parallel (weak scaling, 16 frames per rank) with a
single all-gather on the output grids per step (BASELINE config 4).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG256 = dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12])
CFG512 = dict(input_size=(512, 512), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12])
METRIC = "occupancy_flow_frames_per_sec"
UNIT = "frames/s"
# BASELINE.json configs that fit one GPU (config 4 = config 3 sharded: --gpus N).  GFLOP per frame: nominal, reference
# formulation.
CONFIGS = {
    3: dict(cfg=CFG256, S=256, large_ogm=False, batch=16, dtype="bf16", gflop=200.75,
            name="STrajNet cfg256 (window 8, dims 96/192/384, depths 2/2/2) fg_msa+fg forward -> [B,256,256,32], 8 waypoints, "
                 "BASELINE config 3 (headline)"),
    5: dict(cfg=CFG512, S=512, large_ogm=True, batch=4, dtype="bf16", gflop=226.00,
            name="STrajNet cfg512 (512x512 rasters, large_ogm) fg_msa+fg forward -> [B,256,256,32], 8 waypoints, "
                 "BASELINE config 5 (high-res stress)"),
    2: dict(cfg=CFG256, S=256, large_ogm=False, batch=1, dtype="fp32", gflop=200.75,
            name="STrajNet cfg256 fg_msa+fg forward, batch 1, fp32 (BASELINE config 2: the 1e-3 parity configuration)"),
}
# dominant kernel: the 96->48 up-convolution @256^2 x 8 waypoints fused with the head projection (tc_upconv4h_kernel), ONE
# of its two launches per step.  Algorithmic FLOPs per frame: nominal = the reference's formulation (nearest x2 upsample +
# 3x3 conv 96->48, 9 taps per output pixel, + the 3x3 48->2 head); executed = after the exact sub-pixel
# folding (4 taps per output pixel) + the 48->18 pointwise projection.
PROBE_ROLE = "dec.upconv3"
PROBE_KERNEL = "tc_upconv4h_kernel"
PROBE_GFLOP_PER_FRAME_NOMINAL = 2 * 8 * 65536 * (9 * 96 * 48 + 9 * 48 * 2) / 1e9
PROBE_GFLOP_PER_FRAME = 2 * 8 * 65536 * (4 * 96 * 48 + 48 * 18) / 1e9
# algorithmic HBM bytes of that launch per frame: bf16 input [8,128,128,96] read once, fp16 projected columns [8,256,256,18] written
PROBE_BYTES_PER_FRAME = 8 * (128 * 128 * 96 * 2 + 256 * 256 * 18 * 2)
# ncu numbers committed under profiles/ (per launch at batch 16): DRAM traffic of the dominant kernel, tensor-pipe % of K1
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")


def ncu_summary():
    """{'probe_dram_bytes_b16': ..., 'k1_tensor_pipe_pct': ..., 'source': ...} from the committed ncu capture, or {}."""
    try:
        with open(NCU_SUMMARY) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def synth_inputs(B, S=256, seed=0):
    """Synthetic inputs with the value ranges of the reference's record decode (inference.py:84-96)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ogm = (rng.random((B, S, S, 11, 2), dtype=np.float32) < 0.03).astype(np.float32)
    map_img = rng.integers(-128, 128, size=(B, 256, 256, 3)).astype(np.float32) / 256.0
    flow = ((rng.random((B, S, S, 2), dtype=np.float32) < 0.03) * rng.uniform(-20, 20, size=(B, S, S, 2))).astype(np.float32)

    def actors(n_max, lo):
        a = np.zeros((B, n_max, 11, 8), np.float32)
        for b in range(B):
            for i in range(int(rng.integers(lo, n_max + 1))):
                st = int(rng.integers(1, 12))
                xy = rng.uniform(-80, 80, size=(st, 2))
                xy[xy == 0] = 1.0
                a[b, i, 11 - st:, 0:2] = xy
                a[b, i, 11 - st:, 2:4] = rng.normal(0, 5, size=(st, 2))
                a[b, i, 11 - st:, 4] = rng.uniform(-np.pi, np.pi, size=st)
                a[b, i, :, 5 + int(rng.integers(0, 3))] = 1.0
        return a

    d = dict(ogm=ogm, map_img=map_img, flow=flow, obs=actors(48, 4), occ=actors(16, 0))
    return {k: torch.from_numpy(v) for k, v in d.items()}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_fps(batch, iters, warmup=1, config=3):
    """The CPU restatement of the reference graph (oracle/), fp32, all host threads."""
    from oracle import strajnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cf = CONFIGS[config]
    w = O.make_weights(cf["cfg"], seed=0)
    inp = O.make_inputs(batch, cf["S"], seed=0)
    kw = dict(large_ogm=cf["large_ogm"])
    with torch.no_grad():
        for _ in range(warmup):
            O.forward_from_inputs(w, cf["cfg"], inp, **kw)
        t0 = time.perf_counter()
        for _ in range(iters):
            O.forward_from_inputs(w, cf["cfg"], inp, **kw)
        dt = time.perf_counter() - t0
    return batch * iters / dt, cores, dt / iters


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  TensorFlow is not installable here
    (SURVEY §8c), so this is the oracle port (kind "port"), on all host cores, rank 0 only.  Each step is one forward
    of the SAME batch the GPU arm runs per GPU (16 frames for config 3): about 3.5 s per step on 16 cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.config == 1:
        return run_block_config(args, reference_only=True)
    cf = CONFIGS[args.config]
    batch = cf["batch"]
    fps, cores, s_per_step = cpu_oracle_fps(batch, args.steps, 1, args.config)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cf["name"], "batch_per_gpu": batch,
                   "sample": f"one forward of batch {batch} per step on the host CPU (1 untimed warm-up forward)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} forwards of batch {batch}, torch-CPU fp32 oracle (not TF: not installable)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_block_config(args, reference_only=False):
    """BASELINE config 1: one SwinTransformerBlock(dim 32, 64x64 tokens, 2 heads, window 8, shift 0 and 4), batch 1, fp32,
    GPU time (CUDA events) next to the CPU oracle on the same host (per-block CPU-vs-GPU timing)."""
    from oracle import strajnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).standard_normal((1, 4096, 32)).astype(np.float32))
    out = {}
    for shift in (0, 4):
        w = O.make_block_weights(32, 2, seed=0)
        with torch.no_grad():
            O.swin_block(x, w, "", 64, 64, 2, 8, shift)
            t0 = time.perf_counter()
            n_cpu = max(args.steps, 20)
            for _ in range(n_cpu):
                ref = O.swin_block(x, w, "", 64, 64, 2, 8, shift)
            cpu_ms = (time.perf_counter() - t0) / n_cpu * 1e3
        rec = {"cpu_ms": cpu_ms}
        if not reference_only:
            import strajnet_b200 as sj
            blk = sj.SwinTransformerBlock(32, (64, 64), 2, window_size=8, shift_size=shift)
            blk.set_weights(w)
            xd = x.cuda()
            for _ in range(max(args.warmup, 3)):
                y = blk(xd)
            torch.cuda.synchronize()
            n_gpu = max(args.steps, 20) * 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_gpu):
                y = blk(xd)
            e1.record()
            torch.cuda.synchronize()
            rec.update(gpu_ms=e0.elapsed_time(e1) / n_gpu, max_abs_err=(y.cpu() - ref).abs().max().item())
        out[f"shift{shift}"] = rec
    cpu_ms = float(np.mean([r["cpu_ms"] for r in out.values()]))
    line = {"metric": "swin_block_forwards_per_sec", "unit": "blocks/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 1: SwinTransformerBlock(dim 32, 64x64, 2 heads, window 8, shift 0/4), "
                                   "x [1,4096,32] fp32; mean over the two shifts; 134.2 MFLOP per call (SURVEY 8d)"},
            "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "blocks/s", "cores": cores, "kind": "port",
                             "sample": "20+ calls of the torch-CPU oracle block per shift"},
            "per_shift": out}
    if reference_only:
        line.update(impl="reference", value=1e3 / cpu_ms, ms_per_step=cpu_ms,
                    e2e={"value": 1e3 / cpu_ms, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
    else:
        gpu_ms = float(np.mean([r["gpu_ms"] for r in out.values()]))
        line.update(value=1e3 / gpu_ms, ms_per_step=gpu_ms, gpu_launches=None,
                    note="latency of one eager call through the Python layer (launch-bound: ~10 kernels of a few us each)")
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 5], help="BASELINE.json configuration (4 = 3 with --gpus N)")
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp32"], help="override the configuration's arithmetic type")
    ap.add_argument("--batch", type=int, default=None, help="override the configuration's frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 1:
        return run_block_config(args)
    cf = dict(CONFIGS[args.config])
    if args.dtype:
        cf["dtype"] = args.dtype
    if args.batch:
        cf["batch"] = args.batch
    args.dtype = cf["dtype"]
    # stdout carries exactly ONE JSON line: anything native libraries print there (e.g. "NCCL version ...") is sent to
    # stderr by pointing fd 1 at fd 2 for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    import strajnet_b200 as sj
    from strajnet_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = {}
    if world > 1 and os.environ.get("SJ_NO_NUMA_BIND") is None:
        from strajnet_b200.parallel import bind_to_local_numa
        numa = bind_to_local_numa(local)  # before any pinned allocation: staging buffers land on the GPU's own node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    B, S = cf["batch"], cf["S"]
    dtype = "bfloat16" if args.dtype == "bf16" else "float32"

    model = sj.STrajNet(cf["cfg"], fg_msa=True, fg=True, large_ogm=cf["large_ogm"], dtype=dtype, device=dev)
    model.build()  # random-init weights of the reference architecture (Keras default initialisers, seeded)
    host = {k: v.pin_memory() for k, v in synth_inputs(B, S, seed=rank).items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    out = torch.empty(B, 256, 256, 32, dtype=torch.float32, device=dev)

    from strajnet_b200.parallel import make_gatherer
    from strajnet_b200.pipeline import InferencePipeline

    # N > 1: the single collective of the path (config 4) runs on a side stream over double-buffered output
    # grids, so the all-gather of step i overlaps the forward of step i+1.  The gatherer moves the shards with the copy
    # engines over NVLink peer memory when every rank can (strajnet_b200/parallel.py), else with NCCL.
    gatherer = make_gatherer(tuple(out.shape), out.dtype, dev, slots=2) if world > 1 else None
    outs = [gatherer.shard(0), gatherer.shard(1)] if world > 1 else [out]
    ev_done = [torch.cuda.Event() for _ in outs]
    ev_gathered = [None for _ in outs]
    state = {"k": 0}

    def step_resident(graph=True):
        # the ~90 launches of one forward over fixed buffers are replayed as a CUDA graph (the serving path does the
        # same per device slot); graph=False issues them as plain stream launches
        s = state["k"] % len(outs)
        state["k"] += 1
        cur = torch.cuda.current_stream()
        if world > 1 and ev_gathered[s] is not None:
            cur.wait_event(ev_gathered[s])  # the previous gather of this slot has read the shard everywhere
        model.forward_into(outs[s], devin["ogm"], devin["map_img"], devin["obs"], devin["occ"], devin["flow"], graph=graph)
        if world > 1:
            ev_done[s].record(cur)
            ev_gathered[s] = gatherer.gather(s, ev_done[s])

    def drain():
        if world > 1:
            torch.cuda.current_stream().wait_stream(gatherer.stream)

    # end to end through the public serving API: every step copies its inputs from pinned host memory to the
    # device and its results back to pinned host memory; copies of neighbouring steps overlap the forward.
    # primary e2e: the record's own input types (bool/uint8 rasters, int8 map; inference.py:91-93) and the fused
    # submission quantisation (uint8/int8 grids; inference.py:160-182) -- what the reference's serving loop moves
    # over PCIe after its host-side casts.  The fp32-in / fp32-logits-out variant is reported beside it.
    # N > 1: the pipeline gathers the result grids of every step on every rank (its own side stream, copy engines).
    host_raw = dict(host)
    # the vehicle plane of the record's [S,S,11,2] raster is the only one the model reads (modules.py:572): the record
    # decoder hands over that plane alone (records.decode_example(vehicle_plane_only=True))
    host_raw["ogm"] = (host["ogm"][..., 0] != 0).to(torch.uint8).contiguous().pin_memory()
    host_raw["map_img"] = torch.round(host["map_img"] * 256).to(torch.int8).pin_memory()
    depth = int(os.environ.get("SJ_E2E_DEPTH", "2"))
    e2e_gather = world > 1 and os.environ.get("SJ_E2E_NO_GATHER") is None  # diagnosis knob: e2e without the all-gather
    pipes = {"raw": (InferencePipeline(model, B, depth=depth, raw_inputs=True, quantized=True, gather=e2e_gather,
                                       vehicle_plane_only=True), host_raw),
             "fp32": (InferencePipeline(model, B, depth=depth, gather=e2e_gather), host)}
    pipe, host_e2e = pipes["raw"]
    pending = []

    def step_e2e():
        pending.append(pipe.submit(host_e2e))
        if len(pending) > depth - 1:
            pending.pop(0).result()  # consume the oldest batch while the newer ones run

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        drain()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    run_dev = torch.cuda.Stream(dev)  # graphs are captured / replayed on a non-default stream
    torch.cuda.set_stream(run_dev)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    lib.sj_launch_count(1)
    step_resident(graph=False)
    launches_per_step = lib.sj_launch_count(1)

    import ctypes
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
        # per-kernel CUDA events cannot be recorded inside a replayed graph: the dominant kernel is timed live in a
        # second pass of the same K steps issued as plain stream launches (its own duration does not depend on how it
        # was launched); that pass also gives the un-graphed step time
        lib.sj_probe_start(PROBE_ROLE.encode())
        ms_eager = timed(lambda: step_resident(graph=False), args.steps)
        pms, pn = ctypes.c_double(0), ctypes.c_int(0)
        lib.sj_probe_stop(ctypes.byref(pms), ctypes.byref(pn))
    fps = world * B * args.steps / (ms / 1e3)

    # ---- DP invariant: the gathered grids == what ONE GPU computes for the same samples -----------------
    dp = None
    if world > 1:
        barrier()
        s = 0
        ev = torch.cuda.Event()
        model.forward_into(gatherer.shard(s), devin["ogm"], devin["map_img"], devin["obs"], devin["occ"], devin["flow"])
        ev.record(torch.cuda.current_stream())
        gatherer.gather(s, ev).synchronize()
        full = gatherer.full[s]
        # (i) every chunk r of the gathered tensor carries rank r's bytes: order-sensitive checksums, exchanged with NCCL
        def checksum(t):
            v = t.reshape(-1).view(torch.int32).to(torch.int64)
            idx = torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 65521 + 1
            return torch.stack([(v * idx).sum(), v.sum()])
        mine = checksum(gatherer.shard(s))
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        chunks_ok = all(torch.equal(checksum(full[r * B:(r + 1) * B]), allc[r]) for r in range(world))
        # (ii) chunk (rank + 1) % world recomputed HERE, on one GPU, from that rank's inputs: bit-identical
        peer = (rank + 1) % world
        pin = {k: v.to(dev) for k, v in synth_inputs(B, S, seed=peer).items()}
        single = torch.empty_like(out)
        model.forward_into(single, pin["ogm"], pin["map_img"], pin["obs"], pin["occ"], pin["flow"])
        torch.cuda.synchronize()
        same = torch.equal(single, full[peer * B:(peer + 1) * B])
        flag = torch.tensor([int(chunks_ok and same)], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dp = {"dp_invariant_ok": bool(flag.item()), "checked": "every gathered chunk r == rank r's shard (position-weighted "
              "checksums); chunk (rank+1) % N == a single-GPU forward of that rank's inputs on this GPU, bit-exact; all ranks agree"}
        del pin, single

    for _ in range(3):
        step_e2e()
    pipe.synchronize()

    def e2e_all():
        for _ in range(args.steps):
            step_e2e()
        while pending:
            pending.pop(0).result()
        pipe.synchronize()

    def time_e2e():
        barrier()
        t0 = time.perf_counter()  # several streams: bracket with host clocks around full synchronisation
        e2e_all()
        barrier()
        ms_t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        return ms_t.item()

    ms_e2e = time_e2e()
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_gather_kind = pipe.gather_kind
    pipe, host_e2e = pipes["fp32"]
    for _ in range(3):
        step_e2e()
    while pending:
        pending.pop(0).result()
    pipe.synchronize()
    ms_e2e_fp32 = time_e2e()
    h2d_fp32, d2h_fp32 = pipe.h2d_bytes, pipe.d2h_bytes
    fps_e2e = world * B * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        ncu = ncu_summary()
        # which measured peak: MEASURED_PEAKS.json holds a burst figure (best of 10 launches at full clock) and a sustained
        # one (4 s back to back, power-limited clock).  A timed region shorter than 1 s never reaches the power-limited
        # state (the clock sample below shows it), so the burst peak is the comparator; longer regions use the sustained one.
        region_s = ms / 1e3
        use_burst = region_s < 1.0
        peak_tf = peaks.get("bf16_tflops", 1590.0) if use_burst else peaks.get("bf16_tflops_sustained", 1400.0)
        peak_kind = "burst bf16 (timed region %.2f s < 1 s)" % region_s if use_burst else "sustained bf16 (timed region %.1f s)" % region_s
        roof = None
        if pn.value > 0 and args.dtype == "bf16":
            per_launch_ms = pms.value / pn.value
            ach = PROBE_GFLOP_PER_FRAME * B / per_launch_ms  # GFLOP / ms = TFLOP/s
            ach_nom = PROBE_GFLOP_PER_FRAME_NOMINAL * B / per_launch_ms
            traffic = ncu.get("probe_dram_bytes_b16")
            roof = {"bound": "tensor", "kernel": PROBE_ROLE + " (" + PROBE_KERNEL + ")", "achieved": ach, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "traffic": traffic * B / 16 if traffic else None,
                    "traffic_source": ncu.get("source") if traffic else None,
                    "peak_source": peak_src + ", " + peak_kind,
                    "launch_ms": per_launch_ms, "launches_timed": pn.value,
                    "timed_in": "second pass of the same steps as plain stream launches (events cannot sit inside the replayed graph)",
                    "flops_counted": "executed = algorithmic after the exact sub-pixel folding (4 of 9 taps per output pixel) + the 48->18 head projection",
                    "algorithmic_gflop_per_launch": PROBE_GFLOP_PER_FRAME * B,
                    "frac_of_sustained_peak": ach / peaks.get("bf16_tflops_sustained", 1400.0),
                    "nominal_gflop_per_launch": PROBE_GFLOP_PER_FRAME_NOMINAL * B, "nominal_tflops": ach_nom,
                    "algorithmic_dram_bytes_per_launch": PROBE_BYTES_PER_FRAME * B,
                    "hbm_frac": PROBE_BYTES_PER_FRAME * B / (per_launch_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                    "forward_nominal_tflops": fps / world * cf["gflop"] / 1e3,
                    "forward_nominal_frac": fps / world * cf["gflop"] / 1e3 / peak_tf,
                    "k1_tensor_pipe_pct": ncu.get("k1_tensor_pipe_pct"), "k1_source": ncu.get("k1_source")}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb = min(B, 16)
            cfps, cores, sps = cpu_oracle_fps(cb, 2, 1, args.config)
            cpu = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"2 forwards of batch {cb} (after 1 warm-up), torch-CPU fp32 restatement of the reference graph "
                             "(oracle/); TensorFlow is not installable here"}
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "ms_per_step_stream_launches": ms_eager / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": cf["name"] + (" / config 4 sharding" if world > 1 else ""),
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
                       "weights": "random init (Keras default initialisers)",
                       "launch": "one CUDA graph replay per step (captured from the library's stream launches)",
                       "l2": "per-step inputs and activations (> 1 GB at batch 16) exceed the 126 MB L2; no explicit flush",
                       "collective": (f"one all-gather of the fp32 output grids per step ({gatherer.kind}), on a side stream "
                                      "overlapping the next forward") if world > 1 else "none"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "InferencePipeline(raw_inputs, quantized, vehicle_plane_only): pinned host -> device (uint8 vehicle-plane raster [B,S,S,11], int8 map, fp32 flow/actors), forward, fused submission quantisation, device -> pinned host (uint8 grids) every "
                           "step; 3 streams, double-buffered; synchronised wall clock, max over ranks"
                           + (f"; every step's uint8 grids all-gathered on every rank ({e2e_gather_kind})" if e2e_gather else ""),
                    "fp32_io": {"value": world * B * args.steps / (ms_e2e_fp32 / 1e3), "unit": UNIT,
                                "h2d_bytes_per_step": h2d_fp32, "d2h_bytes_per_step": d2h_fp32}},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks.summary(),
        }
        if dp is not None:
            line.update(dp)
        if numa:
            line["host_binding"] = numa
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
