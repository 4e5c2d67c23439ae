"""Pipelined inference: overlaps the host->device copy of batch i+1 and the device->host copy of
batch i-1 with the forward of batch i, on three CUDA streams with double-buffered device slots.

This is the call a serving user makes (the reference's `inference.py:186-214` loop copies one batch to
the device, runs the model and pulls the logits back with `.numpy()`, serially):

    pipe = InferencePipeline(model, batch=16)
    for handle in pipe.run(iter_of_host_batches):     # or: h = pipe.submit(batch); ...; y = h.result()
        logits = handle.result()                       # pinned host tensor [B,256,256,32] fp32
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List

import torch

_KEYS = ("ogm", "map_img", "obs", "occ", "flow")


class Handle:
    def __init__(self, pipe: "InferencePipeline", slot: int, generation: int, event: torch.cuda.Event):
        self._pipe, self._slot, self._gen, self._event = pipe, slot, generation, event

    def result(self) -> torch.Tensor:
        """Block until this batch's logits are in host memory and return them (valid until the slot is reused).
        Raises if the slot has already been handed to a later batch: the pinned buffer then holds (or is being
        overwritten with) that batch's data."""
        if self._pipe._generation[self._slot] != self._gen:
            raise RuntimeError("InferencePipeline: this handle's slot was reused by a later submit(); consume results "
                               f"within {self._pipe.depth} submits")
        self._event.synchronize()
        return self._pipe.host_out[self._slot]


class InferencePipeline:
    def __init__(self, model, batch: int, depth: int = 2, raw_inputs: bool = False, quantized: bool = False,
                 graph: bool = True, gather: bool = False, group=None, vehicle_plane_only: bool = False):
        """raw_inputs: ogm arrives as uint8/bool and map_img as int8 (the record's own types, inference.py:91-93);
        quantized: results are the uint8 submission bytes (inference.py:160-182) instead of fp32 logits;
        graph: replay one captured CUDA graph per device slot instead of ~90 stream launches per step;
        vehicle_plane_only: `ogm` arrives as [B,S,S,11], the one plane of the record's [B,S,S,11,2] raster the model reads
        (modules.py:572): half the host->device bytes of the largest input;
        gather (data-parallel serving, one process per GPU): every step's result grids are all-gathered on every rank
        (`gathered(slot)`), by `parallel.make_gatherer` on its own stream -- copy engines over NVLink peer memory when
        available, so no SM and no host thread is taken from the forwards; collective: construct on every rank."""
        self.model, self.B, self.depth, self.graph = model, batch, depth, graph
        dev = model.device
        S = model.cfg["input_size"][0]
        shapes = {"ogm": (batch, S, S, 11) if vehicle_plane_only else (batch, S, S, 11, 2), "map_img": (batch, 256, 256, 3), "obs": (batch, 48, 11, 8),
                  "occ": (batch, 16, 11, 8), "flow": (batch, S, S, 2)}
        dts = {k: torch.float32 for k in _KEYS}
        if raw_inputs:
            dts["ogm"], dts["map_img"] = torch.uint8, torch.int8
        odt = torch.uint8 if quantized else torch.float32
        self.dev_in: List[Dict[str, torch.Tensor]] = [
            {k: torch.empty(shapes[k], dtype=dts[k], device=dev) for k in _KEYS} for _ in range(depth)]
        self.gatherer, self.gather_kind = None, "none"
        if gather:
            from .parallel import make_gatherer
            self.gatherer = make_gatherer((batch, 256, 256, 32), odt, dev, group=group, slots=depth)
            self.gather_kind = self.gatherer.kind
            self.dev_out = [self.gatherer.shard(s) for s in range(depth)]  # the forward writes straight into the shard
        else:
            self.dev_out = [torch.empty(batch, 256, 256, 32, dtype=odt, device=dev) for _ in range(depth)]
        self.ev_gathered = [None] * depth
        self.host_out = [torch.empty(batch, 256, 256, 32, dtype=odt).pin_memory() for _ in range(depth)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]       # inputs of slot landed
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]      # forward of slot done (inputs reusable)
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # logits of slot in host memory
        self._generation = [0] * depth                                # submits seen per slot (stale-handle check)
        self.i = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.dev_in[0].values())
        self.d2h_bytes = self.dev_out[0].numel() * self.dev_out[0].element_size()
        with torch.cuda.device(dev):
            model.packed()  # weights on the device before the first submit
            model._workspace(model.workspace_bytes(batch))
            torch.cuda.synchronize(dev)  # weight uploads ran on the default stream

    def submit(self, host_batch: Dict[str, torch.Tensor]) -> Handle:
        """Enqueue one batch (host tensors, ideally pinned).  Returns immediately."""
        s = self.i % self.depth
        first_use = self.i < self.depth
        self.i += 1
        with torch.cuda.stream(self.s_in):
            if not first_use:
                self.s_in.wait_event(self.ev_run[s])  # the previous forward on this slot has consumed its inputs
            for k in _KEYS:
                src = host_batch[k]
                if src.dtype == torch.bool:
                    src = src.view(torch.uint8)
                self.dev_in[s][k].copy_(src, non_blocking=True)
            self.ev_in[s].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[s])
            if not first_use:
                self.s_run.wait_event(self.ev_out[s])  # the previous logits of this slot have left the device
                if self.ev_gathered[s] is not None:
                    self.s_run.wait_event(self.ev_gathered[s])  # ... and every rank has pulled the shard
            d = self.dev_in[s]
            # fixed device slots: the forward of each slot is a replayed CUDA graph after its first use
            self.model.forward_into(self.dev_out[s], d["ogm"], d["map_img"], d["obs"], d["occ"], d["flow"], graph=self.graph)
            self.ev_run[s].record(self.s_run)
        if self.gatherer is not None:
            self.ev_gathered[s] = self.gatherer.gather(s, self.ev_run[s])
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[s])
            self.host_out[s].copy_(self.dev_out[s], non_blocking=True)
            ev = torch.cuda.Event()  # a fresh event per submit: a stale handle never waits on a later batch's copy
            ev.record(self.s_out)
            self.ev_out[s] = ev
        self._generation[s] += 1
        return Handle(self, s, self._generation[s], ev)

    def run(self, batches: Iterable[Dict[str, torch.Tensor]]) -> Iterator[Handle]:
        """Submit every batch, yielding the handle of batch i once batch i+depth-1 has been enqueued."""
        pending: List[Handle] = []
        for b in batches:
            pending.append(self.submit(b))
            if len(pending) >= self.depth:
                yield pending.pop(0)
        yield from pending

    def gathered(self, slot: int) -> torch.Tensor:
        """[world * B, 256, 256, 32] grids of the batch last submitted to `slot` (valid once its gather event has fired)."""
        if self.gatherer is None:
            raise RuntimeError("InferencePipeline was built without gather=True")
        if self.ev_gathered[slot] is not None:
            self.ev_gathered[slot].synchronize()
        return self.gatherer.full[slot]

    def synchronize(self) -> None:
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()
        if self.gatherer is not None and self.gatherer.stream is not None:
            self.gatherer.stream.synchronize()
