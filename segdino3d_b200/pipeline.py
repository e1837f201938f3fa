"""Host-fed scene streaming: the end-to-end form of the lifting path.

A scene arrives in (pinned) host memory -- points, cameras, depth, DINO-X maps, superpoint ids -- and the
results (points_2dfeats, count, superpoint features) are wanted back in host memory, which is what the
reference's offline lifter / dataset loader pair does through ``features_2d/{scene}.pth``
(/root/reference/segdino3d/datasets/dataset/scannet200.py:219-234). Over PCIe the copies dominate
(~248 MB in, ~103 MB out per ScanNet-sized scene against ~0.35 ms of kernels), so the pipeline keeps
three CUDA streams busy at once:

    copy-in stream : H2D of scene i+1        (pinned -> device slot buffers)
    compute stream : plan + lift of scene i  (libsd3d kernels)
    copy-out stream: D2H of scene i-1        (device -> pinned slot buffers)

with ``depth`` slots of device/host buffers and CUDA events between the stages. PCIe is full duplex, so
steady-state throughput approaches max(H2D, D2H) per scene instead of their sum.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Tuple

import torch

from . import ops

_INPUT_KEYS = ("xyz", "K", "w2c", "depth", "fmap", "sp_ids")


class _Slot:
    def __init__(self):
        self.dev: Dict[str, torch.Tensor] = {}
        self.out_dev = None          # keeps the device outputs alive until their D2H has completed
        self.host: Dict[str, torch.Tensor] = {}
        self.ev_h2d = torch.cuda.Event()
        self.ev_comp = torch.cuda.Event()
        self.ev_d2h = torch.cuda.Event()
        self.busy = False
        self.meta = None
        self.buffers = None          # ops.LiftPoolBuffers of the scene shape this slot last held


class ScenePipeline:
    """``for feat, count, sp_feat in ScenePipeline(dev).run(scenes)`` where every scene is a dict with the
    pinned host tensors ``xyz, K, w2c, depth, fmap, sp_ids`` plus ``n_superpoints`` and ``stride``.
    The yielded tensors are pinned host buffers owned by the pipeline; they stay valid until ``depth`` more
    scenes have been yielded (copy them if they must outlive that)."""

    def __init__(self, device: torch.device, depth: int = 3, run: int = ops.DEFAULT_RUN, variant: int = 0,
                 tau: float = ops.TAU_DEFAULT, z_near: float = ops.Z_NEAR_DEFAULT):
        if torch.device(device).type != "cuda":
            raise ops.Sd3dError("ScenePipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.depth, self.run_len, self.variant, self.tau, self.z_near = depth, run, variant, tau, z_near
        with torch.cuda.device(self.device):
            self.s_in = torch.cuda.Stream()
            self.s_comp = torch.cuda.Stream()
            self.s_out = torch.cuda.Stream()
            self.slots = [_Slot() for _ in range(depth)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # -- stages ------------------------------------------------------------------------------------------
    def _stage_in(self, slot: _Slot, scene: Dict) -> None:
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(slot.ev_comp)  # the previous occupant's kernels have finished reading
            nbytes = 0
            for k in _INPUT_KEYS:
                src = scene[k]
                dst = slot.dev.get(k)
                if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                    dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                    slot.dev[k] = dst
                dst.copy_(src, non_blocking=True)
                nbytes += src.numel() * src.element_size()
            slot.ev_h2d.record(self.s_in)
        slot.meta = (int(scene["n_superpoints"]), float(scene["stride"]))
        self.h2d_bytes = nbytes

    def _stage_compute(self, slot: _Slot) -> None:
        n_sp, stride = slot.meta
        d = slot.dev
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(slot.ev_h2d)
            self.s_comp.wait_event(slot.ev_d2h)  # the previous outputs of this slot have left the device
            key = (d["xyz"].shape[0], d["K"].shape[0], d["fmap"].shape[3], n_sp, self.run_len, self.device)
            if slot.buffers is None or slot.buffers.key != key:  # outputs + workspace live with the slot
                slot.buffers = ops.LiftPoolBuffers(*key)
            feat, count, sp, plan = ops.lift_and_pool(d["xyz"], d["K"], d["w2c"], d["depth"], d["fmap"], d["sp_ids"], n_sp,
                                                      stride=stride, tau=self.tau, z_near=self.z_near, run=self.run_len,
                                                      variant=self.variant, overlap=True, buffers=slot.buffers)
            slot.out_dev = (feat, count, sp, plan)
            slot.ev_comp.record(self.s_comp)

    def _stage_out(self, slot: _Slot) -> None:
        feat, count, sp, _ = slot.out_dev
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot.ev_comp)
            nbytes = 0
            for name, t in (("feat", feat), ("count", count), ("sp_feat", sp)):
                h = slot.host.get(name)
                if h is None or h.shape != t.shape or h.dtype != t.dtype:
                    h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    slot.host[name] = h
                h.copy_(t, non_blocking=True)
                nbytes += t.numel() * t.element_size()
            slot.ev_d2h.record(self.s_out)
        self.d2h_bytes = nbytes

    def _collect(self, slot: _Slot) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        slot.ev_d2h.synchronize()
        slot.busy = False
        return slot.host["feat"], slot.host["count"], slot.host["sp_feat"]

    # -- driver ------------------------------------------------------------------------------------------
    def run(self, scenes: Iterable[Dict]) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Feeds the scenes through the three stages; yields results in submission order."""
        pending = []
        i = 0
        for scene in scenes:
            slot = self.slots[i % self.depth]
            if slot.busy:  # ring is full: hand out the oldest result first
                yield self._collect(pending.pop(0))
            slot.busy = True
            self._stage_in(slot, scene)
            self._stage_compute(slot)
            self._stage_out(slot)
            pending.append(slot)
            i += 1
        while pending:
            yield self._collect(pending.pop(0))
