"""Drop-in hooks with the reference's call patterns (citations relative to /root/reference).

The reference modules cannot be imported here (mmengine / spconv / MinkowskiEngine / torch_scatter are
not installed), so the hooks are written against the exact tensor contracts of their call sites:

* :func:`install_torch_scatter_shim` -- makes ``from torch_scatter import scatter_mean``
  (segdino3d/models/backbone/spconvunet.py:17, minkunet.py:16) resolve to :func:`ops.scatter_mean`.
* :func:`batch_superpoint_ids` / :func:`pool_superpoints` -- the id-offset batching + pooling + split of
  ``SpConvUNet.forward_wrapper`` (spconvunet.py:365-373,390-395) and ``Res16UNetBase.forward_wrapper``
  (minkunet.py:634-650) with ONE sort shared by every pooled tensor of the batch.
* :func:`forward_head_masks` -- the mask part of ``ScanNetQueryDecoder._forward_head``
  (decoder/instance_seg_3d_decoder.py:567-574): logits + attention masks per scene.
* :class:`PointFeatureLifter` -- fills ``targets[i]["extra_features"]["points_2dfeats"]``, the slot the
  backbones read at spconvunet.py:378 / minkunet.py:612-618 (the reference loads it from disk,
  datasets/dataset/scannet200.py:219-234).
INTEGRATION.md shows the three-line patches a maintainer applies on the reference side.
"""
from __future__ import annotations

import sys
import types
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops


def install_torch_scatter_shim(force: bool = False) -> types.ModuleType:
    """Register a module named ``torch_scatter`` whose ``scatter_mean`` is the sm_100a kernel path."""
    if "torch_scatter" in sys.modules and not force:
        mod = sys.modules["torch_scatter"]
        if getattr(mod, "__sd3d_shim__", False):
            return mod
        raise RuntimeError("a real torch_scatter is already imported; pass force=True to override it")
    mod = types.ModuleType("torch_scatter")
    mod.__sd3d_shim__ = True
    mod.scatter_mean = ops.scatter_mean
    sys.modules["torch_scatter"] = mod
    return mod


def batch_superpoint_ids(targets: Sequence[Dict]) -> Tuple[torch.Tensor, List[int]]:
    """spconvunet.py:365-373: per-scene ids + running ``max()+1`` bias -> (hstack ids, batch_offsets)."""
    batch_offsets = [0]
    bias = 0
    ids = []
    for tgt in targets:
        sp = tgt["extra_features"]["super_point_masks"].clone()
        sp += bias
        bias = int(sp.max().item()) + 1
        batch_offsets.append(bias)
        ids.append(sp)
    return torch.hstack(ids), batch_offsets


def pool_superpoints(tensors: Sequence[torch.Tensor], sp_pts_masks: torch.Tensor, batch_offsets: Sequence[int],
                     exact: bool = True) -> List[List[torch.Tensor]]:
    """``scatter_mean(t, sp_pts_masks, dim=0)`` for every tensor in ``tensors`` (backbone features,
    DINO-X features, coordinates: spconvunet.py:390,392,325) sharing one sort, then the per-scene split of
    spconvunet.py:393-395. Returns, per input tensor, the list over scenes of ``[S_i, C]``."""
    plan = ops.sp_sort(sp_pts_masks, batch_offsets[-1])
    out = []
    for t in tensors:
        # differentiable: x.features[inverse_mapping] requires grad in training (spconvunet.py:390)
        pooled = ops.sp_mean_autograd(t.float().contiguous(), sp_pts_masks, plan, exact=exact)
        out.append([pooled[batch_offsets[i]: batch_offsets[i + 1]] for i in range(len(batch_offsets) - 1)])
    return out


def forward_head_masks(norm_queries: Sequence[torch.Tensor], mask_feats: Sequence[torch.Tensor],
                       mask_attention_threshold: Optional[float], precision: str = "fp32"):
    """instance_seg_3d_decoder.py:567-574 for a batch given as python lists (one entry per scene):
    returns (pred_masks, attn_masks or None)."""
    if not (torch.is_grad_enabled() and any(t.requires_grad for t in list(norm_queries) + list(mask_feats))):
        # inference: every scene of the batch in one launch, row reset fused into the GEMM epilogue
        pred, attn = ops.mask_logits_batched(norm_queries, mask_feats, precision=precision,
                                             threshold=mask_attention_threshold)
        return pred, ([a.detach() for a in attn] if attn is not None else None)
    pred_masks, attn_masks = [], []
    for q, mf in zip(norm_queries, mask_feats):
        if mask_attention_threshold is not None:
            pm, am = ops.mask_logits(q, mf, precision=precision, threshold=mask_attention_threshold)
            attn_masks.append(am.detach())
        else:
            pm = ops.mask_logits(q, mf, precision=precision)
        pred_masks.append(pm)
    return pred_masks, (attn_masks if mask_attention_threshold is not None else None)


def fused_out_norm(queries: Sequence[torch.Tensor], norm: torch.nn.LayerNorm) -> List[torch.Tensor]:
    """``self.out_norm(queries[i])`` (instance_seg_3d_decoder.py:558) for every scene of the batch in ONE launch of the
    mask head's operand producer (``sd3d_layernorm_cast`` over the concatenated query rows; the per-scene results are
    views of its output). Under autograd, on CPU tensors or for a norm that is not a plain last-dim LayerNorm it is the
    module itself, scene by scene."""
    queries = list(queries)
    plain = isinstance(norm, torch.nn.LayerNorm) and len(norm.normalized_shape) == 1 and norm.elementwise_affine
    needs_grad = torch.is_grad_enabled() and (any(q.requires_grad for q in queries) or norm.weight.requires_grad)
    if not plain or needs_grad or not queries or any((not q.is_cuda) or q.dtype != torch.float32 or q.dim() != 2
                                                     for q in queries) or queries[0].shape[1] > 1024:
        return [norm(q) for q in queries]
    rows = torch.cat(queries) if len(queries) > 1 else queries[0]
    y32, _ = ops.layernorm_cast(rows, norm.weight, norm.bias, norm.eps, normalize=True, want_f32=True, want_bf16=False)
    return list(y32.split([q.shape[0] for q in queries]))


class PointFeatureLifter:
    """Pre-backbone hook: lifts DINO-X maps to per-point features on the GPU and stores them where the
    reference expects the precomputed ones.

    ``views[i]`` is a dict with ``K [V,4]``, ``w2c [V,3,4]``, ``depth [V,Hd,Wd]`` and ``fmaps`` (list over
    scales of channels-last ``[V,Hl,Wl,C]``); ``samples[i][:, :3]`` must be the RAW world xyz
    (augmentation moves xyz after lifting in the reference pipeline)."""

    def __init__(self, tau: float = ops.TAU_DEFAULT, z_near: float = ops.Z_NEAR_DEFAULT,
                 mode_fuse_multi_scale_2d_feats: str = "mean"):
        if mode_fuse_multi_scale_2d_feats != "mean":
            raise NotImplementedError(mode_fuse_multi_scale_2d_feats)  # scannet200.py:235-236
        self.tau, self.z_near = tau, z_near

    def __call__(self, samples: Sequence[torch.Tensor], targets: Sequence[Dict], views: Sequence[Dict]):
        for pts, tgt, vw in zip(samples, targets, views):
            xyz = pts[:, :3].contiguous()
            sp = tgt["extra_features"].get("super_point_masks")
            plan = ops.sp_sort(sp, xyz=xyz) if sp is not None else None
            feats = ops.lift_features(xyz, vw["K"], vw["w2c"], vw["depth"], vw["fmaps"], tau=self.tau,
                                      z_near=self.z_near, strides=vw.get("strides"), order=plan)
            tgt["extra_features"]["points_2dfeats"] = feats[0] if len(feats) == 1 else ops.scale_mean(feats)
        return targets


# ---------------------------------------------------------------------------------------------------------------
# registered drop-in classes (selected by `type=` in the reference's configs)
# ---------------------------------------------------------------------------------------------------------------
def register_dropins(registry_module=None):
    """Registers ``SpConvUNetB200`` / ``Res16UNet34CB200`` / ``ScanNetQueryDecoderB200`` in the reference's mmengine
    registries (segdino3d/builder.py:3-82), so that a config selects the sm_100a path with ``type='SpConvUNetB200'``
    etc. -- same constructors, same forward signatures, same tensor layouts as the classes they derive from
    (spconvunet.py:102, minkunet.py:692, decoder/instance_seg_3d_decoder.py:437).

    * the backbones inherit ``forward_wrapper`` UNCHANGED; the module-level ``scatter_mean`` they call
      (spconvunet.py:17, minkunet.py:16) is rebound to :func:`ops.scatter_mean` (CUDA tensors; CPU tensors keep the
      original function, so dataset workers are unaffected);
    * the decoder overrides only ``_forward_head`` (instance_seg_3d_decoder.py:532-577): heads as in the parent, the
      einsum + attention-mask epilogue of :567-573 through one batched launch (:func:`forward_head_masks`), ``out_norm``
      (:558) for all scenes in one launch of the operand producer (:func:`fused_out_norm`).

    Needs the reference package and its dependencies to be importable (mmengine, spconv, MinkowskiEngine); raises
    ImportError otherwise. Returns the dict of registered classes."""
    import importlib
    builder = registry_module or importlib.import_module("segdino3d.builder")
    sp_mod = importlib.import_module("segdino3d.models.backbone.spconvunet")
    mk_mod = importlib.import_module("segdino3d.models.backbone.minkunet")
    dec_mod = importlib.import_module("segdino3d.models.decoder.instance_seg_3d_decoder")

    def _route(mod):
        original = getattr(mod, "scatter_mean")
        if getattr(original, "__sd3d_routed__", False):
            return

        def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
            if src.is_cuda:
                return ops.scatter_mean(src, index, dim=dim, out=out, dim_size=dim_size)
            return original(src, index, dim=dim, out=out, dim_size=dim_size)

        scatter_mean.__sd3d_routed__ = True
        mod.scatter_mean = scatter_mean

    _route(sp_mod)
    _route(mk_mod)

    class SpConvUNetB200(sp_mod.SpConvUNet):
        """SpConvUNet whose superpoint pooling runs on the sorted segmented-mean kernels (no global atomics)."""

    class Res16UNet34CB200(mk_mod.Res16UNet34C):
        """Res16UNet34C whose superpoint pooling runs on the sorted segmented-mean kernels."""

    class ScanNetQueryDecoderB200(dec_mod.ScanNetQueryDecoder):
        """ScanNetQueryDecoder whose mask logits + attention masks come from one batched kernel launch."""

        mask_precision = "fp32"  # "bf16": tcgen05 tensor-core kernel (<= 1e-2)

        def _forward_head(self, queries, mask_feats, last_flag):
            norm = fused_out_norm(queries, self.out_norm)   # one launch for the whole batch (inference)
            cls_preds = [self.out_cls(nq) for nq in norm]
            sem_preds = [self.out_sem(nq) for nq in norm] if last_flag else None
            pred_scores = [self.out_score(nq) if self.objectness_flag else None for nq in norm]
            thr = self.mask_attention_threshold if self.attn_mask else None
            pred_masks, attn_masks = forward_head_masks(norm, list(mask_feats), thr, precision=self.mask_precision)
            return cls_preds, sem_preds, pred_scores, pred_masks, attn_masks

    builder.BACKBONES.register_module(module=SpConvUNetB200)
    builder.BACKBONES.register_module(module=Res16UNet34CB200)
    builder.DECODERS.register_module(module=ScanNetQueryDecoderB200)
    return {"SpConvUNetB200": SpConvUNetB200, "Res16UNet34CB200": Res16UNet34CB200,
            "ScanNetQueryDecoderB200": ScanNetQueryDecoderB200}
