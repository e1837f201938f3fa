"""segdino3d_b200 -- B200 (sm_100a) implementation of SegDINO3D's 2D->3D feature-lifting, superpoint
pooling and mask-logit path, behind the reference's own operator interface. See DESIGN.md.

Importing the package does not load the CUDA library; the first op call does, and raises if
libsd3d.so is missing (there is no CPU fallback)."""
from ._lib import Sd3dError  # noqa: F401
from .io import load_points_2dfeats, save_points_2dfeats  # noqa: F401
from .ops import (LiftPoolBuffers, SuperpointPlan, expand_superpoint_masks, lift, lift_and_pool, lift_features,  # noqa: F401
                  lift_finalize, layernorm_cast, mask_logits, mask_logits_batched, mask_logits_bf16, scale_mean, scatter_mean, sp_mean, sp_sort, split_bf16, superpoint_label_masks)

__all__ = ["LiftPoolBuffers", "Sd3dError", "SuperpointPlan", "expand_superpoint_masks", "lift", "lift_and_pool", "lift_features",
           "lift_finalize", "load_points_2dfeats", "layernorm_cast", "mask_logits", "mask_logits_batched", "mask_logits_bf16", "save_points_2dfeats", "scale_mean", "scatter_mean", "split_bf16",
           "sp_mean", "sp_sort", "superpoint_label_masks"]
