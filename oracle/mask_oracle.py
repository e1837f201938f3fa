"""ORACLE (test infrastructure): the decoder's query x superpoint mask-logit einsum and its
attention-mask epilogue, step a-5 (+K6) of SURVEY.md section 8(a).

Follows /root/reference/segdino3d/models/decoder/instance_seg_3d_decoder.py:567-573
(ScanNetQueryDecoder._forward_head) and :339-345 (base decoder):

    pred_mask = torch.einsum('nd,md->nm', norm_query, mask_feats[i])
    attn_mask = (pred_mask.sigmoid() < self.mask_attention_threshold).bool()
    attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False

The arithmetic is torch 2.4 einsum -> SGEMM (third party, installation.md:8); the reference keeps no
golden vectors for it -> pinned by library semantics; a float64 twin is provided for error budgeting.
"""
from __future__ import annotations

import torch


def mask_logits_oracle(q: torch.Tensor, mf: torch.Tensor) -> torch.Tensor:
    return torch.einsum("nd,md->nm", q.float(), mf.float())


def mask_logits_f64(q: torch.Tensor, mf: torch.Tensor) -> torch.Tensor:
    return torch.einsum("nd,md->nm", q.double(), mf.double())


def attn_mask_oracle(pred_mask: torch.Tensor, threshold: float) -> torch.Tensor:
    attn_mask = (pred_mask.sigmoid() < threshold).bool()
    attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
    return attn_mask
