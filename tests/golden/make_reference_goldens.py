"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE TEXT for the pooling / mask-head call patterns.

The reference package cannot be imported here (mmengine / spconv / MinkowskiEngine / torch_scatter are not
installed), but the three methods on the hot path are plain torch code around a handful of injected
collaborators. This script `ast`-extracts their text from /root/reference, compiles it unchanged, and runs it with

  * SpConvUNet.forward_wrapper            segdino3d/models/backbone/spconvunet.py   (scatter_mean call sites :390,:392)
  * Res16UNetBase.forward_wrapper         segdino3d/models/backbone/minkunet.py     (:639,:641,:653,:674)
  * ScanNetQueryDecoder._forward_head     segdino3d/models/decoder/instance_seg_3d_decoder.py (:558-:573)

against stub collaborators: the conv nets are fixed seeded linear maps, `spconv` / `ME` are minimal shape-only stand-ins,
and `torch_scatter.scatter_mean` is torch-scatter 2.1.2's published composite (zeros.scatter_add_ / count clamp /
true_divide_) written here in plain torch, independent of oracle/scatter_oracle.py. Inputs and outputs are stored in
tests/golden/ref_call_patterns.npz; the sha256 of the extracted source text is stored with them so that
tests/test_reference_goldens.py can tell (where /root/reference exists) that the fixture still matches the reference.

    python tests/golden/make_reference_goldens.py        # needs /root/reference; writes the .npz

The fixture travels to the GPU box; /root/reference does not, and nothing at test time on the GPU reads it.
"""
from __future__ import annotations

import ast
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_call_patterns.npz")

SOURCES = {
    "spconv_forward_wrapper": ("segdino3d/models/backbone/spconvunet.py", "SpConvUNet", "forward_wrapper"),
    "mink_forward_wrapper": ("segdino3d/models/backbone/minkunet.py", "Res16UNetBase", "forward_wrapper"),
    "forward_head": ("segdino3d/models/decoder/instance_seg_3d_decoder.py", "ScanNetQueryDecoder", "_forward_head"),
}


def extract(path: str, cls: str, fn: str) -> str:
    """Source text of method `fn` of class `cls` (dedented), verbatim from the reference file."""
    text = open(os.path.join(REF, path)).read()
    tree = ast.parse(text)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == fn:
                    lines = text.splitlines()[item.lineno - 1: item.end_lineno]
                    indent = len(lines[0]) - len(lines[0].lstrip())
                    return "\n".join(ln[indent:] if len(ln) >= indent else ln.lstrip() for ln in lines) + "\n"
    raise LookupError(f"{cls}.{fn} not found in {path}")


POOL_CALLS = []  # (src, index) of every scatter_mean call the reference text makes, in call order


def scatter_mean_published(src, index, dim=-1, out=None, dim_size=None):
    """torch-scatter 2.1.2 `scatter_mean` (torch_scatter/scatter.py), floating-point branch, restated with torch ops."""
    assert out is None
    POOL_CALLS.append((src.detach().clone(), index.detach().clone()))
    d = dim + src.dim() if dim < 0 else dim
    assert d == 0 and index.dim() == 1
    size = int(index.max()) + 1 if dim_size is None else dim_size
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    summed = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype).scatter_add_(0, idx, src)
    count = torch.zeros(size, dtype=src.dtype).scatter_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    count[count < 1] = 1
    return summed.true_divide_(count.view(-1, *([1] * (src.dim() - 1))))


def compile_method(src: str, name: str, globs: dict):
    ns = dict(globs)
    exec(compile(src, f"<reference:{name}>", "exec"), ns)
    return ns[name]


# ---------------------------------------------------------------------------------------------------------------
# stub collaborators
# ---------------------------------------------------------------------------------------------------------------
class _SparseConvTensor:  # spconv.SparseConvTensor(features, coordinates, spatial_shape, batch_size)
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size


class _TensorField:  # ME.TensorField(coordinates=, features=)
    def __init__(self, features=None, coordinates=None):
        self.features, self.coordinates = features, coordinates

    def sparse(self):
        return self


def _batch_sparse_collate(pairs, device=None):
    coords = torch.cat([torch.cat([torch.full((c.shape[0], 1), i, dtype=torch.int32), torch.floor(c).int()], 1)
                        for i, (c, _) in enumerate(pairs)])
    return coords, torch.cat([f for _, f in pairs])


def make_inputs(seed: int = 20251017):
    g = torch.Generator().manual_seed(seed)
    sizes, sps = [900, 400, 700], [37, 12, 29]
    samples, targets = [], []
    for n, s in zip(sizes, sps):
        sp = torch.randint(0, s, (n,), generator=g)
        sp[0] = s - 1  # every scene uses its full id range (ids 3 and 5 of scene 1 stay empty)
        if s == 12:
            sp[sp == 3] = 4
            sp[sp == 5] = 6
        pts = torch.cat([torch.rand(n, 3, generator=g) * 4.0, torch.rand(n, 3, generator=g)], 1)
        samples.append(pts)
        targets.append({"extra_features": {"super_point_masks": sp, "points_2dfeats": torch.randn(n, 64, generator=g)}})
    w_backbone = torch.randn(6 + 64, 32, generator=g) * 0.1   # stand-in for the conv nets: fixed linear map
    inv = [torch.randint(0, n // 2, (n,), generator=g) for n in sizes]  # point -> voxel row (spconv inverse_mapping)
    return samples, targets, w_backbone, inv


def run_spconv(samples, targets, w_backbone, inv):
    src = extract(*SOURCES["spconv_forward_wrapper"])
    fw = compile_method(src, "forward_wrapper", {"torch": torch, "scatter_mean": scatter_mean_published,
                                                 "spconv": types.SimpleNamespace(SparseConvTensor=_SparseConvTensor)})
    # collate stand-in (spconvunet.py:273-362 needs MinkowskiEngine): voxel features = a seeded gather of the point
    # features, inverse_mapping given, no positions
    voxel_rows, offs = [], 0
    inverse = []
    for pts, tgt, iv in zip(samples, targets, inv):
        nvox = int(iv.max()) + 1
        feats = torch.cat([pts, tgt["extra_features"]["points_2dfeats"]], 1)
        vox = torch.zeros(nvox, feats.shape[1]).index_add_(0, iv, feats)
        voxel_rows.append(vox)
        inverse.append(iv + offs)
        offs += nvox
    voxel_feats, inverse = torch.cat(voxel_rows), torch.cat(inverse)

    def collate(points, elastic, rgbfeat, sp_pts_masks=None, batch_offsets=None, return_sp_mean_pos=True):
        dinox = torch.cat(rgbfeat, 0) if rgbfeat is not None else None
        return torch.zeros(voxel_feats.shape[0], 4, dtype=torch.int32), voxel_feats, dinox, inverse, None, None, None

    def forward(x):
        return _SparseConvTensor(x.features @ w_backbone, x.indices, x.spatial_shape, x.batch_size), None

    me = types.SimpleNamespace(collate=collate, input_conv=lambda x: x, forward=forward, output_layer=lambda x: x)
    del POOL_CALLS[:]
    out, x_pos, pos_wo = fw(me, samples, targets, return_sp_mean_pos=True)
    assert len(POOL_CALLS) == 2  # backbone features, DINO-X features (spconvunet.py:390,392)
    return {"spconv_pool_in": POOL_CALLS[0][0], "spconv_pool_ids": POOL_CALLS[0][1], "spconv_out": torch.cat(out),
            "spconv_out_sizes": torch.tensor([o.shape[0] for o in out])}, src


def run_mink(samples, targets, w_backbone):
    src = extract(*SOURCES["mink_forward_wrapper"])
    me_mod = types.SimpleNamespace(utils=types.SimpleNamespace(batch_sparse_collate=_batch_sparse_collate),
                                   TensorField=_TensorField)
    fw = compile_method(src, "forward_wrapper", {"torch": torch, "scatter_mean": scatter_mean_published, "ME": me_mod})

    class _Out:
        def __init__(self, f):
            self.f = f

        def slice(self, field):
            return types.SimpleNamespace(features=self.f)

    me = types.SimpleNamespace(voxel_size=0.02, mode_fuse_2d_feat="early_fusion", add_positional_embedding=True,
                               forward=lambda field: _Out(field.features @ w_backbone[3:]))
    del POOL_CALLS[:]
    feats, pos, pos_wo = fw(me, samples, targets, return_sp_mean_pos=True)
    assert len(POOL_CALLS) == 4  # features, DINO-X features, positions, positions without elastic (minkunet.py:639-674)
    return {"mink_pool_in": POOL_CALLS[0][0], "mink_pool_ids": POOL_CALLS[0][1], "mink_pos_in": POOL_CALLS[2][0],
            "mink_out": torch.cat(feats), "mink_pos": torch.cat(pos), "mink_pos_wo": torch.cat(pos_wo),
            "mink_out_sizes": torch.tensor([f.shape[0] for f in feats])}, src


def run_head(seed: int = 7):
    src = extract(*SOURCES["forward_head"])
    fh = compile_method(src, "_forward_head", {"torch": torch})
    g = torch.Generator().manual_seed(seed)
    d, sizes, nq = 256, [311, 97], [40, 25]
    queries = [torch.randn(n, d, generator=g) for n in nq]
    mask_feats = [torch.randn(s, d, generator=g) * 0.08 for s in sizes]
    ln = torch.nn.LayerNorm(d)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.1 * torch.randn(d, generator=g))
        ln.bias.copy_(0.05 * torch.randn(d, generator=g))
        # scene 1: every superpoint feature has a component against normalised query 3 -> that query's logits are all
        # negative, its attention-mask row is all-true before the reset of :570-571
        nq3 = ln(queries[1][3:4])[0]
        mask_feats[1] = mask_feats[1] - (0.02 + 0.05 * torch.rand(sizes[1], 1, generator=g)) * nq3[None, :]
    w_cls, w_sem, w_score = (torch.randn(d, k, generator=g) * 0.05 for k in (19, 21, 1))
    me = types.SimpleNamespace(out_norm=ln, out_cls=lambda x: x @ w_cls, out_sem=lambda x: x @ w_sem,
                               out_score=lambda x: x @ w_score, objectness_flag=True, attn_mask=True,
                               mask_attention_threshold=0.5)
    with torch.no_grad():
        cls_preds, sem_preds, scores, pred_masks, attn_masks = fh(me, queries, mask_feats, True)
        norm_q = [ln(q) for q in queries]
    out = {}
    for i in range(2):
        out[f"head_q{i}"], out[f"head_mf{i}"], out[f"head_normq{i}"] = queries[i], mask_feats[i], norm_q[i]
        out[f"head_pred{i}"], out[f"head_attn{i}"] = pred_masks[i], attn_masks[i].to(torch.uint8)
    out["head_ln_weight"], out["head_ln_bias"] = ln.weight.detach(), ln.bias.detach()
    return out, src


def generate():
    samples, targets, w_backbone, inv = make_inputs()
    out = {}
    for i, (pts, tgt) in enumerate(zip(samples, targets)):
        out[f"pts{i}"] = pts
        out[f"sp{i}"] = tgt["extra_features"]["super_point_masks"]
        out[f"feat2d{i}"] = tgt["extra_features"]["points_2dfeats"]
    out["w_backbone"] = w_backbone
    hashes = {}
    for name, (res, src) in (("spconv_forward_wrapper", run_spconv(samples, targets, w_backbone, inv)),
                             ("mink_forward_wrapper", run_mink(samples, targets, w_backbone)),
                             ("forward_head", run_head())):
        out.update(res)
        hashes[name] = hashlib.sha256(src.encode()).hexdigest()
    return out, hashes


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} is not available: the goldens can only be regenerated where the reference is")
    out, hashes = generate()
    arrays = {k: v.detach().numpy() for k, v in out.items()}
    for name, h in hashes.items():
        arrays["sha256_" + name] = np.frombuffer(bytes.fromhex(h), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print(OUT, os.path.getsize(OUT), "bytes;", {k: v[:12] for k, v in hashes.items()})


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
