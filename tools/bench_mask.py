"""Mask-logit GEMM timings (a-5): tcgen05 bf16 kernel, fp32 FFMA kernel, torch.einsum (cuBLAS fp32, the
reference's path) at the ScanNet200 decoder shape, the eval shape (n = S) and a large shape. One JSON line each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_decoder_operands

dev = "cuda:0"
PEAK_TF = 1645.7
try:
    PEAK_TF = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass

def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3

for (n, s, d) in [(200, 500, 256), (500, 500, 256), (5000, 5000, 256)]:
    q, mf = make_decoder_operands(n, s, d)
    q, mf = q.to(dev), mf.to(dev)
    flop = 2.0 * n * s * d
    res = {"shape": [n, s, d], "flop": flop}
    _, q16 = sd.layernorm_cast(q, normalize=False, want_f32=False)
    _, mf16 = sd.layernorm_cast(mf, normalize=False, want_f32=False)
    q2, mf2 = sd.split_bf16(q), sd.split_bf16(mf)
    for name, fn in (("tma_tcgen05_bf16_operands", lambda: sd.mask_logits_bf16(q16, mf16)),
                     ("tma_tcgen05_bf16x3_split_operands", lambda: sd.mask_logits_bf16(q2, mf2, split=True)),
                     ("tma_tcgen05_bf16x3_split_operands+attn_mask", lambda: sd.mask_logits_bf16(q2, mf2, threshold=0.5, split=True)),
                     ("tma_tcgen05_bf16_operands+attn_mask", lambda: sd.mask_logits_bf16(q16, mf16, threshold=0.5)),
                     ("tcgen05_bf16", lambda: sd.mask_logits(q, mf, precision="bf16")),
                     ("tcgen05_bf16+attn_mask", lambda: sd.mask_logits(q, mf, precision="bf16", threshold=0.5)),
                     ("fp32_entry(ffma<64 tiles<=bf16x3)", lambda: sd.mask_logits(q, mf, precision="fp32")),
                     ("torch_einsum_fp32", lambda: torch.einsum("nd,md->nm", q, mf)),
                     ("torch_einsum+mask_epilogue", lambda: (lambda pm: ((pm.sigmoid() < 0.5), pm))(torch.einsum("nd,md->nm", q, mf)))):
        t = timeit(fn, 50 if n >= 5000 else 200)
        res[name] = {"us": t * 1e6, "tflops": flop / t / 1e12, "frac_of_bf16_peak": flop / t / 1e12 / PEAK_TF}
    print(json.dumps(res), flush=True)


# a batch of 8 scenes at the decoder shape, with the attention-mask epilogue: ONE batched launch (row reset fused) against
# the reference's per-scene sequence (einsum, sigmoid, compare, row sum, index_put: instance_seg_3d_decoder.py:557-573)
def ref_head(qs, mfs):
    out = []
    for q, mf in zip(qs, mfs):
        pm = torch.einsum("nd,md->nm", q, mf)
        am = (pm.sigmoid() < 0.5).bool()
        am[torch.where(am.sum(-1) == am.shape[-1])] = False
        out.append((pm, am))
    return out

for b in (1, 8):
    ops_ = [make_decoder_operands(200, 500, 256, seed=s_) for s_ in range(b)]
    qs, mfs = [o[0].to(dev) for o in ops_], [o[1].to(dev) for o in ops_]
    res = {"batch": b, "shape": [200, 500, 256]}
    for name, fn in (("batched_tcgen05_bf16+attn", lambda: sd.mask_logits_batched(qs, mfs, precision="bf16", threshold=0.5)),
                     ("batched_ffma_fp32+attn", lambda: sd.mask_logits_batched(qs, mfs, precision="fp32", threshold=0.5)),
                     ("per_scene_tcgen05_bf16+attn", lambda: [sd.mask_logits(q, m, precision="bf16", threshold=0.5) for q, m in zip(qs, mfs)]),
                     ("torch_reference_sequence", lambda: ref_head(qs, mfs))):
        t = timeit(fn, 200)
        res[name] = {"us": t * 1e6}
    print(json.dumps(res), flush=True)
