"""One large mask-logit problem through the TMA-fed tcgen05 kernel (for ncu captures).
usage: exp_mask_tma.py [n] [attn|split]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_decoder_operands
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
mode = sys.argv[2] if len(sys.argv) > 2 else ""
q, mf = make_decoder_operands(n, n, 256)
if mode == "split":
    q16, mf16 = sd.split_bf16(q.cuda()), sd.split_bf16(mf.cuda())
else:
    _, q16 = sd.layernorm_cast(q.cuda(), normalize=False, want_f32=False)
    _, mf16 = sd.layernorm_cast(mf.cuda(), normalize=False, want_f32=False)
for _ in range(4):
    sd.mask_logits_bf16(q16, mf16, threshold=0.5 if mode == "attn" else None, split=mode == "split")
torch.cuda.synchronize()
