"""One large mask-logit problem through the TMA-fed tcgen05 kernel (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_decoder_operands
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
q, mf = make_decoder_operands(n, n, 256)
_, q16 = sd.layernorm_cast(q.cuda(), normalize=False, want_f32=False)
_, mf16 = sd.layernorm_cast(mf.cuda(), normalize=False, want_f32=False)
for _ in range(4):
    sd.mask_logits_bf16(q16, mf16)
torch.cuda.synchronize()
