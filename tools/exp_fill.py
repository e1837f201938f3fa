"""Store-only floor for the [n,S] fp32 logits: time of writing 100 MB (fill / memset / copy) on this GPU."""
import torch
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
x = torch.empty(5000, 5000, device="cuda"); y = torch.empty_like(x)
print("fill_ 100MB us", round(t(lambda: x.fill_(1.0)), 1))
print("zero_ 100MB us", round(t(lambda: x.zero_()), 1))
print("copy_ 100MB->100MB us", round(t(lambda: y.copy_(x)), 1))
m = torch.empty(5000, 5000, device="cuda", dtype=torch.uint8)
print("fill_ 25MB u8 us", round(t(lambda: m.fill_(1)), 1))
