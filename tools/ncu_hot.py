"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` (one kernel), with the dominant reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
body = []
for r in rows[2:]:  # first kernel section only
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print(f"kernel: {rows[0][1][:100]}  instructions: {len(body)}  samples: {tot}")
agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stall_cols}
print("stall totals:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[ci['# Samples']]):6d} ({100 * int(r[ci['# Samples']]) / max(tot, 1):4.1f}%) exec={r[ci['Instructions Executed']]:>9s}  "
          f"{r[ci['Source']].strip()[:70]:70s} {' '.join(f'{n}:{c}' for c, n in reasons if c)}")
