import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        r = d.get("roofline", {})
        print(f"{f}: {d['value']:.1f} {d['unit']}  step={d['ms_per_step']*1e3:.1f}us  kernel={r.get('kernel_ms',0)*1e3:.1f}us  frac={r.get('frac',0):.3f} path_frac={r.get('path_frac',0):.3f} cfg={d['config'].get('run')}/{d['config'].get('variant')} e2e={d.get('e2e',{}).get('value')} host={d.get('host_us_per_step',0):.0f}us")
    except Exception as e:
        print(f, "ERR", e)
