import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d['value']), "ms", round(d['ms_per_step'],3), "launches", d['gpu_launches'], "e2e", d.get('e2e'))
r=d['roofline']; print(r['kernel'], round(r['frac'],3), "events-step", round(r['ms_per_step_with_events'],3))
t=r['ms_per_step_with_events']
for k,v in r['kernel_time_share'].items(): print(f"{k:24s} {v:.4f} {v*t*1000:8.1f} us")
