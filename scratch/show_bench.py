import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d['value']), "ms", round(d['ms_per_step'],3), "launches", d['gpu_launches'], "e2e", d.get('e2e'))
r=d['roofline']; print(r['kernel'], round(r['frac'],3), "events-step", round(r['ms_per_step_with_events'],3))
km=r.get('kernel_ms_per_step')
tot=0
for k,v in km.items():
    tot+=v; print(f"{k:24s} {v*1000:8.1f} us")
print("sum", round(tot,3))
