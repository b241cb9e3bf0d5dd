import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} pairs/s  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value'] if d.get('e2e') else None}  launches {d['gpu_launches']}  flag {d['debug_flag']}  clocks {d['clocks']}")
tot = 0
for k, v in sorted(d["kernel_classes"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    print(f"  {k:12s} n={v['launches']:3d}  {v['ms_per_step']:7.3f} ms   {v['work_per_step'] / v['ms_per_step'] / 1e9:9.1f} G(FLOP|B)/s")
print(f"  classes total {tot:.3f} ms")
