"""Print the headline fields of a bench.py JSON line:  python scripts/show_bench.py file.json [n_kernels]"""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 14
print(f"value {j['value']} {j['unit']}  ms/step {j['ms_per_step']}  e2e {j['e2e']['value']} ({j['e2e'].get('ms_per_step')} ms)  "
      f"launches {j.get('gpu_launches')}  clocks {j.get('clocks')}")
r = j.get("roofline") or {}
print(f"roofline: frac {r.get('frac')} achieved {r.get('achieved')} avg_us {r.get('avg_launch_us')} bwd {r.get('backward')}")
for k in ("grad_exchange", "strong", "cpu_baseline"):
    if j.get(k):
        print(k, j[k])
for k in (j.get("kernels") or [])[:nk]:
    extra = f"  {k['avg_launch_us']:7.1f} us/launch  {k['hbm_frac']:.2f} of HBM peak" if "hbm_frac" in k else ""
    print(f"  {k['ms_per_step']:7.3f} ms  x{k['launches_per_step']:<5g} {k['entry']}{extra}")
