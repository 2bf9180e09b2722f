import json, sys
for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        print(ln); continue
    d = json.loads(ln)
    ph = {k: round(v["ms"], 3) for k, v in d.get("roofline", {}).get("phases", {}).items()}
    print(f"{d['config']['workload'][:12]} value={d['value']/1e9:.3f} G/s ms/step={d['ms_per_step']:.3f} frac={d.get('roofline',{}).get('frac',0):.3f} phases={ph} "
          f"launches={d.get('gpu_launches')} fallback={d.get('knn_fallback_particles')} e2e={d['e2e']['value']/1e9:.3f} G/s clocks={d.get('clocks')}")
