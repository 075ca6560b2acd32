import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d.get("e2e",{}).get("value"))
    print({k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],5))
    if "cpu_baseline" in d: print(d["cpu_baseline"])
