"""Condense the A/B bench lines of a gpurun directory (gpurun_out/<run>/ab_*.json, one bench.py JSON line each) into one table.
  python profiles/collect_ab.py gpurun_out/r2d profiles/r02_ab_<what>.json "<note>" """
import glob, json, os, sys
src, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = []
for f in sorted(glob.glob(os.path.join(src, "ab_*.json"))):
    try:
        d = json.load(open(f))
    except Exception:
        continue
    c, r = d["config"], d["roofline"]
    rows.append({"run": os.path.basename(f)[3:-5], "workload": c["workload"], "tuning": c.get("tuning", []), "kernel": r["kernel"].split(" ")[0],
                 "streams_per_gpu": c["streams_per_gpu"], "block": c["block"], "partitions": c["partitions"], "steps": d["steps"],
                 "ms_per_step": round(d["ms_per_step"], 4), "kernel_gbs": round(r["achieved"], 1), "frac_of_measured_peak": round(r["frac"], 4),
                 "throughput_equivalent_channels": round(c["throughput_equivalent_channels"]), "p99_ms": round(d["latency"]["p99_ms"], 4),
                 "sm_mhz": (d.get("clocks") or {}).get("sm_mhz"), "power_w": (d.get("clocks") or {}).get("power_w_median"), "reasons": (d.get("clocks") or {}).get("reasons")})
json.dump({"note": note, "command": "python bench.py --no-cpu-baseline --no-e2e --no-selfcheck --steps N --warmup 5 --streams S [--block B] [--tune knob=value ...]  (one B200, same box, back to back)",
           "rows": rows}, open(dst, "w"), indent=1)
print(len(rows), "rows ->", dst)
