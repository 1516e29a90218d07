#!/bin/bash
# usage: scripts/sweep.sh "<label>|<bench args>" ...   -> one summary line per configuration (logs in gpurun_out/)
mkdir -p gpurun_out
for spec in "$@"; do
  label="${spec%%|*}"; args="${spec#*|}"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args > "gpurun_out/sweep_${label}.log" 2>&1
  python - "$label" "gpurun_out/sweep_${label}.log" <<'EOF'
import json, sys
label, path = sys.argv[1], sys.argv[2]
lines = open(path).read().strip().splitlines()
try:
    d = json.loads(lines[-1])
    print(label, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "e2e_ms",
          round(d["e2e"]["ms_per_step"], 2), "patch_frac", round(d["roofline"]["frac"], 3), "gemm_ms", round(d.get("gemm_live",{}).get("launch_ms",0),3), "patch_ms", round(d["roofline"]["launch_ms"],3), "path_frac",
          round(d["path_roofline"]["frac_of_hbm_peak"], 3), "host_ms", round(d.get("host_enqueue_ms_per_step", 0), 2), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    print("   per-launch patch ms:", d["roofline"].get("per_launch_ms"))
except Exception as e:
    print(label, "FAILED", e)
    print("\n".join(lines[-12:]))
EOF
done
