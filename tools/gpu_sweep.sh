#!/bin/bash
# Development aid: run a list of bench.py configurations on the GPU box and collect one JSON line per configuration.
#   tools/gpu_sweep.sh <tag> "<name>|<ENV=VAL ...>|<bench args>" ...
# writes gpurun_out/<tag>_sweep.jsonl (one line: {"name":..., "ms_per_step":..., ...}) and per-run stderr in gpurun_out/<tag>_<name>.err
tag=$1; shift
out=gpurun_out/${tag}_sweep.jsonl
: > "$out"
for spec in "$@"; do
  IFS='|' read -r name envs args <<< "$spec"
  line=$(env $envs timeout 300 python bench.py --no-cpu-baseline --no-side $args 2> "gpurun_out/${tag}_${name}.err" | tail -1)
  python - "$name" "$line" >> "$out" <<'PY'
import sys, json
name, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(json.dumps({"name": name, "ms_per_step": round(d["ms_per_step"], 4), "value": d["value"], "mode": d["config"].get("mode"), "e2e_ms": round(d["e2e"]["ms_per_step"], 4),
                      "step_frac": round(d["roofline"]["step_frac"], 4), "sm_mhz": (d.get("clocks") or {}).get("sm_mhz"), "probe": d.get("mode_probe_ms"), "checksum": d.get("checksum")}))
except Exception as ex:
    print(json.dumps({"name": name, "error": str(ex), "raw": line[-300:]}))
PY
done
cat "$out"
