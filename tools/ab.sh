#!/bin/bash
# A/B runs of bench.py's device leg under different tuning environments (developer tool).
#   bash tools/ab.sh name1:ENV=VAL,ENV2=VAL name2: ...
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $(echo $envs | tr ',' ' ') python bench.py --no-e2e --steps 10 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<PY
import json, sys
f = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/ab_{f}.json")); r = j["roofline"]; k = j["config"]["kernel"]
    print(f, round(j["value"]), round(j["ms_per_step"], 3), "exh", round(r.get("exhaustive", {}).get("frac", 0), 3),
          "share", round(r.get("executed_share"), 4), k["threads_per_cta"], k["ctas_per_sm"], k["regs_per_thread"])
except Exception as e:
    print(f, "ERR", e)
PY
done
