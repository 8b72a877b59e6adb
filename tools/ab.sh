#!/bin/bash
# A/B runs of bench.py's device leg under different library builds (developer tool).
#   bash tools/ab.sh name1 name2 ...     (names of tools/build_variant.sh builds; "base" = libicpb.so)
#   AB_ARGS="--workload chain --beams 360" bash tools/ab.sh ...
for name in "$@"; do
  so=icp-slam-with-loop-closure_b200/libicpb_${name}.so
  [ "$name" = base ] && so=icp-slam-with-loop-closure_b200/libicpb.so
  ICPB_SO=$PWD/$so python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 $AB_ARGS > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<PY
import json, sys
f = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/ab_{f}.json")); r = j["roofline"]; k = j["details"]["kernel"]
    print(f, round(j["value"]), round(j["ms_per_step"], 3), "exh", round(r.get("exhaustive", {}).get("frac", 0), 3),
          "share", round(r["algorithmic"]["executed_share"], 4), k["threads_per_cta"], k["ctas_per_sm"], k["regs_per_thread"])
except Exception as e:
    print(f, "ERR", e)
PY
done
