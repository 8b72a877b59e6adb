#!/bin/bash
# One point of a scaling curve (developer tool): bash tools/scaling_run.sh <N> <tag> <workload> [bench flags]
#   -> gpurun_out/<tag>_<workload>_<N>gpu.json   (run it under `gpurun --gpus N`)
N=$1; tag=$2; wl=$3; shift 3
out=gpurun_out/${tag}_${wl}_${N}gpu.json
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --workload $wl "$@" > $out 2> ${out%.json}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      bench.py --gpus $N --workload $wl "$@" > $out 2> ${out%.json}.err
fi
python - $out <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    e = j.get("e2e") or {}
    print(sys.argv[1], "value %.0f %s, %.3f ms/step, e2e %.0f (%.3f ms), pinned-table e2e %.0f, check: %s" % (
        j["value"], j["unit"], j["ms_per_step"], e.get("value", 0), e.get("ms_per_step", 0),
        (e.get("pinned_table") or {}).get("value", 0), j["details"].get("gather_check")))
except Exception as ex:
    print(sys.argv[1], "ERR", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
