B=$PWD/icp-slam-with-loop-closure_b200/bin
run() { # name so extra-env
  env ICPB_SO=$2 $3 python bench.py --no-e2e --steps 10 --warmup 3 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
}
run base $B/libicpb_base.so
run expR4 $B/libicpb_expR4.so
run expR4_t128 $B/libicpb_expR4.so ICPB_THREADS=128
run diffR4 $B/libicpb_diffR4.so
run diffR4_t128 $B/libicpb_diffR4.so ICPB_THREADS=128
python - <<PY
import json
for f in ["base","expR4","expR4_t128","diffR4","diffR4_t128"]:
    try:
        j=json.load(open(f"gpurun_out/ab_{f}.json")); r=j["roofline"]; k=j["config"]["kernel"]
        print(f, round(j["value"]), round(j["ms_per_step"],3), "exh", round(r.get("exhaustive",{}).get("frac",0),3), "share", round(r.get("executed_share"),4), k["threads_per_cta"], k["ctas_per_sm"], k["regs_per_thread"])
    except Exception as e: print(f, "ERR", e)
PY
