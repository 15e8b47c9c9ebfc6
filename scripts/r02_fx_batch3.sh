# GPU batch 3 (last of round 2): A/B of three builds of the FX kernel (exact path inlined = the default library; + the interleaved pair path;
# + the unconditional high-limb add), then with the fastest: parity suites, the default bench line, one ncu --set full capture, the launch list.
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
: > gpurun_out/fx_ab3.jsonl
for lib in baikaldb_b200/libbkgpu.so baikaldb_b200/ab/libbkgpu_B.so baikaldb_b200/ab/libbkgpu_C.so; do
  [ -f $lib ] && BKGPU_LIB=$PWD/$lib timeout 80 python scripts/r02_fx_ab.py --quick --steps 20 >> gpurun_out/fx_ab3.jsonl 2>> gpurun_out/fx_ab3.err
done
WIN=$(python - <<'PY'
import json, sys
best=None
for l in open("gpurun_out/fx_ab3.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    print(d["lib"].split("/")[-1], d["case"], "cas", round(d["cas"]["kernel_ms"],4), "fx", round(d["fx"]["kernel_ms"],4), d["checks"], file=sys.stderr)
    if d["case"].startswith("C2") and d["checks"]["vs_torch"].get("count_exact") and d["checks"]["vs_torch"].get("sum_rel",1)<1e-9:
        if best is None or d["fx"]["kernel_ms"]<best[0]*0.99: best=(d["fx"]["kernel_ms"], d["lib"])
print(best[1] if best else "")
PY
)
echo "WINNER $WIN" | tee gpurun_out/fx_winner3.txt
[ -n "$WIN" ] && export BKGPU_LIB=$WIN
timeout 200 python -m pytest tests/test_gpu_fx.py tests/test_gpu_agg.py tests/test_gpu_join.py tests/test_gpu_merge.py tests/test_gpu_compose.py tests/test_gpu_fullsize.py tests/test_abi_and_plan.py -m gpu -q > gpurun_out/fx_suite3.log 2>&1; tail -4 gpurun_out/fx_suite3.log
timeout 200 python bench.py > gpurun_out/fx_bench3.json 2> gpurun_out/fx_bench3.err; tail -c 300 gpurun_out/fx_bench3.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/fx_bench3.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","parity")}); print(d["roofline"]); print(d["e2e"]["value"], d["e2e"].get("warm"))
    for k,v in d.get("configs",{}).items(): print(k, {x:v[x] for x in v if x in ("ms_per_step","main_kernel","main_kernel_ms","parity")})
except Exception as ex: print("bench err", ex)
PY
timeout 120 ncu --set full --clock-control none --import-source on -f -k regex:k_agg_group_lean -s 1 -c 1 -o gpurun_out/r02_prof_agg_lean_fx python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_fx.log 2>&1; tail -2 gpurun_out/ncu_fx.log | cut -c1-200
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fx.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_fx_l.log 2>&1; tail -c 200 gpurun_out/ncu_fx_l.log
