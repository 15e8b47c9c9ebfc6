# GPU batch 2 for the FX variant: parity tests of the fixed kernel, A/B of five library builds (pair fast path on/off, 640/576/512 threads),
# then with the fastest build: the whole GPU suite with FX forced on, the default bench line, one ncu --set full capture and the launch list.
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_fx.py -q > gpurun_out/fx_tests2.log 2>&1; tail -4 gpurun_out/fx_tests2.log
: > gpurun_out/fx_ab2.jsonl
for lib in baikaldb_b200/libbkgpu.so build/libbkgpu_p0.so build/libbkgpu_p0t576.so build/libbkgpu_p1t576.so build/libbkgpu_p1t512.so; do
  [ -f $lib ] && BKGPU_LIB=$PWD/$lib timeout 90 python scripts/r02_fx_ab.py --quick --steps 20 >> gpurun_out/fx_ab2.jsonl 2>> gpurun_out/fx_ab2.err
done
WIN=$(python - <<'PY'
import json
best=None
for l in open("gpurun_out/fx_ab2.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    print(d["lib"].split("/")[-1], d["case"], "cas", round(d["cas"]["kernel_ms"],4), "fx", round(d["fx"]["kernel_ms"],4), d["checks"], file=__import__("sys").stderr)
    if d["case"].startswith("C2") and d["checks"]["vs_torch"].get("count_exact") and d["checks"]["vs_torch"].get("sum_rel",1)<1e-9:
        if best is None or d["fx"]["kernel_ms"]<best[0]: best=(d["fx"]["kernel_ms"], d["lib"])
print(best[1] if best else "")
PY
)
echo "WINNER $WIN" | tee gpurun_out/fx_winner.txt
[ -n "$WIN" ] && [ "$WIN" != "default" ] && export BKGPU_LIB=$WIN
export BKGPU_LEAN_FX=1
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/fx_suite2.log 2>&1; tail -5 gpurun_out/fx_suite2.log
timeout 240 python bench.py > gpurun_out/fx_bench2.json 2> gpurun_out/fx_bench2.err; tail -c 300 gpurun_out/fx_bench2.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/fx_bench2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","parity")}); print(d["roofline"]); print(d["e2e"]["value"], d["e2e"].get("warm"))
    for k,v in d.get("configs",{}).items(): print(k, {x:v[x] for x in v if x in ("ms_per_step","main_kernel","main_kernel_ms","parity")})
except Exception as ex: print("bench err", ex)
PY
timeout 150 ncu --set full --clock-control none --import-source on -f -k regex:k_agg_group_lean -s 1 -c 1 -o gpurun_out/r02_prof_agg_lean_fx python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_fx.log 2>&1; tail -2 gpurun_out/ncu_fx.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fx.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fx_l.log 2>&1; tail -2 gpurun_out/ncu_fx_l.log
