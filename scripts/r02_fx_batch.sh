# GPU batch for the FX variant of the lean kernel: parity tests, A/B against the CAS variant, the aggregate / join suites with FX forced on,
# then the default bench line with FX on and one ncu capture.  Every step has its own timeout and writes into gpurun_out/.
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_fx.py -q > gpurun_out/fx_tests.log 2>&1; tail -6 gpurun_out/fx_tests.log
timeout 200 python scripts/r02_fx_ab.py --steps 10 > gpurun_out/fx_ab.jsonl 2> gpurun_out/fx_ab.err; cut -c1-700 gpurun_out/fx_ab.jsonl; tail -3 gpurun_out/fx_ab.err
BKGPU_LEAN_FX=1 timeout 240 python -m pytest tests/test_gpu_agg.py tests/test_gpu_join.py tests/test_gpu_merge.py tests/test_gpu_compose.py tests/test_gpu_fullsize.py -q > gpurun_out/fx_suite.log 2>&1; tail -6 gpurun_out/fx_suite.log
BKGPU_LEAN_FX=1 timeout 240 python bench.py > gpurun_out/fx_bench.json 2> gpurun_out/fx_bench.err; tail -c 300 gpurun_out/fx_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/fx_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","parity")}); print(d["roofline"])
    for k,v in d.get("configs",{}).items(): print(k, {x:v[x] for x in v if x in ("ms_per_step","main_kernel","main_kernel_ms","parity")})
except Exception as ex: print("bench err", ex)
PY
BKGPU_LEAN_FX=1 timeout 150 ncu --set full --clock-control none --import-source on -f -k regex:k_agg_group_lean -s 1 -c 1 -o gpurun_out/r02_prof_agg_lean_fx python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-parity > gpurun_out/ncu_fx.log 2>&1; tail -2 gpurun_out/ncu_fx.log
