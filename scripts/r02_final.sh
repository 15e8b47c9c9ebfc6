set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.log 2>&1; tail -4 gpurun_out/r02_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 600 gpurun_out/r02_bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","parity")})
print(d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"])
e=d["e2e"]; print(e["value"], e["pageable"], e["warm"]["ms_per_step"], e["cold_ms"]["calls"])
for k,v in d.get("configs",{}).items(): print(k, {x:v[x] for x in v if x in ("ms_per_step","main_kernel_ms_per_step","parity","frac_of_measured_hbm_kernel")})
try:
    r=json.loads(open("gpurun_out/r02_bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r.get("cpu_baseline"))
except Exception as ex: print("ref err", ex)
PY
