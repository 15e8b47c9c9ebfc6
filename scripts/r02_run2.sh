set -x
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:k_rs_pass -s 5 -c 1 -o gpurun_out/r02_prof_rs_pass_v3 python bench_configs.py c5full --steps 1 --warmup 1 > gpurun_out/ncu2.log 2>&1
timeout 400 python bench.py --no-configs --no-cpu-baseline > gpurun_out/r02_bench_nt.json 2> gpurun_out/r02_bench_nt.err
BKGPU_BENCH_OPTS="no_stream_copy=1" timeout 400 python bench.py --no-configs --no-cpu-baseline --no-parity > gpurun_out/r02_bench_memcpy.json 2> gpurun_out/r02_bench_memcpy.err
python - <<'PY'
import json
for f in ("nt","memcpy"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], json.dumps(d["e2e"])[:700])
    except Exception as e: print(f, "ERR", e)
PY
