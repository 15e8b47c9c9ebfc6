set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_tests2.log 2>&1; tail -12 gpurun_out/r02_tests2.log
timeout 300 python bench_configs.py c1 c3 c5 c5full --steps 10 > gpurun_out/r02_cfg2.json 2> gpurun_out/r02_cfg2.err
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:k_rs_pass -s 5 -c 1 -o gpurun_out/r02_prof_rs_pass_v3 python bench_configs.py c5full --steps 1 --warmup 1 > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:k_agg_group_lean -s 1 -c 1 -o gpurun_out/r02_prof_c3_lean_join_v2 python bench_configs.py c3 --steps 1 --warmup 1 > gpurun_out/ncu3.log 2>&1
timeout 400 python bench.py --no-configs --no-cpu-baseline > gpurun_out/r02_bench_nt.json 2> gpurun_out/r02_bench_nt.err
BKGPU_BENCH_OPTS="lean_bank=1" timeout 400 python bench.py --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_bank.json 2> gpurun_out/r02_bench_bank.err
python -c "import json; d=json.loads(open('gpurun_out/r02_bench_bank.json').read().strip().splitlines()[-1]); print('bank', d['ms_per_step'], d['roofline'], d.get('parity'))"
BKGPU_BENCH_OPTS="no_stream_copy=1" timeout 400 python bench.py --no-configs --no-cpu-baseline --no-parity > gpurun_out/r02_bench_memcpy.json 2> gpurun_out/r02_bench_memcpy.err
python - <<'PY'
import json
for f in ("nt","memcpy"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], json.dumps(d["e2e"])[:700])
    except Exception as e: print(f, "ERR", e)
for l in open("gpurun_out/r02_cfg2.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["main_kernel"], round(d["ms_per_step"],4), round(d["main_kernel_ms_per_step"],4))
PY
