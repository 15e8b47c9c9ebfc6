set -x
NCU="ncu --set full --clock-control none --import-source on -f"
B="--no-configs --no-parity --no-e2e --no-cpu-baseline"
timeout 300 $NCU -k regex:k_agg_group_lean -s 1 -c 1 -o gpurun_out/r02_prof_agg_lean python bench.py --steps 2 --warmup 1 $B > gpurun_out/ncu1.log 2>&1
timeout 300 $NCU -k regex:k_rs_pass -s 5 -c 1 -o gpurun_out/r02_prof_rs_pass_v2 python bench_configs.py c5full --steps 1 --warmup 1 > gpurun_out/ncu2.log 2>&1
timeout 300 $NCU -k regex:k_rs_hist -c 1 -o gpurun_out/r02_prof_rs_hist python bench_configs.py c5full --steps 1 --warmup 1 > gpurun_out/ncu2b.log 2>&1
timeout 300 $NCU -k regex:k_agg_group_lean -s 1 -c 1 -o gpurun_out/r02_prof_c3_lean_join python bench_configs.py c3 --steps 1 --warmup 1 > gpurun_out/ncu3.log 2>&1
timeout 300 $NCU -k regex:k_join_build_fast -c 1 -o gpurun_out/r02_prof_join_build_fast python bench_configs.py c3 --steps 1 --warmup 1 > gpurun_out/ncu4.log 2>&1
timeout 300 $NCU -k regex:k_collect_rows_vec -s 1 -c 1 -o gpurun_out/r02_prof_collect_rows_vec python bench_configs.py c5 --steps 1 --warmup 1 > gpurun_out/ncu5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu6.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg.csv python bench_configs.py c1 c3 c5 --steps 2 --warmup 1 > gpurun_out/ncu7.log 2>&1
timeout 300 python bench_configs.py c1 c3 c5 c5full --steps 10 > gpurun_out/r02_cfg.json 2> gpurun_out/r02_cfg.err
ls -la gpurun_out/
