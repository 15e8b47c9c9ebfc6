for C in 2 10 50; do
  BKGPU_BENCH_CLOCK_MS=$C timeout 200 python bench.py --no-configs --no-e2e --no-parity --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/clk_$C.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/clk_$C.json').read().strip().splitlines()[-1]); print('clock sample every $C ms: step %.4f ms kernel %.4f clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['clocks']))"
done
