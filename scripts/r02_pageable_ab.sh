for T in 8 12 16 24; do for P in 0 300; do
  BKGPU_COPY_THREADS=$T BKGPU_COPY_POLL_US=$P timeout 200 python bench.py --no-configs --no-cpu-baseline --no-parity --steps 5 --warmup 3 > gpurun_out/ab_$T_$P.json 2>/dev/null
  python -c "
import json,sys; d=json.loads(open('gpurun_out/ab_$T_$P.json').read().strip().splitlines()[-1]); e=d['e2e']; print('threads $T poll $P us: pinned %.1f GB/s pageable %.1f GB/s of_pinned %.3f' % (e['h2d_gbs_per_gpu'], e['pageable']['h2d_gbs_per_gpu'], e['pageable']['of_pinned']))"
done; done
cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc
