for mode in peer sync; do
POSEFIT_BENCH_GATHER=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-extra 2>gpurun_out/s22_$mode.err > gpurun_out/s22_n8_$mode.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/s22_n8_$mode.json').read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('$mode', 'ms/step %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'bwd %.3f' % k['fit_backward_kernel']['ms'], d['clocks']['sm_mhz'], 'e2e %.3e' % d['e2e']['value'], d['config']['collective'][:40])"
grep "bench:" gpurun_out/s22_$mode.err | head -2
done
