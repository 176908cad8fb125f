#!/bin/bash
out=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer" 2>&1 | tail -2
for mode in ce kernel; do
RT_B200_PEER_GATHER=$mode BENCH_TRACE_COMBINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --ns-tris 0 > $out/r2g_n2_$mode.json 2> $out/r2g_n2_$mode.err
grep "rank 0: a ->" $out/r2g_n2_$mode.err
python -c "
import json
d=json.loads(open('$out/r2g_n2_$mode.json').read().strip().splitlines()[-1])
print('gather $mode ms/step %.4f e2e %.4f'%(d['ms_per_step'],d['e2e']['ms_per_frame']), d['verify'])
"
done
