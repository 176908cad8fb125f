#!/bin/bash
out=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer" 2>&1 | tail -1
for pb in 8 16 32 128; do
RT_B200_PEER_BLOCKS=$pb BENCH_TRACE_COMBINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --ns-tris 0 > $out/r2n_n2_pb$pb.json 2> $out/r2n_n2_pb$pb.err
grep "rank 0: a ->" $out/r2n_n2_pb$pb.err | tr '\n' ' '; echo
python -c "
import json
d=json.loads(open('$out/r2n_n2_pb$pb.json').read().strip().splitlines()[-1])
print('ce blocks $pb ms/step %.4f e2e %.4f'%(d['ms_per_step'],d['e2e']['ms_per_frame']), d['verify']['combined_frame_equals_single_gpu_spp_N_frame'])
"
done
