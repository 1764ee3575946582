#!/bin/bash
# One multi-GPU session on the box: bash tools/scale_run.sh N  (c2 default line incl. the strong c3 block, c3, c5 with the
# global batch of 2048 utterances cut over the N ranks).  Lines land in gpurun_out/r02/scale_<workload>_n<N>.json.
N=$1
mkdir -p gpurun_out/r02
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29531 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02/scale_c2_n$N.json 2> gpurun_out/r02/scale_c2_n$N.err
$T --master-port 29532 bench.py --gpus $N --steps 50 --warmup 5 --workload c3 --no-extras > gpurun_out/r02/scale_c3_n$N.json 2> gpurun_out/r02/scale_c3_n$N.err
$T --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload c5 --batch $((2048 / N)) --no-extras > gpurun_out/r02/scale_c5_n$N.json 2> gpurun_out/r02/scale_c5_n$N.err
for f in gpurun_out/r02/scale_c*_n$N.json; do python -c "
import json,sys
l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{')]
if not l: print(sys.argv[1], 'NO LINE'); sys.exit()
d=json.loads(l[-1]); print(sys.argv[1], d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), d['bucket_sizes'], d['parity_check'].get('ok'), (d.get('strong') or {}).get('ms_per_step'))" $f; done
tail -3 gpurun_out/r02/scale_c5_n$N.err
