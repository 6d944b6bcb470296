#!/bin/bash
# usage: train_scale.sh "1 2" -- decoder-training scaling lines for the listed GPU counts
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in $1; do
  if [ $n = 1 ]; then python bench.py --workload ffhq_train --steps 100 --warmup 5 --no-cpu-baseline 2>>$O/errt.log | tail -1 > $O/r02_scale_train_1gpu.json
  else $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --workload ffhq_train --steps 100 --warmup 5 2>>$O/errt.log | tail -1 > $O/r02_scale_train_${n}gpu.json; fi
  python -c "import json; d=json.load(open('$O/r02_scale_train_${n}gpu.json')); print($n, round(d['value'],1), round(d['ms_per_step'],3), d.get('allreduce'))"
done
tail -2 $O/errt.log
