#!/bin/bash
# Round-2 multi-GPU evidence on ONE box with N GPUs (usage: tools/multi_gpu_session.sh N): generate scaling, decoder-training
# scaling with the all-reduce timed, and the config-5 dataset sweep through the main.py-compatible CLI.
N=${1:-8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nproc; free -g | head -2
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $O/r02_scale_gen_1gpu.json 2>$O/errm.log
for n in 2 4 8; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 3 > $O/r02_scale_gen_${n}gpu.json 2>>$O/errm.log
done
python bench.py --workload ffhq_train --steps 100 --warmup 5 > $O/r02_scale_train_1gpu.json 2>>$O/errm.log
for n in 2 4 8; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --workload ffhq_train --steps 100 --warmup 5 > $O/r02_scale_train_${n}gpu.json 2>>$O/errm.log
done
# config 5: 10 000 FFHQ image + mask pairs to files (JPEG + PNG), and the same sweep without the encoders
mkdir -p /tmp/exp5
printf 'BASE_DIR: "/tmp/exp5"\nGAN: "ffhq"\nGAN_DIR: "none"\nGAN_GPU_IDS: [0]\nGAN_BATCH_SIZE_PER_GPU: 32\nSOLVER_GPU_IDS: [0]\nANNOTATION: "segmentation"\nGENERATE_NUM: 10000\n' > /tmp/exp5/config.yml
printf 'BASE_DIR: "/tmp/exp5"\nGAN: "ffhq"\nGAN_DIR: "none"\nGAN_GPU_IDS: [0]\nGAN_BATCH_SIZE_PER_GPU: 32\nSOLVER_GPU_IDS: [0]\nANNOTATION: "segmentation"\nGENERATE_NUM: 2000\n' > /tmp/exp5/config1.yml
{
  echo "== config 5 sweep, $(nproc) host cores"
  python -m gan_segmentation_b200.main generate --config /tmp/exp5/config1.yml --random-init --psi 0.7 --no-write 2>>$O/errm.log | tail -1
  python -m gan_segmentation_b200.main generate --config /tmp/exp5/config1.yml --random-init --psi 0.7 2>>$O/errm.log | tail -1
  ls /tmp/exp5/dataset/train_generated | wc -l; du -sh /tmp/exp5/dataset/train_generated | cut -f1; rm -rf /tmp/exp5/dataset
  if [ $N -ge 2 ]; then
    $TR --nproc-per-node $N --master-port 29533 -m gan_segmentation_b200.main generate --config /tmp/exp5/config.yml --random-init --psi 0.7 --no-write 2>>$O/errm.log | tail -1
    $TR --nproc-per-node $N --master-port 29534 -m gan_segmentation_b200.main generate --config /tmp/exp5/config.yml --random-init --psi 0.7 2>>$O/errm.log | tail -1
    ls /tmp/exp5/dataset/train_generated | wc -l; du -sh /tmp/exp5/dataset/train_generated | cut -f1
  fi
} > $O/r02_config5_sweep.txt 2>&1
for f in $O/r02_scale_*.json; do python -c "import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],3), round(d.get('e2e',{}).get('value',0),1), d.get('allreduce'))"; done
cat $O/r02_config5_sweep.txt; tail -3 $O/errm.log
