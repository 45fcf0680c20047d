#!/bin/bash
# 1 -> 8 GPU strong-scaling run of the headline workload (GPU box with 8 GPUs)
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -n 1 gpurun_out/scale_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'ms %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
done
