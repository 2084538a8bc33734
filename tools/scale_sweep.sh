#!/bin/bash
# 1 -> N GPU scaling of bench.py on one box (the driver's SCALE run, for the record in profiles/)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline 2>/dev/null | grep "^{" | tail -1 > gpurun_out/scale_n$n.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 6 --warmup 3 2>/dev/null | grep "^{" | tail -1 > gpurun_out/scale_n$n.json; fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_n$n.json')); print($n, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['clocks'])"
done
