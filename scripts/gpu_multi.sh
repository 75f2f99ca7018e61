#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/multigpu_check_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/multigpu_check_n$N.log
grep -v Warn gpurun_out/multigpu_check_n$N.log | tail -8
for n in 1 $N; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2>gpurun_out/bench_n$n.err; fi
  tail -1 gpurun_out/bench_n$n.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('n_gpus', d['n_gpus'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'ms/step %.3f' % d['ms_per_step'], 'svgd', d.get('svgd'))"
done
