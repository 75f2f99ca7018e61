#!/bin/bash
# bench lines at N GPUs (N = $1): the default W2 line (HMC + svgd block) and the W4 / W7 / W6 headline lines, plus the reference arm at N = 1
N=${1:-2}
TAG=${TAG:-r02}
mkdir -p gpurun_out
run() {   # workload, extra args
  if [ "$N" = "1" ]; then
    python bench.py --workload $1 $2 > gpurun_out/${TAG}_bench_n${N}_$1.json 2> gpurun_out/${TAG}_bench_n${N}_$1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $1 $2 \
      > gpurun_out/${TAG}_bench_n${N}_$1.json 2> gpurun_out/${TAG}_bench_n${N}_$1.err
  fi
  echo "$1 rc=$? $(head -c 300 gpurun_out/${TAG}_bench_n${N}_$1.json)"
}
for w in ${WLS:-W2 W4 W7 W6}; do run $w; done
if [ "${CHECK:-1}" = "1" ] && [ "$N" != "1" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n$N.log 2>&1
  grep -v "^$" gpurun_out/${TAG}_multigpu_check_n$N.log | grep "ranks\|SVGD\|BBB\|MULTIGPU" | tail -14
fi
