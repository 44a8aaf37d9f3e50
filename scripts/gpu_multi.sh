#!/bin/bash
# Multi-GPU call: gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_multi.sh N tag'
N=${1:-2}; TAG=${2:-r1}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus_${TAG}.txt
echo "== multi-GPU parity"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mp_model_check.py > $OUT/mp_check_${TAG}.log 2>&1
echo "mp check exit $?"; tail -n 6 $OUT/mp_check_${TAG}.log | cut -c1-1500
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_${TAG}.json 2> $OUT/bench_n1_${TAG}.err; echo "exit $?"
cut -c1-300 $OUT/bench_n1_${TAG}.json; tail -n 3 $OUT/bench_n1_${TAG}.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 3 > $OUT/bench_n${n}_${TAG}.json 2> $OUT/bench_n${n}_${TAG}.err
    echo "exit $?"; grep '^{' $OUT/bench_n${n}_${TAG}.json | cut -c1-400; tail -n 5 $OUT/bench_n${n}_${TAG}.err
  fi
done
