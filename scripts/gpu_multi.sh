#!/bin/bash
# N-GPU validation: row-tiled field/inflation + query-parallel batch against single-GPU results, then the bench line
N=${1:-2}; TAG=${2:-multi$N}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/tiled_check.py --size 8192 --out $OUT/tiled_check.json > $OUT/tiled_check.log 2>&1; echo "tiled_check rc=$?"; tail -5 $OUT/tiled_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-extras > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; tail -c 1500 $OUT/bench_n$N.json
