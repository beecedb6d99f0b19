#!/bin/bash
# N-GPU run of round 2: the bench line as the driver launches it (per-rank stats, multi-GPU bit-exact checks), the CPU arm
# under torchrun, and the cfg5 line (row-tiled field over N slabs next to rank 0's one-GPU numbers).
N=${1:-2}; TAG=${2:-r02_multi$N}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "ref rc=$?"; tail -c 900 $OUT/bench_ref_n$N.json; echo
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; tail -c 3500 $OUT/bench_n$N.json; echo; tail -3 $OUT/bench_n$N.err
timeout 900 $TR --master-port 29513 bench.py --config cfg5 --gpus $N --steps 2 --warmup 1 > $OUT/bench_cfg5_n$N.json 2> $OUT/bench_cfg5_n$N.err; echo "cfg5 rc=$?"; tail -c 2500 $OUT/bench_cfg5_n$N.json; echo; tail -3 $OUT/bench_cfg5_n$N.err
ls -la $OUT
