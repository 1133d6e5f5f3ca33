#!/bin/bash
# round-2 probe: shard-sized batches on one GPU (what an 8-GPU strong-scaling rank sees)
set -x
mkdir -p gpurun_out/r2p1
for cfg in 2 3; do
  for n in 125000 250000 1000000; do
    python bench.py --config $cfg --n $n --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p1/cfg${cfg}_n${n}.json 2> gpurun_out/r2p1/cfg${cfg}_n${n}.err
  done
done
ZOE_CUDA_DEBUG=1 python bench.py --config 3 --n 125000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p1/cfg3_dbg.json 2> gpurun_out/r2p1/cfg3_dbg.err
python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2p1/cfg1.json 2> gpurun_out/r2p1/cfg1.err
python bench.py --config 3 --mode 3pass --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p1/cfg3_3pass_n125000.json 2>&1
python bench.py --config 3 --mode ranges --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p1/cfg3_ranges_n125000.json 2>&1
nproc > gpurun_out/r2p1/nproc.txt; lscpu | head -30 >> gpurun_out/r2p1/nproc.txt
