#!/bin/bash
mkdir -p gpurun_out/r2p18
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2p18/pytest_all.log 2>&1; tail -3 gpurun_out/r2p18/pytest_all.log
for v in on off; do
  if [ $v = off ]; then export ZOE_CUDA_NO_STREAMED_UPLOAD=1; else unset ZOE_CUDA_NO_STREAMED_UPLOAD; fi
  timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p18/cfg3_1M_$v.json 2> gpurun_out/r2p18/cfg3_1M_$v.err
  timeout 300 python bench.py --config 2 --steps 6 --warmup 3 --no-cpu-baseline --legs none > gpurun_out/r2p18/cfg2_1M_$v.json 2> gpurun_out/r2p18/cfg2_1M_$v.err
  timeout 300 python bench.py --config 3 --mode 3pass --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2p18/cfg3_3pass_$v.json 2> gpurun_out/r2p18/cfg3_3pass_$v.err
  timeout 300 python bench.py --config 3 --n 125000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p18/cfg3_125k_$v.json 2> gpurun_out/r2p18/cfg3_125k_$v.err
done
