#!/bin/bash
mkdir -p gpurun_out/r2p20
timeout 900 python -m pytest tests/test_align_gpu.py tests/test_chunking_gpu.py tests/test_exact_fast_gpu.py -x -q -m gpu > gpurun_out/r2p20/pytest.log 2>&1; tail -3 gpurun_out/r2p20/pytest.log
for v in on off; do
  if [ $v = off ]; then export ZOE_CUDA_NO_PIN_OVERLAP=1; else unset ZOE_CUDA_NO_PIN_OVERLAP; fi
  for n in 125000 250000; do
  timeout 300 python bench.py --config 3 --n $n --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2p20/cfg3_${n}_$v.json 2> gpurun_out/r2p20/cfg3_${n}_$v.err
  done
done
unset ZOE_CUDA_NO_PIN_OVERLAP
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p20/cfg3_1M.json 2> gpurun_out/r2p20/cfg3_1M.err
