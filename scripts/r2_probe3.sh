#!/bin/bash
# literal-kernel rewrite: parity tests, then the tie-heavy workload with the old and the new kernel
mkdir -p gpurun_out/r2p3
timeout 1500 python -m pytest tests/test_exact_fast_gpu.py tests/test_align_gpu.py -x -q -m gpu > gpurun_out/r2p3/pytest.log 2>&1
tail -5 gpurun_out/r2p3/pytest.log
for n in 200000; do
  ZOE_CUDA_EXACT_SLOW=1 python bench.py --config 3 --n $n --scoring 4,-2,-3,-1 --steps 3 --warmup 1 --parity sample > gpurun_out/r2p3/hz_old_n$n.json 2> gpurun_out/r2p3/hz_old_n$n.err
  python bench.py --config 3 --n $n --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p3/hz_new_n$n.json 2> gpurun_out/r2p3/hz_new_n$n.err
done
python bench.py --config 3 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p3/hz_new_1M.json 2> gpurun_out/r2p3/hz_new_1M.err
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p3/cfg3_n125000.json 2> gpurun_out/r2p3/cfg3_n125000.err
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p3/cfg3_1M.json 2> gpurun_out/r2p3/cfg3_1M.err
