#!/bin/bash
mkdir -p gpurun_out/r2p10
timeout 1500 python -m pytest tests/test_exact_fast_gpu.py tests/test_align_gpu.py -x -q -m gpu > gpurun_out/r2p10/pytest.log 2>&1
tail -15 gpurun_out/r2p10/pytest.log
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p10/cfg3_n125000.json 2> gpurun_out/r2p10/cfg3_n125000.err
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p10/cfg3_1M.json 2> gpurun_out/r2p10/cfg3_1M.err
for sc in 2,-3,-4,-1 1,-4,-6,-1 3,-4,-5,-2 2,-5,-5,-1; do
python bench.py --config 3 --n 200000 --scoring $sc --steps 3 --warmup 1 --parity sample > gpurun_out/r2p10/hz_$sc.json 2> gpurun_out/r2p10/hz_$sc.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p10/launches_cfg3_125k.csv python bench.py --config 3 --n 125000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p10/ncu_cfg3.log 2>&1
