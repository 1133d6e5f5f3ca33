#!/bin/bash
mkdir -p gpurun_out/r2p6
timeout 1500 python -m pytest tests/test_exact_fast_gpu.py tests/test_align_gpu.py tests/test_long_ranges_gpu.py -x -q -m gpu > gpurun_out/r2p6/pytest.log 2>&1
tail -15 gpurun_out/r2p6/pytest.log
for sl in 16 64 128; do
python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 --align-opts 0,6,$sl > gpurun_out/r2p6/hz_n200000_slack$sl.json 2> gpurun_out/r2p6/hz_n200000_slack$sl.err
done
python bench.py --config 3 --steps 5 --warmup 2 --no-cpu-baseline --align-opts 0,6,64 > gpurun_out/r2p6/cfg3_slack64.json 2> gpurun_out/r2p6/cfg3_slack64.err
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p6/cfg3_n125000.json 2> gpurun_out/r2p6/cfg3_n125000.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p6/cfg4_3pass_20k.json 2> gpurun_out/r2p6/cfg4_3pass_20k.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p6/launches_cfg4_3pass.csv python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p6/ncu_cfg4.log 2>&1
python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p6/launches_hz_200k.csv python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p6/ncu_hz.log 2>&1
