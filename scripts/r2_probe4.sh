#!/bin/bash
mkdir -p gpurun_out/r2p4
timeout 1500 python -m pytest tests/test_exact_fast_gpu.py tests/test_align_gpu.py -x -q -m gpu > gpurun_out/r2p4/pytest.log 2>&1
tail -3 gpurun_out/r2p4/pytest.log
ZOE_CUDA_EXACT_SLOW=1 python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 --parity sample > gpurun_out/r2p4/hz_old_n200000.json 2> gpurun_out/r2p4/hz_old_n200000.err
python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p4/hz_new_n200000.json 2> gpurun_out/r2p4/hz_new_n200000.err
python bench.py --config 3 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p4/hz_new_1M.json 2> gpurun_out/r2p4/hz_new_1M.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p4/launches_cfg3_125k.csv python bench.py --config 3 --n 125000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p4/ncu_cfg3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p4/launches_hz_200k.csv python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p4/ncu_hz.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p4/pytest_all.log 2>&1
tail -3 gpurun_out/r2p4/pytest_all.log
ZOE_CUDA_SCAN1=1 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p4/cfg3_scan1.json 2> gpurun_out/r2p4/cfg3_scan1.err
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p4/cfg3_scan2.json 2> gpurun_out/r2p4/cfg3_scan2.err
python bench.py --config 3 --mode ranges --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p4/cfg3_ranges_scan2.json 2> gpurun_out/r2p4/cfg3_ranges_scan2.err
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p4/cfg3_n125000.json 2> gpurun_out/r2p4/cfg3_n125000.err
