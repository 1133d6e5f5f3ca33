#!/bin/bash
mkdir -p gpurun_out/r2p14
python bench.py --config 4 --n 20000 --mode ranges --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p14/cfg4_ranges_20k.json 2> gpurun_out/r2p14/cfg4_ranges_20k.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p14/cfg4_3pass_20k.json 2> gpurun_out/r2p14/cfg4_3pass_20k.err
timeout 600 python -m pytest tests/test_long_ranges_gpu.py -x -q -m gpu > gpurun_out/r2p14/pytest.log 2>&1; tail -3 gpurun_out/r2p14/pytest.log
timeout 400 python scripts/soak.py 240 101 > gpurun_out/r2p14/soak_seed101.txt 2>&1; tail -3 gpurun_out/r2p14/soak_seed101.txt
timeout 300 python scripts/soak.py 150 202 > gpurun_out/r2p14/soak_seed202.txt 2>&1; tail -3 gpurun_out/r2p14/soak_seed202.txt
ZOE_CUDA_DEBUG=1 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p14/cfg3_dbg.json 2> gpurun_out/r2p14/cfg3_dbg.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2p14/launches_headline.csv python bench.py --steps 2 --warmup 1 --legs none --no-cpu-baseline > gpurun_out/r2p14/ncu_headline.log 2>&1
