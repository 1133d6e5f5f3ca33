#!/bin/bash
mkdir -p gpurun_out/r2p17
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p17/pytest_all.log 2>&1; tail -3 gpurun_out/r2p17/pytest_all.log
python bench.py --config 3 --n 125000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p17/cfg3_n125000.json 2> gpurun_out/r2p17/cfg3_n125000.err
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p17/cfg3_1M.json 2> gpurun_out/r2p17/cfg3_1M.err
python bench.py --config 3 --mode 3pass --n 125000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p17/cfg3_3pass_125k.json 2> gpurun_out/r2p17/cfg3_3pass_125k.err
python bench.py --config 2 --n 125000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p17/cfg2_125k.json 2> gpurun_out/r2p17/cfg2_125k.err
