#!/bin/bash
mkdir -p gpurun_out/r2p11
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2p11/pytest_all.log 2>&1
tail -5 gpurun_out/r2p11/pytest_all.log
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p11/cfg3_n125000.json 2> gpurun_out/r2p11/cfg3_n125000.err
python bench.py --config 3 --steps 8 --warmup 3 > gpurun_out/r2p11/cfg3_1M.json 2> gpurun_out/r2p11/cfg3_1M.err
python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p11/hz_200k.json 2> gpurun_out/r2p11/hz_200k.err
