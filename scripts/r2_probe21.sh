#!/bin/bash
mkdir -p gpurun_out/r2p21
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2p21/pytest_all.log 2>&1; tail -3 gpurun_out/r2p21/pytest_all.log
timeout 300 python bench.py --config 2 --steps 6 --warmup 3 --legs none > gpurun_out/r2p21/cfg2.json 2> gpurun_out/r2p21/cfg2.err
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p21/cfg3.json 2> gpurun_out/r2p21/cfg3.err
timeout 300 python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2p21/cfg1.json 2> gpurun_out/r2p21/cfg1.err
timeout 200 python scripts/soak.py 100 404 > gpurun_out/r2p21/soak_seed404.txt 2>&1; tail -1 gpurun_out/r2p21/soak_seed404.txt
