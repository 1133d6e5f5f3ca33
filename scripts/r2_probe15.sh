#!/bin/bash
mkdir -p gpurun_out/r2p15
timeout 900 python -m pytest tests/test_sneaky_snake.py tests/test_protein_align_scale_gpu.py tests/test_width_policy_gpu.py -x -q -m gpu > gpurun_out/r2p15/pytest.log 2>&1; tail -5 gpurun_out/r2p15/pytest.log
timeout 400 python scripts/soak.py 240 101 > gpurun_out/r2p15/soak_seed101.txt 2>&1; tail -2 gpurun_out/r2p15/soak_seed101.txt
timeout 300 python scripts/soak.py 150 303 > gpurun_out/r2p15/soak_seed303.txt 2>&1; tail -2 gpurun_out/r2p15/soak_seed303.txt
python bench.py --config 1 --steps 20 --warmup 3 --legs sneaky > gpurun_out/r2p15/cfg1_sneaky.json 2> gpurun_out/r2p15/cfg1_sneaky.err; tail -2 gpurun_out/r2p15/cfg1_sneaky.err
