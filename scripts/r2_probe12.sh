#!/bin/bash
mkdir -p gpurun_out/r2p12
timeout 1200 python -m pytest tests/test_align_gpu.py tests/test_chunking_gpu.py tests/test_fullscale_gpu.py -x -q -m gpu > gpurun_out/r2p12/pytest.log 2>&1
tail -3 gpurun_out/r2p12/pytest.log
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p12/cfg3_n125000.json 2> gpurun_out/r2p12/cfg3_n125000.err
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p12/cfg3_1M.json 2> gpurun_out/r2p12/cfg3_1M.err
