#!/bin/bash
mkdir -p gpurun_out/r2p19
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p19/smoke.log 2>&1; tail -4 gpurun_out/r2p19/smoke.log
for opt in 0,5,16 0,6,16 0,7,16; do
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline --align-opts $opt > gpurun_out/r2p19/cfg3_opts_$opt.json 2> gpurun_out/r2p19/cfg3_opts_$opt.err
done
# share of the table build in the headline kernel (north_star (2)): source-level counters
python bench.py --config 2 --n 200000 --steps 1 --warmup 1 --no-cpu-baseline --legs none > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:sw_score_kernel -s 1 -c 1 -o gpurun_out/r2p19/score_src python bench.py --config 2 --n 200000 --steps 1 --warmup 1 --no-cpu-baseline --legs none > gpurun_out/r2p19/ncu_score_src.log 2>&1
ncu -i gpurun_out/r2p19/score_src.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2p19/score_src.csv 2>/dev/null
rm -f gpurun_out/r2p19/score_src.ncu-rep
