#!/bin/bash
# Round-2 evidence: default bench, its ncu launch list, and one `ncu --set full` capture per kernel that changed.
O=gpurun_out/r2ev; mkdir -p $O
(time python bench.py --steps 10 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --config 3 --impl reference --steps 2 --warmup 1 > $O/bench_reference_cfg3.json 2> $O/bench_reference_cfg3.err
# launch list of the default command (short: 1 step), only after the command ran clean above
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, bench args...
  name=$1; k=$2; skip=$3; shift 3
  ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -o $O/$name python bench.py "$@" --no-cpu-baseline --legs none > $O/ncu_$name.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  python scripts/ncu_summary.py $O/$name.raw.csv > $O/$name.txt 2>/dev/null
}
cap score_cfg2 sw_score_kernel 1 --config 2 --steps 1 --warmup 1
cap scan_cfg3 sw_align_scan_kernel 1 --config 3 --steps 1 --warmup 1
cap winfill_cfg3 'sw_align_winfill_kernel<8, 19, 1>' 2 --config 3 --steps 1 --warmup 1
cap exact_fast_tieheavy sw_exact_fast_kernel 1 --config 3 --n 200000 --scoring 2,-3,-4,-1 --steps 1 --warmup 1
cap ends_long_cfg4 'sw_ends_long_kernel<24, 1, 0>' 1 --config 4 --n 20000 --mode ranges --steps 1 --warmup 1
cap tp_band_cfg4 tp_band_warp_kernel 1 --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1
cap rows_stream_cfg5 sw_score_rows_stream_kernel 1 --config 5 --steps 1 --warmup 1
cap long_cfg4 sw_score_long_kernel 1 --config 4 --n 20000 --steps 1 --warmup 1
rm -f $O/*.ncu-rep
ls -la $O
