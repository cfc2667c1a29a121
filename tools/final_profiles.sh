#!/bin/bash
# Everything profiles/ holds for a round, from ONE gpurun call on one B200 (run from the repo root):
#   bash tools/final_profiles.sh r02_final
# Writes into gpurun_out/<prefix>_*; copy the text / json / csv files into profiles/ afterwards (the .ncu-rep files are
# deleted at the end: gpurun merges at most 64 MiB back, and everything needed has been extracted from them by then).
P=${1:-r02_final}
O=gpurun_out
mkdir -p $O
# bench lines (never under a profiler)
python bench.py > $O/${P}_bench_c2_n1.json 2> $O/${P}_bench_c2_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_reference_arm.json 2>/dev/null
python bench.py --config c3 --no-cpu-baseline > $O/${P}_bench_c3_n1.json 2>/dev/null
python bench.py --config c4 > $O/${P}_bench_c4_n1.json 2>/dev/null
python bench.py --config c5 --steps 5 > $O/${P}_bench_c5_n1.json 2>/dev/null
python bench.py --config c5 --impl reference --steps 2 --warmup 0 > $O/${P}_bench_c5_reference_arm.json 2>/dev/null
# launch list of a short bench run (per-launch device times; shares must agree with the CUDA-event timing)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > /dev/null 2>&1
# full captures of the three kernels, config 2 and config 4
for c in c2 c4; do
  ncu --set full --clock-control none --import-source on -k regex:"prune_kernel|beam_kernel|finalize_kernel" -c 3 -o $O/${P}_$c python tools/profile_c2.py --config $c --iters 1 > $O/ncu_$c.log 2>&1
  python tools/ncu_summary.py $O/${P}_$c.ncu-rep $O/${P}_${c}_ncu_full.txt "ncu --set full --clock-control none, tools/profile_c2.py --config $c (prune, beam, finalize kernel of one decode)" > /dev/null 2>&1
  python tools/ncu_traffic.py $O/${P}_$c.ncu-rep $c $O/traffic.json > /dev/null 2>&1
done
python tools/ncu_lines.py $O/${P}_c2.ncu-rep 80 > $O/${P}_c2_beam_source_lines.txt 2>&1
ncu -i $O/${P}_c2.ncu-rep --page raw --csv > $O/${P}_c2_raw.csv 2>/dev/null
rm -f $O/${P}_c2.ncu-rep $O/${P}_c4.ncu-rep
# one frame of the scorer path (config 5, one launch per frame so that the profiler can replay it) + its region cycles
ncu --set full --clock-control none --import-source on -k regex:beam_kernel -s 600 -c 1 -o $O/${P}_c5_frame python tools/profile_lm_frame.py > $O/ncu_c5.log 2>&1
python tools/ncu_lines.py $O/${P}_c5_frame.ncu-rep 60 > $O/${P}_c5_frame_source_lines.txt 2>&1
python tools/ncu_summary.py $O/${P}_c5_frame.ncu-rep $O/${P}_c5_frame_ncu_full.txt "ncu --set full, one frame launch of the scorer path (config 5, CTCDEC_LM_PER_FRAME=1)" > /dev/null 2>&1
rm -f $O/${P}_c5_frame.ncu-rep
python tools/lm_timing.py quick > $O/${P}_region_cycles_c5.txt 2>&1
# per-region cycles of the instrumented beam kernel
python tools/region_timing.py > $O/${P}_region_cycles_c2.txt 2>&1
python tools/region_timing.py --config c4 --batch 256 > $O/${P}_region_cycles_c4.txt 2>&1
# racecheck over every decode path
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_run.py > $O/${P}_racecheck.txt 2>&1
tail -3 $O/${P}_racecheck.txt > $O/${P}_racecheck_summary.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/${P}_memcheck.txt 2>&1
tail -3 $O/${P}_memcheck.txt > $O/${P}_memcheck_summary.txt
# tie report on the device itself
python tools/tie_report.py --gpu > $O/${P}_tie_report_c2.txt 2>&1
