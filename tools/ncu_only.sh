P=r02_final; O=gpurun_out
for c in c2 c4; do
  rm -f $O/${P}_$c.ncu-rep
  ncu --set full --clock-control none --import-source on -k regex:"prune_kernel|beam_kernel|finalize_kernel" -c 3 -o $O/${P}_$c python tools/profile_c2.py --config $c --iters 1 > $O/ncu_$c.log 2>&1
  python tools/ncu_summary.py $O/${P}_$c.ncu-rep $O/${P}_${c}_ncu_full.txt "ncu --set full --clock-control none, tools/profile_c2.py --config $c (prune, beam, finalize kernel of one decode)" > /dev/null 2>&1
  python tools/ncu_traffic.py $O/${P}_$c.ncu-rep $c $O/traffic_r02.json > /dev/null 2>&1
done
python tools/ncu_lines.py $O/${P}_c2.ncu-rep 80 > $O/${P}_c2_beam_source_lines.txt 2>&1
ncu -i $O/${P}_c2.ncu-rep --page raw --csv > $O/${P}_c2_raw.csv 2>/dev/null
# the reports themselves stay on the box (gpurun merges at most 64 MiB back): everything needed has been extracted
rm -f $O/${P}_c2.ncu-rep $O/${P}_c4.ncu-rep
