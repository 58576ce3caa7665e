#!/bin/bash
# A/B of the residual kernel variants (one gpurun call): per-call times, kernel-only durations under ncu.
O=gpurun_out
: > $O/kres_ab.txt
for v in 1 2 4; do PCS_RES_OPT=$v python tools/kres_ab.py >> $O/kres_ab.txt 2>&1; done
for v in 1 2 4; do KRES_RIG=128,2500,dome,0.5 PCS_RES_OPT=$v python tools/kres_ab.py >> $O/kres_ab.txt 2>&1; done
for v in 1 2 4; do
  PCS_RES_OPT=$v ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_residual -c 2 -s 6 --csv --log-file $O/kres_ncu_$v.csv python tools/kres_ab.py > /dev/null 2>&1
done
cat $O/kres_ab.txt; for v in 1 2 4; do grep -o 'k_residual[^(]*\|"gpu__time_duration.sum","ns","[0-9]*"\|"sm__warps_active[^,]*","%","[0-9.]*"' $O/kres_ncu_$v.csv | paste - - - - - | head -2; done
