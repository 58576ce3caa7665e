#!/bin/bash
# K_res on config 4 and on a 128-camera dome x 2500 poses (one gpurun call): per-call time, kernel-only duration under ncu.
# (The variants this script compared in round 2 -- a row cache in shared memory, 2 / 4 observations per thread -- were
# not kept: profiles/r2_kres_variants.txt.)
O=gpurun_out
python tools/kres_ab.py > $O/kres_ab.txt 2>&1
KRES_RIG=128,2500,dome,0.5 python tools/kres_ab.py >> $O/kres_ab.txt 2>&1
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_residual -c 2 -s 6 --csv --log-file $O/kres_ncu.csv python tools/kres_ab.py > /dev/null 2>&1
cat $O/kres_ab.txt; tail -8 $O/kres_ncu.csv
