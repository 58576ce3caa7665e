#!/bin/bash
# A/B of the K_jac copy-out (bulk asynchronous copy vs load / store loop), per-call rates + kernel-only durations under ncu
O=gpurun_out
: > $O/kjac_ab.txt
for v in 0 1 0 1; do echo "PCS_JAC_BULK=$v" >> $O/kjac_ab.txt; PCS_JAC_BULK=$v python tools/kernel_rates.py >> $O/kjac_ab.txt 2>&1; done
for v in 0 1; do
  PCS_JAC_BULK=$v ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_jacobian -c 2 -s 4 --csv --log-file $O/kjac_ncu_$v.csv python tools/kernel_rates.py > /dev/null 2>&1
  grep -o '"[a-z_0-9.]*","[%a-z]*","[0-9.,]*"$' $O/kjac_ncu_$v.csv | tail -6 >> $O/kjac_ab.txt
done
cat $O/kjac_ab.txt
