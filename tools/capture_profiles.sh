#!/bin/bash
# Round-end evidence: plain bench run first, then the ncu launch list and full captures of the same commands.
set -x
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
tail -1 gpurun_out/bench_final.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|gemv|Kernel2" -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-lm > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|gemv|Kernel2" -c 400 --csv --log-file gpurun_out/launches_lm.csv python tools/lm_profile.py 6 > gpurun_out/ncu_lm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_normal -c 1 -s 3 -o gpurun_out/prof_r1_final_kne -f python bench.py --steps 3 --warmup 3 --no-cpu --no-lm > gpurun_out/ncu_kne.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_schur_syrk -c 1 -s 2 -o gpurun_out/prof_r1_final_schur -f python tools/lm_profile.py 4 > gpurun_out/ncu_schur.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chol_solve -c 1 -s 2 -o gpurun_out/prof_r1_final_chol -f python tools/lm_profile.py 4 > gpurun_out/ncu_chol.log 2>&1
tools/chol_trace 480 > gpurun_out/chol_trace_480.txt 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
