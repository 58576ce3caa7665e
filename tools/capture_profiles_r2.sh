#!/bin/bash
# Round-2 evidence (one gpurun call, 1 GPU): plain runs first, then the ncu launch lists and full captures of the same commands.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2_bench_final.json 2> $O/r2_bench_final.err || exit 1
tail -c 600 $O/r2_bench_final.json
python tools/lm_profile.py 6 ring32 > $O/r2_lm_ring32.txt 2>&1 || exit 1
python tools/lm_profile.py 6 ccube_selfcal > $O/r2_lm_selfcal.txt 2>&1 || exit 1
python tools/lm_profile.py 4 dome128:2500 > $O/r2_lm_dome.txt 2>&1 || exit 1
python tools/lm_profile.py 6 ring32 mixed > $O/r2_lm_ring32_mixed.txt 2>&1 || exit 1
K='regex:^k_|gemv|potrf|trsm|syrk|getrf|gemm'
ncu --metrics gpu__time_duration.sum --clock-control none -k $K -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-lm --no-config5 --no-lm-e2e > $O/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k $K -c 600 --csv --log-file $O/r2_launches_lm_ring32.csv python tools/lm_profile.py 6 ring32 > $O/ncu_lm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k $K -c 600 --csv --log-file $O/r2_launches_lm_selfcal.csv python tools/lm_profile.py 6 ccube_selfcal > $O/ncu_lm2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k $K -c 600 --csv --log-file $O/r2_launches_lm_dome.csv python tools/lm_profile.py 4 dome128:2500 > $O/ncu_lm3.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^k_normal$' -c 1 -s 3 -o $O/prof_r2_kne_fp64 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-lm --no-config5 --no-lm-e2e > $O/ncu_kne.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_normal_mixed -c 1 -s 3 -o $O/prof_r2_kne_mixed -f python bench.py --steps 3 --warmup 3 --no-cpu --no-lm --no-config5 --no-lm-e2e > $O/ncu_kne_mixed.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_residual -c 1 -s 3 -o $O/prof_r2_kres -f python bench.py --steps 3 --warmup 3 --no-cpu --no-lm --no-config5 --no-lm-e2e > $O/ncu_kres.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_jacobian -c 1 -s 3 -o $O/prof_r2_kjac -f python bench.py --steps 3 --warmup 3 --no-cpu --no-lm --no-config5 --no-lm-e2e > $O/ncu_kjac.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_schur_syrk -c 1 -s 2 -o $O/prof_r2_syrk -f python tools/lm_profile.py 4 ring32 > $O/ncu_syrk.log 2>&1
ls -la $O/*.ncu-rep | tail -6
