#!/bin/bash
# Development aid: the round's bench record + ncu launch list + ncu --set full capture, one gpurun call
# (gpurun -- 'bash tests/profile_round.sh'); summaries: python tests/ncu_summarize.py gpurun_out/r2f_prof.ncu-rep
cd /root/repo
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo bench rc=$?
export AW_B200_NO_GRAPH=1
A="--steps 1 --warmup 1 --clips 64 --iters 40 --no-cpu-baseline --no-e2e --no-alt --parity-clips 0 --no-phases"
python bench.py $A > gpurun_out/r2f_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py $A > gpurun_out/r2f_ncu1.log 2>&1; echo launches rc=$?
B="--steps 1 --warmup 1 --clips 64 --iters 6 --no-cpu-baseline --no-e2e --no-alt --parity-clips 0 --no-phases"
python bench.py $B > gpurun_out/r2f_plain2.log 2>&1 && ncu --set full --clock-control none -k regex:'k_gemm_tc|k_gemm_bwd64|k_norm_rows|k_tc_update' -s 57 -c 22 -o gpurun_out/r2f_prof python bench.py $B > gpurun_out/r2f_ncu2.log 2>&1; echo full rc=$?
ls -la gpurun_out | tail -8
