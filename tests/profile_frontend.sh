#!/bin/bash
# Development aid: ncu --set full of the front-end / adjoint streaming kernels at the bench batch size
# (gpurun -- 'bash tests/profile_frontend.sh'); read with python tests/ncu_summarize.py gpurun_out/r2j_prof.ncu-rep 256
cd /root/repo
export AW_B200_NO_GRAPH=1
B="--steps 1 --warmup 1 --clips 256 --iters 4 --no-cpu-baseline --no-e2e --no-alt --parity-clips 0 --no-phases"
timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:'k_p0_bwd_apply|k_p0_bwd_reduce|k_mel|k_tc_dsprep' -s 12 -c 5 -o gpurun_out/r2j_prof python bench.py $B > gpurun_out/r2j_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out | grep r2j
