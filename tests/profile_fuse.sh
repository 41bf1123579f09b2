#!/bin/bash
# Development aid: ncu --set full of the K = 1024 forward GEMM with InstanceNorm fused into its epilogue
# (EPI_FWD_FUSE, AW_B200_FUSE_NORM=1) and of the ordinary EPI_FWD GEMM + k_norm_rows it replaces, same reduced
# bench command (64 clips, 6 iterations); gpurun -- 'bash tests/profile_fuse.sh'
cd /root/repo
export AW_B200_NO_GRAPH=1
B="--steps 1 --warmup 1 --clips 64 --iters 6 --no-cpu-baseline --no-e2e --no-alt --parity-clips 0 --no-phases"
for f in 1 0; do
  export AW_B200_FUSE_NORM=$f
  [ $f -eq 1 ] && C=5 || C=7       # first iteration: the two K >= 512 forward layers (+ their k_norm_rows passes)
  timeout 200 python bench.py $B > gpurun_out/r2i_plain_f$f.log 2>&1; rc=$?; echo "plain fuse=$f rc=$rc"
  [ $rc -eq 0 ] && timeout 400 ncu --set full --clock-control none \
    -k regex:'k_gemm_tc|k_norm_rows' -s 3 -c $C -o gpurun_out/r2i_prof_f$f python bench.py $B > gpurun_out/r2i_ncu_f$f.log 2>&1
  echo "ncu fuse=$f rc=$?"
done
ls -la gpurun_out | grep r2i
