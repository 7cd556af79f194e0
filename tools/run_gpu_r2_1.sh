set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest1.log
V=sp-nerf_b200/lib/variants
for t in r96 r112; do SPNERF_LIB=$PWD/$V/libspnerf_$t.so timeout 300 python tools/ab_mlp.py $t 0,1 12000,2 24000,2 24000,8 48000,8 96000,16 > gpurun_out/ab_$t.log 2>&1; done
for t in r112w4 r112w5; do SPNERF_LIB=$PWD/$V/libspnerf_$t.so timeout 300 python tools/ab_mlp.py $t 0,1 24000,8 > gpurun_out/ab_$t.log 2>&1; done
cat gpurun_out/r2_pytest1.log; tail -n 20 gpurun_out/ab_*.log
