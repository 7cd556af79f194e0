set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest2.log
V=sp-nerf_b200/lib/variants
SPNERF_LIB=$PWD/$V/libspnerf_exp.so timeout 300 python tools/ab_mlp.py exp "0;0" "0;1024" "0;2048" "0;3072" "0;4096" "0;6144" > gpurun_out/ab_exp.log 2>&1
cat gpurun_out/r2_pytest2.log; tail -n 20 gpurun_out/ab_exp.log
