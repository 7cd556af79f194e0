set -x
cd $GRAFT_REPO_ROOT
V=sp-nerf_b200/lib/variants
for t in ws1 ws3w2 ws3w4 ws3w8; do SPNERF_LIB=$PWD/$V/libspnerf_$t.so timeout 300 python tools/ab_mlp.py $t "0;0" > gpurun_out/ab_$t.log 2>&1; tail -n 2 gpurun_out/ab_$t.log; done
for t in ws3w4; do SPNERF_LIB=$PWD/$V/libspnerf_$t.so timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wgrad_kernel -c 2 --csv --log-file gpurun_out/ncu_dram_$t.csv python tools/profile_step.py 8192 2 > gpurun_out/ncu_$t.log 2>&1; tail -n 7 gpurun_out/ncu_dram_$t.csv | cut -d, -f13-; done
SPNERF_LIB=$PWD/$V/libspnerf_ws3w4.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -15
