set -x
cd $GRAFT_REPO_ROOT
V=sp-nerf_b200/lib/variants
for t in exp ld1 ld2; do SPNERF_LIB=$PWD/$V/libspnerf_$t.so timeout 300 python tools/ab_mlp.py $t "0;0" > gpurun_out/ab_$t.log 2>&1; tail -n 2 gpurun_out/ab_$t.log; done
