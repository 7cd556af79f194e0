cd $GRAFT_REPO_ROOT
V=$PWD/sp-nerf_b200/lib/variants
for i in 1 2 3 4; do
  for t in $ABTAGS; do SPNERF_LIB=$V/libspnerf_$t.so timeout 100 python tools/ab_mlp.py $t 0,1 2>&1 | grep "round 1\|Error\|error" | head -3; done
done
