set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --skip-extra > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --skip-extra > gpurun_out/ncu_launches.log 2>&1
python tools/profile_step.py 8192 3 > gpurun_out/ncu_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mlp_fwd_kernel|mlp_bwd_kernel|wgrad_kernel" -s 3 -c 3 -o gpurun_out/prof_r02_mlp -f python tools/profile_step.py 8192 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
tail -3 gpurun_out/ncu_full.log
