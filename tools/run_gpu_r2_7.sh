set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
cat gpurun_out/allreduce_check.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --skip-extra > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; tail -c 600 gpurun_out/bench_r2_2gpu.err; cut -c1-700 gpurun_out/bench_r2_2gpu.json
