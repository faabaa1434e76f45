cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02_final_n2
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/r02_final_n2/bench_n2.json 2> gpurun_out/r02_final_n2/bench_n2.err
tail -c 600 gpurun_out/r02_final_n2/bench_n2.json
