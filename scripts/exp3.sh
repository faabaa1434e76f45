# multi-GPU evaluation: usage  bash scripts/exp3.sh N
set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp3_n$N
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpu.txt
nvidia-smi topo -m >> $O/gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export SM_KERNEL=warp
( timeout 900 $TR --master-port 29701 tests/multigpu_check.py 1024 8000 rockgravelpebblessand 2 ) > $O/check_warp.log 2>&1
( SM_EXACT=1 timeout 900 $TR --master-port 29702 tests/multigpu_check.py 1024 8000 rockgravelpebblessand 1 ) > $O/check_warp_exact.log 2>&1
( timeout 1200 $TR --master-port 29703 bench.py --gpus $N --steps 3 --warmup 3 ) > $O/bench_warp.json 2> $O/bench_warp.err
( SM_EXACT=1 timeout 1200 $TR --master-port 29704 bench.py --gpus $N --steps 3 --warmup 3 --no-extra ) > $O/bench_warp_exact.json 2> $O/bench_warp_exact.err
( SM_KERNEL=thread timeout 1200 $TR --master-port 29705 bench.py --gpus $N --steps 2 --warmup 2 --no-extra ) > $O/bench_thread.json 2> $O/bench_thread.err
tail -2 $O/*.log; cat $O/*.json | cut -c1-600
