# multi-GPU evaluation, short form: usage  bash scripts/exp3b.sh N
set -x
N=${1:-8}
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp3_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29701 tests/multigpu_check.py 2048 16000 rockgravelpebblessand 1 ) > $O/check_warp.log 2>&1
( timeout 900 $TR --master-port 29703 bench.py --gpus $N --steps 3 --warmup 3 ) > $O/bench_warp.json 2> $O/bench_warp.err
grep multigpu_check $O/check_warp.log; cut -c1-300 $O/bench_warp.json
