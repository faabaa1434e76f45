set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp7_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( for n in 1 64 2000; do timeout 300 $TR --master-port 2971$((n%7)) tests/multigpu_check.py 1024 $n rocksand 1 2>&1 | grep multigpu_check; done ) > $O/small.log 2>&1
cat $O/small.log
