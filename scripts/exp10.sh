set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp10_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( SM_SWEEPSTAT=1 SM_LIB_PATH=$PWD/soilmachine_b200/lib/libsm_prof.so timeout 600 $TR --master-port 29721 tests/multigpu_check.py 4096 25000 rockgravelpebblessand 1 2>&1 | grep "rank\|multigpu_check" ) > $O/stat.log 2>&1
cat $O/stat.log
