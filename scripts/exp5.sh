set -x
cd $GRAFT_REPO_ROOT
L=$PWD/soilmachine_b200/lib
O=gpurun_out/r02_exp5
mkdir -p $O
export SM_KERNEL=warp
Q='not config3 and not config4 and not ipc'
( timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_EXACT=1 timeout 300 python tests/gpu_probe.py cfg3:water 2
  SM_LIB_PATH=$L/libsm_mb4.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_nopf.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  timeout 300 python tests/gpu_probe.py single
  SM_EXACT=1 timeout 300 python tests/gpu_probe.py big ) > $O/timing.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "$Q" 2>&1 | tail -15 ) > $O/tests_warp.log 2>&1
( SM_EXACT=1 timeout 900 python -m pytest tests -m gpu -x -q -k "$Q and not lbm and not facade" 2>&1 | tail -15 ) > $O/tests_warp_exact.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q -k "ipc" 2>&1 | tail -25 ) > $O/tests_ipc.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "config3" 2>&1 | tail -15 ) > $O/tests_cfg3.log 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu ) > $O/bench_warp.json 2> $O/bench_warp.err
tail -3 $O/*.log
