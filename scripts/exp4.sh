set -x
cd $GRAFT_REPO_ROOT
L=$PWD/soilmachine_b200/lib
O=gpurun_out/r02_exp4
mkdir -p $O
export SM_KERNEL=warp
Q='not config3 and not config4 and not ipc'
( timeout 900 python -m pytest tests -m gpu -x -q -k "$Q" 2>&1 | tail -15 ) > $O/tests_warp.log 2>&1
( SM_EXACT=1 timeout 900 python -m pytest tests -m gpu -x -q -k "$Q and not lbm and not facade" 2>&1 | tail -15 ) > $O/tests_warp_exact.log 2>&1
( timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_EXACT=1 timeout 300 python tests/gpu_probe.py cfg3:water 2
  SM_LIB_PATH=$L/libsm_nopf.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_sl32.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_sl100.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_mb4.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  timeout 300 python tests/gpu_probe.py single
  SM_EXACT=1 timeout 300 python tests/gpu_probe.py big ) > $O/timing.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q -k "ipc" 2>&1 | tail -25 ) > $O/tests_ipc.log 2>&1
( SM_EXACT=1 timeout 900 python -m pytest tests -m gpu -x -q -k "config3" 2>&1 | tail -15 ) > $O/tests_cfg3_exact.log 2>&1
( timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -o $O/kwind python bench.py --steps 1 --warmup 1 --no-cpu --no-extra ) > $O/ncu_wind.log 2>&1
tail -3 $O/*.log
