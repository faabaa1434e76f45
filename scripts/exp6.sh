set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp6
mkdir -p $O
( timeout 300 python tests/gpu_probe.py cfg3:both 2
  timeout 300 python tests/gpu_probe.py single ) > $O/timing.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "not config3 and not config4 and not ipc and not lbm and not facade and not hydro" 2>&1 | tail -8 ) > $O/tests.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "config3" 2>&1 | tail -8 ) > $O/tests_cfg3.log 2>&1
tail -3 $O/*.log
