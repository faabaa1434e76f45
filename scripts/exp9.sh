set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp9
mkdir -p $O
( SM_EXACT=2 timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_EXACT=3 timeout 300 python tests/gpu_probe.py cfg3:both 2
  timeout 300 python tests/gpu_probe.py cfg3:both 1 ) > $O/timing.log 2>&1
( SM_EXACT=3 timeout 900 python -m pytest tests -m gpu -x -q -k "not config3 and not config4 and not ipc and not lbm and not facade and not hydro" 2>&1 | tail -8 ) > $O/tests_exact3.log 2>&1
( SM_EXACT=3 timeout 900 python -m pytest tests -m gpu -x -q -k "config3" 2>&1 | tail -8 ) > $O/tests_cfg3_exact3.log 2>&1
grep -v "^+" $O/timing.log; cat $O/tests_exact3.log $O/tests_cfg3_exact3.log
