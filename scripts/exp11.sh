set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp11
mkdir -p $O
L=$PWD/soilmachine_b200/lib
(
timeout 300 python tests/gpu_probe.py cfg3:both 2
SM_LIB_PATH=$L/libsm_acqf.so timeout 300 python tests/gpu_probe.py cfg3:both 2
SM_LIB_PATH=$L/libsm_prof.so timeout 300 python tests/gpu_probe.py sweepstat
) > $O/timing.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config3" > $O/tests_cfg3.log 2>&1
grep -v "^+" $O/timing.log | grep "cfg3\|phases\|avg step"; tail -3 $O/tests_cfg3.log
