set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=$PWD/soilmachine_b200/lib/libsm_prof.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( SM_LIB_PATH=$P timeout 300 python tests/gpu_probe.py sweepcurve
  SM_LIB_PATH=$P timeout 300 python tests/gpu_probe.py tail
  SM_LANES=32 SM_LIB_PATH=$P timeout 300 python tests/gpu_probe.py tail
  SM_LANES=4 SM_LIB_PATH=$P timeout 300 python tests/gpu_probe.py tail
  for L in 2 4 8 16 32; do SM_LANES=$L timeout 300 python tests/gpu_probe.py cfg3:both 1; done
  SM_LIB_PATH=$P timeout 300 python tests/gpu_probe.py profile
  SM_LANES=32 SM_LIB_PATH=$P timeout 300 python tests/gpu_probe.py profile
) > gpurun_out/r02_exp1.log 2>&1
tail -5 gpurun_out/r02_exp1.log
