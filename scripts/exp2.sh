set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/soilmachine_b200/lib
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( export SM_KERNEL=warp
  timeout 900 python -m pytest tests -m gpu -x -q -k "not config3 and not config4 and not ipc" 2>&1 | tail -15
  timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_mb2.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_mb4.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_KERNEL=thread timeout 300 python tests/gpu_probe.py cfg3:both 2
  timeout 300 python tests/gpu_probe.py single
  SM_KERNEL=thread timeout 300 python tests/gpu_probe.py single
  SM_KERNEL=thread timeout 900 python -m pytest tests -m gpu -x -q -k "not config3 and not config4 and not ipc" 2>&1 | tail -5
  timeout 600 python -m pytest tests -m gpu -x -q -k "ipc" 2>&1 | tail -15
  timeout 900 python -m pytest tests -m gpu -x -q -k "config3" 2>&1 | tail -15
  timeout 900 python -m pytest tests -m gpu -x -q -k "config4" 2>&1 | tail -15
) > gpurun_out/r02_exp2.log 2>&1
tail -5 gpurun_out/r02_exp2.log
