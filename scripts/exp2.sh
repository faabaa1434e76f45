# one GPU slot, everything we need from it; every step has its own timeout and log
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/soilmachine_b200/lib
O=gpurun_out/r02_exp2
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
export SM_KERNEL=warp SM_HYDRO=warp
Q='not config3 and not config4 and not ipc and not facade_per_particle and not lbm'
( timeout 900 python -m pytest tests -m gpu -x -q -k "$Q" 2>&1 | tail -15 ) > $O/tests_warp.log 2>&1
( SM_EXACT=1 timeout 900 python -m pytest tests -m gpu -x -q -k "$Q" 2>&1 | tail -15 ) > $O/tests_warp_exact.log 2>&1
( timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_EXACT=1 timeout 300 python tests/gpu_probe.py cfg3:water 2
  SM_LIB_PATH=$L/libsm_mb2.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_mb4.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_w24.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_LIB_PATH=$L/libsm_w4.so timeout 300 python tests/gpu_probe.py cfg3:both 2
  SM_KERNEL=thread timeout 300 python tests/gpu_probe.py cfg3:both 2
  timeout 300 python tests/gpu_probe.py single
  SM_KERNEL=thread timeout 300 python tests/gpu_probe.py single
  timeout 300 python tests/gpu_probe.py big
  timeout 300 python tests/gpu_probe.py hydro:cfg3
  SM_HYDRO=thread timeout 300 python tests/gpu_probe.py hydro:cfg3 ) > $O/timing.log 2>&1
( SM_KERNEL=thread SM_HYDRO=thread timeout 900 python -m pytest tests -m gpu -x -q -k "$Q" 2>&1 | tail -5 ) > $O/tests_thread.log 2>&1
( timeout 600 python -m pytest tests -m gpu -q -k "lbm or facade_per_particle or budget" 2>&1 | tail -25 ) > $O/tests_new.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q -k "ipc" 2>&1 | tail -15 ) > $O/tests_ipc.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "config3" 2>&1 | tail -15 ) > $O/tests_cfg3.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q -k "config4" 2>&1 | tail -15 ) > $O/tests_cfg45.log 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 ) > $O/bench_warp.json 2> $O/bench_warp.err
( SM_KERNEL=thread timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra ) > $O/bench_thread.json 2> $O/bench_thread.err
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra ) > $O/ncu_launches.log 2>&1
( timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 2 -o $O/ksweep_full python bench.py --steps 1 --warmup 1 --no-cpu --no-extra ) > $O/ncu_full.log 2>&1
tail -3 $O/*.log
