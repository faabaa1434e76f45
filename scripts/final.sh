set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_final
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/gpu.txt
( time timeout 230 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 1500 $O/bench.json
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 230 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/tests_parity.log 2>&1; tail -3 $O/tests_parity.log
