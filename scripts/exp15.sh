set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp15
mkdir -p $O
(
timeout 300 python tests/gpu_probe.py cfg3:both 2
SM_EXACT=0 timeout 300 python tests/gpu_probe.py cfg3:both 2
for e in 1 0; do for cf in 4 5; do
SM_EXACT=$e timeout 600 python bench.py --config $cf --steps 2 --warmup 1 --no-extra 2>$O/b$cf.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config $cf exact=$e', d['value'], d['ms_per_step'], d.get('parity'))"
done; done
) > $O/timing.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config3" > $O/tests_cfg3.log 2>&1
SM_EXACT=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config3" > $O/tests_cfg3_e0.log 2>&1
grep -v "^+" $O/timing.log | grep "cfg3\|^config"; tail -2 $O/tests_cfg3.log $O/tests_cfg3_e0.log
