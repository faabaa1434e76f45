set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp12
mkdir -p $O
(
SM_EXACT=3 timeout 300 python tests/gpu_probe.py cfg3:both 2
SM_EXACT=0 timeout 300 python tests/gpu_probe.py cfg3:water 2
for e in 0 1; do for cf in 4 5; do
SM_EXACT=$e timeout 600 python bench.py --config $cf --steps 2 --warmup 1 --no-extra 2>$O/b$cf_$e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config $cf exact=$e', d['value'], d['ms_per_step'], d.get('parity'))"
done; done
) > $O/timing.log 2>&1
SM_EXACT=3 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/tests_exact3.log 2>&1
grep -v "^+" $O/timing.log | grep "cfg3\|config"; tail -3 $O/tests_exact3.log
