set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02_exp14
mkdir -p $O
(
timeout 300 python tests/gpu_probe.py cfg3:water 2
for cf in 4 5; do
timeout 600 python bench.py --config $cf --steps 2 --warmup 1 --no-extra 2>$O/b$cf.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config $cf', d['value'], d['ms_per_step'], d.get('parity'))"
done
) > $O/timing.log 2>&1
grep -v "^+" $O/timing.log | grep "cfg3\|^config"
