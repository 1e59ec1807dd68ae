#!/bin/bash
# Development helper for one gpurun call: GPU test suite, the headline bench (summary), optional extra configs with parity.
#   tools/gpu_check.sh [cad] [sphere] [block]
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee /tmp/pytest_tail.log
echo "== bench cessna 256/16"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/model', round(d['ms_per_step'],4), 'G tests/s', round(d['value'],1), 'e2e ms', round(d['e2e']['ms_per_model'],3), 'launches/step', d['gpu_launches']//d['steps']); print(d['phase_ms'])"
for m in "$@"; do
  case $m in cad) a="--mesh cad --l1 1024 --l2 2";; sphere) a="--mesh sphere --l1 512 --l2 8";; block) a="--mesh block --l1 64 --l2 4";; *) continue;; esac
  echo "== run_config $a"
  timeout 400 python tools/run_config.py $a --reps 6 --check > /tmp/rc.log 2>&1
  python - <<'P'
import json
L=open('/tmp/rc.log').read().strip().splitlines()
for l in L:
    if l.startswith('{'):
        d=json.loads(l); print(d['mesh'], 'min wall ms', min(d['wall_ms']), d['phase_ms'])
    elif l.startswith('parity'): print(l)
    elif 'Error' in l or 'error' in l: print(l)
P
done
echo "== pytest again: $(tail -1 /tmp/pytest_tail.log)"
