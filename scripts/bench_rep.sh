for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-prefetch 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no-prefetch', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-prefetch --no-stage 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('plain', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
