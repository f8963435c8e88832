for i in 1 2 3; do for e in 0 1; do SCN_TILE_BOOK_EAGER=$e python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('eager=$e', d['ms_per_step'], d['inference']['scenes_per_sec'], d['inference']['ms_per_scene_one_thread'])"; done; done
