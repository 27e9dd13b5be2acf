python -m pytest tests/test_gpu_opus.py -m gpu -x -q 2>&1 | tail -2
python tools/stage_times.py silk
python bench.py --device-only --no-overlap --steps 5 --workload silk | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('silk', round(d['value']), d['ms_per_step'])"
