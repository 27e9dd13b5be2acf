python -m pytest tests/test_gpu_opus.py tests/test_gpu_celt.py tests/test_gpu_player.py -m gpu -x -q 2>&1 | tail -2
python tools/stage_times.py silk
python tools/stage_times.py hybrid
python tools/stage_times.py celt
