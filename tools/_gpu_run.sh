python -m pytest tests -m gpu -x -q 2>&1 | tail -2
OPUSB_SORT_FRAMES=1 python -m pytest tests/test_gpu_celt.py tests/test_gpu_opus.py -m gpu -x -q 2>&1 | tail -1
python tools/stage_times.py celt
python tools/stage_times.py silk
python tools/stage_times.py hybrid
