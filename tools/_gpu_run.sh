python -m pytest tests/test_gpu_celt.py tests/test_gpu_opus.py -m gpu -x -q 2>&1 | tail -3
python tools/stage_times.py celt
for v in 3 4 5; do echo "KIL_BLOCKS=$v"; OPUSB_LIB=esp32-opus-player_b200/lib/variants/kil$v.so python tools/stage_times.py celt; done
