python -m pytest tests/test_gpu_celt.py tests/test_gpu_opus.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
python tools/stage_times.py celt
OPUSB_LIB=esp32-opus-player_b200/lib/variants/nostage.so python tools/stage_times.py celt
done
python tools/stage_times.py silk
OPUSB_LIB=esp32-opus-player_b200/lib/variants/nostage.so python tools/stage_times.py silk
python bench.py --device-only --no-overlap --steps 5 --payload 639 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('640B staged', d['ms_per_step'], d['roofline']['stage_ms']['entropy'])"
OPUSB_LIB=esp32-opus-player_b200/lib/variants/nostage.so python bench.py --device-only --no-overlap --steps 5 --payload 639 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('640B nostage', d['ms_per_step'], d['roofline']['stage_ms']['entropy'])"
