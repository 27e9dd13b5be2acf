echo K1_BLOCKS=5; OPUSB_LIB=esp32-opus-player_b200/lib/variants/k1b5.so python tools/stage_times.py celt
echo K1_BLOCKS=7; OPUSB_LIB=esp32-opus-player_b200/lib/variants/k1b7.so python tools/stage_times.py celt
echo K1_BLOCKS=8; OPUSB_LIB=esp32-opus-player_b200/lib/variants/k1b8.so python tools/stage_times.py celt
python bench.py --device-only --no-overlap --steps 1 --warmup 3 --workload celt > gpurun_out/ncu_celt_plain.json 2> gpurun_out/ncu_celt_plain.err && ncu --set full --clock-control none --import-source on -k regex:'^k_(entropy|sequence|leaf|bands|imdct|postfilter|synth|deemph|silk)' -s 51 -c 17 -f -o gpurun_out/ncu_celt python bench.py --device-only --no-overlap --steps 1 --warmup 3 --workload celt > gpurun_out/ncu_celt.log 2>&1
echo ncu rc=$?
