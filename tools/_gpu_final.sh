# final GPU check of the round (gpurun --timeout 560 -- bash tools/_gpu_final.sh): all GPU tests, the headline bench line,
# then the ncu capture of one step that bench.py takes its instruction / DRAM-byte counts from
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_v5_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r02_v5_gpu_tests.log)"
python bench.py > gpurun_out/r02_v5_bench_celt.json 2> gpurun_out/r02_v5_bench_celt.err; echo "bench rc=$? $(head -c 300 gpurun_out/r02_v5_bench_celt.json)"
python bench.py --device-only --no-overlap --steps 1 --warmup 3 --workload celt > gpurun_out/ncu_celt_plain.json 2> gpurun_out/ncu_celt_plain.err && timeout 300 ncu --set full --clock-control none --import-source on -k regex:'^k_(entropy|sequence|leaf|bands|imdct|postfilter|synth|deemph|silk)' -s 54 -c 18 -f -o gpurun_out/ncu_celt python bench.py --device-only --no-overlap --steps 1 --warmup 3 --workload celt > gpurun_out/ncu_celt.log 2>&1
echo "ncu celt rc=$? $(ls -la gpurun_out/ncu_celt.ncu-rep 2>/dev/null)"
