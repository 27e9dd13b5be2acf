# final GPU check of the round (gpurun --timeout 640 -- bash tools/_gpu_final.sh): all GPU tests, the headline bench line,
# compute-sanitizer memcheck + racecheck over the small parity tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_v4_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r02_v4_gpu_tests.log)"
python bench.py > gpurun_out/r02_v4_bench_celt.json 2> gpurun_out/r02_v4_bench_celt.err; echo "bench rc=$? $(head -c 300 gpurun_out/r02_v4_bench_celt.json)"
bash tools/sanitize_gpu.sh memcheck racecheck
