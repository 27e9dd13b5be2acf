python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
for w in mix silk hybrid hybrid_swb celt10 celt5 celt2_5; do python bench.py --workload $w --steps 5 > gpurun_out/r02g_$w.json 2> gpurun_out/r02g_$w.err; echo "$w rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err; echo "ref rc=$?"
