python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02f_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
for w in mix silk hybrid hybrid_swb celt10 celt5 celt2_5; do python bench.py --workload $w --steps 5 > gpurun_out/r02f_$w.json 2> gpurun_out/r02f_$w.err; echo "$w rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err; echo "ref rc=$?"
