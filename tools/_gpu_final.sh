# final GPU check of the round (gpurun --timeout 230 -- bash tools/_gpu_final.sh): all GPU tests, the headline bench line,
# then an ncu pass over one step with exactly the metrics tools/ncu_capture.py reads (a few passes instead of --set full)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_v6_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r02_v6_gpu_tests.log)"
python bench.py > gpurun_out/r02_v6_bench_celt.json 2> gpurun_out/r02_v6_bench_celt.err; echo "bench rc=$? $(head -c 300 gpurun_out/r02_v6_bench_celt.json)"
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio
timeout 80 ncu --metrics $M --clock-control none -k regex:'^k_(entropy|sequence|leaf|bands|imdct|postfilter|synth|deemph|silk)' -s 54 -c 18 -f -o gpurun_out/ncu_celt python bench.py --device-only --no-overlap --steps 1 --warmup 3 --workload celt > gpurun_out/ncu_celt.log 2>&1
echo "ncu celt rc=$? $(ls -la gpurun_out/ncu_celt.ncu-rep 2>/dev/null)"
