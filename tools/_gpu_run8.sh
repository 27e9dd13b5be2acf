TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1
nproc > gpurun_out/r02_host_8gpu.txt; free -g | head -2 >> gpurun_out/r02_host_8gpu.txt; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" >> gpurun_out/r02_host_8gpu.txt
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port 2951$n tools/pcie_probe.py > gpurun_out/r02_probe_${n}gpu.json 2> gpurun_out/r02_probe_${n}gpu.err; echo "probe $n rc=$?"
done
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "celt8 rc=$?"
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --workload mix > gpurun_out/r02_bench_mix_8gpu.json 2> gpurun_out/r02_bench_mix_8gpu.err; echo "mix8 rc=$?"
$TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --scaling strong > gpurun_out/r02_bench_strong_8gpu.json 2> gpurun_out/r02_bench_strong_8gpu.err; echo "strong8 rc=$?"
$TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo "celt4 rc=$?"
$TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "celt2 rc=$?"
$TR --nproc-per-node 4 --master-port 29526 bench.py --gpus 4 --scaling strong --no-e2e > gpurun_out/r02_bench_strong_4gpu.json 2> gpurun_out/r02_bench_strong_4gpu.err; echo "strong4 rc=$?"
$TR --nproc-per-node 2 --master-port 29527 bench.py --gpus 2 --scaling strong --no-e2e > gpurun_out/r02_bench_strong_2gpu.json 2> gpurun_out/r02_bench_strong_2gpu.err; echo "strong2 rc=$?"
python bench.py > gpurun_out/r02_bench_1of8.json 2> gpurun_out/r02_bench_1of8.err; echo "celt1 rc=$?"
