set -x
mkdir -p gpurun_out
python tools/trace_e2e.py threads > gpurun_out/r02k_gather_threads.txt 2>&1
./tools/pipe_bench > gpurun_out/r02_pipe_bench.txt 2>&1 || (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_bench tools/pipe_bench.cu && /tmp/pipe_bench > gpurun_out/r02_pipe_bench.txt 2>&1)
python bench.py --path snpcluster --steps 3 --warmup 1 > gpurun_out/r02k_snpcluster_1gpu.json 2> gpurun_out/r02k_snpcluster.err
