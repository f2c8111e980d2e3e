set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "stage3 or twin or multi or named" 2>&1 | tail -6 > gpurun_out/r02l_tests.log
./tools/pipe_bench > gpurun_out/r02_pipe_bench.txt 2>&1
python bench.py --path snpcluster --steps 3 --warmup 1 > gpurun_out/r02l_snpcluster_1gpu.json 2> gpurun_out/r02l_snpcluster.err
python tools/trace_e2e.py > gpurun_out/r02l_e2e_threads.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02l_ont16s_1gpu.json 2> gpurun_out/r02l_ont16s_1gpu.err
