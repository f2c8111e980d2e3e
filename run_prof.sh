set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "stage3 or snpmer_cluster" 2>&1 | tail -8 > gpurun_out/r02q6_tests.log
timeout 600 python bench.py --path snpcluster --steps 5 --warmup 3 > gpurun_out/r02q6_snpcluster_1gpu.json 2> gpurun_out/r02q6_snpcluster.err
