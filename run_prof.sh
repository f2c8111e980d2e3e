set -x
mkdir -p gpurun_out
SAVONT_B200_SNPCLUSTER_PROF=1 python bench.py --path snpcluster --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02o_snpcluster_1gpu.json 2> gpurun_out/r02o_snpcluster.err
