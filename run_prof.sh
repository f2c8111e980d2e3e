set -x
mkdir -p gpurun_out
SAVONT_B200_FILTER_TIGHT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-twin --no-cpu-baseline > gpurun_out/r02u_tight.json 2> gpurun_out/r02u.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-twin --no-cpu-baseline > gpurun_out/r02u_base.json 2>> gpurun_out/r02u.err
