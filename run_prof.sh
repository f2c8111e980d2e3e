set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02a_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_ont16s_1gpu.json 2> gpurun_out/r02a_ont16s_1gpu.err
SAVONT_B200_DP_TAIL=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-twin > gpurun_out/r02a_ont16s_notail.json 2>/dev/null
python bench.py --workload hifi16s --steps 5 --warmup 3 --no-cpu-baseline --no-twin > gpurun_out/r02a_hifi16s_1gpu.json 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
