set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02p_tests.log
python bench.py --steps 5 --warmup 3 --no-twin > gpurun_out/r02p_ont16s_1gpu.json 2> gpurun_out/r02p_ont16s_1gpu.err
python bench.py --workload pooled --reads 1250000 --steps 3 --warmup 3 --no-twin --no-cpu-baseline > gpurun_out/r02p_pooled_1gpu.json 2>/dev/null
python bench.py --workload operon --reads 400000 --steps 3 --warmup 3 --no-twin --no-cpu-baseline > gpurun_out/r02p_operon_1gpu.json 2>/dev/null
