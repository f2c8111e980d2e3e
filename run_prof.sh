set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02e_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_ont16s_1gpu.json 2> gpurun_out/r02e_ont16s_1gpu.err
