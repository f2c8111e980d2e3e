set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02r_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02r_ont16s_1gpu.json 2> gpurun_out/r02r_ont16s_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02r_reference.json 2> gpurun_out/r02r_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1
