set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "stage3" 2>&1 | tail -6 > gpurun_out/r02m_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1
python bench.py --path snpcluster --steps 3 --warmup 1 > gpurun_out/r02m_snpcluster_1gpu.json 2> gpurun_out/r02m_snpcluster.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02m_ont16s_1gpu.json 2> gpurun_out/r02m_ont16s_1gpu.err
