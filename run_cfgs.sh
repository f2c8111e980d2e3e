set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r01n_ont16s.json 2> gpurun_out/r01n_ont16s.err
python bench.py --steps 3 --warmup 3 --workload hifi16s > gpurun_out/r01n_hifi16s.json 2> gpurun_out/r01n_hifi16s.err
python bench.py --steps 3 --warmup 3 --workload operon --reads 400000 --cpu-reads 30000 > gpurun_out/r01n_operon.json 2> gpurun_out/r01n_operon.err
python bench.py --steps 3 --warmup 3 --workload pooled --reads 2000000 --cpu-reads 50000 > gpurun_out/r01n_pooled.json 2> gpurun_out/r01n_pooled.err
tail -c 400 gpurun_out/r01n_*.err
