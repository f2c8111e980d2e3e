set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r01d_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01d_smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_reference_1gpu.json 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/r01d_ont16s_1gpu.json 2> gpurun_out/r01d_ont16s_1gpu.err
python bench.py --path pileup --steps 5 --warmup 3 > gpurun_out/r01d_pileup_1gpu.json 2> gpurun_out/r01d_pileup_1gpu.err
python bench.py --path cluster --steps 3 --warmup 3 > gpurun_out/r01d_cluster_1gpu.json 2> gpurun_out/r01d_cluster_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01dc.csv python bench.py --path cluster --reads 200000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_dc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rep_query_kernel -s 150 -c 1 -f -o gpurun_out/prof_r01dc_rep_query_kernel python bench.py --path cluster --reads 200000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dc_rep_query.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_walk_kernel -s 9 -c 1 -f -o gpurun_out/prof_r01dp_pileup_walk_kernel python bench.py --path pileup --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dp_walk.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lsh or kmer_clustering_and or pileups_many" 2>&1 | tail -6 > gpurun_out/r01d_sanitizer.log
