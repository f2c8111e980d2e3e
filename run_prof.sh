set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r01e_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01e_smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01e_reference_1gpu.json 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/r01e_ont16s_1gpu.json 2> gpurun_out/r01e_ont16s_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-twin > gpurun_out/b_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:snpmer_compare_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01e_snpmer_compare_kernel python bench.py --path reassign --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e_cmp.log 2>&1
