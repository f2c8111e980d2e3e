set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r01d_tests.log
python bench.py --path pileup --steps 5 --warmup 3 > gpurun_out/r01d_pileup_1gpu.json 2> gpurun_out/r01d_pileup_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d_pileup.csv python bench.py --path pileup --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_d.log 2>&1
for k in align_tb_kernel pileup_walk_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r01d_$k python bench.py --path pileup --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d_$k.log 2>&1
done
