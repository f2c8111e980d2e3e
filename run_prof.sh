set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-twin --e2e-steps 1 > gpurun_out/b_ncu_c.log 2>&1
for k in sketch_fast_kernel filter_kernel align_nm_kernel; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r02c_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-twin --e2e-steps 1 > gpurun_out/ncu_c_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
