set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r01c_ont16s_1gpu.json 2> gpurun_out/r01c_ont16s_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01c_reference_1gpu.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01t.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-twin > gpurun_out/b_ncu_t.log 2>&1
for k in sketch_fast_kernel filter_kernel align_nm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r01t_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-twin > gpurun_out/ncu_t_$k.log 2>&1
done
python bench.py --steps 3 --warmup 3 --workload hifi16s > gpurun_out/r01c_hifi16s_1gpu.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --workload operon --reads 400000 --cpu-reads 30000 > gpurun_out/r01c_operon_1gpu.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --workload pooled --reads 2000000 --cpu-reads 50000 > gpurun_out/r01c_pooled_1gpu.json 2>/dev/null
python bench.py --path kmercount --steps 3 --warmup 3 > gpurun_out/r01c_kmercount_1gpu.json 2>/dev/null
for k in kmer_scatter1_kernel kmer_count_part_kernel; do ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_r01t_$k python bench.py --path kmercount --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_t_$k.log 2>&1; done
