# The ncu evidence of a round, on one B200 (gpurun -- 'bash run_prof.sh'); then, back in the container:
#   python profiles/refresh.py r02t r02t      (launch list + per-kernel summaries -> profiles/)
# and set PROFILE_TAG in bench.py to the same tag: bench.py takes roofline.traffic from these summaries only while the capture's
# kernel duration is within 10 % of the run's CUDA-event time.  Numbers printed under ncu are never bench values.
set -x
TAG=${TAG:-r02t}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-twin --no-cigar"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/${TAG}_launch_run.log 2>&1
for k in sketch_fast_kernel filter_kernel align_nm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$k $B > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
S="python bench.py --path snpcluster --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_snpcluster.csv $S > gpurun_out/${TAG}_snp_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:snp_round_kernel -s 30 -c 1 -f -o gpurun_out/prof_${TAG}_snp_round_kernel $S > gpurun_out/${TAG}_ncu_snp_round.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:snp_resolve_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_snp_resolve_kernel $S > gpurun_out/${TAG}_ncu_snp_resolve.log 2>&1
