set -x
mkdir -p gpurun_out
python tools/trace_e2e.py sweep > gpurun_out/r02i_gather_sweep.txt 2>&1
