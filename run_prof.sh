set -x
mkdir -p gpurun_out
nproc > gpurun_out/r02x_multi.txt
timeout 900 python tools/trace_e2e.py multi >> gpurun_out/r02x_multi.txt 2> gpurun_out/r02x_multi.err
