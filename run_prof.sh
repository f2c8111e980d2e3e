set -x
mkdir -p gpurun_out
python tools/trace_e2e.py > gpurun_out/r02g_e2e_threads.txt 2>&1
SAVONT_B200_TRACE=1 python tools/trace_e2e.py > gpurun_out/r02g_e2e_trace.txt 2>&1
nproc > gpurun_out/r02g_nproc.txt; free -g >> gpurun_out/r02g_nproc.txt; lscpu | head -25 >> gpurun_out/r02g_nproc.txt
