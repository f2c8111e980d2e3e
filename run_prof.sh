set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02j_smi.txt; nproc >> gpurun_out/r02j_smi.txt; nvidia-smi topo -m >> gpurun_out/r02j_smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02j_ont16s_8gpu.json 2> gpurun_out/r02j_ont16s_8gpu.err
tail -30 gpurun_out/r02j_ont16s_8gpu.err
