set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02d_smi.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02d_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_ont16s_1gpu.json 2> gpurun_out/r02d_ont16s_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_ont16s_2gpu.json 2> gpurun_out/r02d_ont16s_2gpu.err
