N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r01c_ont16s_${N}gpu.json 2> gpurun_out/r01c_ont16s_${N}gpu.err
tail -c 400 gpurun_out/r01c_ont16s_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r01c_ref_${N}gpu.json 2>/dev/null
tail -c 600 gpurun_out/r01c_ont16s_${N}gpu.json
