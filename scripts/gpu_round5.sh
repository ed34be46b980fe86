#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --workload convection_pt --gpus $N --rows 8192 --cols 65536 --iterations 50 --steps 2 --warmup 1 > gpurun_out/scale_convection_big_$N.json 2> gpurun_out/scale_convection_big_$N.err; tail -3 gpurun_out/scale_convection_big_$N.err | cut -c1-300; cat gpurun_out/scale_convection_big_$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload fdtd --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_fdtd_$N.json 2> gpurun_out/scale_fdtd_$N.err; tail -3 gpurun_out/scale_fdtd_$N.err | cut -c1-300; cat gpurun_out/scale_fdtd_$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --workload fdtd --gpus 4 --steps 3 --warmup 3 > gpurun_out/scale_fdtd_4.json 2> gpurun_out/scale_fdtd_4.err; cat gpurun_out/scale_fdtd_4.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --workload fdtd --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale_fdtd_2.json 2> gpurun_out/scale_fdtd_2.err; cat gpurun_out/scale_fdtd_2.json
