#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GL_TRACE=1 timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/stream_trace_n$N.json 2> gpurun_out/stream_trace_n$N.err; echo rc=$?
grep "streamed coset" gpurun_out/stream_trace_n$N.err | tail -$N
