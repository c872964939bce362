#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
GL_CHECK_MODES=stream timeout 300 $TR --master-port 29511 tests/check_sharded.py > gpurun_out/stream_check_n8.txt 2>&1; echo "check rc=$?"; grep -c ": ok" gpurun_out/stream_check_n8.txt
timeout 300 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8_stream.json 2> gpurun_out/r02_bench_n8_stream.err; echo "bench rc=$?"
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 2 --log-n 22 --cols 256 --rate-bits 3 > gpurun_out/r02_bench_cfg5_r3_n8_stream.json 2> gpurun_out/r02_bench_cfg5_r3_n8_stream.err; echo "cfg5 rc=$?"
GL_TRACE=1 timeout 200 $TR --master-port 29514 bench.py --gpus 8 --steps 2 --warmup 2 > /dev/null 2> gpurun_out/stream_trace_n8.err; grep "streamed coset" gpurun_out/stream_trace_n8.err | tail -8 | sort > gpurun_out/r02_n8_stream_trace.txt
for f in r02_bench_n8_stream r02_bench_cfg5_r3_n8_stream; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step"), d["e2e"]["api"][-12:], d["parity"])
except Exception as e:
    print("$f parse failed", e, open("gpurun_out/$f.err").read()[-800:])
PY
done
cut -c1-150 gpurun_out/r02_n8_stream_trace.txt
