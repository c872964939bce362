#!/bin/bash
# usage: tools/run_stream.sh N   — parity check of all exchange plans, then bench at N GPUs with the host plan auto (= streamed) and p2p
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GL_CHECK_MODES=${MODES:-coset,stream,p2p,nccl,auto} timeout 600 $TR --master-port 29511 tests/check_sharded.py > gpurun_out/stream_check_n$N.txt 2>&1; echo "check rc=$?" >> gpurun_out/stream_check_n$N.txt
grep -c ": ok" gpurun_out/stream_check_n$N.txt; grep "MISMATCH\|rc=\|Error\|error" gpurun_out/stream_check_n$N.txt | head -20
for ex in auto p2p; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --exchange $ex > gpurun_out/stream_bench_n${N}_$ex.json 2> gpurun_out/stream_bench_n${N}_$ex.err; echo "bench $ex rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/stream_bench_n${N}_$ex.json").read().strip().splitlines()[-1])
    print("$ex", d["value"], d["e2e"]["value"], d["e2e"].get("ms_per_step"), d["e2e"]["api"][-12:], d["parity"], d["e2e"].get("stage_ms_last_call_rank0"))
except Exception as e:
    print("parse failed", e)
PY
done
