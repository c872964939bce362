set -x
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
(python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -4)
$TR 29521 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_n8_auto.json 2> gpurun_out/r02_bench_n8_auto.err
$TR 29522 bench.py --gpus 8 --steps 8 --warmup 3 --exchange p2p > gpurun_out/r02_bench_n8_p2p.json 2> gpurun_out/r02_bench_n8_p2p.err
GL_TRACE=1 $TR 29523 bench.py --gpus 8 --steps 1 --warmup 1 --no-e2e > /dev/null 2> gpurun_out/r02_n8_trace.err
grep "coset plan" gpurun_out/r02_n8_trace.err | tail -8 > gpurun_out/r02_n8_coset_pull_trace.txt
$TR 29524 bench.py --gpus 8 --steps 4 --warmup 2 --log-n 22 --cols 256 --rate-bits 3 > gpurun_out/r02_bench_cfg5_r3_n8.json 2> gpurun_out/r02_bench_cfg5_r3_n8.err
$TR 29525 bench.py --gpus 8 --steps 4 --warmup 2 --log-n 22 --cols 256 --rate-bits 1 > gpurun_out/r02_bench_cfg5_r1_n8.json 2> gpurun_out/r02_bench_cfg5_r1_n8.err
python bench.py --gpus 8 --single-process --steps 6 --warmup 3 > gpurun_out/r02_bench_n8_sp.json 2> gpurun_out/r02_bench_n8_sp.err
for f in r02_bench_n8_auto r02_bench_n8_p2p r02_bench_cfg5_r3_n8 r02_bench_cfg5_r1_n8 r02_bench_n8_sp; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], (d.get("e2e") or {}).get("value"), d.get("stage_ms"), d["parity"].get("match"))
except Exception as e:
    print("$f", "ERR", e, open("gpurun_out/$f.err").read()[-600:])
PY
done
cat gpurun_out/r02_n8_coset_pull_trace.txt
