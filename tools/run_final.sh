#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.txt; cat gpurun_out/r02_pytest_gpu_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
timeout 300 python bench.py --workload wrapper > gpurun_out/r02_bench_wrapper_final.json 2> gpurun_out/r02_bench_wrapper_final.err; echo "wrapper rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n1_final", "r02_bench_wrapper_final"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], (d.get("e2e") or {}).get("value"), d.get("parity"), (d.get("roofline") or {}).get("frac"), list((d.get("gate_eval") or {}).keys()))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-500:])
PY
