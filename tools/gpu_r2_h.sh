#!/bin/bash
# round 2, GPU call H: whole GPU test-suite, smoke(), both bench arms on the tree with the new kernel selection
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
print({n: v["tflops"] for n, v in d["config"]["per_n"].items()})
PY
