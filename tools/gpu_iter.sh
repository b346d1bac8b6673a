#!/bin/bash
# quick iteration: parity tests, timeline trace, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 300 python tools/trace_ws.py 16384 > gpurun_out/trace16k.log 2>&1
tail -26 gpurun_out/trace16k.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], {k:(v['ms'],v['tflops']) for k,v in d['config']['per_n'].items()}, d['clocks'])
PY
