#!/bin/bash
# full validation as the driver runs it: GPU test-suite, smoke(), bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_validate_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2_validate_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_validate_bench.json 2> gpurun_out/r2_validate_bench.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_validate_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['executed_frac'], d['clocks'])
print('second', d['second_headline']['value'], d['second_headline']['ms_per_step'], d['second_headline']['roofline']['frac'])
for k,v in d['extra'].items(): print(k, {a:b for a,b in v.items() if a in ('ms','fp64_frac','permanents_per_s','seconds','e2e_ms','cdf_and_search_ms')})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
