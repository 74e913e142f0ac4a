#!/bin/bash
# One GPU-box visit: parity tests, smoke, the headline bench (all arms), the other configs. Usage: gpurun -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-chk}
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json'))
print('$tag settled', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'dram_frac', d['roofline'].get('dram_frac'))
print('$tag landing', round(d['landing']['ms_per_step'],3), 'c5', round(d['c5']['ms_per_step'],3), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))"
python tools/bench_configs.py c1 c2 c4 c5 --settle 150 > gpurun_out/configs_$tag.jsonl 2>&1
python -c "
import json
for l in open('gpurun_out/configs_$tag.jsonl'):
    try: d=json.loads(l); print(d['config'], round(d['ms_per_tick'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, d['contact_colours'])
    except Exception as e: print(l[:200])"
