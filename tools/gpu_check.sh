#!/bin/bash
# One GPU-box visit: parity tests, the headline bench, the other configs, the stage trace and (if a -DSFG_PROBE build sits next to
# the library) the contact-path latency probe. Usage: gpurun -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-chk}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('$tag', d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
python tools/bench_configs.py c1 c2 c4 c5 > gpurun_out/configs_$tag.jsonl 2>&1
python -c "
import json
for l in open('gpurun_out/configs_$tag.jsonl'):
    try: d=json.loads(l); print(d['config'], round(d['ms_per_tick'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, d['contact_colours'])
    except Exception as e: print(l[:200])"
if [ -f moleengine_b200/libsfg_probe.so ]; then
SFG_LIB=$PWD/moleengine_b200/libsfg_probe.so python tools/diag_probe.py 300 2>&1 | tail -22
python tools/read_trace.py gpurun_out/trace_a.bin | tail -40
python tools/read_trace.py gpurun_out/trace_b.bin | tail -6
fi
