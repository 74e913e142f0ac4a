for rep in 1 2; do for v in A B; do
SFG_LIB=$PWD/moleengine_b200/libsfg_$v.so python bench.py --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_ab.json')); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['phases_ms'].items()}, 'settled', round(d['settled']['ms_per_step'],3), round(d['settled']['phases_ms']['solver'],3))"
done; done
for v in A B; do SFG_LIB=$PWD/moleengine_b200/libsfg_$v.so python tools/bench_configs.py c2 c5 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l); print('$v', d['config'], round(d['ms_per_tick'],3), {k: round(x,3) for k,x in d['phases_ms'].items()})
    except Exception as e: print(l[:200])"; done
SFG_LIB=$PWD/moleengine_b200/libsfg_B.so python -m pytest tests -m gpu -q 2>&1 | tail -3
