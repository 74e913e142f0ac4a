set -x
timeout 600 python -m pytest tests/test_gpu_first.py tests/test_gpu_scenes.py tests/test_gpu_slab.py -m gpu -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02q_pytest.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02q_bench_$tag.json 2> gpurun_out/r02q_bench_$tag.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02q_bench_$tag.json'))
print("$tag settled", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phases_ms"].items()}, "landing", round(d["landing"]["ms_per_step"],3), round(d["landing"]["phases_ms"]["solver"],3), "c5", round(d["c5"]["device_tick_ms_rank0"],2))
PY
}
run steal A=1
run unroll1 SFG_LIB=$PWD/moleengine_b200/libsfg_unroll1.so
run inline SFG_ISLANDS_INLINE=1
run steal2 A=1
