set -x
timeout 600 python -m pytest tests/test_gpu_first.py tests/test_gpu_scenes.py tests/test_gpu_narrowphase.py tests/test_gpu_rope_edits.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02o_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02o_bench.json'))
print("settled", d["ms_per_step"], d["phases_ms"], "launches/tick", d["gpu_launches"]/d["steps"])
print("landing", d["landing"]["ms_per_step"], d["landing"]["phases_ms"])
print("c5", d["c5"]["ms_per_step"], d["c5"]["device_tick_ms_rank0"], d["c5"]["cast_ms_incl_pcie_rank0"])
print("e2e", d["e2e"]["ms_per_step"])
PY
