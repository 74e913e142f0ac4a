set -x
python -m pytest tests/test_gpu_first.py tests/test_gpu_scenes.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02k_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench.json'))
print("settled", d["ms_per_step"], d["phases_ms"], "frac", d["roofline"]["frac"], d["roofline"].get("dram_frac"))
print("landing", d["landing"]["ms_per_step"], d["landing"]["phases_ms"])
print("c5", d["c5"]["ms_per_step"], d["c5"]["device_tick_ms_rank0"], d["c5"]["cast_ms_incl_pcie_rank0"])
print("e2e", d["e2e"]["ms_per_step"])
PY
