set -x
python -m pytest tests/test_gpu_rope_edits.py tests/test_gpu_islands.py tests/test_gpu_scenes.py tests/test_gpu_slab.py tests/test_gpu_longrun.py -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r02d_pytest.log
