set -x
timeout 600 python -m pytest tests/test_gpu_rope_edits.py tests/test_gpu_tour.py tests/test_gpu_scenes.py tests/test_gpu_islands.py tests/test_host_cpp.py -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02s_pytest.log
