set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02n_pytest.log
cat gpurun_out/fullsize_settled_stats.json
