set -x
python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02c_pytest.log
python tools/diag_settled.py c3 300 gpurun_out/t.bin > gpurun_out/r02c_settled.json 2>&1; rm -f gpurun_out/t.bin; cat gpurun_out/r02c_settled.json
