set -x
python -m pytest tests/test_gpu_slab.py -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02h_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-c5 --no-landing > gpurun_out/r02h_bench2.json 2> gpurun_out/r02h_bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/r02h_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench2.json'))
print("settled", d["ms_per_step"], d["value"], d["phases_ms"])
print("slab", d.get("slab"))
PY
