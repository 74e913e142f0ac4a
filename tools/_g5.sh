set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench2.json 2> gpurun_out/r02g_bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/r02g_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench2.json'))
print("settled", d["ms_per_step"], d["value"], d["phases_ms"])
print("e2e", d["e2e"]["ms_per_step"])
print("c5", {k: d["c5"][k] for k in ("value","ms_per_step","worlds_this_rank","device_tick_ms_rank0","cast_ms_incl_pcie_rank0")})
print("slab", d.get("slab"))
PY
