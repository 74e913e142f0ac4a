set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02m_bench8.json 2> gpurun_out/r02m_bench8.err; echo "bench rc=$?"; tail -5 gpurun_out/r02m_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench8.json'))
print("settled", d["ms_per_step"], d["value"], d["phases_ms"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("landing", d["landing"]["ms_per_step"], d["landing"]["value"])
print("c5", {k: d["c5"][k] for k in ("value","ms_per_step","worlds_this_rank","device_tick_ms_rank0","cast_ms_incl_pcie_rank0")})
print("slab", d.get("slab"))
PY
