set -x
python tools/diag_manifold_dist.py > gpurun_out/r02_manifold_dist.json 2> gpurun_out/r02_manifold_dist.err; tail -2 gpurun_out/r02_manifold_dist.err
python tools/bench_configs.py c1 c2 c4 c5 c5d --settle 150 --cpu > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err; tail -2 gpurun_out/r02_configs.err
# ncu: casts + the kernels of a C5 tick
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:k_cast_batch -o gpurun_out/r02_cast -f python tools/prof_ticks.py c5 150 1 > gpurun_out/r02_ncu_cast.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02_cast.ncu-rep gpurun_out/r02_cast_full_c5.txt "k_cast_batch, C5 (8192 worlds x 256 bodies, 524 288 rays), after tick 151"
rm -f gpurun_out/r02_cast.ncu-rep
# ncu: the streaming stages (integrate = k_stage<0>, derive = k_stage<5>) and the per-colour stages as separate launches, settled C3
SFG_PROF_FLAGS=2 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:k_stage -c 75 -o gpurun_out/r02_stages -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02_ncu_stages.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02_stages.ncu-rep gpurun_out/r02_stages_full_settled.txt "one substep of a settled C3 tick as separate launches (SFG_TUNE_SEPARATE_LAUNCHES): k_stage<0> integrate, <4> contact colours, <5> derive, <6> contact velocity colours"
rm -f gpurun_out/r02_stages.ncu-rep
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu-baseline --no-landing --no-c5 --steps 3 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
python tools/summarise_ncu.py launches gpurun_out/r02_launches.csv gpurun_out/r02_launches_bench.txt "python bench.py --no-cpu-baseline --no-landing --no-c5 --steps 3 --warmup 3 (settled C3: 300 settle + 3 warm-up + 3 timed + 3 e2e warm-up + 3 e2e + 2 counted ticks)"
rm -f gpurun_out/r02_launches.csv
du -sh gpurun_out
