set -x
python tools/diag_settled.py c3 300 gpurun_out/r02a_trace_settled.bin > gpurun_out/r02a_settled.json 2> gpurun_out/r02a_settled.err
python tools/read_trace.py gpurun_out/r02a_trace_settled.bin > gpurun_out/r02a_trace_settled.txt 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02a_settled_tick -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02a_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
