set -x
python tools/diag_settled.py c3 300 gpurun_out/r02a_trace_settled.bin > gpurun_out/r02a_settled.json 2> gpurun_out/r02a_settled.err
python tools/read_trace.py gpurun_out/r02a_trace_settled.bin > gpurun_out/r02a_trace_settled.txt 2>&1
rm -f gpurun_out/r02a_trace_settled.bin
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_solve_persistent -o gpurun_out/r02a_settled_solver -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02a_ncu.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02a_settled_solver.ncu-rep gpurun_out/r02a_solver_full_settled.txt "k_solve_persistent, C3 settled (tick 301)"
timeout 600 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r02a_settled_tick -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02a_ncu2.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02a_settled_tick.ncu-rep gpurun_out/r02a_tick_full_settled.txt "all kernels of one settled C3 tick (tick 301)"
rm -f gpurun_out/r02a_settled_tick.ncu-rep
du -sh gpurun_out
