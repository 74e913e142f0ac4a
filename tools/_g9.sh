set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02p_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02p_smoke.log 2>&1; tail -1 gpurun_out/r02p_smoke.log
timeout 900 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02p_bench.err
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_solve_persistent -o gpurun_out/r02p_solver_settled -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02p_ncu1.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02p_solver_settled.ncu-rep gpurun_out/r02p_solver_full_settled.txt "k_solve_persistent, C3 (1 M bodies) settled pile, tick 301, final round-2 build"
python tools/ncu_lines.py gpurun_out/r02p_solver_settled.ncu-rep 40 > gpurun_out/r02p_solver_lines_settled.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:k_solve_persistent -o gpurun_out/r02p_solver_landing -f python tools/prof_ticks.py c3 15 1 > gpurun_out/r02p_ncu2.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02p_solver_landing.ncu-rep gpurun_out/r02p_solver_full_landing.txt "k_solve_persistent, C3 (1 M bodies) landing phase, tick 16, final round-2 build"
rm -f gpurun_out/r02p_solver_landing.ncu-rep
timeout 600 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r02p_tick -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02p_ncu3.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02p_tick.ncu-rep gpurun_out/r02p_tick_full_settled.txt "every kernel of one settled C3 tick (tick 301), final round-2 build"
rm -f gpurun_out/r02p_tick.ncu-rep
python tools/diag_settled.py c3 300 gpurun_out/t.bin > gpurun_out/r02p_settled.json 2>&1; python tools/read_trace.py gpurun_out/t.bin > gpurun_out/r02p_trace_settled.txt 2>&1; rm -f gpurun_out/t.bin
du -sh gpurun_out
