set -x
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02f_pytest.log
# ncu: solver kernel at the settled window (tick 301) and at the landing window (tick 15), full sets with source
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_solve_persistent -o gpurun_out/r02f_solver_settled -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02f_ncu1.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02f_solver_settled.ncu-rep gpurun_out/r02f_solver_full_settled.txt "k_solve_persistent, C3 (1 M bodies) settled pile, tick 301"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:k_solve_persistent -o gpurun_out/r02f_solver_landing -f python tools/prof_ticks.py c3 15 1 > gpurun_out/r02f_ncu2.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02f_solver_landing.ncu-rep gpurun_out/r02f_solver_full_landing.txt "k_solve_persistent, C3 (1 M bodies) landing phase, tick 16"
rm -f gpurun_out/r02f_solver_landing.ncu-rep
# every kernel of one settled tick (summary only) + launch list
timeout 600 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r02f_tick -f python tools/prof_ticks.py c3 300 1 > gpurun_out/r02f_ncu3.log 2>&1
python tools/summarise_ncu.py full gpurun_out/r02f_tick.ncu-rep gpurun_out/r02f_tick_full_settled.txt "every kernel of one settled C3 tick (tick 301)"
rm -f gpurun_out/r02f_tick.ncu-rep
du -sh gpurun_out
