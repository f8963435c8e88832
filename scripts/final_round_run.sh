# One gpurun call: GPU tests, smoke, bench (both arms), config-3 sweep, config-4 timing, ncu launch list.
python -m pytest tests -m gpu -x -q > gpurun_out/r1v_tests.log 2>&1; tail -2 gpurun_out/r1v_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1v_smoke.log 2>&1; tail -2 gpurun_out/r1v_smoke.log
python bench.py > gpurun_out/r1v_bench1.log 2>&1; tail -1 gpurun_out/r1v_bench1.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1v_ref.log 2>&1; tail -1 gpurun_out/r1v_ref.log | cut -c1-300
python scripts/time_raw.py > gpurun_out/r1v_raw.log 2>&1; cat gpurun_out/r1v_raw.log
python scripts/time_mask_network.py > gpurun_out/r1v_mask.log 2>&1; tail -7 gpurun_out/r1v_mask.log
python scripts/sweep_conv.py > gpurun_out/r1v_sweep.md 2>&1; tail -21 gpurun_out/r1v_sweep.md
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1v_pre_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 800 --csv --log-file gpurun_out/launches_r1v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1v_ncu_bench.log 2>&1; tail -1 gpurun_out/r1v_ncu_bench.log | cut -c1-100
