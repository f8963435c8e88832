python -m pytest tests -m gpu -x -q > gpurun_out/r1n_tests.log 2>&1; tail -2 gpurun_out/r1n_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1n_smoke.log 2>&1; tail -2 gpurun_out/r1n_smoke.log
python bench.py > gpurun_out/r1n_bench1.log 2>&1; tail -1 gpurun_out/r1n_bench1.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1n_ref.log 2>&1; tail -1 gpurun_out/r1n_ref.log | cut -c1-300
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1n_pre_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 800 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1n_ncu_bench.log 2>&1
python scripts/time_raw.py 0 > gpurun_out/r1n_raw0.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 10 -c 1 -f -o gpurun_out/conv_l0_skip python scripts/time_raw.py 0 > gpurun_out/r1n_ncu_full.log 2>&1; tail -2 gpurun_out/r1n_ncu_full.log
