# One gpurun call at the end of round 2: GPU tests, smoke, bench (both arms), per-launch ncu csv of one training step and one
# inference scene (scripts/ncu_step.py), the plain ncu launch list of the bench command itself.
python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; tail -2 gpurun_out/r2j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
python bench.py > gpurun_out/r2j_bench1.log 2>gpurun_out/r2j_bench1.err; tail -1 gpurun_out/r2j_bench1.log | cut -c1-400
python bench.py > gpurun_out/r2j_bench2.log 2>gpurun_out/r2j_bench2.err; tail -1 gpurun_out/r2j_bench2.log | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_ref.log 2>&1; tail -1 gpurun_out/r2j_ref.log | cut -c1-300
python bench.py --geometry-ahead off --no-cpu-baseline > gpurun_out/r2j_bench_noahead.log 2>/dev/null; tail -1 gpurun_out/r2j_bench_noahead.log | cut -c1-200
python scripts/ncu_step.py train > /dev/null 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r2j_train_step_launches.csv python scripts/ncu_step.py train > gpurun_out/r2j_ncu_train.log 2>&1; tail -1 gpurun_out/r2j_ncu_train.log
python scripts/ncu_step.py infer > /dev/null 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r2j_infer_scene_launches.csv python scripts/ncu_step.py infer > gpurun_out/r2j_ncu_infer.log 2>&1; tail -1 gpurun_out/r2j_ncu_infer.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 330 --csv --log-file gpurun_out/r2j_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu_bench.log 2>&1; tail -1 gpurun_out/r2j_ncu_bench.log | cut -c1-100
