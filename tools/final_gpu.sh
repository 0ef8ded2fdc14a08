# Round-end evidence run (one GPU): tests, bench (both arms), launch list and ncu --set full captures, each ncu pass only
# after the same command exited 0 without ncu.  Scratch output in gpurun_out/; tools/refresh_profiles.py turns it into profiles/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 400 gpurun_out/bench_r2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/plain_r2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_list_r2.log 2>&1
timeout 300 python tools/prof_k1.py > gpurun_out/plain_k1_r2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 6 --launch-count 3 -f -o gpurun_out/prof_r2 python tools/prof_k1.py > gpurun_out/ncu_full_r2.log 2>&1
tail -2 gpurun_out/ncu_full_r2.log
timeout 300 python tools/prof_k1_small.py > gpurun_out/plain_small_r2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stft --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r2_small python tools/prof_k1_small.py > gpurun_out/ncu_small_r2.log 2>&1
tail -2 gpurun_out/ncu_small_r2.log
timeout 300 python tools/gpu_time_c2.py > gpurun_out/configs_timing_r2.log 2>&1; tail -12 gpurun_out/configs_timing_r2.log
