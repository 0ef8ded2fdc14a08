set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r1_s3.json 2> gpurun_out/bench_r1_s3.err; tail -c 600 gpurun_out/bench_r1_s3.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_s3.json 2> gpurun_out/bench_ref_s3.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/plain_s3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1_s3.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_list_s3.log 2>&1
timeout 300 python tools/prof_k1.py > gpurun_out/plain_k1_s3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 6 --launch-count 3 -f -o gpurun_out/prof_r1_s3 python tools/prof_k1.py > gpurun_out/ncu_full_s3.log 2>&1
tail -3 gpurun_out/ncu_full_s3.log
