set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; echo bench rc=$?; cat gpurun_out/bench_r01f.json
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r01f.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r01f.log 2>&1; echo ncu rc=$?
timeout 200 python tools/step_profile.py > gpurun_out/step_profile_k.txt 2>&1
