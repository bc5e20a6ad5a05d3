# Round-end validation on one B200 (run through gpurun): GPU tests, smoke(), bench, CUPTI step profile, ncu launch list.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
timeout 200 python tools/step_profile.py > gpurun_out/step_profile_final.txt 2>&1
if [ "$1" = "ncu" ]; then
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/plain_final.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_final.log 2>&1; echo ncu rc=$?
fi
