timeout 300 python -m pytest tests/test_gpu_fuse.py tests/test_gpu_thin.py -x -q 2>&1 | tail -3
timeout 200 python tools/step_profile.py > gpurun_out/step_profile_l.txt 2>&1; grep -E "total|thin_down" gpurun_out/step_profile_l.txt
