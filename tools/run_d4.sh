timeout 300 python -m pytest tests/test_gpu_thin.py tests/test_gpu_fuse.py -x -q 2>&1 | tail -5
timeout 200 python tools/step_profile.py > gpurun_out/step_profile_i.txt 2>&1; grep -E "total|latent|thin_down" gpurun_out/step_profile_i.txt
