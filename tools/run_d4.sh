timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fuse.py tests/test_gpu_step.py -x -q 2>&1 | tail -2
timeout 200 python tools/step_profile.py > gpurun_out/step_profile_p.txt 2>&1; grep -E "total|bn_act" gpurun_out/step_profile_p.txt
