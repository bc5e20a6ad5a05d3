timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fuse.py tests/test_gpu_step.py -x -q 2>&1 | tail -4
timeout 60 python tools/one_kernel.py d1_up 512 stats time
timeout 200 python tools/step_profile.py > gpurun_out/step_profile_h.txt 2>&1; head -40 gpurun_out/step_profile_h.txt
