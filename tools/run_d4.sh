for d in 0 1; do echo DEEP=$d; B200GAN_WGRAD_DEEP=$d timeout 60 python tools/wgrad_bench.py; done
B200GAN_WGRAD_DEEP=1 timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k wgrad 2>&1 | tail -2
