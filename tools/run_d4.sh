B200GAN_UP4_YREG=0 timeout 60 python tools/one_kernel.py d1_up 512 mask time
B200GAN_UP4_YREG=1 timeout 60 python tools/one_kernel.py d1_up 512 mask time
