#!/usr/bin/env python
"""Development aid: does the captured iteration (bucket all-reduces inside the CUDA graph) run on 2 ranks?  Prints progress markers;
a watchdog dumps the Python stacks and exits if a rank makes no progress."""
import faulthandler, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, port):
    faulthandler.dump_traceback_later(40, exit=True)
    import torch.distributed as dist
    import gan_enhanced_pneumonia_classifier_b200 as pkg
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    torch.manual_seed(0)
    G, D = pkg.Generator(16, 1, 8).cuda(), pkg.Discriminator(1, 8).cuda()
    tr = DCGANTrainer(G, D, dtype=torch.float32, use_graph=True)
    real = torch.rand(4, 1, 224, 224, device='cuda') * 2 - 1
    t0 = time.time()
    for it in range(5):
        out = tr.step(real, torch.randn(4, 16, 1, 1, device='cuda'))
        torch.cuda.synchronize()
        print(f'[rank {rank}] step {it} done at {time.time() - t0:.2f}s graphs={len(tr._graphs)} hist={out.cpu().tolist()[:2]}', flush=True)
    w = G.main[0].weight.detach().clone()
    ref = w.clone(); dist.broadcast(ref, 0)
    print(f'[rank {rank}] replicas identical: {torch.equal(ref, w)} collectives={tr.comm.collectives}', flush=True)
    tr.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    import torch.multiprocessing as mp
    mp.start_processes(worker, args=(2, int(sys.argv[1]) if len(sys.argv) > 1 else 29533), nprocs=2, join=True, start_method='spawn')
