"""Data-parallel layer on CPU: two `gloo` ranks (world_size 2) run the gradient-bucket layer of the product
(gan-enhanced-pneumonia-classifier_b200/dp.py, the same class the CUDA trainer drives over NCCL) around the numpy oracle's
per-rank gradients, and must land on the weights of the single-process DP emulation `dcgan_oracle.train_iteration_dp`
(mean of per-rank gradients, local BatchNorm statistics).  Also checks the bucket plan (order, coverage, early launch)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dcgan_oracle as orc
from gan_enhanced_pneumonia_classifier_b200.dp import GradBuckets
from parity_utils import synthetic_noise, synthetic_real

NZ, NC, FM, B, WORLD = 8, 1, 4, 2, 2


def _nets(seed=11):
    rng = np.random.RandomState(seed)
    sdG = orc.init_state(orc.generator_plan(NZ, NC, FM), True, rng)
    sdD = orc.init_state(orc.discriminator_plan(NC, FM), False, rng)
    G = orc.GeneratorOracle(NZ, NC, FM, {k: v.copy() for k, v in sdG.items()})
    D = orc.DiscriminatorOracle(NC, FM, {k: v.copy() for k, v in sdD.items()})
    return G, D, orc.AdamOracle(orc.param_keys(G.plan), 2e-4, 0.5), orc.AdamOracle(orc.param_keys(D.plan), 2e-4, 0.5)


def _shard(rank):
    return synthetic_real(100 + rank, B, NC), synthetic_noise(200 + rank, B, NZ)


def _arena(keys, sd):
    """flat fp32 arena in parameter order with 4-element aligned slices, like trainer._Arena"""
    slices, o = [], 0
    for k in keys:
        n = sd[k].size
        slices.append((o, o + n))
        o += (n + 3) // 4 * 4
    return torch.zeros(o, dtype=torch.float32), slices


def _allreduce_mean(keys, sd, grads, bucket_numel):
    arena, slices = _arena(keys, sd)
    for k, (lo, hi) in zip(keys, slices):
        arena[lo:hi] = torch.from_numpy(np.ascontiguousarray(grads[k]).reshape(-1))
    bk = GradBuckets(arena, slices, bucket_numel=bucket_numel)
    bk.begin()
    launched_before_finish = 0
    for j in reversed(range(len(keys))):                    # backward order: last layer's parameters first
        bk.ready(j)
        launched_before_finish = bk.collectives
    bk.finish()
    assert bk.collectives == len(bk.buckets)
    if len(bk.buckets) > 1:
        assert launched_before_finish == len(bk.buckets)    # every bucket went out as soon as it was complete, not at finish()
    scale = 1.0 / dist.get_world_size()
    return {k: (arena[lo:hi].numpy().reshape(sd[k].shape) * np.float32(scale)).astype(np.float32) for k, (lo, hi) in zip(keys, slices)}


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    try:
        G, D, optG, optD = _nets()
        real, noise = _shard(rank)
        kD, kG = orc.param_keys(D.plan), orc.param_keys(G.plan)
        # D update (train_gan.py:122-141) with the gradient exchange in the middle
        p_real, c_real = D.probs(real, train=True)
        gD, _ = D.backward_from_probs(c_real, orc.bce_bwd(p_real, orc.REAL_LABEL), need_input_grad=False)
        fake, c_g = G.forward(noise, train=True)
        p_fake, c_fake = D.probs(fake, train=True)
        g2, _ = D.backward_from_probs(c_fake, orc.bce_bwd(p_fake, orc.FAKE_LABEL), need_input_grad=False)
        gD = orc.accumulate(gD, g2)
        optD.step(D.sd, _allreduce_mean(kD, D.sd, gD, bucket_numel=64))           # many small buckets
        # G update (train_gan.py:144-150)
        p2, c2 = D.probs(fake, train=True)
        _, dfake = D.backward_from_probs(c2, orc.bce_bwd(p2, orc.REAL_LABEL), need_input_grad=True)
        gG, _ = G.backward(c_g, dfake, need_input_grad=False)
        optG.step(G.sd, _allreduce_mean(kG, G.sd, gG, bucket_numel=1 << 20))       # one bucket
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), **{f'G.{k}': v for k, v in G.sd.items()}, **{f'D.{k}': v for k, v in D.sd.items()})
    finally:
        dist.destroy_process_group()


def test_bucket_plan_covers_arena_in_backward_order():
    slices = [(0, 100), (100, 104), (104, 108), (108, 5108), (5108, 5112), (5112, 5116), (5116, 5216)]
    bk = GradBuckets(torch.zeros(5216), slices, bucket_numel=1000)
    assert bk.buckets[0][2][0] == len(slices) - 1                                   # first bucket starts at the last parameter
    covered = sorted(i for _, _, idx in bk.buckets for i in idx)
    assert covered == list(range(len(slices)))
    for lo, hi, idx in bk.buckets:
        assert lo == slices[min(idx)][0] and hi == slices[max(idx)][1]
    assert bk.world == 1
    bk.begin(); bk.ready(6); bk.finish()                                            # world 1: no collective is issued
    assert bk.collectives == 0


def test_two_gloo_ranks_match_dp_emulation(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.start_processes(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True, start_method='spawn')
    # single-process emulation of the same two ranks
    reps = [_nets() for _ in range(WORLD)]
    shards = [_shard(r) for r in range(WORLD)]
    orc.train_iteration_dp([x[0] for x in reps], [x[1] for x in reps], [x[2] for x in reps], [x[3] for x in reps],
                           [s[0] for s in shards], [s[1] for s in shards])
    for r in range(WORLD):
        got = np.load(os.path.join(str(tmp_path), f'rank{r}.npz'))
        for tag, net in (('G', reps[r][0]), ('D', reps[r][1])):
            for k, v in net.sd.items():
                np.testing.assert_allclose(got[f'{tag}.{k}'], v, rtol=1e-5, atol=1e-7, err_msg=f'rank {r} {tag}.{k}')
    # replicas stay in lock-step on the weights; BatchNorm buffers are rank-local (different shards -> different statistics)
    a, b = np.load(os.path.join(str(tmp_path), 'rank0.npz')), np.load(os.path.join(str(tmp_path), 'rank1.npz'))
    assert np.array_equal(a['D.main.2.weight'], b['D.main.2.weight'])
    assert not np.array_equal(a['D.main.3.running_mean'], b['D.main.3.running_mean'])
