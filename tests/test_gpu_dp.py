"""Two-GPU data-parallel parity: DCGANTrainer on two NCCL ranks (one process per GPU, bucketed gradient all-reduce on a
communication stream, graph replay) against the CPU oracle's DP emulation.  Needs two GPUs: skipped on a one-GPU box
(run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
from parity_utils import close, synthetic_noise, synthetic_real, weights_close

pytestmark = pytest.mark.gpu
NZ, NC, FM, B, WORLD, STEPS = 16, 1, 8, 4, 2, 3


def _frac(key):
    """Share of entries that must sit within (rtol, atol) after STEPS Adam steps.  BatchNorm biases start at 0 and are a sum of
    three ~lr-sized normalised Adam updates; the entries whose rank-averaged gradient nearly cancels are the least
    well-conditioned numbers of the whole step (measured: 92 % of G.main.1.bias within 1e-5, worst entry 2.8e-5 = 0.14 lr,
    with first-iteration gradients, history and every other tensor inside the single-GPU tolerances; all inside the 2*lr
    per-step envelope that weights_close always enforces)."""
    return 0.5 if key.endswith('.bias') else 0.97         # biases: every entry within 0.25 lr (checked below) is the criterion


def _state(seed=21):
    rng = np.random.RandomState(seed)
    return (orc.init_state(orc.generator_plan(NZ, NC, FM), True, rng), orc.init_state(orc.discriminator_plan(NC, FM), False, rng))


def _shard(rank, it):
    return synthetic_real(300 + 10 * it + rank, B, NC), synthetic_noise(400 + 10 * it + rank, B, NZ)


def _worker(rank, port, out_dir, use_graph, sync_bn=False):
    import torch.distributed as dist
    import gan_enhanced_pneumonia_classifier_b200 as pkg
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=WORLD, device_id=torch.device('cuda', rank))
    try:
        sdG, sdD = _state()
        G, D = pkg.Generator(NZ, NC, FM), pkg.Discriminator(NC, FM)
        G.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdG.items()})
        D.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdD.items()})
        tr = DCGANTrainer(G.cuda(), D.cuda(), dtype=torch.float32, use_graph=use_graph, sync_bn=sync_bn)
        hist = []
        for it in range(STEPS):
            real, noise = _shard(rank, it)
            hist.append(tr.step(torch.from_numpy(real).cuda(), torch.from_numpy(noise).cuda()).cpu().numpy())
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), hist=np.stack(hist), collectives=tr.bucketsD.collectives + tr.bucketsG.collectives,
                 **{f'G.{k}': v.cpu().numpy() for k, v in G.state_dict().items()}, **{f'D.{k}': v.cpu().numpy() for k, v in D.state_dict().items()})
        tr.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('use_graph', [False, True])
def test_two_nccl_ranks_match_dp_emulation(tmp_path, use_graph):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.start_processes(_worker, args=(port, str(tmp_path), use_graph), nprocs=WORLD, join=True, start_method='spawn')
    sdG, sdD = _state()
    Gs = [orc.GeneratorOracle(NZ, NC, FM, {k: v.copy() for k, v in sdG.items()}) for _ in range(WORLD)]
    Ds = [orc.DiscriminatorOracle(NC, FM, {k: v.copy() for k, v in sdD.items()}) for _ in range(WORLD)]
    oG = [orc.AdamOracle(orc.param_keys(g.plan), 2e-4, 0.5) for g in Gs]
    oD = [orc.AdamOracle(orc.param_keys(d.plan), 2e-4, 0.5) for d in Ds]
    ref_hist = [[] for _ in range(WORLD)]
    for it in range(STEPS):
        shards = [_shard(r, it) for r in range(WORLD)]
        out, _, _ = orc.train_iteration_dp(Gs, Ds, oG, oD, [s[0] for s in shards], [s[1] for s in shards])
        for r in range(WORLD):
            ref_hist[r].append([out[r][k] for k in ('errD', 'errG', 'D_x', 'D_G_z1', 'D_G_z2')])
    for r in range(WORLD):
        got = np.load(os.path.join(str(tmp_path), f'rank{r}.npz'))
        assert use_graph or int(got['collectives']) > 0        # eager: bucketed all-reduces on the communication stream
        close(got['hist'][0], np.array(ref_hist[r][0]), rtol=1e-4, atol=1e-6, what=f'rank {r} first-iteration history')
        close(got['hist'], np.array(ref_hist[r]), rtol=5e-3, atol=1e-4, what=f'rank {r} history')
        for tag, net in (('G', Gs[r]), ('D', Ds[r])):
            for k, v in net.sd.items():
                if k.endswith('num_batches_tracked'):
                    assert int(got[f'{tag}.{k}']) == int(v)
                elif 'running' in k:
                    close(got[f'{tag}.{k}'], v, rtol=2e-3, atol=1e-4, what=f'rank {r} {tag}.{k}')
                else:
                    weights_close(got[f'{tag}.{k}'], v, what=f'rank {r} {tag}.{k}', steps=STEPS, rtol=1e-3, atol=1e-5, frac=_frac(k))
                    assert np.abs(got[f'{tag}.{k}'] - v).max() < (5e-5 if k.endswith('.bias') else 1.3e-3), k
    a, b = np.load(os.path.join(str(tmp_path), 'rank0.npz')), np.load(os.path.join(str(tmp_path), 'rank1.npz'))
    assert np.array_equal(a['G.main.0.weight'], b['G.main.0.weight']), 'replicas must stay bit-identical on the weights'


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('use_graph', [False, True])
def test_two_nccl_ranks_with_synchronised_batchnorm_equal_one_process_on_the_concatenated_batch(tmp_path, use_graph):
    """--sync-bn (optional): BatchNorm statistics and BatchNorm-backward reductions summed over the ranks on the library's
    communicator.  Two ranks of batch B then ARE the single-process iteration on the 2B batch -- the plain oracle, no DP emulation: the mean of the
    ranks' loss rows is the oracle's row, the weights and the running statistics (global mean / unbiased global variance) are the oracle's."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.start_processes(_worker, args=(port, str(tmp_path), use_graph, True), nprocs=WORLD, join=True, start_method='spawn')
    sdG, sdD = _state()
    G = orc.GeneratorOracle(NZ, NC, FM, {k: v.copy() for k, v in sdG.items()})
    D = orc.DiscriminatorOracle(NC, FM, {k: v.copy() for k, v in sdD.items()})
    oG, oD = orc.AdamOracle(orc.param_keys(G.plan), 2e-4, 0.5), orc.AdamOracle(orc.param_keys(D.plan), 2e-4, 0.5)
    ref = []
    for it in range(STEPS):
        shards = [_shard(r, it) for r in range(WORLD)]
        out = orc.train_iteration(G, D, oG, oD, np.concatenate([s[0] for s in shards]), np.concatenate([s[1] for s in shards]))
        ref.append([out[k] for k in ('errD', 'errG', 'D_x', 'D_G_z1', 'D_G_z2')])
    got = [np.load(os.path.join(str(tmp_path), f'rank{r}.npz')) for r in range(WORLD)]
    mean_hist = np.mean([g['hist'] for g in got], axis=0)
    close(mean_hist[0], np.array(ref[0]), rtol=1e-4, atol=1e-6, what='first-iteration history (mean over ranks)')
    close(mean_hist, np.array(ref), rtol=5e-3, atol=1e-4, what='history (mean over ranks)')
    for tag, net in (('G', G), ('D', D)):
        for k, v in net.sd.items():
            if k.endswith('num_batches_tracked'):
                assert int(got[0][f'{tag}.{k}']) == int(v)
            elif 'running' in k:
                close(got[0][f'{tag}.{k}'], v, rtol=2e-3, atol=1e-4, what=f'{tag}.{k}')
            else:
                weights_close(got[0][f'{tag}.{k}'], v, what=f'{tag}.{k}', steps=STEPS, rtol=1e-3, atol=1e-5, frac=_frac(k))
    for k in got[0].files:
        if k[:2] in ('G.', 'D.'):
            assert np.array_equal(got[0][k], got[1][k]), f'{k}: with synchronised BatchNorm the replicas agree on every tensor, buffers included'


def test_two_replicas_on_one_gpu_match_dp_emulation():
    """The same data-parallel semantics without a second GPU: two trainer replicas on one device are stepped in lock-step
    through DCGANTrainer._segments (the generator the real step is built from) and their gradient arenas are summed by hand
    where the NCCL all-reduce would run.  Checked against the oracle's DP emulation."""
    import gan_enhanced_pneumonia_classifier_b200 as pkg
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    sdG, sdD = _state()
    trs = []
    for r in range(WORLD):
        G, D = pkg.Generator(NZ, NC, FM), pkg.Discriminator(NC, FM)
        G.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdG.items()})
        D.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdD.items()})
        tr = DCGANTrainer(G.cuda(), D.cuda(), dtype=torch.float32, use_graph=False)
        tr.world = WORLD                                     # Adam scales the summed gradients by 1/world
        trs.append(tr)
    Gs = [orc.GeneratorOracle(NZ, NC, FM, {k: v.copy() for k, v in sdG.items()}) for _ in range(WORLD)]
    Ds = [orc.DiscriminatorOracle(NC, FM, {k: v.copy() for k, v in sdD.items()}) for _ in range(WORLD)]
    oG = [orc.AdamOracle(orc.param_keys(g.plan), 2e-4, 0.5) for g in Gs]
    oD = [orc.AdamOracle(orc.param_keys(d.plan), 2e-4, 0.5) for d in Ds]
    for it in range(STEPS):
        shards = [_shard(r, it) for r in range(WORLD)]
        gens = [tr._segments(torch.from_numpy(s[0]).cuda(), torch.from_numpy(s[1]).cuda(), overlap=False) for tr, s in zip(trs, shards)]
        hist = [None] * WORLD
        for phase in ('D', 'G', None):
            for r, g in enumerate(gens):
                try:
                    assert next(g) == phase
                except StopIteration as e:
                    hist[r] = e.value.cpu().numpy()
            if phase is not None:
                arenas = [(tr.arenaD if phase == 'D' else tr.arenaG).grad for tr in trs]
                total = torch.stack(arenas).sum(0)
                for a in arenas:
                    a.copy_(total)
        out, gD_mean, gG_mean = orc.train_iteration_dp(Gs, Ds, oG, oD, [s[0] for s in shards], [s[1] for s in shards])
        for r in range(WORLD):
            want = np.array([out[r][k] for k in ('errD', 'errG', 'D_x', 'D_G_z1', 'D_G_z2')])
            close(hist[r], want, what=f'rank {r} history it{it}', **(dict(rtol=1e-4, atol=1e-6) if it == 0 else dict(rtol=5e-3, atol=1e-4)))
        if it == 0:                                          # summed gradient arenas / world == the emulation's mean gradients
            for tr_arena, keys, mean in ((trs[0].arenaD, orc.param_keys(Ds[0].plan), gD_mean), (trs[0].arenaG, orc.param_keys(Gs[0].plan), gG_mean)):
                for k, (lo, hi) in zip(keys, tr_arena.slices):
                    from parity_utils import grad_close
                    grad_close(tr_arena.grad[lo:hi].cpu().numpy().reshape(mean[k].shape) / WORLD, mean[k], what=f'mean gradient {k}')
    for r in range(WORLD):
        for tag, net, o in (('G', trs[r].netG, Gs[r]), ('D', trs[r].netD, Ds[r])):
            for k, v in net.state_dict().items():
                if k.endswith('num_batches_tracked'):
                    assert int(v) == int(o.sd[k])
                elif 'running' in k:
                    close(v.cpu().numpy(), o.sd[k], rtol=2e-3, atol=1e-4, what=f'rank {r} {tag}.{k}')
                else:
                    weights_close(v.cpu().numpy(), o.sd[k], what=f'rank {r} {tag}.{k}', steps=STEPS, rtol=1e-3, atol=1e-5, frac=_frac(k))
                    assert np.abs(v.cpu().numpy() - o.sd[k]).max() < (5e-5 if k.endswith('.bias') else 1.3e-3), k


def _wgan_worker(rank, port, out_dir):
    import torch.distributed as dist
    from gan_enhanced_pneumonia_classifier_b200 import wggan
    from gan_enhanced_pneumonia_classifier_b200.wgan_trainer import WGANGPTrainer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=WORLD, device_id=torch.device('cuda', rank))
    try:
        torch.manual_seed(7)                                   # same initial weights on both ranks
        G, D = wggan.Generator(NZ, NC, FM).cuda(), wggan.Discriminator(NC, FM).cuda()
        tr = WGANGPTrainer(G, D, critic_iters=2, dtype=torch.float32)
        torch.manual_seed(100 + rank)                          # different shards / noise / interpolation draws per rank
        real = torch.rand(B, NC, 224, 224, device='cuda') * 2 - 1
        hist = torch.stack([tr.step(real) for _ in range(2)]).cpu().numpy()
        np.savez(os.path.join(out_dir, f'wgan{rank}.npz'), hist=hist, **{f'G.{k}': v.cpu().numpy() for k, v in G.state_dict().items()},
                 **{f'D.{k}': v.cpu().numpy() for k, v in D.state_dict().items()})
        tr.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_nccl_ranks_wgan_gp_replicas_stay_identical(tmp_path):
    """WGANGPTrainer under data parallelism: both ranks see different data, so their losses and BatchNorm buffers differ, but every weight
    (critic and generator, after 4 critic + 2 generator updates on summed gradients) must stay bit-identical across the replicas."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.start_processes(_wgan_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True, start_method='spawn')
    a, b = np.load(os.path.join(str(tmp_path), 'wgan0.npz')), np.load(os.path.join(str(tmp_path), 'wgan1.npz'))
    assert np.isfinite(a['hist']).all() and np.isfinite(b['hist']).all() and not np.array_equal(a['hist'], b['hist'])
    for k in a.files:
        if k == 'hist' or 'running' in k or k.endswith('num_batches_tracked'):
            continue
        assert np.array_equal(a[k], b[k]), f'replicas diverged on {k}'
    assert not np.array_equal(a['D.main.3.running_mean'], b['D.main.3.running_mean'])


def _cgan_worker(rank, port, out_dir):
    import json
    import torch.distributed as dist
    from conftest import GOLDEN
    from gan_enhanced_pneumonia_classifier_b200 import cgan
    from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
    from test_oracle_golden import cgan_state
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=WORLD, device_id=torch.device('cuda', rank))
    try:
        g = np.load(os.path.join(GOLDEN, 'cgan_step_nc3.npz'))
        m = json.loads(str(g['meta']))
        G, D = cgan.Generator(m['nz'], 2, m['nc'], m['nf']), cgan.Discriminator(2, m['nc'], m['nf'])
        G.load_state_dict({k: torch.from_numpy(v) for k, v in cgan_state(g, 'G').items()})
        D.load_state_dict({k: torch.from_numpy(v) for k, v in cgan_state(g, 'D').items()})
        tr = CGANTrainer(G.cuda(), D.cuda(), perceptual_weight=0.0, dtype=torch.float32)
        torch.manual_seed(100 + rank)                          # each rank its own images, labels and random draws
        rows, stepped = [], []
        for it, epoch in enumerate((0, 5, 5)):                 # epoch 5: the D-step skip rule is live (rank-averaged probabilities decide)
            real = torch.rand(3, m['nc'], 224, 224, device='cuda') * 2 - 1
            labels = torch.randint(0, 2, (3,), device='cuda')
            before = tr.d_steps
            rows.append(tr.step(real, labels, epoch=epoch).cpu().numpy())
            stepped.append(tr.d_steps - before)
        np.savez(os.path.join(out_dir, f'cgan{rank}.npz'), rows=np.stack(rows), stepped=np.array(stepped), collectives=tr.comm.collectives,
                 **{f'G.{k}': v.cpu().numpy() for k, v in G.state_dict().items() if 'running' not in k and 'num_batches' not in k},
                 **{f'D.{k}': v.cpu().numpy() for k, v in D.state_dict().items() if 'running' not in k and 'num_batches' not in k})
        tr.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_nccl_ranks_cgan_replicas_stay_identical_and_agree_on_the_skip_rule(tmp_path):
    """CGANTrainer under data parallelism: different shards and random draws per rank, gradient arenas summed on the library's communicator, the
    D-step skip rule decided on rank-averaged probabilities -- so the replicas take the same branch, issue the same collectives and hold
    bit-identical weights afterwards (per-rank BatchNorm buffers excluded: statistics are local)."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.start_processes(_cgan_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True, start_method='spawn')
    a, b = (np.load(os.path.join(str(tmp_path), f'cgan{r}.npz')) for r in range(WORLD))
    assert np.array_equal(a['stepped'], b['stepped']) and int(a['collectives']) == int(b['collectives']) > 0
    assert np.isfinite(a['rows']).all() and np.isfinite(b['rows']).all() and not np.array_equal(a['rows'], b['rows'])
    for k in a.files:
        if k[:2] in ('G.', 'D.'):
            assert np.array_equal(a[k], b[k]), f'{k} differs between the replicas'
