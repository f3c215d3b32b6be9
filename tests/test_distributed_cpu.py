"""world_size-2 gloo tests (CPU) of the data-parallel host logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_buckets(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pianobart_b200.engine import ParamLayout
    from pianobart_b200.parallel import BucketReducer
    lay = ParamLayout(64, 2, 2, 128, 32, True)
    g = torch.full((lay.size,), float(rank + 1))
    norm_sq = [0.0]

    def partial(lo, hi):     # the clip-norm partial of an all-reduced bucket (pretrain.PretrainStep._bucket_norm on the GPU)
        norm_sq[0] += float((g[lo:hi].double() ** 2).sum())
    red = BucketReducer(g, None, target_bytes=64 * 1024, after_reduce=partial)
    # marker order of the backward plan: heads, decoder layers (last first), decoder front, encoder layers, encoder front, front
    order = ['heads', 'decoder.layers.1', 'decoder.layers.0', 'decoder.front', 'encoder.layers.1', 'encoder.layers.0',
             'encoder.front', 'front']
    for k in order:
        red.on_final('grads_final', *lay.ranges[k])
    issued = red.finish()
    cover = np.zeros(lay.size, dtype=np.int32)
    for lo, hi in issued:
        cover[lo:hi] += 1
    ok = bool((cover == 1).all()) and bool((g == sum(range(1, world + 1))).all()) and len(issued) < len(order)
    ok = ok and abs(norm_sq[0] - float((g.double() ** 2).sum())) <= 1e-9 * norm_sq[0]   # partials = norm of the reduced buffer
    out[rank] = ok
    dist.destroy_process_group()


def test_bucket_reducer_covers_flat_buffer_once_and_sums():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_buckets, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _worker_dp_math(rank, world, port, out):
    """Sharding the batch + all-reducing the mask sums BEFORE backward makes sum-of-rank-gradients equal the
    reference's full-batch gradient (pretrain.py:112-118,179-189 normalise by the full-batch mask sum)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import params as P
    from oracle import pianobart_oracle as O
    torch.set_num_threads(2)
    cfgt = (32, 1, 1, 2, 64, 16)
    cfg = O.Cfg(*cfgt)
    B, S = 4, 16
    ids = torch.from_numpy(P.synth_ids(B, S, 5, padded=True))
    dec = O.shift_right(ids, P.SOS)
    keep = (ids[:, :, 0] != 256).float()
    dkeep = (dec[:, :, 0] != 256).float()
    rs = np.random.RandomState(1)
    lm = torch.from_numpy((rs.rand(B, S, 1) < 0.4).astype(np.float32).repeat(8, 2))

    def make():
        prm = P.make_params(cfgt[0], cfgt[1], cfgt[2], cfgt[4], cfgt[5], 7)
        p = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in prm.items() if not k.startswith('decoder_linear')}
        p['decoder_linear.weight'], p['decoder_linear.bias'] = p['encoder_linear.weight'], p['encoder_linear.bias']
        return p

    W = torch.tensor(O.N_TOKENS_KEY_ORDER, dtype=torch.float64)
    # full batch (what the reference computes on device 0)
    p = make()
    h, _ = O.pianobart_forward(p, cfg, ids, dec, keep, dkeep)
    total, _ = O.pretrain_loss(O.lm_heads(p, h), ids, lm.double())
    total.backward()
    full = {k: v.grad.clone() for k, v in p.items() if v.grad is not None}
    # this rank's shard, denominators all-reduced first
    sl = slice(rank * B // world, (rank + 1) * B // world)
    p = make()
    h, _ = O.pianobart_forward(p, cfg, ids[sl], dec[sl], keep[sl], dkeep[sl])
    logits = O.lm_heads(p, h)
    den = lm[sl].double().sum((0, 1))
    dist.all_reduce(den)
    loss = 0
    for i in range(8):
        lse = torch.logsumexp(logits[i], -1)
        picked = logits[i].gather(-1, ids[sl][..., i][..., None]).squeeze(-1)
        loss = loss + W[i] * ((lse - picked) * lm[sl][..., i].double()).sum() / den[i]
    loss = loss / W.sum()
    loss.backward()
    ok = True
    for k, v in p.items():
        if v.grad is None or k.startswith('decoder_linear'):
            continue
        gsum = v.grad.clone()
        dist.all_reduce(gsum)
        ok &= bool(torch.allclose(gsum, full[k], rtol=1e-9, atol=1e-12))
    out[rank] = ok
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_full_batch_gradient():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_dp_math, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
