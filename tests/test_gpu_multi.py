"""2-GPU NCCL test of the data-parallel pretraining step (skipped on a single-GPU box):
all-reduced gradients of two half-batches == gradients of the full batch on one GPU."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from util import build_cuda_model, golden_inputs, load_golden
    from pianobart_b200.pretrain import PretrainStep
    g = load_golden('fwd_tiny')
    dev = 'cuda:%d' % rank
    enc, dec, ori, lmask, em, dm = [t[:4] for t in golden_inputs(g, dev)]
    pb, lm = build_cuda_model(g['cfg'], int(g['seed']), 'fp32', device=dev)
    S = enc.shape[1]
    # full batch on every rank (single-process reference of the same kernels)
    full = PretrainStep(lm, 4, S, None, 0.15)
    full.set_device_batch(enc, dec, ori, lmask, em, dm)
    full.run(train=True)
    t_full, l_full, a_full = full.fetch_stats()
    g_full = pb._grad.clone()
    # sharded: rank r takes samples [2r, 2r+2)
    sl = slice(2 * rank, 2 * rank + 2)
    st = PretrainStep(lm, 2, S, None, 0.15, dist.group.WORLD)
    st.set_device_batch(enc[sl], dec[sl], ori[sl], lmask[sl], em[sl], dm[sl])
    st.run(train=True)
    t_dp, l_dp, a_dp = st.fetch_stats()
    torch.cuda.synchronize()
    err = ((pb._grad - g_full).abs().max() / g_full.abs().max()).item()
    out[rank] = (err, abs(t_dp - t_full) / abs(t_full), float(np.abs(a_dp - a_full).max()))
    dist.destroy_process_group()


def test_two_gpu_data_parallel_step_matches_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        err, lerr, aerr = out[r]
        assert err < 1e-4 and lerr < 1e-5 and aerr < 1e-6, (r, err, lerr, aerr)
