"""Pins the oracle (oracle/*.py, CPU restatement) against outputs of the REAL reference that
tools/make_golden.py recorded in tests/golden/ (the reference has no tests of its own)."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import noising_oracle as NO
from oracle import params as P
from oracle import pianobart_oracle as O


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def oracle_run(g, dtype=torch.float32, need_grad=True):
    d, el, dl, heads, ffn, max_pos = [int(x) for x in g['cfg']]
    cfg = O.Cfg(d, el, dl, heads, ffn, max_pos)
    prm = P.make_params(d, el, dl, ffn, max_pos, int(g['seed']))
    p = {k: torch.tensor(v, dtype=dtype) for k, v in prm.items() if not k.startswith('decoder_linear')}
    if need_grad:
        for v in p.values():
            v.requires_grad_(True)
    p['decoder_linear.weight'] = p['encoder_linear.weight']
    p['decoder_linear.bias'] = p['encoder_linear.bias']
    enc = torch.from_numpy(g['enc'].astype(np.int64))
    dec = torch.from_numpy(g['dec'].astype(np.int64))
    ori = torch.from_numpy(g['ori'].astype(np.int64))
    lm = torch.from_numpy(g['loss_mask'].astype(np.float32)).to(dtype)
    h, eh = O.pianobart_forward(p, cfg, enc, dec, torch.from_numpy(g['enc_mask']), torch.from_numpy(g['dec_mask']))
    logits = O.lm_heads(p, h)
    total, losses = O.pretrain_loss(logits, ori, lm)
    accs = O.pretrain_accuracy(logits, ori, lm)
    return p, h, eh, logits, total, losses, accs


def test_forward_tiny_matches_reference():
    g = load('fwd_tiny')
    p, h, eh, logits, total, losses, accs = oracle_run(g)
    assert np.allclose(h.detach().numpy(), g['last_hidden'], atol=2e-5)
    assert np.allclose(eh.detach().numpy(), g['enc_hidden'], atol=2e-5)
    assert np.allclose(torch.cat(logits, -1).detach().numpy(), g['logits'], atol=5e-5)
    assert abs(total.item() - float(g['total'])) < 1e-5 * abs(float(g['total']))
    assert np.allclose([l.item() for l in losses], g['losses'], rtol=1e-5)
    assert np.allclose([a.item() for a in accs], g['accs'], atol=1e-7)
    total.backward()
    for k in g.files:
        if k.startswith('grad:'):
            name = k[5:]
            assert np.allclose(p[name].grad.numpy(), g[k], atol=1e-6 + 1e-4 * np.abs(g[k]).max()), name
    names = [str(x) for x in g['grad_norm_names']]
    for n, v in zip(names, g['grad_norm_vals']):
        if n.startswith('decoder_linear'):
            continue
        assert abs(p[n].grad.double().norm().item() - v) <= 1e-4 * v + 1e-7, n


def test_forward_mid_matches_reference():
    g = load('fwd_mid')
    p, h, eh, logits, total, losses, accs = oracle_run(g)
    st = int(g['logit_stride'])
    assert np.allclose(torch.cat(logits, -1).detach().numpy()[:, ::st], g['logits_sub'], atol=2e-4)
    assert abs(total.item() - float(g['total'])) < 1e-5 * abs(float(g['total']))
    total.backward()
    gn = np.sqrt(sum(p[n].grad.double().norm().item() ** 2 for n in p if not n.startswith('decoder_linear')))
    assert abs(gn - float(g['grad_total_norm'])) < 1e-4 * float(g['grad_total_norm'])


def _check_noise(ori, enc, lm, S, choice=None, mask_percent=0.15):
    out, loss, c = NO.gen_mask(ori.astype(np.int64), S, mask_percent, choice)
    assert np.array_equal(out, enc.astype(np.int64))
    assert np.array_equal(loss.astype(np.uint8), lm)
    return c


@pytest.mark.parametrize('S', [1024, 64])
def test_noising_batches_bit_exact(S):
    g = load('noising')
    ori, enc, lm, ch = g['S%d_ori' % S], g['S%d_enc' % S], g['S%d_loss_mask' % S], g['S%d_choices' % S]
    for seed in range(ori.shape[0]):
        random.seed(seed)
        np.random.seed(seed)
        for b in range(ori.shape[1]):
            c = _check_noise(ori[seed, b], enc[seed, b], lm[seed, b], S)
            assert c == ch[seed, b]


def test_noising_direct_choices_bit_exact():
    g = load('noising')
    for S in (1024, 10):
        for choice in (1, 2, 3, 4, 5):
            for seed in (11, 12, 13):
                key = 'direct_S%d_c%d_s%d' % (S, choice, seed)
                random.seed(seed)
                np.random.seed(seed)
                _check_noise(g[key + '_ori'], g[key + '_enc'], g[key + '_loss_mask'], S, choice,
                             0.15 if S > 10 else 0.5)


def test_infilling_failure_branch():
    g = load('noising')
    random.seed(5)
    real = np.random.poisson
    np.random.poisson = lambda lam: 0
    try:
        out, loss, c = NO.gen_mask(g['infill_fail_ori'].astype(np.int64), 32, 0.9, 4)
    finally:
        np.random.poisson = real
    assert np.array_equal(out, g['infill_fail_enc'].astype(np.int64))
    assert loss.shape == (32, 8) and not loss.any()
    assert np.array_equal(loss.astype(np.uint8), g['infill_fail_loss_mask'])
    assert random.random() == float(g['infill_fail_state_after'][0])


def test_generation_finetune_loss_pinned_to_executed_reference():
    """Row A16: oracle.generation_finetune_loss == GenerationTrainer.iteration of the reference (finetune_generation.py:
    140-258) executed by tools/make_golden.py (genft_tiny.npz): per-head losses, total, gradients."""
    g = load('genft_tiny')
    d, el, dl, heads, ffn, max_pos = [int(x) for x in g['cfg']]
    prm = P.make_params(d, el, dl, ffn, max_pos, int(g['seed']))
    p = {k: torch.tensor(v).requires_grad_(True) for k, v in prm.items() if not k.startswith('decoder_linear')}
    p['decoder_linear.weight'], p['decoder_linear.bias'] = p['encoder_linear.weight'], p['encoder_linear.bias']
    x = torch.from_numpy(g['x'].astype(np.int64))
    y = torch.from_numpy(g['y'].astype(np.int64))
    keep = (x[:, :, 0] != 256).float()
    h, _ = O.pianobart_forward(p, O.Cfg(d, el, dl, heads, ffn, max_pos), x, x, keep, keep)   # y_shift = x (:155)
    total, losses = O.generation_finetune_loss(O.lm_heads(p, h), y, keep)
    assert abs(total.item() - float(g['train_total'])) < 1e-5 * float(g['train_total'])
    assert np.allclose([l.item() for l in losses], g['train_losses'], rtol=1e-5)
    total.backward()
    for k in g.files:
        if k.startswith('grad:'):
            assert np.allclose(p[k[5:]].grad.numpy(), g[k], atol=1e-6 + 1e-4 * np.abs(g[k]).max()), k
    gn = np.sqrt(sum(v.grad.double().norm().item() ** 2 for n, v in p.items() if not n.startswith('decoder_linear')))
    assert abs(gn - float(g['grad_norm_preclip'])) < 1e-4 * float(g['grad_norm_preclip'])


def test_generate_default_fixture_matches_oracle_sampler():
    """The default-size decode fixture: oracle.sample_step on the reference's recorded logits reproduces the tokens the
    reference's sampler picked (model.py:68-107), consuming numpy's stream in the same order (8 draws per step)."""
    g = load('generate_default')
    steps = {int(t): i for i, t in enumerate(g['steps'])}
    seg = np.cumsum([0, 262, 134, 135, 262, 134, 38, 260, 55])
    for key in ('A', 'B'):
        rs = np.random.RandomState(0)
        mism = 0
        for t in range(1024):
            if t not in steps:
                rs.random_sample(8)
                continue
            row = torch.from_numpy(g['tf_logits_' + key][steps[t]])
            tok = O.sample_step([row[seg[j]:seg[j + 1]] for j in range(8)], rs)
            mism += int((np.asarray(tok) != g['ref_sampled_' + key][steps[t]]).sum())
        assert mism == 0, (key, mism)


def test_truncation_oracle_pinned_to_executed_reference():
    """oracle.postprocess_oracle.octuple_truncate == demo.Octuple2Midi (demo.py:72-102) executed with the MIDI codec stubbed."""
    from oracle import postprocess_oracle as PO
    g = load('truncate')
    for x, edited, n in zip(g['inputs'], g['edited'], g['lengths']):
        out, length = PO.octuple_truncate(x)
        assert np.array_equal(out, edited.astype(np.int64))
        assert (length if length else -1) == int(n) or (length == 0 and n == -1)
