"""GPU parity of the finetune wrappers against fixtures recorded from the reference
(tests/golden/cls_tiny.npz) and against the oracle."""
import numpy as np
import pytest
import torch

from oracle import params as P
from util import build_cuda_model, load_golden

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _load_extra(module, prefix, names, seed):
    sd = module.state_dict()
    for k in names:
        sd[k] = torch.from_numpy(P.gen_tensor(prefix + k, tuple(sd[k].shape), seed)).to(sd[k].device)
    module.load_state_dict(sd)


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_sequence_classification_matches_reference(dtype):
    from pianobart_b200.modules import SequenceClassification
    g = load_golden('cls_tiny')
    pb, _ = build_cuda_model(g['cfg'], int(g['seed']), dtype, lm=False)
    sc = SequenceClassification(pb, class_num=4, hs=int(g['cfg'][0])).cuda()
    _load_extra(sc, 'seqcls.', ['attention.ws1.weight', 'attention.ws2.weight', 'classifier.1.weight', 'classifier.1.bias',
                                'classifier.3.weight', 'classifier.3.bias'], 9)
    sc.eval()
    ids = torch.from_numpy(g['ids'].astype(np.int64)).cuda()
    mask = (ids[:, :, 0] != pb.bar_pad_word).float()
    with torch.no_grad():
        out = sc(ids, mask)
    assert _rel(out.cpu().numpy(), g['seqcls_logits']) < (1e-4 if dtype == 'fp32' else 3e-2)


@pytest.mark.parametrize('cn', [4, 8])
def test_token_classification_matches_reference(cn):
    """cn=4: decoder ids = encoder ids; cn=8: PianoBart.change_decoder_embedding path (label embedding 8x64 + Linear)."""
    from pianobart_b200.modules import TokenClassification
    g = load_golden('cls_tiny')
    pb, _ = build_cuda_model(g['cfg'], int(g['seed']), 'fp32', lm=False)
    tc = TokenClassification(pb, class_num=cn, hs=int(g['cfg'][0])).cuda()
    names = ['classifier.1.weight', 'classifier.1.bias', 'classifier.3.weight', 'classifier.3.bias']
    if cn >= 5:
        names += ['pianobart.decoder_emb.lut.weight', 'pianobart.decoder_linear.weight', 'pianobart.decoder_linear.bias']
    _load_extra(tc, 'tokcls%d.' % cn, names, 9)
    tc.eval()
    ids = torch.from_numpy(g['ids'].astype(np.int64)).cuda()
    mask = (ids[:, :, 0] != pb.bar_pad_word).float()
    if cn >= 5:
        dec_in = torch.from_numpy(g['tokcls%d_dec_in' % cn].astype(np.int64)).cuda()
        out = tc(ids, dec_in, mask, mask)
        out.sum().backward()                       # gradient reaches the replacement decoder front end
        assert tc.pianobart.decoder_emb.lut.weight.grad.abs().sum().item() > 0
    else:
        with torch.no_grad():
            out = tc(ids, ids, mask, mask)
    assert _rel(out.detach().cpu().numpy(), g['tokcls%d_logits' % cn]) < 1e-4


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
@pytest.mark.parametrize('task', ['seqcls', 'tokcls4', 'tokcls8'])
def test_finetune_training_gradients_match_reference(task, dtype):
    """Rows A14 / A15: loss and gradients of one finetune step (backbone through the kernel path, head parameters included)
    against the executed reference (cls_tiny.npz: reference SequenceClassification / TokenClassification + the loss of
    finetune.py:125-132, eval-mode arithmetic so that nn.Dropout is off)."""
    from pianobart_b200.modules import SequenceClassification, TokenClassification
    g = load_golden('cls_tiny')
    d = int(g['cfg'][0])
    pb, _ = build_cuda_model(g['cfg'], int(g['seed']), dtype, lm=False)
    ids = torch.from_numpy(g['ids'].astype(np.int64)).cuda()
    mask = (ids[:, :, 0] != pb.bar_pad_word).float()
    if task == 'seqcls':
        m = SequenceClassification(pb, class_num=4, hs=d).cuda()
        _load_extra(m, 'seqcls.', ['attention.ws1.weight', 'attention.ws2.weight', 'classifier.1.weight', 'classifier.1.bias',
                                   'classifier.3.weight', 'classifier.3.bias'], 9)
        m.eval()
        y = torch.from_numpy(g['seqcls_labels']).cuda()
        loss = torch.nn.functional.cross_entropy(m(ids, mask), y, reduction='none').sum() / ids.shape[0]
    elif task == 'tokcls8':
        # class_num >= 5: label embedding + decoder_linear replace the decoder front end (finetune.py:194-198); the gradient
        # wrt those decoder input embeddings leaves the backbone through its own buffer
        m = TokenClassification(pb, class_num=8, hs=d).cuda()
        _load_extra(m, 'tokcls8.', ['classifier.1.weight', 'classifier.1.bias', 'classifier.3.weight', 'classifier.3.bias',
                                    'pianobart.decoder_emb.lut.weight', 'pianobart.decoder_linear.weight',
                                    'pianobart.decoder_linear.bias'], 9)
        m.eval()
        y = torch.from_numpy(g['tokcls8_labels']).cuda()
        y_shift = torch.from_numpy(g['tokcls8_dec_in'].astype(np.int64)).cuda()
        attn_shift = torch.zeros_like(mask)
        attn_shift[:, 1:] = mask[:, :-1]
        attn_shift[:, 0] = mask[:, 0]
        lg = m(ids, y_shift, mask, attn_shift)
        loss = (torch.nn.functional.cross_entropy(lg.permute(0, 2, 1), y, reduction='none') * mask).sum() / mask.sum()
    else:
        m = TokenClassification(pb, class_num=4, hs=d).cuda()
        _load_extra(m, 'tokcls4.', ['classifier.1.weight', 'classifier.1.bias', 'classifier.3.weight', 'classifier.3.bias'], 9)
        m.eval()
        y = torch.from_numpy(g['tokcls4_labels']).cuda()
        lg = m(ids, ids, mask, mask)
        loss = (torch.nn.functional.cross_entropy(lg.permute(0, 2, 1), y, reduction='none') * mask).sum() / mask.sum()
    m.zero_grad()
    loss.backward()
    ref = float(g[task + '_loss'])
    assert abs(loss.item() - ref) / ref < (1e-5 if dtype == 'fp32' else 1e-2)
    sd = dict(m.named_parameters())
    n = 0
    for k in g.files:
        if k.startswith(task + '_grad:'):
            name = k.split(':', 1)[1]
            got, want = sd[name].grad.cpu().numpy().astype(np.float64), g[k].astype(np.float64)
            if dtype == 'fp32':
                assert _rel(got, want) < 2e-4, name
            else:
                # tiny d = 64 model: bf16 rounding noise is large relative to its small gradients - direction and scale
                cos = float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
                assert cos > 0.99 and _rel(got, want) < 0.2, (name, cos, _rel(got, want))
            n += 1
    assert n >= 8


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
@pytest.mark.parametrize('task', ['seqcls', 'tokcls8'])
def test_default_model_finetune_step_matches_reference(task, dtype):
    """BASELINE configs[3] / [4] at the default model size (d 1024, 8 + 8 layers, S 1024, batch 2): logits, loss and
    gradients of the composer sequence task (8 classes) and the velocity token task (class_num 8: label embedding +
    decoder_linear as the decoder front end, shifted labels, finetune.py:194-198) against the executed reference
    (tests/golden/cls_default.npz, tools/make_golden.py::golden_cls_default) - the trainer's own loss path
    (FinetuneTrainer._loss -> pb_heads_ce) included."""
    from pianobart_b200 import heads
    from pianobart_b200.modules import SequenceClassification, TokenClassification
    g = load_golden('cls_default')
    d = int(g['cfg'][0])
    pb, _ = build_cuda_model(g['cfg'], int(g['seed']), dtype, lm=False)
    ids = torch.from_numpy(g['ids'].astype(np.int64)).cuda()
    mask = (ids[:, :, 0] != pb.bar_pad_word).float()
    if task == 'seqcls':
        m = SequenceClassification(pb, class_num=8, hs=d).cuda()
        _load_extra(m, 'seqcls.', ['attention.ws1.weight', 'attention.ws2.weight', 'classifier.1.weight', 'classifier.1.bias',
                                   'classifier.3.weight', 'classifier.3.bias'], 12)
        m.eval()
        y = torch.from_numpy(g['seqcls_labels']).cuda()
        logits = m(ids, mask)
        loss, correct, am = heads.masked_ce(logits, y, None)
        want_logits, got_logits = g['seqcls_logits'], logits.detach().cpu().numpy()
    else:
        m = TokenClassification(pb, class_num=8, hs=d).cuda()
        _load_extra(m, 'tokcls8.', ['classifier.1.weight', 'classifier.1.bias', 'classifier.3.weight', 'classifier.3.bias',
                                    'pianobart.decoder_emb.lut.weight', 'pianobart.decoder_linear.weight',
                                    'pianobart.decoder_linear.bias'], 12)
        m.eval()
        y = torch.from_numpy(g['tokcls8_labels'].astype(np.int64)).cuda()
        y_shift = torch.from_numpy(g['tokcls8_dec_in'].astype(np.int64)).cuda()
        attn_shift = torch.zeros_like(mask)
        attn_shift[:, 1:] = mask[:, :-1]
        attn_shift[:, 0] = mask[:, 0]
        logits = m(ids, y_shift, mask, attn_shift)
        M = logits.shape[0] * logits.shape[1]
        loss, correct, am = heads.masked_ce(logits.reshape(M, -1), y.reshape(M), mask.reshape(M))
        want_logits, got_logits = g['tokcls8_logits'], logits.detach().cpu().numpy()[:, ::8]
    m.zero_grad()
    loss.backward()
    ref = float(g[task + '_loss'])
    assert abs(loss.item() - ref) / ref < (1e-5 if dtype == 'fp32' else 1e-2)
    assert _rel(got_logits, want_logits) < (2e-4 if dtype == 'fp32' else 5e-2)
    sd = dict(m.named_parameters())
    n = 0
    for k in g.files:
        if k.startswith(task + '_grad:'):
            name = k.split(':', 1)[1]
            got, want = sd[name].grad.cpu().numpy().astype(np.float64), g[k].astype(np.float64)
            wmax = float(np.abs(want).max())
            if wmax < 1e-4:
                # The pooling-attention weights: with these weights the sequence softmax is nearly uniform and the decoder
                # outputs of neighbouring positions nearly equal, so p * (dp - sum p dp) cancels to ~1e-6 of its terms - the
                # reference's own value is rounding noise of the backbone output (our pooling kernels agree with fp64 to 1e-6
                # on well-conditioned inputs, tests/test_gpu_heads.py).  Require the right order of magnitude only.
                assert float(np.abs(got).max()) < 50 * wmax + 1e-7, (name, float(np.abs(got).max()), wmax)
            elif dtype == 'fp32':
                # (5e-3 of the largest entry: a ReLU pre-activation within rounding distance of zero flips its gate)
                assert float(np.abs(got - want).max()) < 5e-3 * wmax + 1e-8, (name, _rel(got, want))
            else:
                # bf16 through 16 layers: direction and scale.  (Batch 2: one ReLU gate of the pooled classifier flipping under
                # bf16 noise switches a whole row of classifier.1.weight's gradient on or off - hence the loose max-norm bound)
                cos = float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
                assert cos > 0.98 and _rel(got, want) < 0.5, (name, cos, _rel(got, want))
            n += 1
    assert n >= 6


def test_generation_finetune_loss_matches_oracle():
    """finetune_generation.py:140-258 semantics (y_shift = x, decoder-mask loss, extra per-head factors)."""
    from oracle import pianobart_oracle as O
    from pianobart_b200.finetune_generation import GenerationTrainer
    g = load_golden('fwd_tiny')
    cfgt = [int(v) for v in g['cfg']]
    pb, lm = build_cuda_model(g['cfg'], int(g['seed']), 'fp32')
    tr = GenerationTrainer(pb, None, None, None, 1e-4, None, False, [0], model=lm, verbose=False)
    x = torch.from_numpy(g['ori'].astype(np.int64))
    y = torch.from_numpy(P.synth_ids(x.shape[0], x.shape[1], 77))
    lm.eval()
    loss, losses, accs = tr.step(x, y, train=False)
    prm = P.make_params(cfgt[0], cfgt[1], cfgt[2], cfgt[4], cfgt[5], int(g['seed']))
    p = {k: torch.from_numpy(v) for k, v in prm.items()}
    keep = (x[:, :, 0] != 256).float()
    h, _ = O.pianobart_forward(p, O.Cfg(*cfgt), x, x, keep, keep)
    ref_total, ref_losses = O.generation_finetune_loss(O.lm_heads(p, h), y, keep)
    assert abs(loss - ref_total.item()) / abs(ref_total.item()) < 1e-5
    assert np.allclose(losses, [l.item() * e for l, e in zip(ref_losses, O.GEN_EXTRA_W)], rtol=1e-5)


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_generation_trainer_matches_executed_reference(dtype):
    """Row A16 pinned to the REAL reference: tests/golden/genft_tiny.npz was recorded by running
    GenerationTrainer.iteration of /root/reference/finetune_generation.py:140-258 (valid and train mode, dropout 0, lr 0;
    tools/make_golden.py golden_genft).  Loss, per-head losses, pre-clip gradient norm and gradients must agree."""
    from pianobart_b200.finetune_generation import GenerationTrainer
    g = load_golden('genft_tiny')
    pb, lm = build_cuda_model(g['cfg'], int(g['seed']), dtype, dropout=0.0)
    tr = GenerationTrainer(pb, None, None, None, 0.0, None, False, [0], model=lm, verbose=False)
    x = torch.from_numpy(g['x'].astype(np.int64))
    y = torch.from_numpy(g['y'].astype(np.int64))
    tol = 1e-5 if dtype == 'fp32' else 1e-2
    lm.eval()
    loss, losses, accs = tr.step(x, y, train=False)
    extra = np.array([1, 1, 0.3, 1.5, 1, 1, 0.3, 0.3])
    assert abs(loss - float(g['valid_total'])) / float(g['valid_total']) < tol
    assert np.allclose(losses, g['valid_losses'] * extra, rtol=tol * 10)
    if dtype == 'fp32':
        assert np.allclose(accs, g['valid_acc_rounded'], atol=1e-4)
    lm.train()
    loss, losses, accs = tr.step(x, y, train=True)
    torch.cuda.synchronize()
    assert abs(loss - float(g['train_total'])) / float(g['train_total']) < tol
    gt = tol * 10 if dtype == 'fp32' else 5e-2
    for k in g.files:
        if k.startswith('grad:'):
            assert _rel(pb.flat_grad(k[5:]).cpu().numpy(), g[k]) < gt, k
    gn = float(torch.sqrt((pb._grad.double() ** 2).sum()).item())
    assert abs(gn - float(g['grad_norm_preclip'])) / float(g['grad_norm_preclip']) < (1e-4 if dtype == 'fp32' else 2e-2)


@pytest.mark.parametrize('seq', [True, False])
def test_finetune_trainer_steps_reduce_loss(seq):
    """FinetuneTrainer mirror (finetune.py:152-256): a few optimisation steps on a fixed batch reduce the loss; the
    backbone gradient flows through the kernel path (fp32 mode for a deterministic check)."""
    from pianobart_b200.finetune import FinetuneTrainer
    g = load_golden('cls_tiny')
    pb, _ = build_cuda_model(g['cfg'], int(g['seed']), 'fp32', lm=False)
    d = int(g['cfg'][0])
    tr = FinetuneTrainer(pb, None, None, None, 1e-3, 4 if seq else 3, d, None, False, [0], SeqClass=seq)
    tr.model.eval()   # deterministic (no dropout) for the monotonicity check
    ids = torch.from_numpy(g['ids'].astype(np.int64))
    rs = np.random.RandomState(0)
    y = torch.from_numpy(rs.randint(0, 4 if seq else 3, size=(ids.shape[0],) if seq else ids.shape[:2]))
    losses = []
    for _ in range(4):
        loss, correct, count, out = tr.step(ids, y, mode=0)
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert pb.flat_grad('bart.encoder.layers.0.fc1.weight').abs().sum().item() > 0


def test_pretrain_entry_point_runs_and_checkpoint_roundtrips(tmp_path, monkeypatch):
    """main.pretrain() mirror: one epoch on synthetic data with a small model, then the checkpoint written in the
    reference's layout loads back into a fresh PianoBart with identical parameters."""
    from pianobart_b200 import main as M
    from pianobart_b200.modules import BartConfig, PianoBart
    from pianobart_b200.vocab import build_octuple_vocab
    monkeypatch.chdir(tmp_path)
    M.pretrain(['--synthetic', '8', '--epochs', '1', '--batch_size', '4', '--max_seq_len', '64', '--hs', '64', '--layers', '1',
                '--ffn_dims', '128', '--heads', '4', '--num_workers', '0', '--name', 't', '--dtype', 'bf16', '--lr', '1e-3'])
    ck = torch.load(tmp_path / 'result' / 'pretrain' / 't' / 'model.ckpt', map_location='cpu', weights_only=False)
    assert set(ck.keys()) == {'epoch', 'state_dict', 'best_acc', 'valid_acc', 'valid_loss', 'train_loss', 'optimizer'}
    e2w, w2e = build_octuple_vocab()
    pb = PianoBart(BartConfig(max_position_embeddings=64, d_model=64, encoder_layers=1, decoder_layers=1, encoder_ffn_dim=128,
                              decoder_ffn_dim=128, encoder_attention_heads=4, decoder_attention_heads=4), e2w, w2e)
    pb.load_state_dict(ck['state_dict'])      # strict, reference key names (main.py:167-168)
    assert torch.equal(pb.state_dict()['bart.encoder.layers.0.fc1.weight'], ck['state_dict']['bart.encoder.layers.0.fc1.weight'])
    assert np.isfinite(ck['train_loss']) and (tmp_path / 'result' / 'pretrain' / 't' / 'log').exists()


def test_pretrainer_iteration_prefetch_and_pipelining_are_transparent():
    """`Pretrainer.iteration` draws the noise plan of batch i+1 on a host thread and stages its H2D copies while
    step i runs.  The statistics must be exactly those of the plain sequential loop (upload -> noise -> run -> fetch)."""
    import random
    from pianobart_b200.pretrain import Pretrainer
    g = load_golden('noising')
    ori = g['S64_ori'].astype(np.int64)
    pb, _ = build_cuda_model((64, 1, 1, 2, 64, 1024), 3, 'fp32', lm=False)
    tr = Pretrainer(pb, None, None, 1e-4, ori.shape[1], 64, 0.15, False, [0], verbose=False)
    tr.model.eval()
    data = [torch.from_numpy(ori[i]) for i in range(5)] + [torch.from_numpy(ori[5][:3])]      # ragged last batch
    out = []
    for prefetch in (True, False):
        tr.prefetch = prefetch
        random.seed(5); np.random.seed(5)
        out.append(tr.iteration(data, 64, train=False))
    assert out[0] == out[1]
    # sequential reference loop on the same stream
    random.seed(5); np.random.seed(5)
    tot = 0.0
    for batch in data:
        st = tr._step(batch.shape[0], batch.shape[1])
        st.upload(batch); st.noise(); st.run(train=False)
        tot += st.fetch_stats()[0]
    assert round(tot / len(data), 3) == out[0][0]


TINY = ['--max_seq_len', '64', '--hs', '64', '--layers', '1', '--ffn_dims', '128', '--heads', '4', '--num_workers', '0',
        '--dtype', 'bf16', '--nopretrain']


@pytest.mark.parametrize('task', ['composer', 'velocity', 'melody'])
def test_finetune_entry_point_runs(task, tmp_path, monkeypatch):
    """main.finetune() mirror (reference main.py:103-211) on synthetic data with a small model: the three task shapes - sequence
    labels (composer), token labels through the label-embedding decoder front end (velocity: class_num 7 + 1) and token
    labels with decoder ids = encoder ids (melody with --class_num 3) - train, validate, test and checkpoint."""
    from pianobart_b200 import main as M
    monkeypatch.chdir(tmp_path)
    extra = ['--class_num', '3'] if task == 'melody' else []
    tr = M.finetune(['--task', task, '--dataset', 'Pianist8', '--synthetic', '16', '--epochs', '2', '--batch_size', '4',
                     '--lr', '1e-3', '--name', 't'] + TINY + extra)
    ck = torch.load(tmp_path / 'result' / 'finetune' / (task + '_t') / 'model.ckpt', map_location='cpu', weights_only=False)
    assert {'epoch', 'state_dict', 'valid_acc', 'valid_loss', 'train_loss', 'train_acc', 'optimizer'} <= set(ck.keys())
    assert np.isfinite(ck['train_loss']) and np.isfinite(ck['valid_loss'])
    log = (tmp_path / 'result' / 'finetune' / (task + '_t') / 'log').read_text()
    assert 'Epoch 2' in log
    if task == 'velocity':
        assert any(k.startswith('pianobart.decoder_emb') for k in ck['state_dict'])


def test_generation_entry_points_run(tmp_path, monkeypatch):
    """main.finetune_generation() (reference main.py:214-321) then main.eval_generation() (eval_generation.py:49-115) on the
    checkpoint it wrote: the .npy has the reference's shape / dtype, rows after the stop step are <PAD>, and the truncated
    copy obeys the Octuple2Midi rules."""
    from pianobart_b200 import main as M
    monkeypatch.chdir(tmp_path)
    M.finetune_generation(['--synthetic', '8', '--epochs', '1', '--batch_size', '4', '--lr', '1e-3', '--name', 'g'] + TINY)
    ck_path = tmp_path / 'result' / 'finetune' / 'generation_g' / 'model.ckpt'
    assert ck_path.exists()
    args = [a for a in TINY if a != '--nopretrain']
    out = M.eval_generation(['--ckpt', str(ck_path), '--synthetic', '3', '--batch_size', '1', '--output', 'gen.npy', '--truncate'] + args)
    arr = np.load(tmp_path / 'gen.npy')
    assert arr.shape == (3, 64, 8) and arr.dtype == np.float32 and np.array_equal(arr, out.numpy())
    trunc, lens = np.load(tmp_path / 'gen.trunc.npy'), np.load(tmp_path / 'gen.len.npy')
    assert trunc.shape == (3, 64, 8) and lens.shape == (3,) and (lens >= 0).all() and (lens < 64).all()
