"""GPU parity tests of the KV-cache decode path against the reference's generate() fixtures
(tests/golden/generate_tiny.npz, recorded by running reference model.py:28-107)."""
import ctypes as C

import numpy as np
import pytest
import torch

from util import build_cuda_model, load_golden

pytestmark = pytest.mark.gpu


def _setup():
    g = load_golden('generate_tiny')
    pb, lm = build_cuda_model(g['cfg'], int(g['seed']), 'bf16', suppress_specials=True)
    lm.eval()
    enc = torch.from_numpy(g['enc'].astype(np.int64)).cuda()
    mask = (enc[:, :, 0] != pb.bar_pad_word).float()
    return g, pb, lm, enc, mask


def test_teacher_forced_step_logits_match_reference():
    """Position t of the reference's full re-forward == step t of the KV-cache decode given the same prefix."""
    from pianobart_b200.generate import Generator
    g, pb, lm, enc, mask = _setup()
    S = enc.shape[1]
    res = g['result_seed0'].astype(np.int64)
    n = int((res[0, :, 0] != 256).sum())
    forced = torch.from_numpy(res)                      # token fed after step t = reference's token t
    gen = Generator(lm, 1, S, S, use_graph=False)
    gen.start(enc, mask, np.zeros((1, S, 8)), forced)
    ref = g['tf_logits'][0]
    worst = 0.0
    for t in range(min(n, S - 1)):
        gen.run_steps(1)
        torch.cuda.synchronize()
        got = gen.logits[0].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref[t]).max() / np.abs(ref[t]).max()))
    assert worst < 3e-2, worst


def test_sampler_reproduces_reference_tokens_from_reference_logits():
    """Feed the reference's own logits + numpy's uniform stream: the device sampler must pick the reference's tokens."""
    from pianobart_b200 import _lib as L
    from pianobart_b200.generate import SAMPLE_P, SAMPLE_T
    g = load_golden('generate_tiny')
    lib = L.lib()
    ref_logits = g['tf_logits'][0]                      # (S, 1280) logits the reference sampled from at step t
    res = g['result_seed0'].astype(np.int64)[0]
    S = ref_logits.shape[0]
    n = int((res[:, 0] != 256).sum())
    np.random.seed(0)
    uni = np.random.random_sample((1, S, 8))
    dev = 'cuda:0'
    d_uni = torch.from_numpy(uni).to(dev)
    t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    cur = torch.zeros(1, 8, dtype=torch.int32, device=dev)
    sampled = torch.zeros(1, S, 8, dtype=torch.int32, device=dev)
    seg = (C.c_int * 8)(262, 134, 135, 262, 134, 38, 260, 55)
    tt = (C.c_float * 8)(*[float(x) for x in SAMPLE_T])
    pp = (C.c_float * 8)(*[float(x) for x in SAMPLE_P])
    P = C.c_void_p
    mism = 0
    for t in range(min(n, S - 1)):
        lg = torch.from_numpy(ref_logits[t:t + 1].copy()).to(dev)
        t_dev.fill_(t)
        L.check(lib.pb_decode_sample(P(lg.data_ptr()), P(d_uni.data_ptr()), P(None), P(t_dev.data_ptr()), P(cur.data_ptr()),
                                     P(sampled.data_ptr()), 1, S, seg, tt, pp, L.stream_ptr()), 'sample')
        torch.cuda.synchronize()
        mism += int((cur[0].cpu().numpy() != res[t]).sum())
    # exact up to float-rounding ties at the nucleus threshold: allow a vanishing number of flips
    assert mism <= 2, mism


def test_generate_api_shape_stop_rule_and_rng_advance():
    g, pb, lm, enc, mask = _setup()
    S = enc.shape[1]
    np.random.seed(3)
    out = lm(enc, encoder_attention_mask=mask, generate=True)
    assert out.shape == (1, S, 8) and out.dtype == torch.int64
    o = out[0].cpu().numpy()
    pad = pb.pad_word_np
    valid = (o < pad).all(1)
    n = int(valid.sum())
    assert valid[:n].all() and (o[n:] == pad).all()     # PAD-filled tail, no special tokens inside
    after = np.random.random_sample()
    np.random.seed(3)
    executed = n + (1 if n < S else 0)
    np.random.random_sample(8 * min(executed, S))
    assert after == np.random.random_sample()


def test_batched_decode_teacher_forced():
    """New capability (reference exits unless batch == 1): every row of a batch-3 decode reproduces the
    reference's per-step logits for its own prompt (teacher-forced; CUDA-graph replay path)."""
    from pianobart_b200.generate import Generator
    g, pb, lm, enc, mask = _setup()
    S = enc.shape[1]
    res = g['result_seed0'].astype(np.int64)
    n = int((res[0, :, 0] != 256).sum())
    B = 3
    gen = Generator(lm, B, S, S, use_graph=True, force_gemm=True)   # tcgen05 split-K decode path (used for large batches)
    gen.start(enc.expand(B, S, 8).contiguous(), mask.expand(B, S).contiguous(), np.zeros((B, S, 8)),
              torch.from_numpy(res).expand(B, S, 8).contiguous())
    ref = g['tf_logits'][0]
    worst = 0.0
    for t in range(min(n, S - 1)):
        gen.run_steps(1)
        torch.cuda.synchronize()
        got = gen.logits.cpu().numpy()
        for b in range(B):
            worst = max(worst, float(np.abs(got[b] - ref[t]).max() / np.abs(ref[t]).max()))
    assert worst < 3e-2, worst
    result, n_written, done = gen.finish()
    assert (result[:, :n_written[0]].cpu().numpy() == res[:, :n_written[0]]).all()


# --------------------------------------------------------------------------- BASELINE.json configs[2]: default model
def _setup_default():
    g = load_golden('generate_default')
    pb, lm = build_cuda_model(g['cfg'], int(g['seed']), 'bf16', suppress_specials=True)
    lm.eval()
    return g, pb, lm


@pytest.mark.parametrize('B', [1, 64])
def test_decode_default_model_teacher_forced_vs_reference(B):
    """KV-cache decode at the configuration the bench quotes (d=1024, hd=128, 8 layers, S_enc=1024, batch 1 and 64)
    against the executed reference (tests/golden/generate_default.npz, tools/make_golden.py golden_generate_default):
    per-step logits for the reference's prefix at 54 steps spread over the whole 1024-step range (long caches, every
    key split of decode attention, the ticket combine), and the tokens the reference's sampler picked.
    Rows alternate between a padded and a full-length prompt, so per-row key masks and cache rows are exercised."""
    from pianobart_b200.generate import Generator
    g, pb, lm = _setup_default()
    S = 1024
    keys = ['A'] if B == 1 else ['A', 'B']
    rows = [keys[b % len(keys)] for b in range(B)]
    enc = torch.from_numpy(np.concatenate([g['enc_' + k].astype(np.int64) for k in rows])).cuda()
    forced = torch.from_numpy(np.concatenate([g['forced_' + k].astype(np.int64) for k in rows])).cuda()
    mask = (enc[:, :, 0] != pb.bar_pad_word).float()
    np.random.seed(0)
    uni = np.tile(np.random.random_sample((1, S, 8)), (B, 1, 1))
    gen = Generator(lm, B, S, S)
    gen.start(enc, mask, uni, forced)
    steps = [int(t) for t in g['steps']]
    done, worst, mism, total = 0, 0.0, 0, 0
    for i, t in enumerate(steps):
        gen.run_steps(t + 1 - done)
        done = t + 1
        torch.cuda.synchronize()
        got = gen.logits.cpu().numpy()
        smp = gen.sampled[:, t].cpu().numpy()
        for b in (range(B) if B <= 4 else (0, 1, B // 2, B - 2, B - 1)):
            ref = g['tf_logits_' + rows[b]][i]
            worst = max(worst, float(np.abs(got[b] - ref).max() / np.abs(ref).max()))
            mism += int((smp[b] != g['ref_sampled_' + rows[b]][i]).sum())
            total += 8
    assert worst < 5e-2, worst
    # Sampled tokens from OUR bf16 logits vs the reference's tokens from its fp32 logits: with random-init weights the
    # distributions are flat (a 0.9-nucleus of ~230 candidates), so a logit perturbation of bf16 size moves the inverse-CDF
    # pick for the three nucleus heads most of the time and flips greedy near-ties; the sampler itself is pinned exactly
    # on the reference's logits in test_sampler_on_reference_logits_default_vocab.  Here: a loose sanity bound only.
    assert mism <= 0.35 * total, (mism, total)


def test_sampler_on_reference_logits_default_vocab():
    """Device sampler fed with the REFERENCE's logits (default model) and numpy's uniform stream must pick the
    reference's tokens (model.py:68-107 executed: greedy for p == 1, nucleus 0.9 otherwise)."""
    from pianobart_b200 import _lib as L
    from pianobart_b200.generate import SAMPLE_P, SAMPLE_T
    g = load_golden('generate_default')
    lib = L.lib()
    S = 1024
    np.random.seed(0)
    uni = np.random.random_sample((1, S, 8))
    dev = 'cuda:0'
    d_uni = torch.from_numpy(uni).to(dev)
    t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    cur = torch.zeros(1, 8, dtype=torch.int32, device=dev)
    sampled = torch.zeros(1, S, 8, dtype=torch.int32, device=dev)
    seg = (C.c_int * 8)(262, 134, 135, 262, 134, 38, 260, 55)
    tt = (C.c_float * 8)(*[float(x) for x in SAMPLE_T])
    pp = (C.c_float * 8)(*[float(x) for x in SAMPLE_P])
    P = C.c_void_p
    mism = 0
    for key in ('A', 'B'):
        for i, t in enumerate(g['steps']):
            lg = torch.from_numpy(g['tf_logits_' + key][i:i + 1].copy()).to(dev)
            t_dev.fill_(int(t))
            L.check(lib.pb_decode_sample(P(lg.data_ptr()), P(d_uni.data_ptr()), P(None), P(t_dev.data_ptr()), P(cur.data_ptr()),
                                         P(sampled.data_ptr()), 1, S, seg, tt, pp, L.stream_ptr()), 'sample')
            torch.cuda.synchronize()
            mism += int((cur[0].cpu().numpy() != g['ref_sampled_' + key][i]).sum())
    assert mism <= 2, mism


def test_generate_loop_default_width_vs_reference_loop():
    """The reference's real generate loop (model.py:28-66) was executed at the default model width on a 64-token prompt:
    our KV-cache decode reproduces its per-step logits along its own trajectory, and the public generate=True call returns
    a well-formed result of the same layout."""
    from pianobart_b200.generate import Generator
    g, pb, lm = _setup_default()
    enc = torch.from_numpy(g['enc_short'].astype(np.int64)).cuda()
    S = enc.shape[1]
    mask = (enc[:, :, 0] != pb.bar_pad_word).float()
    res = g['result_short_seed0'].astype(np.int64)
    n = int((res[0, :, 0] != 256).sum())
    gen = Generator(lm, 1, S, S)
    np.random.seed(0)
    uni = np.random.random_sample((1, S, 8))
    gen.start(enc, mask, uni, torch.from_numpy(res))
    ref = g['tf_logits_short'][0]
    worst, mism = 0.0, 0
    for t in range(min(n, S - 1)):
        gen.run_steps(1)
        torch.cuda.synchronize()
        got = gen.logits[0].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref[t]).max() / np.abs(ref[t]).max()))
        mism += int((gen.sampled[0, t].cpu().numpy() != res[0, t]).sum())
    assert worst < 5e-2, worst
    assert mism <= 0.35 * 8 * n, (mism, n)          # (see the note in the teacher-forced test above)
    np.random.seed(0)
    out = lm(enc, encoder_attention_mask=mask, generate=True)
    assert out.shape == (1, S, 8) and out.dtype == torch.int64
    o = out[0].cpu().numpy()
    valid = (o < pb.pad_word_np).all(1)
    k = int(valid.sum())
    assert valid[:k].all() and (o[k:] == pb.pad_word_np).all()


def test_octuple_truncate_matches_reference_and_oracle():
    """Device post-processing: pb_octuple_truncate vs the executed reference fixture (demo.py:72-102) and the
    oracle on random generated-looking batches, int32 and int64 inputs."""
    from oracle import postprocess_oracle as PO
    from oracle import params as P
    from pianobart_b200.postprocess import octuple_truncate
    g = load_golden('truncate')
    x = torch.from_numpy(g['inputs'].astype(np.int64)).cuda()
    out, ln = octuple_truncate(x)
    assert np.array_equal(out.cpu().numpy(), g['edited'].astype(np.int64))
    want = np.where(g['lengths'] < 0, 0, g['lengths'])
    assert np.array_equal(ln.cpu().numpy(), want)
    rs = np.random.RandomState(3)
    ids = P.synth_ids(37, 200, 900)
    for b in range(37):
        if b % 3:
            ids[b, rs.randint(0, 200), rs.randint(0, 8)] = 300
    out, ln = octuple_truncate(torch.from_numpy(ids).int().cuda())
    for b in range(37):
        ref, n = PO.octuple_truncate(ids[b])
        assert np.array_equal(out[b].cpu().numpy(), ref) and int(ln[b]) == (n or 0)
    o1, l1 = octuple_truncate(torch.from_numpy(ids[5]).cuda())
    assert o1.shape == (200, 8) and np.array_equal(o1.cpu().numpy(), out[5].cpu().numpy())


def test_demo_entry_point_midi_in_midi_out(tmp_path, monkeypatch):
    """main.demo() mirror of reference demo.py: a MIDI file becomes the (1, 1024, 8) prompt (Midi2Octuple), the model generates
    with the KV-cache decode, and the result is truncated, decoded and written as a MIDI file (Octuple2Midi) that reads back
    as notes - the codec and MIDI I/O are pianobart_b200/codec.py."""
    from pianobart_b200 import codec as K, main as M
    from pianobart_b200.postprocess import midi_to_octuple, octuple_to_midi, octuple_truncate
    monkeypatch.chdir(tmp_path)
    rs = np.random.RandomState(0)
    notes = [K.Note(int(t), int(t) + 240, int(rs.randint(50, 80)), 80) for t in range(0, 480 * 4 * 6, 120)]
    K.write_midi(K.Score(480, [K.Instrument(0, False, 'PIANO', notes)], [K.TimeSignature(4, 4, 0)], [K.TempoChange(120.0, 0)]),
                 'in.mid')
    x = midi_to_octuple('in.mid')
    assert tuple(x.shape) == (1, 1024, 8) and int((x[0, :, 0] < 256).sum()) == len(notes) and int(x[0, -1, 0]) == 256
    torch.manual_seed(3)
    np.random.seed(3)
    xin, y = M.demo(['--input', 'in.mid', '--output', 'out.mid', '--nopretrain', '--hs', '64', '--layers', '1', '--ffn_dims', '128',
                     '--heads', '4'])
    assert tuple(y.shape) == (1, 1024, 8) and torch.equal(xin.cpu(), x.long().cpu())
    trunc, ln = octuple_truncate(y)
    if int(ln[0]) > 0:                     # (a random-init model may stop at once: then the reference prints "Generate Fail")
        back = K.read_midi('out.mid')
        assert sum(len(i.notes) for i in back.instruments) > 0
    # the decoder on a sequence with known content: the prompt itself -> MIDI -> the same Octuple rows
    assert octuple_to_midi(x, ['rt.mid']) == ['rt.mid']
    rows = K.score_to_octuple(K.read_midi('rt.mid'))
    want = [tuple(r) for r in x[0].cpu().tolist() if r[0] < 256]
    assert [r[:4] + r[5:7] for r in rows] == [r[:4] + r[5:7] for r in want]
