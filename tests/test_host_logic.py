"""CPU tests of the host-side logic: noise-plan generator (bit-exact against the reference fixtures),
parameter layout / state_dict compatibility, vocabulary structure."""
import json
import os
import random

import numpy as np
import torch

from conftest import GOLDEN
from pianobart_b200 import noising as N

PAD = np.array([256, 128, 129, 256, 128, 32, 254, 49])
MASK = PAD + 1


def apply_plan_host(ori, plan):
    """numpy emulation of csrc/noise.cu (the device half) for CPU-side checking of the plan."""
    B, S, _ = ori.shape
    out = np.empty_like(ori)
    lm = np.zeros((B, S, 8), np.uint8)
    rt = np.array(plan.rand_tok).reshape(-1, 8)
    for b in range(B):
        src = plan.src[b]
        rows = np.where((src >= 0)[:, None], ori[b][np.maximum(src, 0)], 0)
        rows = np.where((src == -1)[:, None], PAD, rows)
        rows = np.where((src == -2)[:, None], MASK, rows)
        for s in np.flatnonzero(src <= -3):
            rows[s] = rt[-3 - src[s]]
        out[b] = rows
        if plan.loss_mode[b] == 0:
            lm[b] = plan.loss[b][:, None]
        elif plan.loss_mode[b] == 1:
            lm[b] = (out[b] != ori[b]).any(1)[:, None]
    return out, lm


def test_noise_plan_bit_exact_against_reference_fixtures():
    g = np.load(os.path.join(GOLDEN, 'noising.npz'))
    for S in (1024, 64):
        ori, enc, lmk, ch = (g['S%d_ori' % S].astype(np.int64), g['S%d_enc' % S], g['S%d_loss_mask' % S],
                             g['S%d_choices' % S])
        for seed in range(ori.shape[0]):
            random.seed(seed)
            np.random.seed(seed)
            plan = N.make_plan(ori[seed], S)
            out, lm = apply_plan_host(ori[seed], plan)
            assert np.array_equal(out, enc[seed])
            assert np.array_equal(lm, lmk[seed])
            assert np.array_equal(plan.choices, ch[seed])


def test_noise_plan_direct_choices():
    g = np.load(os.path.join(GOLDEN, 'noising.npz'))
    for S, mp in ((1024, 0.15), (10, 0.5)):
        for choice in (1, 2, 3, 4, 5):
            for seed in (11, 12, 13):
                key = 'direct_S%d_c%d_s%d' % (S, choice, seed)
                random.seed(seed)
                np.random.seed(seed)
                ori = g[key + '_ori'].astype(np.int64)[None]
                plan = N.make_plan(ori, S, mp, [choice])
                out, lm = apply_plan_host(ori, plan)
                assert np.array_equal(out[0], g[key + '_enc']), key
                assert np.array_equal(lm[0], g[key + '_loss_mask']), key


def test_state_dict_keys_match_reference_fixture():
    from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
    from pianobart_b200.vocab import build_octuple_vocab
    with open(os.path.join(GOLDEN, 'state_dict_keys.json')) as f:
        ref = json.load(f)
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=32, d_model=64, encoder_layers=2, decoder_layers=2, encoder_ffn_dim=128,
                    decoder_ffn_dim=128, encoder_attention_heads=4, decoder_attention_heads=4)
    pb = PianoBart(bc, e2w, w2e)
    mine_pb = {k: list(v.shape) for k, v in pb.state_dict().items()}
    assert mine_pb == ref['PianoBart']
    lm = PianoBartLM(pb)
    mine_lm = {k: list(v.shape) for k, v in lm.state_dict().items()}
    assert mine_lm == ref['PianoBartLM']


def test_vocab_structure():
    from pianobart_b200.vocab import build_octuple_vocab
    e2w, w2e = build_octuple_vocab()
    assert list(e2w.keys()) == ['Bar', 'Position', 'Pitch', 'Duration', 'Velocity', 'Instrument', 'Tempo', 'TimeSig']
    classes = ['Bar', 'Position', 'Instrument', 'Pitch', 'Duration', 'Velocity', 'TimeSig', 'Tempo']
    assert [len(e2w[k]) for k in classes] == [262, 134, 135, 262, 134, 38, 260, 55]
    assert [e2w[k]['%s <PAD>' % k] for k in classes] == [256, 128, 129, 256, 128, 32, 254, 49]
    assert e2w['Bar']['Bar <SOS>'] == 258 and e2w['Tempo']['Tempo <EOS>'] == 52


def test_param_layout_fused_groups_are_contiguous():
    from pianobart_b200.engine import ParamLayout
    lay = ParamLayout(64, 2, 2, 128, 32, True)
    for fname, (off, shape) in lay.fused.items():
        n = int(np.prod(shape))
        members = sorted((o, s, k) for k, (o, s) in lay.entries.items() if off <= o < off + n)
        pos = off
        for o, s, k in members:
            assert o == pos, (fname, k)
            pos += int(np.prod(s))
        assert pos == off + n
    los = sorted(lay.ranges.values())
    assert los[0][0] == 0 and los[-1][1] == lay.size
    for (a, b), (c, d) in zip(los[:-1], los[1:]):
        assert b == c


def test_param_layout_ranges_tile_the_flat_buffer():
    """The 'grads_final' ranges of the backward plan (gradient buckets, clip-norm partials) must cover the flat parameter
    buffer exactly once - PretrainStep checks the same property on the recorded plan before it trusts the partial norms."""
    from pianobart_b200.engine import ParamLayout
    for args in ((64, 2, 2, 128, 32, True), (1024, 8, 8, 2048, 1024, True), (128, 1, 3, 256, 64, False)):
        lay = ParamLayout(*args)
        pos = 0
        for lo, hi in sorted(lay.ranges.values()):
            assert lo == pos and hi > lo
            pos = hi
        assert pos == lay.size


def test_wgrad_split_k_wave_model():
    """engine.choose_split_k picks the split measured fastest on a 148-SM B200 (tools/gpu_wgrad_split.py) and never
    exceeds a quarter of the k-blocks."""
    from pianobart_b200.engine import choose_split_k
    M = 16 * 1024
    for (n_out, n_in), want in {(1024, 1024): 4, (3072, 1024): 3, (2048, 1024): 2, (1024, 2048): 2, (1280, 1024): 7}.items():
        assert choose_split_k(n_out, n_in, M, 256, True, 148) == want
    assert choose_split_k(64, 64, 128, 256, False, 148) == 1          # 2 k-blocks: no split
    assert 1 <= choose_split_k(256, 256, 4096, 128, False, 148) <= 16


def test_global_plan_slices_equal_sequential_reference_stream():
    """under data parallelism every rank draws the plan of the GLOBAL batch in sample order and keeps
    its slice - the union of the rank slices is bit-exactly the reference's single-process corruption of that batch."""
    g = np.load(os.path.join(GOLDEN, 'noising.npz'))
    S = 1024
    ori, enc, lmk = g['S%d_ori' % S].astype(np.int64), g['S%d_enc' % S], g['S%d_loss_mask' % S]
    for seed in range(6):
        for world in (2, 4):
            per = ori.shape[1] // world
            for rank in range(world):
                random.seed(seed)
                np.random.seed(seed)
                plan = N.slice_plan(N.make_plan(ori[seed], S), rank * per, (rank + 1) * per)
                out, lm = apply_plan_host(ori[seed][rank * per:(rank + 1) * per], plan)
                assert np.array_equal(out, enc[seed][rank * per:(rank + 1) * per])
                assert np.array_equal(lm, lmk[seed][rank * per:(rank + 1) * per])


def test_plan_prefetcher_preserves_rng_order():
    """The prefetch thread draws plans one step ahead; the decisions must be the sequential ones."""
    from pianobart_b200.pretrain import PlanPrefetcher
    g = np.load(os.path.join(GOLDEN, 'noising.npz'))
    S = 64
    ori, enc, lmk = g['S%d_ori' % S].astype(np.int64), g['S%d_enc' % S], g['S%d_loss_mask' % S]
    # the fixture re-seeds per batch; emulate one long stream instead: sequential plans == prefetched plans
    random.seed(123); np.random.seed(123)
    seq = [N.make_plan(ori[i], S) for i in range(8)]
    random.seed(123); np.random.seed(123)
    got = list(PlanPrefetcher(iter([torch.from_numpy(ori[i]) for i in range(8)]), 0.15))
    assert len(got) == 8
    for a, (o, b) in zip(seq, got):
        assert np.array_equal(a.src, b.src) and np.array_equal(a.loss, b.loss) and np.array_equal(a.loss_mode, b.loss_mode)
        assert a.rand_tok == b.rand_tok or np.array_equal(np.array(a.rand_tok), np.array(b.rand_tok))


def test_bucket_reducer_arms_buckets_until_the_next_attention_backward():
    """parallel.BucketReducer with aligned launches: a complete bucket is held until engine.Plan.run reaches the next attention
    backward (on_attention) or the end of the pass (finish); every element is still reduced exactly once, in marker order."""
    import numpy as np
    import torch
    from pianobart_b200.engine import ParamLayout
    from pianobart_b200.parallel import BucketReducer
    lay = ParamLayout(64, 2, 2, 128, 32, True)
    g = torch.zeros(lay.size)
    red = BucketReducer(g, None, target_bytes=64 * 1024)
    red.align = True                       # (normally: CUDA comm stream present and PIANOBART_B200_COMM_ALIGN != 0)
    launched = []
    red._launch = lambda lo, hi: launched.append((lo, hi)) or red.issued.append((lo, hi))
    order = ['heads', 'decoder.layers.1', 'decoder.layers.0', 'decoder.front', 'encoder.layers.1', 'encoder.layers.0',
             'encoder.front', 'front']
    seen_at_attention = []
    for k in order:
        if 'layers' in k:                  # each layer's backward contains attention backward launches before its marker
            red.on_attention('attn_bwd')
            seen_at_attention.append(len(launched))
        n_before = len(launched)
        red.on_final('grads_final', *lay.ranges[k])
        assert len(launched) == n_before   # markers only arm
    assert len(red.armed) >= 1             # the last bucket waits for finish()
    issued = red.finish()
    cover = np.zeros(lay.size, dtype=np.int32)
    for lo, hi in issued:
        cover[lo:hi] += 1
    assert (cover == 1).all() and issued == launched and not red.armed
    assert seen_at_attention[-1] >= 1      # earlier buckets went out next to an attention backward, not at the end
