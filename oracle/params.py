"""ORACLE support - test infrastructure only.

Deterministic parameter generator shared by tools/make_golden.py (which loads these values
into the *real* reference modules) and by the tests (which load the same values into the
CUDA implementation and into the oracle restatement).  numpy's legacy RandomState stream is
stable across numpy versions, so fixtures only need to store (config, seed), not weights.

Key names / shapes follow the reference state_dict: PianoBart.py:19-54,
model.py:109-117 (mask_lm.proj.N), HF BartModel parameter names.
"""
import zlib

import numpy as np

N_TOKENS = [262, 134, 135, 262, 134, 38, 260, 55]


def param_shapes(d_model, enc_layers, dec_layers, ffn, max_pos, lm_heads=True, vocab=50265):
    s = {}
    for i, n in enumerate(N_TOKENS):
        s['word_emb.%d.lut.weight' % i] = (n, 256)
    s['encoder_linear.weight'] = (d_model, 2048)
    s['encoder_linear.bias'] = (d_model,)
    for side, nl in (('encoder', enc_layers), ('decoder', dec_layers)):
        s['bart.%s.embed_positions.weight' % side] = (max_pos + 2, d_model)
        s['bart.%s.layernorm_embedding.weight' % side] = (d_model,)
        s['bart.%s.layernorm_embedding.bias' % side] = (d_model,)
        for l in range(nl):
            pre = 'bart.%s.layers.%d.' % (side, l)
            attns = ['self_attn'] + (['encoder_attn'] if side == 'decoder' else [])
            for a in attns:
                for pr in ('k_proj', 'v_proj', 'q_proj', 'out_proj'):
                    s[pre + a + '.' + pr + '.weight'] = (d_model, d_model)
                    s[pre + a + '.' + pr + '.bias'] = (d_model,)
                s[pre + a + '_layer_norm.weight'] = (d_model,)
                s[pre + a + '_layer_norm.bias'] = (d_model,)
            s[pre + 'fc1.weight'] = (ffn, d_model)
            s[pre + 'fc1.bias'] = (ffn,)
            s[pre + 'fc2.weight'] = (d_model, ffn)
            s[pre + 'fc2.bias'] = (d_model,)
            s[pre + 'final_layer_norm.weight'] = (d_model,)
            s[pre + 'final_layer_norm.bias'] = (d_model,)
    if lm_heads:
        for i, n in enumerate(N_TOKENS):
            s['mask_lm.proj.%d.weight' % i] = (n, d_model)
            s['mask_lm.proj.%d.bias' % i] = (n,)
    return s


def gen_tensor(name, shape, seed):
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7fffffff)
    x = rs.standard_normal(int(np.prod(shape))).astype(np.float32).reshape(shape)
    if 'lut.weight' in name:
        return x * 0.5                      # reference: nn.Embedding N(0,1), then *16; keep activations O(10)
    if name.endswith('layer_norm.weight') or name.endswith('layernorm_embedding.weight'):
        return 1.0 + 0.1 * x
    if name.endswith('.bias'):
        return 0.05 * x
    if 'embed_positions' in name:
        return 0.05 * x
    if name.startswith('encoder_linear'):
        return x * (1.0 / 45.0)             # ~ default nn.Linear(2048, d) scale
    if name.startswith('mask_lm'):
        return x * 0.05
    return x * 0.04                         # BART linear weights (HF init_std 0.02; a bit larger to exercise softmax)


def make_params(d_model, enc_layers, dec_layers, ffn, max_pos, seed, lm_heads=True, extra=None):
    """dict name -> np.float32 array.  `decoder_linear.*` aliases `encoder_linear.*` (PianoBart.py:51-52)."""
    shapes = param_shapes(d_model, enc_layers, dec_layers, ffn, max_pos, lm_heads)
    if extra:
        shapes.update(extra)
    p = {k: gen_tensor(k, v, seed) for k, v in shapes.items()}
    p['decoder_linear.weight'] = p['encoder_linear.weight']
    p['decoder_linear.bias'] = p['encoder_linear.bias']
    return p


REAL = [256, 128, 129, 256, 128, 32, 254, 49]
PAD = [256, 128, 129, 256, 128, 32, 254, 49]
MASK = [x + 1 for x in PAD]
SOS = [x + 2 for x in PAD]
EOS = [x + 3 for x in PAD]


def synth_ids(batch, seq, seed, padded=False, min_len=None):
    """Synthetic Octuple ids per SURVEY.md section 8(d): column i ~ U{0..real_i-1}, bars ascending;
    padded variant: valid length L ~ U{seq/2..seq-1}, row L = EOS, rows after = PAD."""
    rs = np.random.RandomState(seed)
    ids = np.stack([rs.randint(0, REAL[i], size=(batch, seq)) for i in range(8)], axis=-1).astype(np.int64)
    ids[:, :, 0] = np.sort(ids[:, :, 0], axis=1)
    if padded:
        lo = seq // 2 if min_len is None else min_len
        for b in range(batch):
            L = int(rs.randint(lo, seq))
            ids[b, L] = EOS
            ids[b, L + 1:] = PAD
    return ids


def suppress_specials(p):
    """Make the LM heads never pick a special token (bias -30 on ids >= PAD) so that generation
    fixtures run for many steps instead of stopping at the first special id (model.py:63-64)."""
    for i in range(8):
        p['mask_lm.proj.%d.bias' % i] = p['mask_lm.proj.%d.bias' % i].copy()
        p['mask_lm.proj.%d.bias' % i][PAD[i]:] = -30.0
    return p
