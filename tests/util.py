"""Shared helpers for the tests: build the CUDA implementation with oracle.params weights."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import params as P


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def build_cuda_model(cfg, seed, dtype, lm=True, suppress_specials=False, device='cuda:0', dropout=None):
    from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
    from pianobart_b200.vocab import build_octuple_vocab
    d, el, dl, heads, ffn, max_pos = [int(x) for x in cfg]
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=max_pos, d_model=d, encoder_layers=el, decoder_layers=dl,
                    encoder_ffn_dim=ffn, decoder_ffn_dim=ffn, encoder_attention_heads=heads,
                    decoder_attention_heads=heads, vocab_size=64, **({} if dropout is None else {'dropout': dropout}))
    pb = PianoBart(bc, e2w, w2e, dtype=dtype)
    model = PianoBartLM(pb) if lm else pb
    prm = P.make_params(d, el, dl, ffn, max_pos, seed)
    if suppress_specials:
        P.suppress_specials(prm)
    sd = model.state_dict()
    for k, v in prm.items():
        kk = k if (k.startswith('mask_lm') or not lm) else 'pianobart.' + k
        if kk in sd:
            assert tuple(sd[kk].shape) == v.shape, (kk, sd[kk].shape, v.shape)
            sd[kk] = torch.from_numpy(v.copy())
    model.load_state_dict(sd)
    model.to(device)
    model.eval()   # parity fixtures are eval-mode arithmetic; dropout tests call .train() explicitly
    return pb, model


def golden_inputs(g, device='cuda:0'):
    enc = torch.from_numpy(g['enc'].astype(np.int64)).to(device)
    dec = torch.from_numpy(g['dec'].astype(np.int64)).to(device)
    ori = torch.from_numpy(g['ori'].astype(np.int64)).to(device)
    lm = torch.from_numpy(g['loss_mask'].astype(np.float32)).to(device)
    em = torch.from_numpy(g['enc_mask'].astype(np.float32)).to(device)
    dm = torch.from_numpy(g['dec_mask'].astype(np.float32)).to(device)
    return enc, dec, ori, lm, em, dm
