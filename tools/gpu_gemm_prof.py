"""Per-GEMM breakdown of one pretraining step (B=16 x S=1024, bf16): every GEMM launch of the plan is bracketed by
CUDA events (engine.Plan.run(profile=...)) and aggregated by call site.  Run on a B200: python tools/gpu_gemm_prof.py"""
import os
import re
import sys
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from oracle import params as P
    from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
    from pianobart_b200.pretrain import FusedAdamW, PretrainStep
    from pianobart_b200.vocab import build_octuple_vocab
    import random
    B = int(os.environ.get('B', 16))
    S = 1024
    torch.manual_seed(0)
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=1024, d_model=1024, encoder_layers=8, decoder_layers=8, encoder_ffn_dim=2048,
                    decoder_ffn_dim=2048, encoder_attention_heads=8, decoder_attention_heads=8)
    dev = torch.device('cuda', 0)
    pb = PianoBart(bc, e2w, w2e, dtype='bf16')
    lm = PianoBartLM(pb).to(dev)
    lm.train()
    opt = FusedAdamW(pb, lr=2e-5, weight_decay=0.01)
    step = PretrainStep(lm, B, S, opt, 0.15, None)
    random.seed(1)
    np.random.seed(1)
    step.upload(P.synth_ids(B, S, 7))
    for _ in range(3):
        step.noise(); step.run(train=True)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for rep in range(3):
        prof = []
        step.noise()
        step.run(train=True, profile=prof)
        torch.cuda.synchronize()
        for name, flops, e0, e1 in prof:
            key = re.sub(r'L\d+', 'L#', name)
            a = agg[key]
            a[0] += 1; a[1] += flops; a[2] += e0.elapsed_time(e1)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][2])
    tot_ms = sum(v[2] for v in agg.values()) / 3
    tot_fl = sum(v[1] for v in agg.values()) / 3
    print('%-28s %5s %9s %9s %7s' % ('site', 'n', 'ms/step', 'TFLOP/s', 'share'))
    for k, (n, fl, ms) in rows:
        print('%-28s %5d %9.3f %9.1f %6.1f%%' % (k, n // 3, ms / 3, fl / ms / 1e9, 100 * ms / 3 / tot_ms))
    print('total GEMM %.3f ms/step, %.1f TFLOP/s average' % (tot_ms, tot_fl / tot_ms / 1e9))


if __name__ == '__main__':
    main()
