"""Phase trace of the batch-64 persistent decode kernel (developer tool)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from oracle import params as P
    from pianobart_b200 import engine as E
    from pianobart_b200.generate import Generator
    from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
    from pianobart_b200.vocab import build_octuple_vocab
    c = bench.default_cfg()
    torch.manual_seed(2023)
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=c['max_pos'], d_model=c['d_model'], encoder_layers=c['layers'],
                    decoder_layers=c['layers'], encoder_ffn_dim=c['ffn'], decoder_ffn_dim=c['ffn'],
                    encoder_attention_heads=c['heads'], decoder_attention_heads=c['heads'])
    dev = torch.device('cuda', 0)
    pb = PianoBart(bc, e2w, w2e, dtype='bf16')
    lm = PianoBartLM(pb).to(dev)
    lm.eval()
    S, B = 1024, 64
    gen = Generator(lm, B, S, S)
    ids = torch.from_numpy(P.synth_ids(B, S, 4321)).to(dev)
    forced = torch.from_numpy(P.synth_ids(B, S, 99)).to(dev)
    gen.start(ids, torch.ones(B, S, device=dev), np.random.RandomState(0).random_sample((B, S, 8)), forced)
    gen.run_steps(int(os.environ.get('TRACE_AT', '100')))
    torch.cuda.synchronize()
    tr = torch.zeros(148, 160, dtype=torch.int64, device=dev)
    gen.pdesc.trace = E._ptr(tr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gen.run_steps(4)
    e1.record()
    torch.cuda.synchronize()
    print('4 steps: %.1f us/step' % (e0.elapsed_time(e1) * 1e3 / 4))
    t = tr.cpu().numpy()
    names = ['embed', 'in_lin'] + [n for l in range(8) for n in ('qkv', 'sattn', 'wo', 'qc', 'cattn', 'woc', 'fc1', 'fc2')] + ['heads', 'sample']
    for cta in (0, 70, 147):
        x = t[cta]
        n = int((x != 0).sum())
        x = x[:n].reshape(-1, 2)          # (arrive at barrier, leave barrier)
        work = x[:, 0] - np.concatenate([[x[0, 0]], x[:-1, 1]])
        wait = x[:, 1] - x[:, 0]
        agg = {}
        for i in range(len(x)):
            nm = names[i] if i < len(names) else 'x'
            a = agg.setdefault(nm, [0, 0, 0])
            a[0] += work[i]; a[1] += wait[i]; a[2] += 1
        print('CTA %d: %d barriers, span %d cycles' % (cta, len(x), x[-1, 1] - x[0, 0]))
        print('  phase     n   work  barrier_wait (avg cycles)')
        for nm, a in agg.items():
            print('  %-7s %3d %7.0f %7.0f' % (nm, a[2], a[0] / a[2], a[1] / a[2]))


if __name__ == '__main__':
    main()
