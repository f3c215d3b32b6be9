"""Per-hop clock64 trace of the persistent decode kernel (developer tool; csrc/decode_persist.cu `trace`)."""
import os
import sys
import time

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
    S = 1024
    gen = Generator(lm, 1, S, S)
    ids = torch.from_numpy(P.synth_ids(1, S, 4321)).to(dev)
    forced = torch.from_numpy(P.synth_ids(1, S, 99)).to(dev)
    gen.start(ids, torch.ones(1, S, device=dev), np.random.RandomState(0).random_sample((1, S, 8)), forced)
    gen.run_steps(int(os.environ.get('TRACE_AT', '500')))
    torch.cuda.synchronize()
    tr = torch.zeros(148, 6 * 96, dtype=torch.int64, device=dev)
    gen.pdesc.dbg_flags = int(os.environ.get('DBG_FLAGS', '0'))
    gen.pdesc.trace = E._ptr(tr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gen.run_steps(8)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = tr.cpu().numpy()
    np.save('gpurun_out/r2_decode_trace.npy', t)
    print('8 steps: %.1f us/step' % (ms * 1e3 / 8))
    names = ['L%d.%s' % (l, n) for l in range(8) for n in ('qkv', 'spart', 'wo', 'qc', 'cpart', 'woc', 'fc1', 'fc2')] + ['heads']
    full = [c for c in range(148) if int((t[c] != 0).sum()) == 396]
    for cta in (1, 0, 100):
        x = t[cta][:396]
        print('CTA %d: token span %d cycles; front %s' % (cta, x[-1] - x[0], np.diff(x[:6]).tolist()))
        s = x[6:].reshape(-1, 6)
        prev = np.concatenate([[x[5]], s[:-1, 5]])
        agg = {}
        for i in range(len(s)):
            nm = names[i].split('.')[-1]
            a = agg.setdefault(nm, np.zeros(7))
            a[0] += s[i, 0] - prev[i]
            a[1:6] += np.diff(s[i])
            a[6] += 1
        print('  hop      gap own_word vec+LN weights compute  store | period   (avg cycles)')
        for nm, a in agg.items():
            v = a[:6] / a[6]
            print('  %-6s' % nm, ' '.join('%7.0f' % q for q in v), '| %7.0f' % v.sum())
    S = np.stack([t[c][6:396].reshape(-1, 6) for c in full])
    own = S[:, :, 1] - S[:, :, 0]
    work = S[:, :, 5] - S[:, :, 1]
    print('per hop over %d CTAs (layer 1): own_word min / median / max | work-after-input median / max' % len(full))
    for h in range(8, 16):
        o, w = own[:, h], work[:, h]
        print('  %-9s %6d %6d %6d | %6d %6d' % (names[h], o.min(), int(np.median(o)), o.max(), int(np.median(w)), w.max()))


if __name__ == '__main__':
    main()
