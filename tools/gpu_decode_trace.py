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
    tr = torch.zeros(148, 4 * 96, dtype=torch.int64, device=dev)
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
    names = ['front'] + ['L%d.%s' % (l, n) for l in range(8) for n in ('qkv', 'spart', 'wo', 'qc', 'cpart', 'woc', 'fc1', 'fc2')] + ['heads']
    for cta in (1, 0, 100):
        x = t[cta]
        n = int((x != 0).sum())
        x = x[:n]
        print('CTA %d: %d stamps, token span %d cycles' % (cta, n, x[-1] - x[0]))
        if (n - 1) % 4 != 0:
            continue
        hops = (n - 1) // 4
        s = x[1:].reshape(hops, 4)
        prev = np.concatenate([[x[0]], s[:-1, 3]])
        agg = {}
        for i in range(hops):
            nm = names[i].split('.')[-1] if i < len(names) else 'x'
            a = agg.setdefault(nm, [0, 0, 0, 0, 0])
            a[0] += s[i, 0] - prev[i]; a[1] += s[i, 1] - s[i, 0]; a[2] += s[i, 2] - s[i, 1]; a[3] += s[i, 3] - s[i, 2]; a[4] += 1
        print('  hop     n  wait_input  wait_weights  compute  store   (avg cycles)')
        for nm, a in agg.items():
            print('  %-6s %2d  %10.0f  %12.0f  %7.0f  %5.0f' % (nm, a[4], a[0] / a[4], a[1] / a[4], a[2] / a[4], a[3] / a[4]))
        tot = [sum(a[k] for a in agg.values()) for k in range(4)]
        print('  total cycles: wait_input %d  wait_weights %d  compute %d  store %d' % tuple(tot))


if __name__ == '__main__':
    main()
