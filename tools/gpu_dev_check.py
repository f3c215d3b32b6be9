"""Developer bring-up script (GPU): module forward/backward vs the golden fixtures, verbose."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from util import load_golden, build_cuda_model, golden_inputs
from oracle import pianobart_oracle as O

def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))

def run(name, dtype):
    g = load_golden(name)
    pb, lm = build_cuda_model(g['cfg'], int(g['seed']), dtype)
    lm.eval()
    enc, dec, ori, lmask, em, dm = golden_inputs(g)
    t0 = time.time()
    y = lm(enc, dec, em, dm)
    torch.cuda.synchronize()
    logits = torch.cat(y, -1)
    total, losses = O.pretrain_loss(y, ori, lmask)
    print('[%s %s] fwd %.2fs loss %.6f ref %.6f rel %.2e' % (name, dtype, time.time() - t0, total.item(), float(g['total']),
          abs(total.item() - float(g['total'])) / float(g['total'])))
    if 'logits' in g.files:
        print('   logits rel err %.3e' % rel(logits.detach().cpu().numpy(), g['logits']))
    else:
        st = int(g['logit_stride'])
        print('   logits(sub) rel err %.3e' % rel(logits.detach().cpu().numpy()[:, ::st], g['logits_sub']))
    lm.zero_grad()
    total.backward()
    torch.cuda.synchronize()
    sd = dict(lm.named_parameters())
    worst = (0, '')
    names = [str(x) for x in g['grad_norm_names']]
    for n, v in zip(names, g['grad_norm_vals']):
        if n.startswith('decoder_linear'): continue
        kk = n if n.startswith('mask_lm') else 'pianobart.' + n
        p = sd[kk]
        gn = p.grad.double().norm().item() if p.grad is not None else float('nan')
        e = abs(gn - v) / (v + 1e-12)
        if not (e <= worst[0]): worst = (e, n + ' got %.4e want %.4e' % (gn, v))
    print('   worst grad-norm rel err %.3e (%s)' % worst)
    for k in g.files:
        if k.startswith('grad:'):
            n = k[5:]
            kk = n if n.startswith('mask_lm') else 'pianobart.' + n
            print('   grad %-55s rel err %.3e' % (n, rel(sd[kk].grad.cpu().numpy(), g[k])))

which = sys.argv[1:] or ['fwd_tiny', 'fwd_mid']
for nm in which:
    for dt in ('fp32', 'bf16'):
        try:
            run(nm, dt)
        except Exception as e:
            import traceback; traceback.print_exc()
