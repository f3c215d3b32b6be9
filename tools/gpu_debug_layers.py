"""Developer tool (GPU): per-op localisation of forward errors - every recorded buffer of the
encoder's first layers is recomputed with torch fp64 FROM THE PRECEDING BUFFER and compared."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from util import load_golden, build_cuda_model, golden_inputs

name = sys.argv[1] if len(sys.argv) > 1 else 'fwd_tiny'
dtype = sys.argv[2] if len(sys.argv) > 2 else 'fp32'
g = load_golden(name)
pb, lm = build_cuda_model(g['cfg'], int(g['seed']), dtype)
lm.eval()
enc, dec, ori, lmask, em, dm = golden_inputs(g)
with torch.no_grad():
    y = lm(enc, dec, em, dm)
torch.cuda.synchronize()
gr = pb._live_graph
d, H, F = gr.d, gr.H, gr.F
hd = d // H
B = gr.B
sd = {k: v.detach().double() for k, v in pb.named_parameters()}

def buf(nm, *shape):
    t = gr.ws.t[nm]
    n = int(np.prod(shape))
    return t[:n].view(*shape).double()

def cmp(tag, got, want):
    err = (got - want).abs().max().item(); sc = want.abs().max().item() + 1e-12
    print('%-28s rel err %.3e  (max |want| %.3e)' % (tag, err / sc, sc))

def ln(x, w, b):
    mu = x.mean(-1, keepdim=True); var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * w + b

for side, ids, keep, S, nl in (('encoder', enc, em, gr.Se, pb.layout.enc_layers), ('decoder', dec, dm, gr.Sd, pb.layout.dec_layers)):
    M = B * S
    X = torch.cat([sd['word_emb.%d.lut.weight' % i][ids[..., i]] * 16.0 for i in range(8)], -1).view(M, 2048)
    cmp(side + ' X', buf(side + '.X', M, 2048), X)
    pos = sd['bart.%s.embed_positions.weight' % side][2:2 + S]
    Y0 = buf(side + '.X', M, 2048) @ sd['encoder_linear.weight'].t() + sd['encoder_linear.bias']
    Y0 = (Y0.view(B, S, d) + pos).view(M, d)
    cmp(side + ' Y0', buf(side + '.Y0', M, d), Y0)
    H0 = ln(buf(side + '.Y0', M, d), sd['bart.%s.layernorm_embedding.weight' % side], sd['bart.%s.layernorm_embedding.bias' % side])
    cmp(side + ' H0', buf(side + '.H0', M, d), H0)
    h_in = buf(side + '.H0', M, d)
    enc_out_buf = None
    for l in range(nl):
        lp = 'bart.%s.layers.%d' % (side, l); L = '%s.L%d.' % (side, l)
        Wqkv = torch.cat([sd[lp + '.self_attn.%s_proj.weight' % n] for n in 'qkv'], 0)
        bqkv = torch.cat([sd[lp + '.self_attn.%s_proj.bias' % n] for n in 'qkv'], 0)
        cmp(L + 'QKV', buf(L + 'QKV', M, 3 * d), h_in @ Wqkv.t() + bqkv)
        QKV = buf(L + 'QKV', B, S, 3, H, hd)
        q, k, v = [QKV[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
        s = (q @ k.transpose(-1, -2)) * hd ** -0.5
        allow = (keep != 0)[:, None, None, :].expand(B, H, S, S)
        if side == 'decoder':
            allow = allow & torch.ones(S, S, dtype=torch.bool, device=s.device).tril()
        p = torch.softmax(s.masked_fill(~allow, float('-inf')), -1)
        cmp(L + 'P', buf(L + 'P', B, H, S, S), p)
        o = (buf(L + 'P', B, H, S, S) @ v).permute(0, 2, 1, 3).reshape(M, d)
        cmp(L + 'O', buf(L + 'O', M, d), o)
        A = buf(L + 'O', M, d) @ sd[lp + '.self_attn.out_proj.weight'].t() + sd[lp + '.self_attn.out_proj.bias'] + h_in
        cmp(L + 'A', buf(L + 'A', M, d), A)
        H1 = ln(buf(L + 'A', M, d), sd[lp + '.self_attn_layer_norm.weight'], sd[lp + '.self_attn_layer_norm.bias'])
        cmp(L + 'H1', buf(L + 'H1', M, d), H1)
        hm = buf(L + 'H1', M, d)
        if side == 'decoder':
            Se = gr.Se; Me = B * Se
            eo = gr.enc_out[:Me * d].view(Me, d).double()
            ca = lp + '.encoder_attn'
            cmp(L + 'Qc', buf(L + 'Qc', M, d), hm @ sd[ca + '.q_proj.weight'].t() + sd[ca + '.q_proj.bias'])
            Wkv = torch.cat([sd[ca + '.k_proj.weight'], sd[ca + '.v_proj.weight']], 0)
            bkv = torch.cat([sd[ca + '.k_proj.bias'], sd[ca + '.v_proj.bias']], 0)
            cmp(L + 'KVc', buf(L + 'KVc', Me, 2 * d), eo @ Wkv.t() + bkv)
            qc = buf(L + 'Qc', B, S, H, hd).permute(0, 2, 1, 3)
            KV = buf(L + 'KVc', B, Se, 2, H, hd)
            kc, vc = KV[:, :, 0].permute(0, 2, 1, 3), KV[:, :, 1].permute(0, 2, 1, 3)
            sc = (qc @ kc.transpose(-1, -2)) * hd ** -0.5
            pc = torch.softmax(sc.masked_fill(~(em != 0)[:, None, None, :].expand(B, H, S, Se), float('-inf')), -1)
            cmp(L + 'Pc', buf(L + 'Pc', B, H, S, Se), pc)
            oc = (buf(L + 'Pc', B, H, S, Se) @ vc).permute(0, 2, 1, 3).reshape(M, d)
            cmp(L + 'Oc', buf(L + 'Oc', M, d), oc)
            Ac = buf(L + 'Oc', M, d) @ sd[ca + '.out_proj.weight'].t() + sd[ca + '.out_proj.bias'] + hm
            cmp(L + 'Ac', buf(L + 'Ac', M, d), Ac)
            cmp(L + 'Hc', buf(L + 'Hc', M, d), ln(buf(L + 'Ac', M, d), sd[lp + '.encoder_attn_layer_norm.weight'], sd[lp + '.encoder_attn_layer_norm.bias']))
            hm = buf(L + 'Hc', M, d)
        Z = hm @ sd[lp + '.fc1.weight'].t() + sd[lp + '.fc1.bias']
        Zg = Z.detach().clone().requires_grad_(True)
        torch.nn.functional.gelu(Zg).sum().backward()
        cmp(L + 'Z (gelu\')', buf(L + 'Z', M, F), Zg.grad)      # the fc1 epilogue stores gelu'(pre-activation)
        cmp(L + 'G', buf(L + 'G', M, F), torch.nn.functional.gelu(Z))
        A2 = buf(L + 'G', M, F) @ sd[lp + '.fc2.weight'].t() + sd[lp + '.fc2.bias'] + hm
        cmp(L + 'A2', buf(L + 'A2', M, d), A2)
        cmp(L + 'Hn', buf(L + 'Hn', M, d), ln(buf(L + 'A2', M, d), sd[lp + '.final_layer_norm.weight'], sd[lp + '.final_layer_norm.bias']))
        h_in = buf(L + 'Hn', M, d)
print('err flag', gr.err_flag.item())
