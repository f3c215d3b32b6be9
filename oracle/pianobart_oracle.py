"""ORACLE - test infrastructure only (never imported by the product package).

CPU restatement of the PianoBART forward / loss arithmetic, written directly from the
reference's call sites and the published BART algorithm, independent of both the
reference's nn.Modules and of HuggingFace `transformers`:

  * Octuple front end          reference PianoBart.py:9-16 (Embeddings: lut(x)*sqrt(256)),
                               PianoBart.py:60-71 (8 gathers, cat, encoder_linear shared by both streams)
  * learned positions          HF transformers modeling_bart.py BartLearnedPositionalEmbedding (offset 2)
  * encoder / decoder stacks   HF modeling_bart.py BartEncoder.forward / BartDecoder.forward
                               (x + pos -> layernorm_embedding -> layers), post-LN BartEncoderLayer /
                               BartDecoderLayer, BartAttention (q*hd^-0.5, additive key mask, softmax)
                               `transformers` is a third-party dependency of the reference
                               (environment.yml:197 pins 4.29.2; 5.5.0 is installed here) - its BART
                               algorithm is restated, anchored on the reference call PianoBart.py:76.
  * LM heads                   reference model.py:109-126
  * pretrain loss              reference pretrain.py:112-118 (masked CE) and :179-189 (weights read in
                               e2w *pickle key order* while losses are in `classes` order - reproduced)
  * generation-finetune loss   reference finetune_generation.py:238-250
  * sequence / token heads     reference model.py:128-143,165-218,236-272

Parity pinning: the reference ships no tests or golden vectors for this path ("parity
unpinned" by the reference itself).  This restatement is pinned instead against outputs of
the real reference executed in the build container (tools/make_golden.py imports
/root/reference and writes tests/golden/*.npz; tests/test_oracle_golden.py checks them).

Everything is plain torch CPU tensor arithmetic (matmul / softmax written out), in the dtype
of the parameters (fp32, or fp64 for tight gradient references); autograd of this restatement
provides gradient references.
"""
import math

import torch

CLASSES = ['Bar', 'Position', 'Instrument', 'Pitch', 'Duration', 'Velocity', 'TimeSig', 'Tempo']
# pickle key order of Data/Octuple.pkl (make_dict.py:28) - the order pretrain.py:184-189 reads n_tok in
E2W_KEY_ORDER = ['Bar', 'Position', 'Pitch', 'Duration', 'Velocity', 'Instrument', 'Tempo', 'TimeSig']
N_TOKENS = [262, 134, 135, 262, 134, 38, 260, 55]          # classes order (PianoBart.py:29-31)
N_TOKENS_KEY_ORDER = [262, 134, 262, 134, 38, 135, 55, 260]  # what multiplies loss i in pretrain.py:188
EMB = 256


class Cfg:
    def __init__(self, d_model=1024, enc_layers=8, dec_layers=8, heads=8, ffn=2048, max_pos=1024):
        self.d_model, self.enc_layers, self.dec_layers = d_model, enc_layers, dec_layers
        self.heads, self.ffn, self.max_pos = heads, ffn, max_pos


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def octuple_embed(p, ids, prefix='word_emb'):
    """PianoBart.py:60-67 - eight table gathers scaled by sqrt(256), concatenated."""
    embs = [p['%s.%d.lut.weight' % (prefix, i)][ids[..., i]] * math.sqrt(EMB) for i in range(8)]
    return torch.cat(embs, dim=-1)


def attention(p, pre, x_q, x_kv, heads, key_keep, causal):
    """HF BartAttention: q,k,v,out projections with bias, scores scaled by head_dim^-0.5.
    key_keep: (B, S_k) bool, True = key may be attended.  causal: query i sees keys j <= i."""
    B, Sq, d = x_q.shape
    Sk = x_kv.shape[1]
    hd = d // heads
    q = linear(x_q, p[pre + '.q_proj.weight'], p[pre + '.q_proj.bias']).view(B, Sq, heads, hd).transpose(1, 2)
    k = linear(x_kv, p[pre + '.k_proj.weight'], p[pre + '.k_proj.bias']).view(B, Sk, heads, hd).transpose(1, 2)
    v = linear(x_kv, p[pre + '.v_proj.weight'], p[pre + '.v_proj.bias']).view(B, Sk, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    allow = torch.ones(B, 1, Sq, Sk, dtype=torch.bool)
    if key_keep is not None:
        allow = allow & key_keep[:, None, None, :]
    if causal:
        allow = allow & torch.ones(Sq, Sk, dtype=torch.bool).tril()[None, None]
    s = s.masked_fill(~allow, float('-inf'))
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Sq, d)
    return linear(o, p[pre + '.out_proj.weight'], p[pre + '.out_proj.bias'])


def _drop(x, masks, key):
    """Training-mode dropout with an explicit, already scaled mask (mask = keep / (1 - p)).  HF applies
    nn.functional.dropout(p=config.dropout) at exactly these sites (Bart{Encoder,Decoder}.forward after
    layernorm_embedding; Bart*Layer.forward after each attention block and after fc2)."""
    if masks is None or key not in masks:
        return x
    return x * masks[key].to(x.dtype)


def encoder(p, cfg, x, keep, masks=None):
    """HF BartEncoder.forward: embeds + positions(offset 2) -> LN -> post-LN layers."""
    S = x.shape[1]
    h = x + p['bart.encoder.embed_positions.weight'][2:2 + S]
    h = layer_norm(h, p['bart.encoder.layernorm_embedding.weight'], p['bart.encoder.layernorm_embedding.bias'])
    h = _drop(h, masks, ('encoder', -1, 0))
    for l in range(cfg.enc_layers):
        pre = 'bart.encoder.layers.%d' % l
        a = _drop(attention(p, pre + '.self_attn', h, h, cfg.heads, keep, False), masks, ('encoder', l, 1))
        h = layer_norm(h + a, p[pre + '.self_attn_layer_norm.weight'], p[pre + '.self_attn_layer_norm.bias'])
        f = linear(gelu_erf(linear(h, p[pre + '.fc1.weight'], p[pre + '.fc1.bias'])), p[pre + '.fc2.weight'],
                   p[pre + '.fc2.bias'])
        f = _drop(f, masks, ('encoder', l, 3))
        h = layer_norm(h + f, p[pre + '.final_layer_norm.weight'], p[pre + '.final_layer_norm.bias'])
    return h


def decoder(p, cfg, y, enc_out, enc_keep, dec_keep, masks=None):
    """HF BartDecoder.forward: causal self-attention, cross-attention on the encoder output."""
    S = y.shape[1]
    h = y + p['bart.decoder.embed_positions.weight'][2:2 + S]
    h = layer_norm(h, p['bart.decoder.layernorm_embedding.weight'], p['bart.decoder.layernorm_embedding.bias'])
    h = _drop(h, masks, ('decoder', -1, 0))
    for l in range(cfg.dec_layers):
        pre = 'bart.decoder.layers.%d' % l
        a = _drop(attention(p, pre + '.self_attn', h, h, cfg.heads, dec_keep, True), masks, ('decoder', l, 1))
        h = layer_norm(h + a, p[pre + '.self_attn_layer_norm.weight'], p[pre + '.self_attn_layer_norm.bias'])
        c = _drop(attention(p, pre + '.encoder_attn', h, enc_out, cfg.heads, enc_keep, False), masks, ('decoder', l, 2))
        h = layer_norm(h + c, p[pre + '.encoder_attn_layer_norm.weight'], p[pre + '.encoder_attn_layer_norm.bias'])
        f = linear(gelu_erf(linear(h, p[pre + '.fc1.weight'], p[pre + '.fc1.bias'])), p[pre + '.fc2.weight'],
                   p[pre + '.fc2.bias'])
        f = _drop(f, masks, ('decoder', l, 3))
        h = layer_norm(h + f, p[pre + '.final_layer_norm.weight'], p[pre + '.final_layer_norm.bias'])
    return h


def pianobart_forward(p, cfg, enc_ids, dec_ids=None, enc_mask=None, dec_mask=None, dec_embeds=None, masks=None):
    """PianoBart.forward (PianoBart.py:56-80).  Returns (last_hidden_state, encoder_last_hidden_state).
    Masks are (B,S) with non-zero = keep (pretrain.py:151-153 builds them as float 0/1).
    dec_embeds: pre-computed decoder input embeddings (the change_decoder_embedding path)."""
    enc_keep = None if enc_mask is None else (enc_mask != 0)
    dec_keep = None if dec_mask is None else (dec_mask != 0)
    x = linear(octuple_embed(p, enc_ids), p['encoder_linear.weight'], p['encoder_linear.bias'])
    enc_out = encoder(p, cfg, x, enc_keep, masks)
    if dec_ids is None and dec_embeds is None:
        return enc_out, enc_out
    if dec_embeds is None:
        y = linear(octuple_embed(p, dec_ids), p['decoder_linear.weight'], p['decoder_linear.bias'])
    else:
        y = dec_embeds
    dec_out = decoder(p, cfg, y, enc_out, enc_keep, dec_keep, masks)
    return dec_out, enc_out


def lm_heads(p, h, prefix='mask_lm.proj'):
    """MLM.forward (model.py:119-126): list of 8 logits in `classes` order."""
    return [linear(h, p['%s.%d.weight' % (prefix, i)], p['%s.%d.bias' % (prefix, i)]) for i in range(8)]


def masked_ce(logits, target, mask):
    """Pretrainer.compute_loss (pretrain.py:112-118): sum(CE * mask) / sum(mask)."""
    lse = torch.logsumexp(logits, dim=-1)
    picked = logits.gather(-1, target[..., None]).squeeze(-1)
    return ((lse - picked) * mask).sum() / mask.sum()


def pretrain_loss(logits, targets, loss_mask):
    """pretrain.py:179-189: per-attribute masked CE, weighted by n_tok read in e2w key order."""
    losses = [masked_ce(logits[i], targets[..., i], loss_mask[..., i]) for i in range(8)]
    total = sum(l * w for l, w in zip(losses, N_TOKENS_KEY_ORDER)) / sum(N_TOKENS_KEY_ORDER)
    return total, losses


def pretrain_accuracy(logits, targets, loss_mask):
    """pretrain.py:163-176: argmax accuracy over masked positions, per attribute."""
    accs = []
    for i in range(8):
        pred = logits[i].argmax(-1)
        accs.append(((pred == targets[..., i]).to(loss_mask.dtype) * loss_mask[..., i]).sum() / loss_mask[..., i].sum())
    return accs


GEN_EXTRA_W = [1.0, 1.0, 0.3, 1.5, 1.0, 1.0, 0.3, 0.3]


def generation_finetune_loss(logits, targets, dec_keep_mask):
    """finetune_generation.py:238-250: same masked CE with extra per-head factors, mask = decoder non-pad."""
    losses = [masked_ce(logits[i], targets[..., i], dec_keep_mask) for i in range(8)]
    ws = [n * e for n, e in zip(N_TOKENS_KEY_ORDER, GEN_EXTRA_W)]
    # the reference multiplies the extra factor into the loss and still divides by sum(n_tok)
    total = sum(l * w for l, w in zip(losses, ws)) / sum(N_TOKENS_KEY_ORDER)
    return total, losses


def shift_right(ids, sos):
    """pretrain.py:132-139: dec[b,0]=SOS, dec[b,1:]=orig[b,:-1]."""
    out = torch.empty_like(ids)
    out[:, 1:] = ids[:, :-1]
    out[:, 0] = torch.as_tensor(sos, dtype=ids.dtype)
    return out


def seq_cls_head(p, h, prefix=''):
    """model.py:140-143,214-217: softmax over the *sequence* dim (no pad mask), r=4 aspects."""
    a = torch.tanh(h @ p[prefix + 'attention.ws1.weight'].t()) @ p[prefix + 'attention.ws2.weight'].t()  # (B,S,r)
    a = torch.softmax(a, dim=1).permute(0, 2, 1)
    m = torch.bmm(a, h).reshape(h.shape[0], -1)
    z = torch.relu(linear(m, p[prefix + 'classifier.1.weight'], p[prefix + 'classifier.1.bias']))
    return linear(z, p[prefix + 'classifier.3.weight'], p[prefix + 'classifier.3.bias'])


def token_cls_head(p, h, prefix=''):
    """model.py:247-253,271."""
    z = torch.relu(linear(h, p[prefix + 'classifier.1.weight'], p[prefix + 'classifier.1.bias']))
    return linear(z, p[prefix + 'classifier.3.weight'], p[prefix + 'classifier.3.bias'])


# ----------------------------------------------------------------------------- sampling
SAMPLE_T = [1.2, 1.2, 5, 1, 2, 5, 5, 1.2]   # model.py:70
SAMPLE_P = [1, 1, 1, 0.9, 0.9, 1, 1, 0.9]   # model.py:71


def nucleus_candidates(probs, p):
    """model.py:84-96 up to (not including) the random draw: returns (candidate ids, normalised probs).
    numpy semantics (float32 probs, argsort descending via [::-1])."""
    import numpy as np
    probs = probs / (sum(probs) + 1e-5)
    sorted_probs = np.sort(probs)[::-1]
    sorted_index = np.argsort(probs)[::-1]
    cus = np.cumsum(sorted_probs)
    after = cus > p
    if sum(after) > 0:
        last = np.where(after)[0][0] + 1
        cand = sorted_index[:last]
    else:
        cand = sorted_index[0:1]
    cp = np.array([probs[i] for i in cand])
    cp = cp / sum(cp)
    return cand, cp


def sample_step(logits_at_step, rng):
    """PianoBartLM.sample (model.py:68-78) for one position: logits_at_step = list of 8 1-D tensors.
    rng: numpy RandomState-like with .choice (the reference uses the global np.random)."""
    out = []
    for j in range(8):
        probs = torch.softmax(logits_at_step[j] / SAMPLE_T[j], dim=-1).detach().numpy()
        cand, cp = nucleus_candidates(probs, SAMPLE_P[j])
        out.append(int(rng.choice(cand, size=1, p=cp)[0]))
    return out
