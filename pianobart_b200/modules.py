"""Host-side mirror of the reference's module API for the hot path.

Same class names, constructor arguments, forward signatures, attributes and state_dict keys as
the reference (`PianoBart` PianoBart.py:19-91, `PianoBartLM` / `MLM` model.py:14-126,
`SequenceClassification` model.py:165-218, `TokenClassification` model.py:236-272), but every
device computation of the backbone is a launch into libpianobart_b200.so (engine.py).
There is no PyTorch/CPU fallback: on a machine without the CUDA library forward() raises.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E

_DTYPES = {'fp32': E.PB_F32, 'bf16': E.PB_BF16}


class BartConfig:
    """Minimal stand-in for transformers.BartConfig (only the fields PianoBART reads:
    main.py:39-47).  A real transformers.BartConfig is accepted everywhere as well."""

    def __init__(self, max_position_embeddings=1024, d_model=1024, encoder_layers=12, decoder_layers=12,
                 encoder_ffn_dim=4096, decoder_ffn_dim=4096, encoder_attention_heads=16,
                 decoder_attention_heads=16, vocab_size=50265, pad_token_id=1, init_std=0.02, dropout=0.1, **kw):
        self.max_position_embeddings = max_position_embeddings
        self.d_model = d_model
        self.encoder_layers, self.decoder_layers = encoder_layers, decoder_layers
        self.encoder_ffn_dim, self.decoder_ffn_dim = encoder_ffn_dim, decoder_ffn_dim
        self.encoder_attention_heads, self.decoder_attention_heads = encoder_attention_heads, decoder_attention_heads
        self.vocab_size, self.pad_token_id, self.init_std, self.dropout = vocab_size, pad_token_id, init_std, dropout
        for k, v in kw.items():
            setattr(self, k, v)


class _Holder(nn.Module):
    """Anonymous container used to reproduce the reference's parameter names."""


def _register(root, dotted, param):
    parts = dotted.split('.')
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, _Holder())
        m = m._modules[p]
    m.register_parameter(parts[-1], param)


class ModelOutput:
    """Fields of HF Seq2SeqModelOutput / BaseModelOutput that the reference's callers read
    (model.py:121,207,265; PianoBart.py:130)."""

    def __init__(self, last_hidden_state, encoder_last_hidden_state=None):
        self.last_hidden_state = last_hidden_state
        self.encoder_last_hidden_state = encoder_last_hidden_state

    def __getitem__(self, i):
        return (self.last_hidden_state, self.encoder_last_hidden_state)[i]


class Embeddings(nn.Module):
    """PianoBart.py:9-16 (used by TokenClassification's replacement decoder front end)."""

    def __init__(self, n_token, d_model):
        super().__init__()
        self.lut = nn.Embedding(n_token, d_model)
        self.d_model = d_model

    def forward(self, x):
        from . import heads
        return heads.embed_rows(x, self.lut.weight, math.sqrt(self.d_model))     # pb_rows_gather / pb_rows_scatter_add


class _BackboneFn(torch.autograd.Function):
    """Autograd bridge: forward/backward are replays of the engine's launch plans; parameter
    gradients are accumulated by the kernels directly into the flat fp32 gradient buffer that
    backs every Parameter.grad."""

    @staticmethod
    def forward(ctx, anchor, dec_embeds, owner, graph, want_logits):
        ctx.owner, ctx.graph = owner, graph
        # the graph's activation buffers (and its dropout seed) belong to the MOST RECENT forward of this shape: remember
        # which forward this is so that backward can refuse to differentiate an overwritten one
        graph.fwd_generation = getattr(graph, 'fwd_generation', 0) + 1
        ctx.generation = graph.fwd_generation
        ctx.has_dec_embeds = dec_embeds is not None
        if dec_embeds is not None:
            n = dec_embeds.numel()
            graph.dec_in[:n].copy_(dec_embeds.reshape(-1))
        graph.forward()
        src = graph.logits if want_logits else graph.out
        width = E.VOCAB if want_logits else graph.d
        out = src[:graph.Mo * width].view(graph.B, -1, width).to(torch.float32, copy=True)
        enc = graph.enc_out[:graph.B * graph.Se * graph.d].view(graph.B, graph.Se, graph.d).to(torch.float32, copy=True)
        ctx.want_logits = want_logits
        ctx.mark_non_differentiable(enc)
        return out, enc

    @staticmethod
    def backward(ctx, g_out, g_enc):
        graph, owner = ctx.graph, ctx.owner
        if graph.bwd is None:
            raise RuntimeError('forward was run without gradient support (torch.no_grad)')
        if owner._live_graph is not graph or graph.fwd_generation != ctx.generation:
            raise RuntimeError('pianobart_b200: backward through a stale forward (another forward ran in between and '
                               'overwrote the saved activations); only the most recent forward can be differentiated - '
                               'call backward() before the next forward, or run the extra forward under torch.no_grad()')
        owner._prepare_grads()
        dst = graph.dlogits if ctx.want_logits else graph.d_out
        n = g_out.numel()
        dst[:n].copy_(g_out.reshape(-1))
        graph.backward()
        g_dec = None
        if ctx.has_dec_embeds:
            M = graph.B * graph.Sd
            g_dec = graph.d_dec_in[:M * graph.d].view(graph.B, graph.Sd, graph.d).to(torch.float32, copy=True)
        return None, g_dec, None, None, None


class PianoBart(nn.Module):
    """Drop-in for reference PianoBart (PianoBart.py:19-91)."""

    def __init__(self, bartConfig, e2w, w2e, dtype=None):
        super().__init__()
        self.hidden_size = bartConfig.d_model
        self.bartConfig = bartConfig
        self.n_tokens = []
        self.classes = ['Bar', 'Position', 'Instrument', 'Pitch', 'Duration', 'Velocity', 'TimeSig', 'Tempo']
        for key in self.classes:
            self.n_tokens.append(len(e2w[key]))
        if self.n_tokens != E.N_TOKENS:
            raise ValueError('pianobart_b200 kernels are specialised for the Octuple vocabulary sizes %s, got %s'
                             % (E.N_TOKENS, self.n_tokens))
        self.emb_sizes = [256] * 8
        self.e2w, self.w2e = e2w, w2e
        self.bar_pad_word = self.e2w['Bar']['Bar <PAD>']
        self.mask_word_np = np.array([self.e2w[t]['%s <MASK>' % t] for t in self.classes], dtype=np.int64)
        self.pad_word_np = np.array([self.e2w[t]['%s <PAD>' % t] for t in self.classes], dtype=np.int64)
        self.sos_word_np = np.array([self.e2w[t]['%s <SOS>' % t] for t in self.classes], dtype=np.int64)
        self.eos_word_np = np.array([self.e2w[t]['%s <EOS>' % t] for t in self.classes], dtype=np.int64)
        self.decoder_emb = None

        c = bartConfig
        if c.encoder_ffn_dim != c.decoder_ffn_dim or c.encoder_attention_heads != c.decoder_attention_heads:
            raise ValueError('encoder/decoder ffn and head counts must match (the reference always sets them equal)')
        self.heads = c.encoder_attention_heads
        self.layout = E.ParamLayout(c.d_model, c.encoder_layers, c.decoder_layers, c.encoder_ffn_dim,
                                    c.max_position_embeddings, with_heads=True)
        self.dtype_name = dtype or os.environ.get('PIANOBART_B200_DTYPE', 'bf16')
        self.pb_dtype = _DTYPES[self.dtype_name]

        # ---- parameters, reference names; storage re-pointed into flat buffers by _repack()
        std = getattr(c, 'init_std', 0.02)
        self._flat_names = []
        for name, (off, shape) in self.layout.entries.items():
            if name.startswith('mask_lm'):
                continue
            t = torch.empty(shape)
            if name.startswith('word_emb'):
                t.normal_(0.0, 1.0)                        # nn.Embedding default init
            elif name.startswith('encoder_linear'):
                t.uniform_(-1.0 / math.sqrt(2048), 1.0 / math.sqrt(2048))  # nn.Linear default init
            elif name.endswith('layer_norm.weight') or name.endswith('layernorm_embedding.weight'):
                t.fill_(1.0)
            elif name.endswith('.bias'):
                t.zero_()
            else:
                t.normal_(0.0, std)                        # HF BartPreTrainedModel._init_weights
            _register(self, name, nn.Parameter(t))
            self._flat_names.append(name)
        # same module registered twice in the reference (PianoBart.py:51-52) -> duplicate state_dict keys
        self.add_module('decoder_linear', self._modules['encoder_linear'])
        # bart.shared / embed_tokens: present in the reference checkpoint, never used with inputs_embeds
        shared = torch.empty(getattr(c, 'vocab_size', 50265), c.d_model).normal_(0.0, std)
        pad_id = getattr(c, 'pad_token_id', 1)
        if pad_id is not None and pad_id < shared.shape[0]:
            shared[pad_id].zero_()
        sp = nn.Parameter(shared)
        _register(self, 'bart.shared.weight', sp)
        _register(self, 'bart.encoder.embed_tokens.weight', sp)
        _register(self, 'bart.decoder.embed_tokens.weight', sp)

        self._flat = None       # fp32 master
        self._grad = None       # fp32 gradients
        self._wact = None       # working weights in activation dtype
        self._extra = {}        # name -> Parameter living in the flat buffer but owned by a wrapper (LM heads)
        self._graphs = {}
        self._live_graph = None
        self._wver = None
        self._anchor = None
        self._drop_seed = None
        self._pack_gen = 0      # bumped by every _repack(): objects holding raw pointers into the buffers check it

    # ------------------------------------------------------------------ flat storage
    def _named_flat_params(self):
        sd = dict(self.named_parameters())
        for n in self._flat_names:
            yield n, sd[n]
        for n, p in self._extra.items():
            yield n, p

    def _repack(self, device):
        """(Re)build the flat fp32 master/grad buffers on `device` and point every Parameter into them."""
        lay = self.layout
        flat = torch.zeros(lay.size, device=device, dtype=torch.float32)
        grad = torch.zeros(lay.size, device=device, dtype=torch.float32)
        for name, p in self._named_flat_params():
            off, shape = lay.entries[name]
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = flat[off:off + n].view(shape)
            p.grad = None
        self._flat, self._grad = flat, grad
        self._wact = torch.zeros(lay.size, device=device, dtype=torch.bfloat16 if self.pb_dtype == E.PB_BF16 else torch.float32)
        self._graphs.clear()
        self._live_graph = None
        self._wver = None
        self._anchor = torch.zeros(1, device=device, requires_grad=True)
        self._pack_gen += 1

    def check_pack_generation(self, gen, what):
        """Launch plans record raw device pointers into the flat buffers: an object built before the module was moved
        (.to / .cuda / .float) or before extra parameters were attached must not run on the old memory."""
        self._ensure_packed()
        if gen != self._pack_gen:
            raise L.PBError('%s was built for an earlier layout of the parameter buffers (the module was moved or re-packed '
                            'since); build a new one' % what)

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._flat = None  # storage moved: re-pack lazily on next use
        return r

    def _ensure_packed(self):
        dev = self.bart.shared.weight.device
        if dev.type != 'cuda':
            raise L.PBError('pianobart_b200 runs on CUDA devices only (sm_100a); move the module with .to("cuda") - '
                            'there is no CPU path')
        if self._flat is None or self._flat.device != dev:
            self._repack(dev)
        else:
            # a Parameter whose storage was replaced (e.g. load_state_dict(assign=True)) is copied back in
            for name, p in self._named_flat_params():
                off, shape = self.layout.entries[name]
                if p.data.data_ptr() != self._flat.data_ptr() + off * 4:
                    self._flat[off:off + p.numel()].copy_(p.data.reshape(-1))
                    p.data = self._flat[off:off + p.numel()].view(shape)
                    self._wver = None

    def attach_extra_params(self, named):
        """Let a wrapper (PianoBartLM) place its parameters (LM heads) in this module's flat buffers."""
        self._extra.update(named)
        self._flat = None

    def _weights_version(self):
        return sum(p._version for _, p in self._named_flat_params())

    def mark_weights_dirty(self):
        self._wver = None

    def _sync_weights(self):
        """Refresh the working-dtype copy of the weights (emb tables pre-scaled by sqrt(256)=16,
        PianoBart.py:16) when any fp32 master changed."""
        ver = self._weights_version()
        if self._wver == ver:
            return
        lib = L.lib()
        lay = self.layout
        s = L.stream_ptr()
        L.check(lib.pb_cast_from_f32(E.C.c_void_p(self._flat.data_ptr()), E.C.c_void_p(self._wact.data_ptr()),
                                     E.C.c_longlong(lay.emb_end), E.C.c_float(16.0), self.pb_dtype, s), 'cast emb')
        n = lay.size - lay.emb_end
        L.check(lib.pb_cast_from_f32(E.C.c_void_p(self._flat.data_ptr() + lay.emb_end * 4),
                                     E.C.c_void_p(self._wact.data_ptr() + lay.emb_end * self._wact.element_size()),
                                     E.C.c_longlong(n), E.C.c_float(1.0), self.pb_dtype, s), 'cast weights')
        self._wver = ver

    def _prepare_grads(self):
        """Before a backward replay: Parameter.grad views of the flat gradient buffer; a grad that is
        None (zero_grad(set_to_none=True)) is zeroed and re-attached."""
        for name, p in self._named_flat_params():
            off, shape = self.layout.entries[name]
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self._grad.data_ptr() + off * 4:
                view = self._grad[off:off + n].view(shape)
                if p.grad is None:
                    view.zero_()
                else:
                    view.copy_(p.grad)
                p.grad = view

    def flat_grad(self, name):
        """View of the flat fp32 gradient buffer for one reference-named parameter (fused trainer path, where
        Parameter.grad is not populated)."""
        off, shape = self.layout.entries[name]
        n = 1
        for x in shape:
            n *= x
        return self._grad[off:off + n].view(shape)

    def zero_grad_flat(self):
        self._ensure_packed()
        self._grad.zero_()

    # ------------------------------------------------------------------ graphs
    def dropout_p(self):
        """HF BartConfig.dropout (0.1 by default) is active in train() mode, exactly like the reference's BartModel."""
        return float(getattr(self.bartConfig, 'dropout', 0.0)) if self.training else 0.0

    def _graph(self, B, Se, Sd, with_heads, need_bwd, drop_p=None, custom_dec=False):
        drop_p = self.dropout_p() if drop_p is None else float(drop_p)
        key = (B, Se, Sd, with_heads, need_bwd, drop_p, custom_dec)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 3:
                self._graphs.clear()  # bound activation memory
                self._live_graph = None
            if self._drop_seed is None or self._drop_seed.device != self._flat.device:
                self._drop_seed = torch.tensor([torch.initial_seed() & 0x7fffffffffffffff], dtype=torch.int64,
                                               device=self._flat.device)
            g = E.BackboneGraph(self.layout, self.heads, self.pb_dtype, self._flat.device, B, Se, Sd, self._wact,
                                self._flat, self._grad, with_heads, need_backward=need_bwd, drop_p=drop_p,
                                drop_seed=self._drop_seed, dec_embed=custom_dec)
            self._graphs[key] = g
        return g

    def _run(self, input_ids_encoder, input_ids_decoder, encoder_attention_mask, decoder_attention_mask, want_logits):
        self._ensure_packed()
        dec_embeds = None
        if self.decoder_emb is not None and input_ids_decoder is not None:
            # PianoBart.py:63-66,71: decoder stream = decoder_linear(decoder_emb(ids)) (class_num x 64 table, 64 -> d
            # projection): gather kernel + GEMM (heads.py); the result enters the backbone plan as decoder input embeddings
            from . import heads
            lin = self.decoder_linear
            dec_embeds = heads.linear(self.decoder_emb(input_ids_decoder), lin.weight, lin.bias, self.pb_dtype)
        B, Se = input_ids_encoder.shape[0], input_ids_encoder.shape[1]
        Sd = 0 if input_ids_decoder is None else input_ids_decoder.shape[1]
        if max(Se, Sd) + 2 > self.layout.max_pos + 2:
            raise ValueError('sequence length exceeds max_position_embeddings')
        need_bwd = torch.is_grad_enabled()
        g = self._graph(B, Se, Sd, want_logits, need_bwd, custom_dec=dec_embeds is not None)
        self._sync_weights()
        g.set_inputs(input_ids_encoder, encoder_attention_mask,
                     None if dec_embeds is not None else input_ids_decoder, decoder_attention_mask)
        self._live_graph = g
        out, enc = _BackboneFn.apply(self._anchor, dec_embeds, self, g, want_logits)
        return out, enc

    def forward(self, input_ids_encoder, input_ids_decoder=None, encoder_attention_mask=None,
                decoder_attention_mask=None, output_hidden_states=True, generate=False):
        out, enc = self._run(input_ids_encoder, input_ids_decoder, encoder_attention_mask, decoder_attention_mask, False)
        return ModelOutput(out, enc)

    def get_rand_tok(self):
        """PianoBart.py:82-86 (consumes Python's `random` stream exactly like the reference)."""
        import random
        rand = [0] * 8
        for i in range(8):
            rand[i] = random.choice(range(self.n_tokens[i]))
        return np.array(rand)

    def change_decoder_embedding(self, new_embedding, new_linear=None):
        """PianoBart.py:88-91."""
        self.decoder_emb = new_embedding
        if new_linear is not None:
            self.decoder_linear = new_linear


class MLM(nn.Module):
    """model.py:109-126 - eight Linear heads `proj.i`; evaluated as one N=1280 GEMM inside the backbone plan."""

    def __init__(self, e2w, n_tokens, hidden_size):
        super().__init__()
        self.e2w = e2w
        self.proj = _Holder()
        bound = 1.0 / math.sqrt(hidden_size)
        for i in range(len(n_tokens)):
            h = _Holder()
            h.register_parameter('weight', nn.Parameter(torch.empty(n_tokens[i], hidden_size).uniform_(-bound, bound)))
            h.register_parameter('bias', nn.Parameter(torch.empty(n_tokens[i]).uniform_(-bound, bound)))
            self.proj.add_module(str(i), h)


class PianoBartLM(nn.Module):
    """Drop-in for reference PianoBartLM (model.py:14-78)."""

    def __init__(self, pianobart):
        super().__init__()
        self.pianobart = pianobart
        self.mask_lm = MLM(pianobart.e2w, pianobart.n_tokens, pianobart.hidden_size)
        extra = {}
        for i in range(8):
            extra['mask_lm.proj.%d.weight' % i] = self.mask_lm.proj._modules[str(i)].weight
            extra['mask_lm.proj.%d.bias' % i] = self.mask_lm.proj._modules[str(i)].bias
        pianobart.attach_extra_params(extra)
        self._generator = None

    def forward(self, input_ids_encoder, input_ids_decoder=None, encoder_attention_mask=None,
                decoder_attention_mask=None, generate=False, device_num=-1):
        if not generate:
            logits, _ = self.pianobart._run(input_ids_encoder, input_ids_decoder, encoder_attention_mask,
                                            decoder_attention_mask, True)
            return list(torch.split(logits, self.pianobart.n_tokens, dim=-1))
        from .generate import generate as _generate
        return _generate(self, input_ids_encoder, encoder_attention_mask)


class SelfAttention(nn.Module):
    """model.py:128-143: softmax over the sequence axis of ws2(tanh(ws1(h))), returned as [B, r, S]."""

    def __init__(self, input_dim, da, r):
        super().__init__()
        self.ws1 = nn.Linear(input_dim, da, bias=False)
        self.ws2 = nn.Linear(da, r, bias=False)

    def probs(self, h, pb_dtype):
        """[B, S, r] attention weights: tcgen05 GEMM (ws1), tanh folded into the r-output projection (ws2), sequence softmax"""
        from . import heads
        a1 = heads.linear(h, self.ws1.weight, None, pb_dtype)
        a2 = heads.linear(a1, self.ws2.weight, None, pb_dtype, act_in=heads.ACT_TANH)
        return heads.seq_softmax(a2)

    def forward(self, h, pb_dtype=None):
        return self.probs(h, E.PB_BF16 if pb_dtype is None else pb_dtype).permute(0, 2, 1)


def _classifier(seq, x, pb_dtype, training, seeds):
    """nn.Sequential(Dropout(0.1), Linear(k, 256), ReLU, Linear(256, class_num)) on the kernel path (model.py:173-178,
    244-249); the modules only hold the parameters (state_dict keys classifier.1.* / classifier.3.*)."""
    from . import heads
    drop, lin1, _, lin2 = seq[0], seq[1], seq[2], seq[3]
    if lin2.out_features > heads.SMALL_N:
        raise ValueError('pianobart_b200 classifier heads support class_num <= %d' % heads.SMALL_N)
    x = heads.dropout(x, drop.p, training, seeds)
    h1 = heads.linear(x, lin1.weight, lin1.bias, pb_dtype)
    return heads.linear(h1, lin2.weight, lin2.bias, pb_dtype, act_in=heads.ACT_RELU)   # ReLU folded into the operand load


class SequenceClassification(nn.Module):
    """model.py:165-218: backbone called with decoder ids = encoder ids (model.py:204), then the self-attentive pooled head
    (K15) - every op a library launch (heads.py, csrc/cls_heads.cu)."""

    def __init__(self, pianobart, class_num, hs, da=128, r=4):
        super().__init__()
        self.pianobart = pianobart
        self.attention = SelfAttention(hs, da, r)
        self.classifier = nn.Sequential(nn.Dropout(0.1), nn.Linear(hs * r, 256), nn.ReLU(), nn.Linear(256, class_num))
        self._seeds = None

    def forward(self, input_ids_encoder, encoder_attention_mask=None):
        from . import heads
        x = self.pianobart(input_ids_encoder=input_ids_encoder, input_ids_decoder=input_ids_encoder,
                           encoder_attention_mask=encoder_attention_mask,
                           decoder_attention_mask=encoder_attention_mask).last_hidden_state
        if self._seeds is None:
            self._seeds = heads.DropSeeds(x.device)
        dt = self.pianobart.pb_dtype
        m = heads.attn_pool(self.attention.probs(x, dt), x)          # == torch.bmm(attn_mat, x), [B, r, hs]
        return _classifier(self.classifier, m.reshape(m.shape[0], -1), dt, self.training, self._seeds)


class TokenClassification(nn.Module):
    """model.py:236-272 (K16)."""

    def __init__(self, pianobart, class_num, hs, d_model=64):
        super().__init__()
        self.pianobart = pianobart
        if class_num >= 5:
            self.pianobart.change_decoder_embedding(Embeddings(n_token=class_num, d_model=d_model),
                                                    nn.Linear(d_model, pianobart.bartConfig.d_model))
        self.classifier = nn.Sequential(nn.Dropout(0.1), nn.Linear(hs, 256), nn.ReLU(), nn.Linear(256, class_num))
        self._seeds = None

    def forward(self, input_ids_encoder, input_ids_decoder, encoder_attention_mask=None, decoder_attention_mask=None):
        from . import heads
        x = self.pianobart(input_ids_encoder, input_ids_decoder, encoder_attention_mask,
                           decoder_attention_mask).last_hidden_state
        if self._seeds is None:
            self._seeds = heads.DropSeeds(x.device)
        return _classifier(self.classifier, x, self.pianobart.pb_dtype, self.training, self._seeds)
