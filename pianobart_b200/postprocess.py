"""Generation post-processing on the device: the truncation rules of reference `Octuple2Midi`
(demo.py:72-102) for a batch of generated sequences, up to (not including) the MIDI writer, which needs `miditoolkit` and is
out of scope."""
import ctypes as C

import torch

from . import _lib as L

PAD = [256, 128, 129, 256, 128, 32, 254, 49]


def octuple_truncate(octuple):
    """octuple: (B,S,8) or (S,8) integer CUDA tensor as returned by `PianoBartLM.forward(generate=True)`.
    Returns (truncated int64 tensor of the same shape, lengths int64 [B]): rows [0, length) are what the reference passes to
    `encoding_to_MIDI`; length 0 is the reference's "Generate Fail! (empty)"."""
    if not octuple.is_cuda:
        raise L.PBError('octuple_truncate runs on CUDA tensors only (no CPU path)')
    squeeze = octuple.dim() == 2
    x = octuple.unsqueeze(0) if squeeze else octuple
    if x.dtype not in (torch.int32, torch.int64):
        x = x.long()
    x = x.contiguous()
    B, S = x.shape[0], x.shape[1]
    out = torch.empty(B, S, 8, dtype=torch.int64, device=x.device)
    ln = torch.empty(B, dtype=torch.int64, device=x.device)
    pad = (C.c_int * 8)(*PAD)
    L.check(L.lib().pb_octuple_truncate(C.c_void_p(x.data_ptr()), 1 if x.dtype == torch.int64 else 0,
                                        C.c_void_p(out.data_ptr()), C.c_void_p(ln.data_ptr()), B, S, pad, L.stream_ptr()),
            'octuple_truncate')
    return (out[0], ln) if squeeze else (out, ln)
