"""Generation post-processing: the truncation rules of reference `Octuple2Midi` (demo.py:72-102) for a batch of
generated sequences on the device, then - on the host - the Octuple -> MIDI decoder and a Standard MIDI File writer
(codec.py; the reference uses `miditoolkit`, absent from this image) and the matching MIDI -> prompt direction
(`Midi2Octuple`, demo.py:60-67)."""
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


def octuple_to_midi(octuple, midi_paths):
    """demo.py:72-105 (`Octuple2Midi`) for a batch: truncate on the device, decode every non-empty sequence with
    codec.octuple_to_score and write it as a Standard MIDI File.  midi_paths: one path per sequence (a str for a single
    sequence).  Returns the list of written paths (None where the reference prints "Generate Fail! (empty)")."""
    from . import codec
    trunc, lens = octuple_truncate(octuple)
    if trunc.dim() == 2:
        trunc, midi_paths = trunc.unsqueeze(0), [midi_paths]
    rows, lens = trunc.cpu().numpy(), lens.cpu().numpy()
    written = []
    for b, path in enumerate(midi_paths):
        n = int(lens[b])
        if n == 0:
            written.append(None)
            continue
        codec.write_midi(codec.octuple_to_score(rows[b, :n].tolist()), path)
        written.append(path)
    return written


def midi_to_octuple(midi_path, device='cuda'):
    """demo.py:60-67 (`Midi2Octuple`): a MIDI file as a (1, 1024, 8) int32 prompt - its LAST 1023 notes + <EOS> when it is
    longer than the window, <PAD>-filled otherwise."""
    from . import codec
    rows = codec.score_to_octuple(codec.read_midi(midi_path), task='pretrain')
    rows = codec.pad_segment(rows, window=1024, last=True)
    return torch.tensor([rows], dtype=torch.int32, device=device)
