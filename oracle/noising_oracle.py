"""ORACLE - test infrastructure only (never imported by the product package).

CPU restatement of the BART noising of the PianoBART pretrainer, reference
pretrain.py:211-546 (`Pretrainer.gen_mask`, dispatch :519-546) and of the batch prologue
pretrain.py:128-153 (decoder shift-right, loss-mask broadcast, attention masks).

It consumes the SAME two random streams as the reference, in the same order - Python's
`random` module (random.randint / shuffle / sample / choice / random) and numpy's global
legacy RandomState (np.random.poisson) - so that after `random.seed(s); np.random.seed(s)`
it reproduces the reference bit for bit, including the CPython set-iteration order that
pretrain.py:281,283,385 depend on.  The arithmetic is restated with numpy index operations
instead of the reference's np.delete / torch.cat loops.

Pinned against the real reference by tests/golden/noising_*.npz (tools/make_golden.py).
Only the five corruptions reachable under the flags pinned at pretrain.py:530-531,542
(octuple-level mask, octuple-level infilling) are restated; the bar-/element-level branches
are dead code in the reference.
"""
import random

import numpy as np

PAD = np.array([256, 128, 129, 256, 128, 32, 254, 49], dtype=np.int64)
MASK = PAD + 1
SOS = PAD + 2
N_TOKENS = [262, 134, 135, 262, 134, 38, 260, 55]


def token_deletion(ids, mask_percent=0.15):
    """pretrain.py:218-236."""
    S = ids.shape[0]
    n_del = int(S * mask_percent)
    flags = [1 if i < n_del else 0 for i in range(S)]
    random.shuffle(flags)
    flags = np.array(flags)
    out = np.concatenate([ids[flags == 0], np.tile(PAD, (n_del, 1))], axis=0)
    loss = flags.copy()
    hit = np.where(flags == 1)[0]
    if len(hit) > 0:
        loss[hit[0]:] = 1
    return out.astype(np.int64), loss.astype(np.float32)


def get_rand_tok():
    """PianoBart.py:82-86 - eight random.choice draws over the FULL vocab (specials included)."""
    return np.array([random.choice(range(N_TOKENS[i])) for i in range(8)], dtype=np.int64)


def token_mask(ids, max_seq_len, mask_percent=0.15):
    """pretrain.py:277-295 (n=0, octuple level).  Index lists range over max_seq_len."""
    lseq = list(range(max_seq_len))
    mask_ind = random.sample(lseq, round(max_seq_len * mask_percent))
    mask80 = random.sample(mask_ind, round(len(mask_ind) * 0.8))
    left = list(set(mask_ind) - set(mask80))           # CPython set order matters for the next draw
    rand10 = random.sample(left, round(len(mask_ind) * 0.1))
    out = ids.copy()
    loss = np.zeros(max_seq_len, dtype=np.float32)
    out[mask80] = MASK
    for i in rand10:
        out[i] = get_rand_tok()
    loss[mask_ind] = 1.0
    return out, loss


def sentence_permutation(ids):
    """pretrain.py:368-397 - shuffle bars (groups of equal column-0 value)."""
    bars = [int(b) for b in ids[:, 0]]
    order = list(set(bars))                            # CPython set order of small ints
    random.shuffle(order)
    groups = {}
    for r, b in enumerate(bars):
        groups.setdefault(b, []).append(r)
    src = [r for b in order for r in groups[b]]
    out = ids[src]
    loss = (out != ids).any(axis=1).astype(np.float32)
    return out, loss


def token_infilling(ids, mask_percent=0.15, lamda=3):
    """pretrain.py:402-436 (n=0).  Returns (rows, loss, failed)."""
    S = ids.shape[0]
    thr = mask_percent / max(1, lamda)
    for attempt in range(10):
        src = []  # >=0: source row, -1: MASK row
        i = 0
        while i < S:
            if random.random() < thr:
                p = np.random.poisson(lamda)
                if p == 0:
                    src.append(i)
                    src.append(-1)
                    i += 1
                else:
                    src.append(-1)
                    i += p
            else:
                src.append(i)
                i += 1
        if len(src) <= S:
            break
        if attempt >= 9:
            # pretrain.py:429-430 - unchanged input and an all-zero (S,8) loss mask
            return ids.copy(), np.zeros((S, 8), dtype=np.float32), True
    src = np.array(src + [-2] * (S - len(src)), dtype=np.int64)  # -2: PAD row
    out = np.where((src >= 0)[:, None], ids[np.maximum(src, 0)], np.where((src == -1)[:, None], MASK, PAD))
    loss = (out != ids).any(axis=1).astype(np.float32)
    return out.astype(np.int64), loss, False


def document_rotation(ids):
    """pretrain.py:508-517."""
    S = ids.shape[0]
    r = random.randint(0, S - 1)
    out = np.concatenate([ids[r:], ids[:r]], axis=0)
    loss = np.full(S, 1.0 if r != 0 else 0.0, dtype=np.float32)
    return out, loss


def gen_mask(ids, max_seq_len, mask_percent=0.15, choice=None):
    """pretrain.py:519-546.  ids: (S,8) int64.  Returns (noised ids (S,8) int64, loss mask (S,8) float32,
    choice)."""
    if choice is None:
        choice = random.randint(1, 5)
    if choice == 1:
        out, loss = token_deletion(ids, mask_percent)
    elif choice == 2:
        out, loss = token_mask(ids, max_seq_len, mask_percent)
    elif choice == 3:
        out, loss = sentence_permutation(ids)
    elif choice == 4:
        out, loss, failed = token_infilling(ids, mask_percent)
        if failed:
            return out, loss, choice
    else:
        out, loss = document_rotation(ids)
    # pretrain.py:141-142: a 1-D mask is repeated over the 8 attributes
    return out, np.repeat(loss[:, None], 8, axis=1), choice


def noise_batch(ori, max_seq_len, mask_percent=0.15, bar_pad=256):
    """pretrain.py:128-153.  ori: (B,S,8) int64.
    Returns dict(enc, dec, loss_mask, enc_mask, dec_mask, choices)."""
    B, S, _ = ori.shape
    enc = np.empty_like(ori)
    dec = np.empty_like(ori)
    loss_mask = np.zeros((B, S, 8), dtype=np.float32)
    choices = []
    for b in range(B):
        dec[b, 1:] = ori[b, :-1]
        dec[b, 0] = SOS
        enc[b], loss_mask[b], c = gen_mask(ori[b], max_seq_len, mask_percent)
        choices.append(c)
    enc_mask = (enc[:, :, 0] != bar_pad).astype(np.float32)   # computed AFTER noising (pretrain.py:151)
    dec_mask = (dec[:, :, 0] != bar_pad).astype(np.float32)
    return dict(enc=enc, dec=dec, loss_mask=loss_mask, enc_mask=enc_mask, dec_mask=dec_mask,
                choices=np.array(choices, dtype=np.int64))
