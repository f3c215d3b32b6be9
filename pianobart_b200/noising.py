"""Host half of the BART noising (reference pretrain.py:211-546): draws the random decisions
and emits a compact per-row plan; the device half (csrc/noise.cu) moves the data.

Bit-exactness contract: after `random.seed(s); np.random.seed(s)` the plan reproduces the
reference's corruption exactly, because the decisions are drawn with the very primitives the
reference calls, in the same order - Python `random` (randint / shuffle / sample / choice /
random) and numpy's global legacy generator (np.random.poisson) - including the CPython
set-iteration order that pretrain.py:281,283,385 depend on.

Plan row codes (int32 per output row): >= 0 source row; -1 PAD row; -2 MASK row;
<= -3 random-token row (-3 - k) of the side table.  loss_mode per sample: 0 = host flags,
1 = "row changed" (device compare), 2 = all zero.
"""
import random

import numpy as np

N_TOKENS = [262, 134, 135, 262, 134, 38, 260, 55]


class NoisePlan:
    def __init__(self, B, S):
        self.src = np.empty((B, S), dtype=np.int32)
        self.loss = np.zeros((B, S), dtype=np.uint8)
        self.loss_mode = np.zeros(B, dtype=np.int32)
        self.rand_tok = []
        self.choices = np.zeros(B, dtype=np.int64)


def _deletion(S, mask_percent, src, loss):
    """pretrain.py:218-236."""
    n_del = int(S * mask_percent)
    flags = [1 if i < n_del else 0 for i in range(S)]
    random.shuffle(flags)
    flags = np.asarray(flags)
    keep = np.flatnonzero(flags == 0)
    src[:len(keep)] = keep
    src[len(keep):] = -1
    hit = np.flatnonzero(flags)
    if len(hit):
        loss[hit[0]:] = 1


def _token_mask(S, max_seq_len, mask_percent, src, loss, rand_tok):
    """pretrain.py:277-295 (octuple level)."""
    mask_ind = random.sample(list(range(max_seq_len)), round(max_seq_len * mask_percent))
    mask80 = random.sample(mask_ind, round(len(mask_ind) * 0.8))
    left = list(set(mask_ind) - set(mask80))
    rand10 = random.sample(left, round(len(mask_ind) * 0.1))
    src[:] = np.arange(S, dtype=np.int32)
    src[mask80] = -2
    for i in rand10:
        # PianoBart.get_rand_tok (PianoBart.py:82-86): 8 x random.choice over the full vocab
        rand_tok.append([random.choice(range(N_TOKENS[a])) for a in range(8)])
        src[i] = -3 - (len(rand_tok) - 1)
    loss[mask_ind] = 1


def _permutation(bars, src):
    """pretrain.py:368-397.  bars: python ints of column 0."""
    order = list(set(bars))
    random.shuffle(order)
    groups = {}
    for r, b in enumerate(bars):
        groups.setdefault(b, []).append(r)
    k = 0
    for b in order:
        g = groups[b]
        src[k:k + len(g)] = g
        k += len(g)


def _infilling(S, mask_percent, src, lamda=3):
    """pretrain.py:402-436 (octuple level).  Returns False on the 10-failures branch."""
    thr = mask_percent / max(1, lamda)
    rnd = random.random
    for attempt in range(10):
        out = []
        i = 0
        while i < S:
            if rnd() < thr:
                p = np.random.poisson(lamda)
                if p == 0:
                    out.append(i)
                    out.append(-2)
                    i += 1
                else:
                    out.append(-2)
                    i += p
            else:
                out.append(i)
                i += 1
        if len(out) <= S:
            src[:len(out)] = out
            src[len(out):] = -1
            return True
    return False


def make_plan(ori, max_seq_len, mask_percent=0.15, choices=None):
    """ori: (B,S,8) integer numpy array on the host.  Consumes the global RNG streams exactly like
    `for b in range(batch): gen_mask(input_ids_encoder[b])` (pretrain.py:131-144)."""
    B, S, _ = ori.shape
    plan = NoisePlan(B, S)
    for b in range(B):
        choice = random.randint(1, 5) if choices is None else int(choices[b])
        plan.choices[b] = choice
        src, loss = plan.src[b], plan.loss[b]
        if choice == 1:
            _deletion(S, mask_percent, src, loss)
        elif choice == 2:
            _token_mask(S, max_seq_len, mask_percent, src, loss, plan.rand_tok)
        elif choice == 3:
            _permutation(ori[b, :, 0].tolist(), src)
            plan.loss_mode[b] = 1
        elif choice == 4:
            if _infilling(S, mask_percent, src):
                plan.loss_mode[b] = 1
            else:
                src[:] = np.arange(S, dtype=np.int32)
                plan.loss_mode[b] = 2
        else:
            r = random.randint(0, S - 1)   # pretrain.py:508-517
            src[:] = (np.arange(S, dtype=np.int32) + r) % S
            if r != 0:
                loss[:] = 1
    return plan


def slice_plan(plan, lo, hi):
    """Rows [lo, hi) of a batch plan as a stand-alone plan (random-token side table re-indexed).  Data-parallel runs draw
    ONE logical plan for the global batch - every rank consumes the RNG streams for all samples in global sample order,
    exactly like the single-process reference (pretrain.py:131-144) - and keep their slice."""
    B, S = hi - lo, plan.src.shape[1]
    out = NoisePlan(B, S)
    out.src[...] = plan.src[lo:hi]
    out.loss[...] = plan.loss[lo:hi]
    out.loss_mode[...] = plan.loss_mode[lo:hi]
    out.choices[...] = plan.choices[lo:hi]
    used = out.src[out.src <= -3]
    if used.size:
        ks = np.unique(-3 - used)                       # side-table rows referenced by the slice, ascending
        remap = {int(k): i for i, k in enumerate(ks)}
        out.rand_tok = [plan.rand_tok[int(k)] for k in ks]
        idx = np.nonzero(out.src <= -3)
        out.src[idx] = [-3 - remap[int(-3 - v)] for v in out.src[idx]]
    return out
