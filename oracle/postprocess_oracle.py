"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatement of the truncation part of reference `Octuple2Midi` (/root/reference/demo.py:72-102): returns the rows the
reference hands to `encoding_to_MIDI`.  Pinned against the executed reference by tests/golden/truncate.npz
(tools/make_golden.py golden_truncate runs demo.Octuple2Midi with `miditoolkit` / the MIDI codec stubbed out)."""
import numpy as np

PAD = np.array([256, 128, 129, 256, 128, 32, 254, 49])


def octuple_truncate(octuple):
    """octuple: (S,8) integer array.  Returns (array after the in-place edits of demo.py:78-89, length of the list passed
    to encoding_to_MIDI or None for "Generate Fail")."""
    x = np.array(octuple, dtype=np.int64).copy()
    S = x.shape[0]
    eos = PAD + 3
    end_flag = False
    for i in range(S):                       # demo.py:78-86
        for j in range(8):
            if x[i, j] >= PAD[j] or (j == 3 and x[i, j] > 127):
                end_flag = True
                x[i] = eos
                x[i + 1:] = PAD
                break
        if end_flag:
            break
    if not end_flag:                         # demo.py:88-89
        x[-1] = eos
    length = None
    for i in range(S):                       # demo.py:92-99
        if x[i, 0] == 259:
            length = i
            break
    return x, length
