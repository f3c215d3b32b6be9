"""Octuple vocabulary helpers.

The reference loads `(e2w, w2e)` from Data/Octuple.pkl (main.py:21-22, built by
Data/data_generation/make_dict.py:28-164).  Users of this package pass their own pickle
exactly as they do to the reference; for synthetic benchmarks and tests the same *structure*
(sizes, key order, special-token names and ids) is generated here without the data file.
"""

CLASSES = ['Bar', 'Position', 'Instrument', 'Pitch', 'Duration', 'Velocity', 'TimeSig', 'Tempo']
# key order of the reference pickle (make_dict.py:28); pretrain.py:184-189 iterates in this order
KEY_ORDER = ['Bar', 'Position', 'Pitch', 'Duration', 'Velocity', 'Instrument', 'Tempo', 'TimeSig']
REAL = {'Bar': 256, 'Position': 128, 'Instrument': 129, 'Pitch': 256, 'Duration': 128, 'Velocity': 32,
        'TimeSig': 254, 'Tempo': 49}
SPECIALS = ['<PAD>', '<MASK>', '<SOS>', '<EOS>', '<CLS>', '<SEP>']


def build_octuple_vocab():
    """(e2w, w2e) with the reference's sizes (262,134,135,262,134,38,260,55 in `classes` order),
    pickle key order and special-token ids (PAD=real, MASK=real+1, SOS=real+2, EOS=real+3)."""
    e2w, w2e = {}, {}
    for key in KEY_ORDER:
        d = {'%s %d' % (key, i): i for i in range(REAL[key])}
        for j, sp in enumerate(SPECIALS):
            d['%s %s' % (key, sp)] = REAL[key] + j
        e2w[key] = d
        w2e[key] = {v: k for k, v in d.items()}
    return e2w, w2e
