"""Host-side data plumbing of the mirrored entry points (reference pretrain.py:548-579, finetune.py:14-72,276-337,
dataset.py:4-33): on-disk `.npy` layout, the 85/15 split, global-batch order under data parallelism, argparse mirrors."""
import os

import numpy as np
import pytest
import torch

from pianobart_b200 import main as M


def _write_pretrain_sets(root, names, n=10, seq=16):
    rs = np.random.RandomState(0)
    total = 0
    for ds in names:
        os.makedirs(os.path.join(root, ds))
        for split, k in (('train', n), ('test', n // 2), ('valid', n // 5)):
            np.save(os.path.join(root, ds, '%s_%s_split.npy' % (ds, split)), rs.randint(0, 30, size=(k, seq, 8)))
            total += k
    return total


def test_load_data_pretrain_concatenates_all_splits_and_cuts_85_15(tmp_path):
    total = _write_pretrain_sets(str(tmp_path), ['a', 'b'])
    tr, va = M.load_data_pretrain(['a', 'b'], root=str(tmp_path), seed=2023)
    assert len(tr) == int(total * 0.85) and len(tr) + len(va) == total and tr.shape[1:] == (16, 8)
    # seeded: every torchrun rank draws the same split (the reference is single-process, its shuffle unseeded)
    tr2, va2 = M.load_data_pretrain(['a', 'b'], root=str(tmp_path), seed=2023)
    assert np.array_equal(tr, tr2) and np.array_equal(va, va2)
    # the split is a permutation of the concatenated files: nothing dropped, nothing duplicated
    allrows = np.concatenate([tr, va]).reshape(total, -1)
    src = np.concatenate([np.load(os.path.join(str(tmp_path), ds, '%s_%s_split.npy' % (ds, s)))
                          for ds in ('a', 'b') for s in ('train', 'test', 'valid')]).reshape(total, -1)
    assert sorted(map(bytes, allrows)) == sorted(map(bytes, src))
    # unseeded (reference behaviour) follows numpy's global stream
    np.random.seed(5)
    t3, _ = M.load_data_pretrain(['a'], root=str(tmp_path))
    np.random.seed(5)
    t4, _ = M.load_data_pretrain(['a'], root=str(tmp_path))
    assert np.array_equal(t3, t4)


def test_global_batch_loader_is_rank_independent_and_reshuffles_per_epoch():
    X = np.arange(37 * 4 * 8).reshape(37, 4, 8)
    a = M.GlobalBatchLoader(X, batch_size=3, world=4, shuffle=True)
    b = M.GlobalBatchLoader(X, batch_size=3, world=4, shuffle=True)
    assert len(a) == 37 // 12
    e0a, e0b = [t.numpy() for t in a], [t.numpy() for t in b]
    assert all(np.array_equal(x, y) for x, y in zip(e0a, e0b)) and all(x.shape == (12, 4, 8) for x in e0a)
    e1a = [t.numpy() for t in a]
    assert not all(np.array_equal(x, y) for x, y in zip(e0a, e1a))           # epoch 1 is reshuffled ...
    assert all(np.array_equal(x, y.numpy()) for x, y in zip(e1a, b))         # ... identically on every rank
    seen = np.concatenate(e0a)[:, 0, 0]
    assert len(set(seen.tolist())) == len(seen)                               # one pass: no sample twice in an epoch
    c = M.GlobalBatchLoader(X, batch_size=3, world=4, shuffle=False)
    assert np.array_equal(next(iter(c)).numpy(), X[:12])


def test_datasets_return_tensors_like_the_reference():
    X = np.arange(2 * 4 * 8).reshape(2, 4, 8)
    y = np.array([[1, 2, 3, 4], [0, 0, 1, 1]])
    d = M.MidiDataset(X)
    assert len(d) == 2 and torch.equal(d[1], torch.tensor(X[1]))
    f = M.FinetuneDataset(X, y)
    xb, yb = f[0]
    assert len(f) == 2 and torch.equal(xb, torch.tensor(X[0])) and torch.equal(yb, torch.tensor(y[0]))


def test_load_data_finetune_file_names_and_answer_suffix(tmp_path, capsys):
    root = str(tmp_path)
    for s, n in (('train', 5), ('valid', 2), ('test', 3)):
        np.save(os.path.join(root, 'Pianist8_%s.npy' % s), np.zeros((n, 16, 8), dtype=np.int64))
        np.save(os.path.join(root, 'Pianist8_%s_ans.npy' % s), np.arange(n))
        np.save(os.path.join(root, 'maestro_%s.npy' % s), np.zeros((n, 16, 8), dtype=np.int64))
        np.save(os.path.join(root, 'maestro_%s_genans.npy' % s), np.ones((n, 16, 8), dtype=np.int64))
    Xtr, Xva, Xte, ytr, yva, yte = M.load_data_finetune('Pianist8', 'composer', root)
    assert (len(Xtr), len(Xva), len(Xte)) == (5, 2, 3) and np.array_equal(yte, np.arange(3))
    out = M.load_data_finetune('maestro', 'gen', root)           # generation answers live in *_genans.npy
    assert out[3].shape == (5, 16, 8) and int(out[3].sum()) == 5 * 16 * 8
    assert 'X_train' in capsys.readouterr().out
    with pytest.raises(SystemExit):
        M.load_data_finetune('nope', 'composer', root)


def test_argparse_mirrors_keep_the_reference_defaults():
    a = M.get_args_pretrain([])
    assert (a.batch_size, a.max_seq_len, a.hs, a.layers, a.heads, a.ffn_dims) == (16, 1024, 1024, 8, 8, 2048)   # main.py / pretrain.py:20-48
    assert abs(a.mask_percent - 0.15) < 1e-12 and abs(a.lr - 2e-5) < 1e-12
    f = M.get_args_finetune(['--task', 'composer', '--dataset', 'Pianist8'])
    assert (f.batch_size, f.class_num, f.epochs) == (8, 8, 50)                                                      # finetune.py:14-72
    assert M.get_args_finetune(['--task', 'velocity', '--dataset', 'GiantMIDI1k']).class_num == 7
    assert M.get_args_finetune(['--task', 'melody', '--dataset', 'POP909']).class_num == 4
    with pytest.raises(SystemExit):
        M.get_args_finetune(['--task', 'composer'])                                                                 # --dataset is required
