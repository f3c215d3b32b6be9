"""Entry points mirroring reference main.py (`pretrain()` :17-100, `finetune()` :103-211, `finetune_generation()`
:214-321), eval_generation.py (:49-115) and pretrain.py (`get_args_pretrain` :18-48, `load_data_pretrain` :548-579) on
top of the kernel path.  Same flags, same result/ layout and log lines.

    torchrun --nproc-per-node 8 -m pianobart_b200.main --batch_size 16          # data parallel, one process per GPU
    python -m pianobart_b200.main --synthetic 64 --epochs 1                      # no dataset on disk: synthetic Octuple ids
    python -m pianobart_b200.main finetune --task composer --dataset Pianist8 --synthetic 32 --epochs 1
    python -m pianobart_b200.main finetune_generation --synthetic 32 --epochs 1
    python -m pianobart_b200.main eval_generation --synthetic 4 --output out.npy

Differences by design: `--cuda_devices` selects the device of THIS process (multi-GPU = torchrun, not nn.DataParallel);
`--dtype {bf16,fp32}`; `--synthetic N` generates N random sequences per split when Data/ is absent.
"""
import argparse
import os
import pickle
import time

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .modules import BartConfig, PianoBart
from .pretrain import Pretrainer
from .vocab import build_octuple_vocab


def get_args_pretrain(argv=None):
    p = argparse.ArgumentParser(description='')
    p.add_argument('--dict_file', type=str, default='./Data/Octuple.pkl')
    p.add_argument('--name', type=str, default='pianobart')
    p.add_argument('--datasets', type=str, nargs='+', default=['asap', 'EMOPIA', 'Pianist8', 'POP1K7', 'POP909'])
    p.add_argument('--num_workers', type=int, default=5)
    p.add_argument('--batch_size', type=int, default=16)
    p.add_argument('--mask_percent', type=float, default=0.15)
    p.add_argument('--max_seq_len', type=int, default=1024)
    p.add_argument('--hs', type=int, default=1024)
    p.add_argument('--layers', type=int, default=8)
    p.add_argument('--ffn_dims', type=int, default=2048)
    p.add_argument('--heads', type=int, default=8)
    p.add_argument('--epochs', type=int, default=500)
    p.add_argument('--lr', type=float, default=2e-5)
    p.add_argument('--cpu', action='store_true')
    p.add_argument('--cuda_devices', type=int, nargs='+', default=[0])
    p.add_argument('--dtype', type=str, default='bf16', choices=['bf16', 'fp32'])
    p.add_argument('--synthetic', type=int, default=0, help='use N synthetic sequences per split instead of Data/')
    return p.parse_args(argv)


class MidiDataset(Dataset):
    """reference dataset.py:4-16."""

    def __init__(self, X):
        self.data = X

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return torch.tensor(self.data[index])


def load_data_pretrain(datasets, mode='pretrain', root='Data/output_pretrain', seed=None):
    """reference pretrain.py:548-579.  seed: under torchrun every rank must draw the SAME 85/15 split (the reference is a
    single process, so its unseeded shuffle is consistent by construction); None keeps the reference's unseeded shuffle."""
    to_concat = []
    for ds in datasets:
        parts = [np.load(os.path.join(root, ds, '%s_%s_split.npy' % (ds, s)), allow_pickle=True) for s in ('train', 'test', 'valid')]
        to_concat.append(np.concatenate(parts, axis=0))
    data = np.vstack(to_concat)
    index = np.arange(len(data))
    (np.random.RandomState(seed) if seed is not None else np.random).shuffle(index)
    data = data[index]
    split = int(len(data) * 0.85)
    return data[:split], data[split:]


def synthetic_data(n, seq, seed):
    real = [256, 128, 129, 256, 128, 32, 254, 49]
    rs = np.random.RandomState(seed)
    ids = np.stack([rs.randint(0, real[i], size=(n, seq)) for i in range(8)], axis=-1).astype(np.int64)
    ids[:, :, 0] = np.sort(ids[:, :, 0], axis=1)
    return ids


def _init_distributed():
    """torchrun: one process per GPU, NCCL.  Returns (process group or None, rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world <= 1:
        return None, 0, 1
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return dist.group.WORLD, rank, world


class GlobalBatchLoader:
    """Data-parallel loader of GLOBAL batches (batch_size x world samples, identical order on every rank, reshuffled
    per epoch from `seed + epoch`): the trainer draws one logical noise plan per global batch and each rank keeps its
    slice (pretrain.Pretrainer(global_batches=True)); an epoch is one pass over the data, not `world` passes."""

    def __init__(self, X, batch_size, world, shuffle, seed=2023):
        self.X, self.gb, self.shuffle, self.seed, self.epoch = X, batch_size * world, shuffle, seed, 0

    def __len__(self):
        return len(self.X) // self.gb

    def __iter__(self):
        idx = np.arange(len(self.X))
        if self.shuffle:
            np.random.RandomState(self.seed + self.epoch).shuffle(idx)
        self.epoch += 1
        for i in range(len(self)):
            yield torch.as_tensor(np.asarray(self.X[idx[i * self.gb:(i + 1) * self.gb]]))


def pretrain(argv=None):
    args = get_args_pretrain(argv)
    pg, rank, world = _init_distributed()
    if os.path.exists(args.dict_file):
        with open(args.dict_file, 'rb') as f:
            e2w, w2e = pickle.load(f)
    else:
        e2w, w2e = build_octuple_vocab()
    if args.synthetic:
        X_train, X_val = synthetic_data(args.synthetic, args.max_seq_len, 1), synthetic_data(max(args.synthetic // 4, args.batch_size * world), args.max_seq_len, 1001)
    else:
        X_train, X_val = load_data_pretrain(args.datasets, seed=2023 if world > 1 else None)
    if world > 1:
        # every rank sees the same split and the same global batches; the trainer slices them (one logical RNG stream)
        import random
        random.seed(2023)
        np.random.seed(2023)
        train_loader = GlobalBatchLoader(X_train, args.batch_size, world, shuffle=True)
        valid_loader = GlobalBatchLoader(X_val, args.batch_size, world, shuffle=False)
        # decorrelate the dropout masks of the ranks (same torch seed would give every rank the same mask per local row)
        torch.manual_seed(torch.initial_seed() ^ (rank << 40))
    else:
        train_loader = DataLoader(MidiDataset(X_train), batch_size=args.batch_size, num_workers=args.num_workers, shuffle=True, drop_last=True)
        valid_loader = DataLoader(MidiDataset(X_val), batch_size=args.batch_size, num_workers=args.num_workers, drop_last=True)
    cfg = BartConfig(max_position_embeddings=args.max_seq_len, d_model=args.hs, encoder_layers=args.layers,
                     encoder_ffn_dim=args.ffn_dims, encoder_attention_heads=args.heads, decoder_layers=args.layers,
                     decoder_ffn_dim=args.ffn_dims, decoder_attention_heads=args.heads)
    pianobart = PianoBart(bartConfig=cfg, e2w=e2w, w2e=w2e, dtype=args.dtype)
    trainer = Pretrainer(pianobart, train_loader, valid_loader, args.lr, args.batch_size, args.max_seq_len,
                         args.mask_percent, args.cpu, args.cuda_devices, process_group=pg, verbose=(rank == 0),
                         global_batches=world > 1)
    save_dir = 'result/pretrain/' + args.name
    os.makedirs(save_dir, exist_ok=True)
    filename = os.path.join(save_dir, 'model.ckpt')
    best_acc, bad_cnt = 0, 0
    start_t = time.time()
    for epoch in range(args.epochs):
        if bad_cnt >= 30:
            print('valid acc not improving for 30 epochs')
            break
        train_loss, train_acc = trainer.train()
        valid_loss, valid_acc = trainer.valid()
        avg_acc = sum(x * y for x, y in zip(valid_acc, pianobart.n_tokens)) / sum(pianobart.n_tokens)
        is_best = avg_acc > best_acc
        best_acc = max(avg_acc, best_acc)
        bad_cnt = 0 if is_best else bad_cnt + 1
        if rank == 0:
            print('epoch: {}/{} | Train Loss: {} | Train acc: {} | Valid Loss: {} | Valid acc: {}'.format(
                epoch + 1, args.epochs, train_loss, train_acc, valid_loss, valid_acc))
            trainer.save_checkpoint(epoch, best_acc, valid_acc, valid_loss, train_loss, is_best, filename)
            with open(os.path.join(save_dir, 'log'), 'a') as outfile:
                outfile.write('Epoch {}: train_loss={}, train_acc={}, valid_loss={}, valid_acc={}\n'.format(
                    epoch + 1, train_loss, train_acc, valid_loss, valid_acc))
    if rank == 0:
        print('Time cost in pretrain of PianoBart is %s' % (time.time() - start_t))


# ----------------------------------------------------------------------------------------------- finetune (main.py:103-211)
class FinetuneDataset(Dataset):
    """reference dataset.py:19-33."""

    def __init__(self, X, y):
        self.data, self.label = X, y

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return torch.tensor(self.data[index]), torch.tensor(self.label[index])


def get_args_finetune(argv=None):
    """reference finetune.py:14-72."""
    p = argparse.ArgumentParser(description='')
    p.add_argument('--task', choices=['melody', 'velocity', 'composer', 'emotion'], required=True)
    p.add_argument('--dataset', type=str, choices=('asap', 'Pianist8', 'POP909', 'EMOPIA', 'GiantMIDI1k'), required=True)
    p.add_argument('--dataroot', type=str, default=None)
    p.add_argument('--dict_file', type=str, default='./Data/Octuple.pkl')
    p.add_argument('--name', type=str, default='pianobart')
    p.add_argument('--ckpt', default='result/pretrain/pianobart/model_best.ckpt')
    p.add_argument('--num_workers', type=int, default=5)
    p.add_argument('--class_num', type=int, default=None)
    p.add_argument('--batch_size', type=int, default=8)
    p.add_argument('--max_seq_len', type=int, default=1024)
    p.add_argument('--hs', type=int, default=1024)
    p.add_argument('--layers', type=int, default=8)
    p.add_argument('--ffn_dims', type=int, default=2048)
    p.add_argument('--heads', type=int, default=8)
    p.add_argument('--epochs', type=int, default=50)
    p.add_argument('--lr', type=float, default=2e-5)
    p.add_argument('--nopretrain', action='store_true')
    p.add_argument('--cpu', action='store_true')
    p.add_argument('--cuda_devices', type=int, nargs='+', default=[0])
    p.add_argument('--weight', type=float, default=None)
    p.add_argument('--error_correction', action='store_true')
    p.add_argument('--dtype', type=str, default='bf16', choices=['bf16', 'fp32'])
    p.add_argument('--synthetic', type=int, default=0, help='use N synthetic sequences per split instead of Data/')
    args = p.parse_args(argv)
    if args.class_num is None:
        args.class_num = {'melody': 4, 'velocity': 7, 'composer': 8, 'emotion': 4}[args.task]
    return args


def load_data_finetune(dataset, task, data_root=None):
    """reference finetune.py:276-337."""
    if data_root is None:
        data_root = 'Data/finetune/others'
    if dataset == 'emotion':
        dataset = 'emopia'
    if dataset not in ['POP909', 'pop909', 'composer', 'EMOPIA', 'asap', 'Pianist8', 'maestro', 'GiantMIDI1k']:
        raise SystemExit('Dataset %s not supported' % dataset)
    ans = 'genans' if task == 'gen' else 'ans'
    X = [np.load(os.path.join(data_root, '%s_%s.npy' % (dataset, s)), allow_pickle=True) for s in ('train', 'valid', 'test')]
    y = [np.load(os.path.join(data_root, '%s_%s_%s.npy' % (dataset, s, ans)), allow_pickle=True) for s in ('train', 'valid', 'test')]
    print('X_train: {}, X_valid: {}, X_test: {}'.format(*[a.shape for a in X]))
    print('y_train: {}, y_valid: {}, y_test: {}'.format(*[a.shape for a in y]))
    return X[0], X[1], X[2], y[0], y[1], y[2]


def _set_seed(seed=2023):
    """main.py:104-110."""
    import random
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def _load_vocab(dict_file):
    if os.path.exists(dict_file):
        with open(dict_file, 'rb') as f:
            return pickle.load(f)
    return build_octuple_vocab()


def _build_pianobart(args, e2w, w2e):
    cfg = BartConfig(max_position_embeddings=args.max_seq_len, d_model=args.hs, encoder_layers=args.layers,
                     encoder_ffn_dim=args.ffn_dims, encoder_attention_heads=args.heads, decoder_layers=args.layers,
                     decoder_ffn_dim=args.ffn_dims, decoder_attention_heads=args.heads)
    return PianoBart(bartConfig=cfg, e2w=e2w, w2e=w2e, dtype=getattr(args, 'dtype', 'bf16'))


def finetune(argv=None):
    """reference main.py:103-211 (sequence / token classification finetuning)."""
    from .finetune import FinetuneTrainer
    _set_seed(2023)
    args = get_args_finetune(argv)
    pg, rank, world = _init_distributed()
    e2w, w2e = _load_vocab(args.dict_file)
    seq_class = args.task in ('composer', 'emotion')
    if args.synthetic:
        rs = np.random.RandomState(7)
        n = args.synthetic
        Xs = [synthetic_data(max(n // d, args.batch_size), args.max_seq_len, 11 + i) for i, d in enumerate((1, 4, 4))]
        if seq_class:
            ys = [rs.randint(0, args.class_num, size=(len(x),)) for x in Xs]
        else:
            ys = [rs.randint(0, args.class_num, size=(len(x), args.max_seq_len)) for x in Xs]
        X_train, X_val, X_test, y_train, y_val, y_test = Xs[0], Xs[1], Xs[2], ys[0], ys[1], ys[2]
    else:
        X_train, X_val, X_test, y_train, y_val, y_test = load_data_finetune(args.dataset, args.task, args.dataroot)
    sampler = None
    if world > 1:
        from torch.utils.data.distributed import DistributedSampler
        sampler = DistributedSampler(FinetuneDataset(X_train, y_train), num_replicas=world, rank=rank, shuffle=True, seed=2023)
    train_loader = DataLoader(FinetuneDataset(X_train, y_train), batch_size=args.batch_size, num_workers=args.num_workers,
                              shuffle=sampler is None, sampler=sampler, drop_last=world > 1)
    valid_loader = DataLoader(FinetuneDataset(X_val, y_val), batch_size=args.batch_size, num_workers=args.num_workers)
    test_loader = DataLoader(FinetuneDataset(X_test, y_test), batch_size=args.batch_size, num_workers=args.num_workers)
    pianobart = _build_pianobart(args, e2w, w2e)
    best_mdl = ''
    if not args.nopretrain and os.path.exists(args.ckpt):
        best_mdl = args.ckpt
        pianobart.load_state_dict(torch.load(best_mdl, map_location='cpu')['state_dict'])
    trainer = FinetuneTrainer(pianobart, train_loader, valid_loader, test_loader, args.lr, args.class_num, args.hs,
                              y_test.shape, args.cpu, args.cuda_devices, None, seq_class, args.error_correction, args.weight,
                              process_group=pg)
    save_dir = os.path.join('result/finetune/', args.task + '_' + args.name)
    os.makedirs(save_dir, exist_ok=True)
    filename = os.path.join(save_dir, 'model.ckpt')
    best_acc, bad_cnt = 0, 0
    with open(os.path.join(save_dir, 'log'), 'a') as outfile:
        outfile.write('Loading pre-trained model from ' + best_mdl.split('/')[-1] + '\n')
        for epoch in range(args.epochs):
            if sampler is not None:
                sampler.set_epoch(epoch)
            train_loss, train_acc = trainer.train()
            valid_loss, valid_acc = trainer.valid()
            test_loss, test_acc, _ = trainer.test()
            is_best = valid_acc >= best_acc
            best_acc = max(valid_acc, best_acc)
            bad_cnt = 0 if is_best else bad_cnt + 1
            if rank == 0:
                print('epoch: {}/{} | Train Loss: {} | Train acc: {} | Valid Loss: {} | Valid acc: {} | Test loss: {} | Test acc: {}'.format(
                    epoch + 1, args.epochs, train_loss, train_acc, valid_loss, valid_acc, test_loss, test_acc))
                trainer.save_checkpoint(epoch, train_acc, valid_acc, valid_loss, train_loss, is_best, filename)
                outfile.write('Epoch {}: train_loss={}, valid_loss={}, test_loss={}, train_acc={}, valid_acc={}, test_acc={}\n'.format(
                    epoch + 1, train_loss, valid_loss, test_loss, train_acc, valid_acc, test_acc))
            if bad_cnt > 3:
                print('valid acc not improving for 3 epochs')
                break
    return trainer


# ----------------------------------------------------------------------------------- generation finetune (main.py:214-321)
def get_args_generation(argv=None):
    """reference finetune_generation.py:15-55."""
    p = argparse.ArgumentParser(description='')
    p.add_argument('--datasets', type=str, default='maestro')
    p.add_argument('--dict_file', type=str, default='./Data/Octuple.pkl')
    p.add_argument('--name', type=str, default='pianobart')
    p.add_argument('--ckpt', default='result/pretrain/pianobart/model_best.ckpt')
    p.add_argument('--num_workers', type=int, default=5)
    p.add_argument('--batch_size', type=int, default=8)
    p.add_argument('--max_seq_len', type=int, default=1024)
    p.add_argument('--hs', type=int, default=1024)
    p.add_argument('--layers', type=int, default=8)
    p.add_argument('--ffn_dims', type=int, default=2048)
    p.add_argument('--heads', type=int, default=8)
    p.add_argument('--epochs', type=int, default=500)
    p.add_argument('--lr', type=float, default=2e-6)
    p.add_argument('--nopretrain', action='store_true')
    p.add_argument('--dataroot', type=str, default=None)
    p.add_argument('--cpu', action='store_true')
    p.add_argument('--cuda_devices', type=int, nargs='+', default=[0])
    p.add_argument('--eval', action='store_true')
    p.add_argument('--dtype', type=str, default='bf16', choices=['bf16', 'fp32'])
    p.add_argument('--synthetic', type=int, default=0)
    return p.parse_args(argv)


def finetune_generation(argv=None):
    """reference main.py:214-321.  The FAD columns of the log are reported as 0 (the `shapesimilarity` metric is out of
    scope)."""
    from .finetune_generation import GenerationTrainer
    from .modules import PianoBartLM
    _set_seed(2023)
    args = get_args_generation(argv)
    pg, rank, world = _init_distributed()
    e2w, w2e = _load_vocab(args.dict_file)
    if args.synthetic:
        n = args.synthetic
        Xs = [synthetic_data(max(n // d, args.batch_size), args.max_seq_len, 21 + i) for i, d in enumerate((1, 4, 4))]
        ys = [synthetic_data(len(x), args.max_seq_len, 31 + i) for i, x in enumerate(Xs)]
        X_train, X_val, X_test, y_train, y_val, y_test = Xs[0], Xs[1], Xs[2], ys[0], ys[1], ys[2]
    else:
        X_train, X_val, X_test, y_train, y_val, y_test = load_data_finetune(dataset=args.datasets, task='gen', data_root=args.dataroot)
    sampler = None
    if world > 1:
        from torch.utils.data.distributed import DistributedSampler
        sampler = DistributedSampler(FinetuneDataset(X_train, y_train), num_replicas=world, rank=rank, shuffle=True, seed=2023)
    train_loader = DataLoader(FinetuneDataset(X_train, y_train), batch_size=args.batch_size, num_workers=args.num_workers,
                              shuffle=sampler is None, sampler=sampler, drop_last=world > 1)
    valid_loader = DataLoader(FinetuneDataset(X_val, y_val), batch_size=args.batch_size, num_workers=args.num_workers)
    test_loader = DataLoader(FinetuneDataset(X_test, y_test), batch_size=args.batch_size, num_workers=args.num_workers)
    pianobart = _build_pianobart(args, e2w, w2e)
    best_mdl, model = '', None
    if args.eval and os.path.exists(args.ckpt):
        best_mdl = args.ckpt
        model = PianoBartLM(pianobart)
        model.load_state_dict(torch.load(best_mdl, map_location='cpu')['state_dict'])
    elif not args.nopretrain and os.path.exists(args.ckpt):
        best_mdl = args.ckpt
        pianobart.load_state_dict(torch.load(best_mdl, map_location='cpu')['state_dict'])
    trainer = GenerationTrainer(pianobart, train_loader, valid_loader, test_loader, args.lr, y_test.shape, args.cpu,
                                args.cuda_devices, model, process_group=pg, verbose=(rank == 0))
    save_dir = os.path.join('result/finetune/generation_' + args.name)
    os.makedirs(save_dir, exist_ok=True)
    filename = os.path.join(save_dir, 'model.ckpt')
    best_acc, bad_cnt = 0, 0
    with open(os.path.join(save_dir, 'log'), 'a') as outfile:
        outfile.write('Loading pre-trained model from ' + best_mdl.split('/')[-1] + '\n')
        for epoch in range(args.epochs):
            if sampler is not None:
                sampler.set_epoch(epoch)
            train_loss, train_acc = trainer.train()
            valid_loss, valid_acc = trainer.valid()
            test_loss, test_acc = trainer.test()
            avg_acc = sum(x * y for x, y in zip(valid_acc, pianobart.n_tokens)) / sum(pianobart.n_tokens)
            is_best = avg_acc > best_acc
            best_acc = max(avg_acc, best_acc)
            bad_cnt = 0 if is_best else bad_cnt + 1
            if rank == 0:
                print('epoch: {}/{} | Train Loss: {} | Train acc: {} | Valid Loss: {} | Valid acc: {} | Test loss: {} | Test acc: {}'.format(
                    epoch + 1, args.epochs, train_loss, train_acc, valid_loss, valid_acc, test_loss, test_acc))
                trainer.save_checkpoint(epoch, train_acc, valid_acc, valid_loss, train_loss, is_best, filename)
                outfile.write('Epoch {}: train_loss={}, valid_loss={}, test_loss={}, train_acc={}, valid_acc={}, test_acc={}, '
                              'train_fad=0, valid_fad=0, test_fad=0\n'.format(epoch + 1, train_loss, valid_loss, test_loss,
                                                                             train_acc, valid_acc, test_acc))
            if bad_cnt > 30:
                print('valid acc not improving for 3 epochs')
                break
    return trainer


# ----------------------------------------------------------------------------------- eval_generation.py:49-115
def get_args_eval_generation(argv=None):
    p = argparse.ArgumentParser(description='')
    p.add_argument('--dict_file', type=str, default='./Data/Octuple.pkl')
    p.add_argument('--ckpt', default='result/finetune/generation_pianobart/model_best.ckpt')
    p.add_argument('--dataset_path', type=str, default='Data/finetune/others')
    p.add_argument('--dataset_name', type=str, default='maestro_test.npy')
    p.add_argument('--output', type=str, default='eval_gen.npy')
    p.add_argument('--num_workers', type=int, default=5)
    p.add_argument('--batch_size', type=int, default=1)
    p.add_argument('--max_seq_len', type=int, default=1024)
    p.add_argument('--hs', type=int, default=1024)
    p.add_argument('--layers', type=int, default=8)
    p.add_argument('--ffn_dims', type=int, default=2048)
    p.add_argument('--heads', type=int, default=8)
    p.add_argument('--nopretrain', action='store_true')
    p.add_argument('--cpu', action='store_true')
    p.add_argument('--cuda_devices', type=int, nargs='+', default=[0])
    p.add_argument('--dtype', type=str, default='bf16', choices=['bf16'])
    p.add_argument('--synthetic', type=int, default=0)
    p.add_argument('--truncate', action='store_true', help='also write <output>.trunc.npy / .len.npy: Octuple2Midi '
                                                           'truncation (demo.py:72-102) applied on the device')
    return p.parse_args(argv)


def eval_generation(argv=None):
    """reference eval_generation.py:49-115: generate for every sequence of a dataset split with the KV-cache decode and
    write the (num, max_seq_len, 8) float32 .npy the reference writes.  Batch sizes > 1 are supported (new capability);
    numpy's global stream is consumed like the reference only for batch 1.  eval() / no-grad arithmetic (demo.py:149-150)."""
    from .modules import PianoBartLM
    from .postprocess import octuple_truncate
    args = get_args_eval_generation(argv)
    if args.cpu or not torch.cuda.is_available():
        raise SystemExit('pianobart_b200 eval_generation has no CPU path (sm_100a kernels only)')
    e2w, w2e = _load_vocab(args.dict_file)
    pianobart = _build_pianobart(args, e2w, w2e)
    model = PianoBartLM(pianobart)
    if not args.nopretrain and os.path.exists(args.ckpt):
        model.load_state_dict(torch.load(args.ckpt, map_location='cpu')['state_dict'], strict=False)
    if args.synthetic:
        data = synthetic_data(args.synthetic, args.max_seq_len, 41)
    else:
        data = np.load(os.path.join(args.dataset_path, args.dataset_name), allow_pickle=True)
    loader = DataLoader(MidiDataset(data), batch_size=args.batch_size, num_workers=args.num_workers, shuffle=False)
    device = torch.device('cuda', args.cuda_devices[0] if args.cuda_devices else 0)
    model = model.to(device)
    model.eval()
    num = len(data)
    output = torch.zeros([num, args.max_seq_len, 8])
    trunc = torch.zeros([num, args.max_seq_len, 8], dtype=torch.int64)
    lens = torch.zeros([num], dtype=torch.int64)
    cnt = 0
    with torch.no_grad():
        for x in loader:
            x = x.long().to(device)
            batch = x.shape[0]
            attn_encoder = (x[:, :, 0] != pianobart.bar_pad_word).float()
            y = model(input_ids_encoder=x, encoder_attention_mask=attn_encoder, generate=True)
            output[cnt:cnt + batch, ...] = y.cpu()
            if args.truncate:
                t, n = octuple_truncate(y)
                trunc[cnt:cnt + batch] = t.cpu()
                lens[cnt:cnt + batch] = n.cpu()
            cnt += batch
    np.save(args.output, output.numpy())
    if args.truncate:
        base = args.output[:-4] if args.output.endswith('.npy') else args.output
        np.save(base + '.trunc.npy', trunc.numpy())
        np.save(base + '.len.npy', lens.numpy())
    return output


# ----------------------------------------------------------------------------------- demo (demo.py:34-170)
def get_args_demo(argv=None):
    p = argparse.ArgumentParser(description='')
    p.add_argument('--dict_file', type=str, default='./Data/Octuple.pkl')
    p.add_argument('--ckpt', default='result/pretrain/pianobart/model_best.ckpt')
    p.add_argument('--input', default='./Data/POP909/POP909/001/001.mid')
    p.add_argument('--output', default='./output.mid')
    p.add_argument('--num_workers', type=int, default=5)
    p.add_argument('--max_seq_len', type=int, default=1024)
    p.add_argument('--hs', type=int, default=1024)
    p.add_argument('--layers', type=int, default=8)
    p.add_argument('--ffn_dims', type=int, default=2048)
    p.add_argument('--heads', type=int, default=8)
    p.add_argument('--nopretrain', action='store_true')
    p.add_argument('--cpu', action='store_true')
    p.add_argument('--cuda_devices', type=int, nargs='+', default=[0])
    p.add_argument('--dtype', type=str, default='bf16', choices=['bf16'])
    return p.parse_args(argv)


def demo(argv=None):
    """reference demo.py: MIDI file -> Octuple prompt (Midi2Octuple) -> KV-cache generation -> truncation + Octuple -> MIDI file
    (Octuple2Midi).  MIDI I/O and the codec are pianobart_b200/codec.py (no miditoolkit needed)."""
    from .modules import PianoBartLM
    from .postprocess import midi_to_octuple, octuple_to_midi
    args = get_args_demo(argv)
    if args.cpu or not torch.cuda.is_available():
        raise SystemExit('pianobart_b200 demo has no CPU path (sm_100a kernels only)')
    e2w, w2e = _load_vocab(args.dict_file)
    pianobart = _build_pianobart(args, e2w, w2e)
    model = PianoBartLM(pianobart)
    if not args.nopretrain and os.path.exists(args.ckpt):
        model.load_state_dict(torch.load(args.ckpt, map_location='cpu')['state_dict'], strict=False)
    device = torch.device('cuda', args.cuda_devices[0] if args.cuda_devices else 0)
    model = model.to(device)
    model.eval()
    x = midi_to_octuple(args.input, device=device)
    if x.shape[1] != args.max_seq_len:
        raise SystemExit('the codec pads prompts to 1024 rows: --max_seq_len must be 1024')
    x = x.long()
    attn_encoder = (x[:, :, 0] != pianobart.bar_pad_word).float()
    with torch.no_grad():
        y = model(input_ids_encoder=x, encoder_attention_mask=attn_encoder, generate=True)
    written = octuple_to_midi(y, [args.output])
    print('Saved to %s' % args.output if written[0] else 'Generate Fail! (empty)')
    return x, y


# ----------------------------------------------------------------------------------- dataset builder (convert.py:583-650)
def get_args_convert(argv=None):
    p = argparse.ArgumentParser(description='MIDI files -> (N, 1024, 8) Octuple .npy blocks (reference Data/data_generation/convert.py)')
    p.add_argument('--input', type=str, required=True, help='directory scanned recursively for .mid / .midi files')
    p.add_argument('--output', type=str, required=True, help='output directory')
    p.add_argument('--dataset', type=str, default='dataset')
    p.add_argument('--task', choices=['pretrain', 'generate', 'melody', 'velocity', 'emotion'], default='pretrain')
    p.add_argument('--nopad', action='store_true', help="pretrain only: concatenate the pieces and cut 1024-row blocks")
    p.add_argument('--seed', type=int, default=None)
    return p.parse_args(argv)


def convert(argv=None):
    """The reference's dataset builder: 80 / 10 / 10 split of the shuffled file list, every piece encoded, cut at the 255-bar
    limit, padded (or packed) to 1024-row blocks; `<dataset>_<split>.npy` (+ `_ans.npy` for the labelled tasks).  Duplicate
    pieces (same (program, pitch) sequence, convert.py:118-122,352-368) are skipped."""
    import hashlib
    import random
    from . import codec
    args = get_args_convert(argv)
    if args.seed is not None:
        random.seed(args.seed)
    files = sorted(os.path.join(r, f) for r, _, fs in os.walk(args.input) for f in fs if f.lower().endswith(('.mid', '.midi')))
    random.shuffle(files)
    os.makedirs(args.output, exist_ok=True)
    n = len(files)
    splits = {'train': files[:80 * n // 100], 'valid': files[80 * n // 100:90 * n // 100], 'test': files[90 * n // 100:]}
    pad = not args.nopad if args.task == 'pretrain' else args.task not in ('melody', 'velocity')
    seen, ok, stats = {}, 0, {}
    for sp in ('test', 'train', 'valid'):
        out, ans = [], []
        for path in splits[sp]:
            try:
                rows = codec.score_to_octuple(codec.read_midi(path), args.task)
                if not rows:
                    print('ERROR(BLANK): ' + path)
                    continue
                digest = hashlib.md5(str(tuple((r[2], r[3]) for r in rows)).encode('ascii')).hexdigest()
                if digest in seen:
                    print('ERROR(DUPLICATED): %s %s == %s' % (digest, path, seen[digest]))
                    continue
                seen[digest] = path
                label = int(os.path.basename(path)[1]) - 1 if args.task == 'emotion' else None
                for item in codec.segments_for_task(rows, args.task, pad=pad, label=label):
                    if args.task == 'pretrain':
                        out.append(item) if pad else out.extend(item)
                    elif args.task == 'generate':
                        out.append(item[0])
                        ans.append(item[1])
                    elif pad:
                        out.append(item[0])
                        ans.append(item[1])
                    else:
                        out.extend(item[0])
                        ans.extend(item[1])
                ok += 1
                print('SUCCESS: ' + path)
            except Exception as ex:                      # (the reference logs and skips a file it cannot process)
                print('ERROR(PROCESS): %s %s' % (path, ex))
        arr = np.array(out)
        if args.task == 'pretrain' and not pad and len(out):
            arr = codec.pack_rows(arr)
        elif args.task in ('melody', 'velocity') and len(out):
            other = codec.MELODY_MAP['OTHER'] if args.task == 'melody' else codec.VELOCITY_MAP['OTHER']
            arr = codec.pack_rows(arr)
            ans = codec.pack_rows(np.array(ans), other, 1)
        if len(out):
            np.save(os.path.join(args.output, '%s_%s.npy' % (args.dataset, sp)), arr)
        if len(ans):
            np.save(os.path.join(args.output, '%s_%s_ans.npy' % (args.dataset, sp)), np.array(ans))
        stats[sp] = (len(splits[sp]), tuple(arr.shape))
    print('%d/%d MIDI files successfully processed' % (ok, n))
    return stats


if __name__ == '__main__':
    import sys
    cmds = {'pretrain': pretrain, 'finetune': finetune, 'finetune_generation': finetune_generation,
            'eval_generation': eval_generation, 'demo': demo, 'convert': convert}
    if len(sys.argv) > 1 and sys.argv[1] in cmds:
        cmds[sys.argv[1]](sys.argv[2:])
    else:
        pretrain()
