"""Entry points mirroring reference main.py (`pretrain()` :17-100) and pretrain.py (`get_args_pretrain` :18-48,
`load_data_pretrain` :548-579) on top of the kernel path.  Same flags, same result/ layout and log lines.

    torchrun --nproc-per-node 8 -m pianobart_b200.main --batch_size 16          # data parallel, one process per GPU
    python -m pianobart_b200.main --synthetic 64 --epochs 1                      # no dataset on disk: synthetic Octuple ids

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


def load_data_pretrain(datasets, mode='pretrain', root='Data/output_pretrain'):
    """reference pretrain.py:548-579."""
    to_concat = []
    for ds in datasets:
        parts = [np.load(os.path.join(root, ds, '%s_%s_split.npy' % (ds, s)), allow_pickle=True) for s in ('train', 'test', 'valid')]
        to_concat.append(np.concatenate(parts, axis=0))
    data = np.vstack(to_concat)
    index = np.arange(len(data))
    np.random.shuffle(index)
    data = data[index]
    split = int(len(data) * 0.85)
    return data[:split], data[split:]


def synthetic_data(n, seq, seed):
    real = [256, 128, 129, 256, 128, 32, 254, 49]
    rs = np.random.RandomState(seed)
    ids = np.stack([rs.randint(0, real[i], size=(n, seq)) for i in range(8)], axis=-1).astype(np.int64)
    ids[:, :, 0] = np.sort(ids[:, :, 0], axis=1)
    return ids


def pretrain(argv=None):
    args = get_args_pretrain(argv)
    pg = None
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
        pg = dist.group.WORLD
    rank = int(os.environ.get('RANK', '0'))
    if os.path.exists(args.dict_file):
        with open(args.dict_file, 'rb') as f:
            e2w, w2e = pickle.load(f)
    else:
        e2w, w2e = build_octuple_vocab()
    if args.synthetic:
        X_train, X_val = synthetic_data(args.synthetic, args.max_seq_len, 1 + rank), synthetic_data(max(args.synthetic // 4, args.batch_size), args.max_seq_len, 1001 + rank)
    else:
        X_train, X_val = load_data_pretrain(args.datasets)
    train_loader = DataLoader(MidiDataset(X_train), batch_size=args.batch_size, num_workers=args.num_workers, shuffle=True, drop_last=True)
    valid_loader = DataLoader(MidiDataset(X_val), batch_size=args.batch_size, num_workers=args.num_workers, drop_last=True)
    cfg = BartConfig(max_position_embeddings=args.max_seq_len, d_model=args.hs, encoder_layers=args.layers,
                     encoder_ffn_dim=args.ffn_dims, encoder_attention_heads=args.heads, decoder_layers=args.layers,
                     decoder_ffn_dim=args.ffn_dims, decoder_attention_heads=args.heads)
    pianobart = PianoBart(bartConfig=cfg, e2w=e2w, w2e=w2e, dtype=args.dtype)
    trainer = Pretrainer(pianobart, train_loader, valid_loader, args.lr, args.batch_size, args.max_seq_len,
                         args.mask_percent, args.cpu, args.cuda_devices, process_group=pg, verbose=(rank == 0))
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


if __name__ == '__main__':
    pretrain()
