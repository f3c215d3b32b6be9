"""Generation finetuning step, mirroring reference finetune_generation.py (`GenerationTrainer`, :58-290).

Reference behaviour reproduced: the decoder input is the ENCODER input (`y_shift = x`,
finetune_generation.py:155), attention masks are Bar != <PAD> of x, the loss mask is the decoder mask for all 8
attributes, per-attribute CE is multiplied by 0.3 (i = 2, 6, 7) / 1.5 (i = 3) / 1 and by n_tok read in e2w key order,
normalised by sum(n_tok) (:238-250), gradients clipped at 3.0, HF AdamW(lr, weight_decay=0.01).
The "FAD" shape-similarity metric (:185-223, needs the `shapesimilarity` pip package) is out of scope.
"""
import shutil
import sys

import numpy as np
import torch

from . import _lib as L
from .modules import PianoBartLM
from .pretrain import LOSS_WEIGHTS, FusedAdamW, PretrainStep

GEN_EXTRA = [1.0, 1.0, 0.3, 1.5, 1.0, 1.0, 0.3, 0.3]


class GenerationTrainer:
    def __init__(self, pianobart, train_dataloader, valid_dataloader, test_dataloader, lr, testset_shape, cpu,
                 cuda_devices=None, model=None, process_group=None, verbose=True):
        if cpu or not torch.cuda.is_available():
            raise L.PBError('pianobart_b200.GenerationTrainer has no CPU path (sm_100a kernels only)')
        dev = 'cuda'
        if process_group is None and cuda_devices is not None and len(cuda_devices) >= 1:
            dev += ':' + str(cuda_devices[0])
        self.device = torch.device(dev)
        self.pianobart = pianobart
        self.model = (model if model is not None else PianoBartLM(pianobart)).to(self.device)
        self.train_data, self.valid_data, self.test_data = train_dataloader, valid_dataloader, test_dataloader
        self.testset_shape = testset_shape
        self.optim = FusedAdamW(self.pianobart, lr=lr, weight_decay=0.01)
        self.pg = process_group
        self.verbose = verbose
        self._steps = {}

    def _step(self, B, S):
        k = (B, S, self.model.training)
        if k not in self._steps:
            w = [e * n for e, n in zip(GEN_EXTRA, LOSS_WEIGHTS)]
            self._steps[k] = PretrainStep(self.model, B, S, self.optim, process_group=self.pg, loss_weights=w,
                                          loss_norm=float(sum(LOSS_WEIGHTS)))
        return self._steps[k]

    def step(self, x, y, train=True):
        """One batch: x, y (B,S,8) integer tensors.  Returns (loss, per-attribute losses incl. the extra factor, accs)."""
        x = x.to(self.device).long()
        y = y.to(self.device).long()
        keep = (x[:, :, 0] != self.pianobart.bar_pad_word)
        st = self._step(x.shape[0], x.shape[1])
        lm = keep.float().unsqueeze(-1).expand(-1, -1, 8).contiguous()
        st.set_device_batch(x, x, y, lm, keep, keep)          # y_shift = x (finetune_generation.py:155)
        st.run(train=train)
        total, losses, accs = st.fetch_stats()
        return total, losses * np.array(GEN_EXTRA), accs

    def train(self):
        self.model.train()
        return self.iteration(self.train_data, 0)

    def valid(self):
        self.model.eval()
        return self.iteration(self.valid_data, 1)

    def test(self):
        """finetune_generation.py:120-124 without the per-batch argmax dump (the loss / accuracy columns only)."""
        self.model.eval()
        return self.iteration(self.test_data, 2)

    def save_checkpoint(self, epoch, train_acc, valid_acc, valid_loss, train_loss, is_best, filename):
        """finetune_generation.py:272-290: 'state_dict' = the PianoBartLM (reference key names)."""
        state = {'epoch': epoch + 1, 'state_dict': {k: v.detach().clone() for k, v in self.model.state_dict().items()},
                 'valid_acc': valid_acc, 'valid_loss': valid_loss, 'train_loss': train_loss, 'train_acc': train_acc,
                 'optimizer': self.optim.state_dict()}
        torch.save(state, filename)
        if is_best:
            shutil.copyfile(filename, filename.split('.')[0] + '_best.ckpt')

    def iteration(self, training_data, mode):
        total_acc, total_loss, nb = np.zeros(8), 0.0, 0
        for x, y in training_data:
            loss, losses, accs = self.step(x, y, train=(mode == 0))
            if self.verbose:
                sys.stdout.write('Loss: {:06f} | loss: {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}\n'.format(loss, *losses))
            total_loss += loss
            total_acc += accs
            nb += 1
        nb = max(nb, 1)
        return round(total_loss / nb, 3), [round(float(a) / nb, 3) for a in total_acc]
