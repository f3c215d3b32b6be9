"""Sequence- / token-classification finetuning, mirroring reference finetune.py (`FinetuneTrainer`, :75-274).

The backbone forward/backward run on the kernel path through the autograd bridge (modules._BackboneFn); the classifier
heads, the replacement decoder front end and the masked cross-entropy with its argmax / accuracy are library
launches as well (heads.py, csrc/cls_heads.cu, pb_heads_ce).  The 174 M backbone parameters are updated by the fused
HF-semantics AdamW kernel over the flat buffer (pretrain.FusedAdamW, clipping off); the handful of head tensors go through
the same kernel, one launch per tensor (HFAdamW below).  Data parallel: one process per GPU, gradients all-reduced (sum)
over NCCL with the loss normalised by the GLOBAL batch / mask count, so the summed gradient is the reference's full-batch
gradient (the reference uses single-process nn.DataParallel, finetune.py:101-103).  Reference behaviour kept: TokenClassification is built
with class_num+1 (finetune.py:98), velocity (class_num >= 5) feeds shifted labels through the replacement decoder front
end (:194-198), otherwise decoder ids = encoder ids (:211-212); loss = CE masked by encoder non-pad / sum(mask) for token
tasks, mean CE for sequence tasks (:125-132); optional L2-norm regulariser (:241-243); NO gradient clipping (:250);
optimizer: HF-semantics AdamW(lr, weight_decay=0.01) over all parameters.
"""
import ctypes as C
import shutil

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import heads
from .modules import SequenceClassification, TokenClassification
from .pretrain import FusedAdamW


class HFAdamW(torch.optim.Optimizer):
    """transformers 4.29 `AdamW` semantics (eps added before bias correction, decay after the update with plain lr;
    torch.optim.AdamW differs) for the few tensors outside the backbone's flat buffer: one pb_adamw
    launch per tensor, no clipping."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self):
        lib, st_ptr = L.lib(), L.stream_ptr()
        for g in self.param_groups:
            b1, b2 = g['betas']
            for p in g['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise L.PBError('HFAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU path)')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                grad = p.grad.contiguous()
                L.check(lib.pb_adamw(C.c_void_p(p.data_ptr()), C.c_void_p(st['exp_avg'].data_ptr()),
                                     C.c_void_p(st['exp_avg_sq'].data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(None),
                                     C.c_longlong(p.numel()), C.c_float(g['lr']), C.c_float(b1), C.c_float(b2),
                                     C.c_float(g['eps']), C.c_float(g['weight_decay']), st['step'], C.c_void_p(None),
                                     C.c_float(0.0), C.c_float(1.0), C.c_float(1.0), st_ptr), 'adamw')


class FinetuneTrainer:
    def __init__(self, pianobart, train_dataloader, valid_dataloader, test_dataloader, lr, class_num, hs, testset_shape,
                 cpu, cuda_devices=None, model=None, SeqClass=False, error=False, weight=None, process_group=None):
        if cpu or not torch.cuda.is_available():
            raise L.PBError('pianobart_b200.FinetuneTrainer has no CPU path (sm_100a kernels only)')
        dev = 'cuda'
        if process_group is not None:
            dev += ':' + str(torch.cuda.current_device())
        elif cuda_devices is not None and len(cuda_devices) >= 1:
            dev += ':' + str(cuda_devices[0])
        self.pg, self.world = process_group, 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.device = torch.device(dev)
        self.pianobart, self.SeqClass, self.class_num = pianobart, SeqClass, class_num
        if model is not None:
            self.model = model.to(self.device)
        elif SeqClass:
            self.model = SequenceClassification(pianobart, class_num, hs).to(self.device)
        else:
            self.model = TokenClassification(pianobart, class_num + 1, hs).to(self.device)
        self.train_data, self.valid_data, self.test_data = train_dataloader, valid_dataloader, test_dataloader
        # backbone: fused kernel over the flat fp32 buffer (no clipping in the reference's finetune, finetune.py:250);
        # head / replacement-front-end tensors: per-tensor HFAdamW
        self.pianobart._ensure_packed()
        flat_ids = {id(p) for _, p in self.pianobart._named_flat_params()}
        head_params = [p for p in self.model.parameters() if p.requires_grad and id(p) not in flat_ids
                       and p is not self.pianobart.bart.shared.weight]
        self.optim_backbone = FusedAdamW(self.pianobart, lr=lr, weight_decay=0.01, max_grad_norm=0.0)
        self.optim = HFAdamW(head_params, lr=lr, weight_decay=0.01)
        self._head_params = head_params
        self.testset_shape = testset_shape if not error else (testset_shape[:-1] if testset_shape is not None else None)
        self.weight, self.error = weight, error

    def compute_loss(self, predict, target, loss_mask, seq):
        """finetune.py:125-132 with the reference's argument layout (token tasks: predict is [B, C, S])."""
        if not seq:
            predict = predict.permute(0, 2, 1)
        return self._loss(predict, target, loss_mask, seq)[0]

    def _loss(self, y_hat, y, attn, seq):
        """-> (loss, #correct, argmax): masked CE / sum(mask) for token tasks, mean CE for sequence tasks - one pb_heads_ce pass"""
        if seq:
            return heads.masked_ce(y_hat, y, None)
        M = y_hat.shape[0] * y_hat.shape[1]
        loss, correct, am = heads.masked_ce(y_hat.reshape(M, -1), y.reshape(M), attn.reshape(M))
        return loss, correct, am.view(y.shape)

    def train(self):
        self.model.train()
        return self.iteration(self.train_data, 0, self.SeqClass)

    def valid(self):
        self.model.eval()
        return self.iteration(self.valid_data, 1, self.SeqClass)

    def test(self):
        self.model.eval()
        return self.iteration(self.test_data, 2, self.SeqClass)

    def step(self, x, y, mode=0):
        """One batch (finetune.py:168-251).  Returns (loss tensor, #correct, #counted, argmax output)."""
        seq = self.SeqClass
        x, y = x.to(self.device).long(), y.to(self.device).long()
        if self.error:
            y = torch.squeeze(y, dim=-1)
        attn = (x[:, :, 0] != self.pianobart.bar_pad_word).float()
        with torch.set_grad_enabled(mode == 0):
            if seq:
                y_hat = self.model(input_ids_encoder=x, encoder_attention_mask=attn)
            else:
                if self.class_num >= 5:
                    y_shift = torch.zeros_like(y) + self.class_num
                    y_shift[:, 1:] = y[:, :-1]
                    attn_shift = torch.zeros_like(attn)
                    attn_shift[:, 1:] = attn[:, :-1]
                    attn_shift[:, 0] = attn[:, 0]
                else:
                    y_shift, attn_shift = x.clone(), attn.clone()
                y_hat = self.model(input_ids_encoder=x, input_ids_decoder=y_shift, encoder_attention_mask=attn,
                                   decoder_attention_mask=attn_shift)
            loss, correct, output = self._loss(y_hat, y, attn, seq)
            output = output.long()
            count = y.shape[0] if seq else torch.sum(attn).item()
            if self.world > 1:
                # global normaliser: loss_r = (local sum) / (global count); the rank gradients then SUM to the full-batch one
                import torch.distributed as dist
                n_loc = torch.tensor([float(count)], device=self.device)
                n_all = n_loc.clone()
                dist.all_reduce(n_all, group=self.pg)
                loss = loss * (n_loc / n_all).squeeze()
            if self.weight is not None:
                reg = sum(torch.norm(param, p=2) for param in self.model.parameters())
                loss = loss + self.weight * reg / self.world
            if mode == 0:
                pb = self.pianobart
                pb._grad.zero_()                      # one memset instead of ~370 per-parameter ones
                for p in self._head_params:
                    p.grad = None
                loss.backward()                       # no clipping in the reference (finetune.py:250)
                if self.world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(pb._grad, group=self.pg)
                    for p in self._head_params:
                        if p.grad is not None:
                            dist.all_reduce(p.grad, group=self.pg)
                self.optim_backbone.step()
                self.optim.step()
        if self.world > 1:
            import torch.distributed as dist
            loss = loss.detach().clone()
            dist.all_reduce(loss, group=self.pg)      # the global loss value (sum of the normalised rank terms)
        return loss.detach(), correct, count, output

    def iteration(self, training_data, mode, seq):
        total_acc, total_cnt, total_loss, nb = 0.0, 0, 0.0, 0
        all_output, cnt = (torch.empty(self.testset_shape), 0) if mode == 2 else (None, 0)
        for x, y in training_data:
            loss, correct, count, output = self.step(x, y, mode)
            if mode == 2:
                all_output[cnt:cnt + x.shape[0]] = output.cpu()
                cnt += x.shape[0]
            total_loss += loss.item()
            total_acc += float(correct)
            total_cnt += count
            nb += 1
        res = (round(total_loss / max(nb, 1), 4), round(total_acc / max(total_cnt, 1), 4))
        return res + (all_output,) if mode == 2 else res

    def save_checkpoint(self, epoch, train_acc, valid_acc, valid_loss, train_loss, is_best, filename):
        state = {'epoch': epoch + 1, 'state_dict': {k: v.detach().clone() for k, v in self.model.state_dict().items()},
                 'valid_acc': valid_acc, 'valid_loss': valid_loss, 'train_loss': train_loss, 'train_acc': train_acc,
                 'optimizer': {'heads': self.optim.state_dict(), 'backbone': self.optim_backbone.state_dict()}}
        torch.save(state, filename)
        if is_best:
            shutil.copyfile(filename, filename.split('.')[0] + '_best.ckpt')
