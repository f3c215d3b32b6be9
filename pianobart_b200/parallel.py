"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce(sum) of the
flat fp32 gradient buffer, issued in buckets on a side stream while backward is still running.

The reference uses single-process torch.nn.DataParallel (pretrain.py:63-65: replicate / gather / reduce-add
to device 0 every step).  Here every rank computes its own loss terms; because the per-attribute
denominators (mask sums, pretrain.py:117) are all-reduced BEFORE backward, summing the rank gradients gives
exactly the gradient of the reference's full-batch loss.

`BucketReducer` is device-agnostic (CUDA + NCCL in production, CPU + gloo in the tests).

Where the all-reduce kernels run matters more than when they are issued: the tcgen05 GEMMs are persistent kernels with a static
tile schedule, so a GEMM CTA whose SM is held by an NCCL CTA starts late and the whole GEMM waits for it (round 1: 1.6 of the
2.0 ms lost at 8 GPUs was GEMM time), while the attention backward kernels are ordinary grids of 1024 CTAs that simply flow
around occupied SMs.  A bucket that is complete is therefore only ARMED at its marker and launched when the main stream
reaches the next attention backward (`on_attention`, called by engine.Plan.run), i.e. next to ~0.25-0.5 ms of attention
kernels; what is still armed at the end of the backward pass is launched by finish().  PIANOBART_B200_COMM_ALIGN=0 launches at
the markers instead.
"""
import os

import torch
import torch.distributed as dist


class BucketReducer:
    def __init__(self, flat_grad, group=None, target_bytes=None, comm_stream=None, after_reduce=None):
        if target_bytes is None:
            target_bytes = int(os.environ.get('PIANOBART_B200_BUCKET_MB', '32')) << 20
        self.g = flat_grad
        self.group = group
        self.target = max(1, target_bytes // flat_grad.element_size())
        self.comm_stream = comm_stream
        self.after_reduce = after_reduce   # called as after_reduce(lo, hi) on the comm stream behind each bucket's all-reduce
        self.pending = None          # (lo, hi) contiguous range whose gradients are final but not yet reduced
        self.issued = []             # ranges handed to all_reduce, in order
        self.works = []
        self.armed = []              # complete buckets waiting for the next attention backward
        self.align = comm_stream is not None and os.environ.get('PIANOBART_B200_COMM_ALIGN', '1') != '0'

    def _issue(self, lo, hi):
        if self.align:
            self.armed.append((lo, hi))
        else:
            self._launch(lo, hi)

    def on_attention(self, name=None):
        """Plan.run hook, called right before an attention backward is launched on the main stream."""
        armed, self.armed = self.armed, []
        for lo, hi in armed:
            self._launch(lo, hi)

    def _launch(self, lo, hi):
        self.issued.append((lo, hi))
        view = self.g[lo:hi]
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, group=self.group)
                if self.after_reduce is not None:
                    self.after_reduce(lo, hi)
        else:
            self.works.append((dist.all_reduce(view, group=self.group, async_op=True), lo, hi))

    def on_final(self, tag, lo, hi):
        """Plan marker callback: gradients of flat[lo:hi] will not be touched again by this backward."""
        if self.pending is None:
            self.pending = (lo, hi)
        elif hi == self.pending[0]:
            self.pending = (lo, self.pending[1])
        elif lo == self.pending[1]:
            self.pending = (self.pending[0], hi)
        else:
            self._issue(*self.pending)
            self.pending = (lo, hi)
        if self.pending[1] - self.pending[0] >= self.target:
            self._issue(*self.pending)
            self.pending = None

    def finish(self):
        if self.pending is not None:
            self._issue(*self.pending)
            self.pending = None
        self.on_attention()
        for w, lo, hi in self.works:
            w.wait()
            if self.after_reduce is not None:
                self.after_reduce(lo, hi)
        self.works = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        issued, self.issued = self.issued, []
        return issued
