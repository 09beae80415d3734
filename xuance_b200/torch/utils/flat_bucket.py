"""One contiguous float32 bucket for a learner's parameters and gradients.

The reference leaves parameters as separate tensors and (for DQN / MARL models only) lets DistributedDataParallel
bucket their gradients.  Here every trainable parameter of the learner is re-homed into ONE
flat HBM array (``flat``) and its gradient into another (``grad``): ``p.data`` / ``p.grad`` become views, autograd
accumulates straight into the bucket, the multi-GPU step is a single NCCL all-reduce over ``grad``, and the
clip-norm + Adam step (K7) streams over the two arrays once."""
import torch

_ALIGN = 32  # elements (128 B): keeps every parameter view 128-byte aligned for cuDNN / cuBLAS vector loads


class FlatBucket:
    def __init__(self, params):
        seen, plist = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                plist.append(p)
        if not plist:
            raise ValueError("FlatBucket: no trainable parameters")
        dev = plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatBucket: parameters must live on a CUDA device (no CPU fallback)")
        self.params = plist
        self.offsets, off = [], 0
        for p in plist:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatBucket: all parameters must be float32 on one device")
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(plist, self.offsets):
            v = self._view(self.flat, p, o)
            v.copy_(p.data)
            p.data = v
            p.grad = self._view(self.grad, p, o)

    @staticmethod
    def _view(buf, p, off):
        chunk = buf[off:off + p.numel()]
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            o, i, h, w = p.shape
            return chunk.view(o, h, w, i).permute(0, 3, 1, 2)
        return chunk.view(p.shape)

    def views(self, buf):
        return [self._view(buf, p, o) for p, o in zip(self.params, self.offsets)]

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):  # re-attach if someone set .grad to None
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self._view(self.grad, p, o)
