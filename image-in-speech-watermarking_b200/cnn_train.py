"""Training-mode execution of ModelA's layers on libwmk kernels (reference step:
`uformerWM/train_modelA.py:402-500`; layers `uformerWM/model.py:3003-3041`).

Every tensor op - convolutions, BatchNorm with batch statistics, activations, pooling, dropout, and
all of their gradients - is a hand-written CUDA kernel reached through the C ABI; torch.autograd is
used only as the tape that orders the backward calls (so the reference's own loop
`loss.backward(); optimizer.step()` keeps working on the drop-in module)."""
import torch
import torch.nn as nn

from . import _lib
from .cnn import ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, _act_of


def _c(t):
    return t.detach().contiguous().float()


def _scratch(C, dev):
    return torch.empty(2 * C, device=dev, dtype=torch.float64)


class _Conv3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        lib = _lib.load()
        x, w = _c(x), _c(w)
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        y = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_conv3x3_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(w), _lib.ptr(_c(b)) if b is not None else None,
                                       None, None, B, Cin, Cout, H, W, 0, Cout, ACT_NONE, 0.0, _lib.stream_ptr()))
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        dy = _c(dy)
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)                                  # correlation with the flipped kernel, weights read in place
            _lib.check(lib.wmk_conv3x3_dgrad_f32(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), B, Cin, Cout, H, W, _lib.stream_ptr()))
        dw = torch.empty_like(w)
        db = torch.empty(Cout, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        _lib.check(lib.wmk_conv3x3_wgrad_f32(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(db), B, Cin, Cout, H, W,
                                             _lib.stream_ptr()))
        return dx, dw, db


class _ConvT2x2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        lib = _lib.load()
        x, w = _c(x), _c(w)
        B, Cin, H, W = x.shape
        Cout = w.shape[1]
        y = torch.empty((B, Cout, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_convT2x2_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(w), _lib.ptr(_c(b)) if b is not None else None,
                                        None, None, B, Cin, Cout, H, W, ACT_NONE, 0.0, _lib.stream_ptr()))
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        dy = _c(dy)
        B, Cin, H, W = x.shape
        Cout = w.shape[1]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.check(lib.wmk_convT2x2_dgrad_f32(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), B, Cin, Cout, H, W, _lib.stream_ptr()))
        dw = torch.empty_like(w)
        db = torch.empty(Cout, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        _lib.check(lib.wmk_convT2x2_wgrad_f32(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(db), B, Cin, Cout, H, W,
                                              _lib.stream_ptr()))
        return dx, dw, db


class _BNAct(torch.autograd.Function):
    """BatchNorm2d (batch statistics, running-stat update) + activation, fused forward and backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, act, slope):
        lib = _lib.load()
        x = _c(x)
        B, C, H, W = x.shape
        y = torch.empty_like(x)
        mr = torch.empty((C, 2), device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_bn_train_fwd_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(_c(gamma)), _lib.ptr(_c(beta)),
                                            _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(mr),
                                            _lib.ptr(_scratch(C, x.device)), B, C, H * W, eps, momentum, act, slope,
                                            _lib.stream_ptr()))
        ctx.save_for_backward(x, y, _c(gamma), mr)
        ctx.act, ctx.slope = act, slope
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, y, gamma, mr = ctx.saved_tensors
        dy = _c(dy)
        B, C, H, W = x.shape
        dx = torch.empty_like(x)
        dg = torch.empty(C, device=x.device, dtype=torch.float32)
        db = torch.empty(C, device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_bn_train_bwd_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(gamma), _lib.ptr(mr),
                                            _lib.ptr(dg), _lib.ptr(db), _lib.ptr(_scratch(C, x.device)), B, C, H * W, ctx.act,
                                            ctx.slope, _lib.stream_ptr()))
        return dx, dg, db, None, None, None, None, None, None


class _BNActPool(torch.autograd.Function):
    """BatchNorm2d (batch statistics) + activation + MaxPool2d(2,2) as one forward pass and one backward pair
    (`wmk_bn_pool_train_fwd/bwd_f32`): the pooling layer's input is written once, its full-resolution gradient never."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, act, slope):
        lib = _lib.load()
        x = _c(x)
        B, C, H, W = x.shape
        y = torch.empty_like(x)
        yp = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=torch.float32)
        mr = torch.empty((C, 2), device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_bn_pool_train_fwd_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(yp), _lib.ptr(_c(gamma)), _lib.ptr(_c(beta)),
                                                 _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(mr),
                                                 _lib.ptr(_scratch(C, x.device)), B, C, H, W, eps, momentum, act, slope,
                                                 _lib.stream_ptr()))
        ctx.save_for_backward(x, y, _c(gamma), mr)
        ctx.act, ctx.slope = act, slope
        return yp

    @staticmethod
    def backward(ctx, dyp):
        lib = _lib.load()
        x, y, gamma, mr = ctx.saved_tensors
        dyp = _c(dyp)
        B, C, H, W = x.shape
        dx = torch.empty_like(x)
        dg = torch.empty(C, device=x.device, dtype=torch.float32)
        db = torch.empty(C, device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_bn_pool_train_bwd_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(dyp), _lib.ptr(dx), _lib.ptr(gamma),
                                                 _lib.ptr(mr), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(_scratch(C, x.device)),
                                                 B, C, H, W, ctx.act, ctx.slope, _lib.stream_ptr()))
        return dx, dg, db, None, None, None, None, None, None


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x = _c(x)
        B, C, H, W = x.shape
        y = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_maxpool2x2_f32(_lib.ptr(x), _lib.ptr(y), B * C, H, W, _lib.stream_ptr()))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        B, C, H, W = x.shape
        dx = torch.empty_like(x)
        _lib.check(lib.wmk_maxpool2x2_bwd_f32(_lib.ptr(x), _lib.ptr(_c(dy)), _lib.ptr(dx), B * C, H, W, _lib.stream_ptr()))
        return dx


class _MaskScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask, scale):
        lib = _lib.load()
        x, mask = _c(x), _c(mask)
        y = torch.empty_like(x)
        _lib.check(lib.wmk_mask_scale_f32(_lib.ptr(x), _lib.ptr(mask), _lib.ptr(y), x.numel(), scale, _lib.stream_ptr()))
        ctx.save_for_backward(mask)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (mask,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _lib.check(lib.wmk_mask_scale_f32(_lib.ptr(dy), _lib.ptr(mask), _lib.ptr(dx), dy.numel(), ctx.scale, _lib.stream_ptr()))
        return dx, None, None


class _ActOnly(torch.autograd.Function):
    """Activation without BatchNorm (ModelA's last `Conv2d(64,1) + ReLU`): runs through the BN kernels'
    activation path with an identity affine would cost a statistics pass; ReLU is a mask multiply."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        mask = (x > 0).float()
        ctx.save_for_backward(mask)
        return _MaskScale.apply(x, mask, 1.0)

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        return _MaskScale.apply(dy, mask, 1.0)


def run_sequential_train(seq, x, dropout_masks=None):
    """Training-mode counterpart of `cnn.run_sequential`: Conv2d / ConvTranspose2d [+ BatchNorm2d (batch
    statistics)] [+ activation] / MaxPool2d / Dropout.  `dropout_masks`: optional iterator of keep masks
    (1 = keep) injected instead of drawing them (parity tests); drawn with torch.rand otherwise."""
    if not x.is_cuda:
        raise _lib.WmkError("the CNN path has no CPU implementation: inputs must be CUDA tensors")
    mods = list(seq)
    masks = iter(dropout_masks) if dropout_masks is not None else None
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv2d):
            if m.kernel_size != (3, 3) or m.padding != (1, 1) or m.stride != (1, 1):
                raise NotImplementedError("conv3x3 kernel: 3x3, stride 1, padding 1 only")
            x = _Conv3x3.apply(x, m.weight, m.bias)
            i += 1
        elif isinstance(m, nn.ConvTranspose2d):
            if m.kernel_size != (2, 2) or m.stride != (2, 2) or m.padding != (0, 0):
                raise NotImplementedError("convT2x2 kernel: kernel 2, stride 2 only")
            x = _ConvT2x2.apply(x, m.weight, m.bias)
            i += 1
        elif isinstance(m, nn.BatchNorm2d):
            act, slope = ACT_NONE, 0.0
            if i + 1 < len(mods) and _act_of(mods[i + 1]) is not None:
                act, slope = _act_of(mods[i + 1])
                i += 1
            momentum = 0.1 if m.momentum is None else float(m.momentum)
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            pool = (isinstance(nxt, nn.MaxPool2d) and nxt.kernel_size in (2, (2, 2)) and nxt.stride in (2, (2, 2)) and
                    nxt.padding in (0, (0, 0)) and x.shape[2] % 2 == 0 and x.shape[3] % 4 == 0)
            if pool:                                   # BatchNorm + activation + MaxPool2d(2,2) fused
                x = _BNActPool.apply(x, m.weight, m.bias, m.running_mean, m.running_var, float(m.eps), momentum, act, slope)
                i += 1
            else:
                x = _BNAct.apply(x, m.weight, m.bias, m.running_mean, m.running_var, float(m.eps), momentum, act, slope)
            if m.num_batches_tracked is not None:
                m.num_batches_tracked += 1
            i += 1
        elif isinstance(m, nn.ReLU):
            x = _ActOnly.apply(x)
            i += 1
        elif isinstance(m, nn.MaxPool2d):
            x = _MaxPool.apply(x)
            i += 1
        elif isinstance(m, nn.Dropout):
            if masks is not None:
                mask = next(masks).to(x.device, torch.float32)
            else:
                mask = (torch.rand(x.shape, device=x.device) >= m.p).float()
            x = _MaskScale.apply(x, mask, 1.0 / (1.0 - m.p))
            i += 1
        else:
            raise NotImplementedError("no CUDA training kernel for %s" % type(m).__name__)
    return x


def mse_loss(pred, target):
    """nn.MSELoss() on libwmk (`train_modelA.py:435-445`): scalar CUDA tensor with a backward."""
    return _MSE.apply(pred, target)


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        lib = _lib.load()
        p, t = _c(pred), _c(target)
        acc = torch.zeros(1, device=p.device, dtype=torch.float64)
        grad = torch.empty_like(p)
        _lib.check(lib.wmk_mse_f32(_lib.ptr(p), _lib.ptr(t), _lib.ptr(grad), p.numel(), 1.0, _lib.ptr(acc), _lib.stream_ptr()))
        ctx.save_for_backward(grad)
        return acc[0].float()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


class FlatAdam:
    """torch.optim.Adam / AdamW semantics (`train_modelA.py:234-236`) as ONE fused kernel over a flat copy of
    the parameters; gradients are gathered into a flat buffer that is also what gets all-reduced across
    ranks (17 655 floats for ModelA: a single latency-bound NCCL call)."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay, self.decoupled = lr, betas, eps, weight_decay, decoupled
        dev = self.params[0].device
        # every parameter starts on a 256-byte boundary of the flat buffer (the dense kernels read weights as 16-byte vectors)
        self.offsets, o = [], 0
        for p in self.params:
            self.offsets.append(o)
            o += (p.numel() + 63) // 64 * 64
        n = o
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1).float())
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros_like(self.grad)
        self.v = torch.zeros_like(self.grad)
        self.t = 0
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)       # completed steps (graph-replayable counter)
        self._pads = [torch.zeros((p.numel() + 63) // 64 * 64 - p.numel(), device=dev, dtype=torch.float32) for p in self.params]
        self._zeros = {}
        # the parameters become views of the flat buffer: the fused Adam kernel updates them in place (no copy back)
        self.views = all(p.dtype == torch.float32 for p in self.params)
        if self.views:
            for p, o in zip(self.params, self.offsets):
                p.data = self.flat[o:o + p.numel()].view_as(p)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def gather_grads(self):
        """All gradients into the flat buffer with ONE concatenation (parameters without a gradient contribute a cached zero
        block: the UformerAudio state_dict holds 63 tensors its forward never touches)."""
        pieces = []
        for i, (p, pad) in enumerate(zip(self.params, self._pads)):
            if p.grad is not None:
                pieces.append(p.grad.reshape(-1))
            else:
                z = self._zeros.get(i)
                if z is None:
                    z = self._zeros[i] = torch.zeros(p.numel(), device=self.grad.device, dtype=torch.float32)
                pieces.append(z)
            if pad.numel():
                pieces.append(pad)
        torch.cat(pieces, out=self.grad)
        return self.grad

    def step(self, grad_scale=1.0):
        """Adam update from self.grad (call gather_grads() - and the all-reduce - first)."""
        lib = _lib.load()
        self.t += 1
        _lib.check(lib.wmk_adam_step_f32(_lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.m), _lib.ptr(self.v),
                                         self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                         self.t, grad_scale, int(self.decoupled), _lib.ptr(self.step_dev), _lib.stream_ptr()))
        if self.views:
            return
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                p.copy_(self.flat[o:o + p.numel()].view_as(p))
