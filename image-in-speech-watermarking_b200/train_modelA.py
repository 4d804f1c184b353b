"""Training step of the CNN baseline (reference `uformerWM/train_modelA.py:402-500`), data parallel.

    audio, wm_decode = model(input_, message)                       # :423
    loss = MSELoss(target, audio) + MSELoss(wm_decode, message)     # :435-447
    loss_scaler(loss, optimizer, parameters=...)                    # :499-500 (timm NativeScaler: GradScaler
                                                                    #  scale -> backward -> unscale -> step; fp32 math)

Here every tensor op of the step is a libwmk kernel (forward, backward, losses, fused Adam); the only
exchange between ranks is ONE all-reduce of the flat 17 655-float gradient buffer per step (NCCL;
BatchNorm statistics stay per rank as in plain DDP).  BASELINE config 5's 'embed+attack+extract' adds
a differentiable attack between encode and decode: additive Gaussian noise (as the reference's own
Uformer variants do, `uformerWM/model.py:1986,2022`)."""
import torch

from . import cnn_train


def gaussian_attack(std, generator=None):
    def attack(encoded):
        return encoded + std * torch.randn(encoded.shape, device=encoded.device, generator=generator)
    return attack


def sync_gradients(flat_grad):
    """Sum the flat gradient buffer over the ranks (one collective: `wmk_grad_allreduce_f32`, NCCL through the
    C ABI); returns the world size (the mean is taken by the Adam kernel's grad_scale)."""
    from . import sharding
    return sharding.allreduce_grads(flat_grad)


def train_step(model, optimizer, input_, message, loss_scale=1.0):
    """One optimisation step on this rank's batch; returns (loss, loss1, loss2) as python floats' device tensors.
    `optimizer` is a `cnn_train.FlatAdam`.  `loss_scale` plays NativeScaler's static role (the scaled
    gradients are unscaled inside the Adam kernel)."""
    model.train()
    optimizer.zero_grad()
    target = input_
    audio, wm_decode = model(input_, message)
    loss1 = cnn_train.mse_loss(audio, target)
    loss2 = cnn_train.mse_loss(wm_decode, message)
    loss = loss1 + loss2
    (loss * loss_scale).backward()
    world = sync_gradients(optimizer.gather_grads())
    optimizer.step(grad_scale=1.0 / (loss_scale * world))
    return loss.detach(), loss1.detach(), loss2.detach()



class HostBatchTrainer:
    """Training loop for HOST batches (what a DataLoader hands over, `uformerWM/train_modelA.py:402-421`): pinned
    (input_, message) pairs are uploaded on a side stream into one of two device slots while the previous step
    computes, and the loss is read back one step late, so neither copy stalls the step.  `submit` returns the
    previous step's loss (None for the first call); `flush` the last one.  Same updates as calling `train_step`
    batch by batch."""

    def __init__(self, model, optimizer, loss_scale=1.0):
        self.model, self.opt, self.loss_scale = model, optimizer, loss_scale
        self.h2d = torch.cuda.Stream()
        self.slots = [None, None]
        self.n = 0
        self.pending = None          # (pinned loss buffer, event) of the last submitted step

    def _slot(self, s, host_x, host_m):
        sl = self.slots[s]
        if sl is None or sl["x"].shape != host_x.shape or sl["m"].shape != host_m.shape:
            dev = torch.device("cuda", torch.cuda.current_device())
            sl = {"x": torch.empty(host_x.shape, dtype=torch.float32, device=dev),
                  "m": torch.empty(host_m.shape, dtype=torch.float32, device=dev),
                  "up": torch.cuda.Event(), "done": torch.cuda.Event(), "loss": torch.empty(3, dtype=torch.float32).pin_memory(),
                  "read": torch.cuda.Event(), "used": False}
            self.slots[s] = sl
        return sl

    def submit(self, host_x, host_m):
        main = torch.cuda.current_stream()
        s = self.n & 1
        sl = self._slot(s, host_x, host_m)
        with torch.cuda.stream(self.h2d):
            if sl["used"]:
                self.h2d.wait_event(sl["done"])             # the step that used this slot two calls ago has finished
            sl["x"].copy_(host_x, non_blocking=True)
            sl["m"].copy_(host_m, non_blocking=True)
            sl["up"].record(self.h2d)
        main.wait_event(sl["up"])
        loss, l1, l2 = train_step(self.model, self.opt, sl["x"], sl["m"], self.loss_scale)
        sl["loss"].copy_(torch.stack([loss, l1, l2]), non_blocking=True)
        sl["done"].record(main)
        sl["used"] = True
        prev, self.pending = self.pending, sl
        self.n += 1
        return self._read(prev)

    @staticmethod
    def _read(sl):
        if sl is None:
            return None
        sl["done"].synchronize()
        return tuple(float(v) for v in sl["loss"])

    def flush(self):
        prev, self.pending = self.pending, None
        return self._read(prev)
