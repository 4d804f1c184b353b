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

