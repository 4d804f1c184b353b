"""Which tensors of the extractor need more than 16 bits?  (CPU study behind the precise-extractor design; not a test.)

    python tests/precision_study.py

Runs the fp32 oracle extractor (`oracle/uformer.py`) on the stress-weight golden clips and re-runs it with ONE
family of tensors rounded to fp16, printing the change of the pre-sigmoid logits (the quantity the north star bounds:
thresholded bits must not flip where |logit| >= 1e-4).  Findings that the CUDA path of `WMK_PREC_MIXED` follows
(csrc/uformer_plan.cu run_block):

  * rounding the WEIGHTS to fp16 moves the logits by ~2.5e-4, rounding every activation by ~9e-5: the weight error is
    the same for every token and does not average out;
  * of the activations, the A operands of the attention projections (LayerNorm-1 output -> QKV, attention output ->
    proj) carry ~8.8e-5; q / k / v / P inside the attention ~2.8e-5; the LeFF tensors (LayerNorm-2 output, the two
    hidden tensors) ~2e-5;
  * hence: weights always as hi + lo (22 / 16 bits), QKV / proj with split-bf16 A operands (three MMAs), the LeFF
    layers with fp16 activations x (hi + lo) fp16 weights (two MMAs), q / k / v / hidden tensors stored in fp16:
    predicted logit deviation 3-4e-5 per 2048 pixels (measured on B200: see DESIGN.md section 2).
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import uformer as O                                    # noqa: E402
from image_in_speech_watermarking_b200 import synthetic as SY      # noqa: E402

CFG = {}


def r(t, key):
    f = CFG.get(key)
    return t if f is None or isinstance(f, str) else t.to(f).float()


def lin(x, w, b, key):
    m = CFG.get(key)
    if m is None:
        return F.linear(x, w, b)
    bd = None if b is None else b.double()
    if m == 'act16':
        return F.linear(x.half().double(), w.double(), bd).float()
    if m == 'w16':
        return F.linear(x.double(), w.half().double(), bd).float()
    return F.linear(r(x, key).double(), r(w, key).double(), bd).float()


def window_attention(sd, p, xx, heads, mask):
    B_, N, C = xx.shape
    hd = C // heads
    q = lin(xx, sd[p + "qkv.to_q.weight"], sd[p + "qkv.to_q.bias"], 'qkv')
    kv = lin(xx, sd[p + "qkv.to_kv.weight"], sd[p + "qkv.to_kv.bias"], 'qkv')
    q = r(q, 'att').reshape(B_, N, 1, heads, hd).permute(2, 0, 3, 1, 4)[0] * (hd ** -0.5)
    kv = r(kv, 'att').reshape(B_, N, 2, heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + "relative_position_bias_table"][O._REL_IDX.view(-1)].view(N, N, -1)
    attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    e = torch.exp(attn - attn.max(-1, keepdim=True).values)
    out = (r(e, 'att') @ v) / e.sum(-1, keepdim=True)
    return lin(out.transpose(1, 2).reshape(B_, N, C), sd[p + "proj.weight"], sd[p + "proj.bias"], 'proj')


def leff(sd, p, xx):
    B, L, C = xx.shape
    hh = int(math.sqrt(L))
    h = r(F.gelu(lin(xx, sd[p + "linear1.0.weight"], sd[p + "linear1.0.bias"], 'l1')), 'h1')
    h = h.view(B, hh, hh, -1).permute(0, 3, 1, 2)
    h = F.gelu(F.conv2d(h, sd[p + "dwconv.0.weight"], sd[p + "dwconv.0.bias"], padding=1, groups=h.shape[1]))
    h = r(h.permute(0, 2, 3, 1).reshape(B, L, -1), 'h2')
    return lin(h, sd[p + "linear2.0.weight"], sd[p + "linear2.0.bias"], 'l2')


def main():
    torch.set_num_threads(os.cpu_count())
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "model_stress.npz")))
    sd = SY.init_state_dict(O.state_dict_schema(), "stress", 0)
    O.window_attention, O.leff = window_attention, leff
    h = torch.float16
    with torch.no_grad():
        for name, clips in (("clean clips", torch.from_numpy(g["x"]) + torch.from_numpy(g["noise"])),
                            ("attacked clips", torch.from_numpy(g["x_att"]))):
            CFG.clear()
            ref = O.wm_decode(sd, clips, return_logits=True)[1]

            def run(cfg, label):
                CFG.clear()
                CFG.update(cfg)
                d = (O.wm_decode(sd, clips, return_logits=True)[1] - ref).abs()
                CFG.clear()
                print("  %-64s max %.2e  mean %.2e" % (label, d.max().item(), d.mean().item()))
            print("%s (%d pixels): change of the logits when ..." % (name, ref.numel()))
            lins = ('qkv', 'proj', 'l1', 'l2')
            run({k: h for k in lins + ('att', 'h1', 'h2')}, "everything in fp16 (operands of all layers)")
            run({k: 'w16' for k in lins}, "only the WEIGHTS of the four dense layers in fp16")
            run({**{k: 'act16' for k in lins}, 'att': h, 'h1': h, 'h2': h}, "only the ACTIVATIONS in fp16 (weights exact)")
            run({'qkv': 'act16', 'proj': 'act16'}, "only the A operands of QKV / proj in fp16")
            run({'att': h}, "only q / k / v / P inside the attention in fp16")
            run({'h1': h, 'h2': h}, "only the two LeFF hidden tensors in fp16")
            run({'att': h, 'h1': h, 'h2': h, 'l1': 'act16', 'l2': 'act16'},
                "the WMK_PREC_MIXED extractor: q/k/v/P, LN2 out, hidden tensors in fp16")


if __name__ == "__main__":
    main()
