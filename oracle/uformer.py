"""Functional PyTorch-CPU fp32 restatement of the reference `UformerAudio`
(`uformerWM/model.py:2225-2511`) for the `Uformer_audio` configuration
(`uformerWM/utils/model_utils.py:83-85`).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Works on a plain ``state_dict``
with the reference's parameter names, so the same weights drive the reference
(in the build container), this oracle (anywhere) and the CUDA plan.

``taps`` (optional dict) receives named intermediates so a GPU parity failure
can be localised to one stage.
"""
import math

import torch
import torch.nn.functional as F

WIN = 8
DEPTHS = [1, 2, 8, 8, 2, 8, 8, 2, 1]
HEADS = [1, 2, 4, 8, 16, 16, 8, 4, 2]
EMBED = 32
IMG = 128
N_FFT = 255
HOP = 63


# --------------------------------------------------------------------------- helpers
def _rel_pos_index():
    """`uformerWM/model.py:496-505`."""
    ch = torch.arange(WIN)
    cw = torch.arange(WIN)
    coords = torch.stack(torch.meshgrid([ch, cw], indexing="ij")).flatten(1)  # 2, 64
    rel = coords[:, :, None] - coords[:, None, :]
    rel = rel.permute(1, 2, 0).contiguous()
    rel[:, :, 0] += WIN - 1
    rel[:, :, 1] += WIN - 1
    rel[:, :, 0] *= 2 * WIN - 1
    return rel.sum(-1)  # 64, 64


_REL_IDX = _rel_pos_index()


def _window_partition(x):
    """`uformerWM/model.py:742-743` (dilation 1)."""
    B, H, W, C = x.shape
    x = x.view(B, H // WIN, WIN, W // WIN, WIN, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, WIN, WIN, C)


def _window_reverse(w, H, W):
    """`uformerWM/model.py:748-754`."""
    B = int(w.shape[0] / (H * W / WIN / WIN))
    x = w.view(B, H // WIN, W // WIN, WIN, WIN, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def _shift_mask(H, W, shift, dtype):
    """`uformerWM/model.py:954-972`: 0 / -100 mask for shifted windows."""
    m = torch.zeros((1, H, W, 1), dtype=dtype)
    slices = (slice(0, -WIN), slice(-WIN, -shift), slice(-shift, None))
    cnt = 0
    for h in slices:
        for w in slices:
            m[:, h, w, :] = cnt
            cnt += 1
    mw = _window_partition(m).view(-1, WIN * WIN)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)  # nW, 64, 64


def window_attention(sd, p, x, heads, mask):
    """`WindowAttention.forward` `uformerWM/model.py:523-551` with
    `LinearProjection.forward` `:460-471`."""
    B_, N, C = x.shape
    hd = C // heads
    q = F.linear(x, sd[p + "qkv.to_q.weight"], sd[p + "qkv.to_q.bias"])
    kv = F.linear(x, sd[p + "qkv.to_kv.weight"], sd[p + "qkv.to_kv.bias"])
    q = q.reshape(B_, N, 1, heads, hd).permute(2, 0, 3, 1, 4)[0]
    kv = kv.reshape(B_, N, 2, heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    q = q * (hd ** -0.5)
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + "relative_position_bias_table"][_REL_IDX.view(-1)].view(N, N, -1)
    attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])


def leff(sd, p, x):
    """`LeFF.forward` `uformerWM/model.py:695-714`."""
    B, L, C = x.shape
    hh = int(math.sqrt(L))
    x = F.gelu(F.linear(x, sd[p + "linear1.0.weight"], sd[p + "linear1.0.bias"]))
    x = x.view(B, hh, hh, -1).permute(0, 3, 1, 2)
    x = F.gelu(F.conv2d(x, sd[p + "dwconv.0.weight"], sd[p + "dwconv.0.bias"], padding=1,
                        groups=x.shape[1]))
    x = x.permute(0, 2, 3, 1).reshape(B, L, -1)
    return F.linear(x, sd[p + "linear2.0.weight"], sd[p + "linear2.0.bias"])


# Train-mode stochastic depth (`DropPath`, `uformerWM/model.py:916,1016-1017`): {block prefix: (2, B) factors in {0, 1 / keep}} for
# the attention / MLP branch; None = eval mode.  Set with `with drop_scales(d): ...` to replay the factors a reference run drew.
_DROP = None


class drop_scales:
    def __init__(self, d):
        self.d = d

    def __enter__(self):
        global _DROP
        self.prev, _DROP = _DROP, self.d
        return self

    def __exit__(self, *a):
        global _DROP
        _DROP = self.prev


def _drop(p, branch, y):
    if _DROP is None or p not in _DROP:
        return y
    return y * torch.as_tensor(_DROP[p][branch], dtype=y.dtype).view(-1, 1, 1)


def lewin_block(sd, p, x, heads, shift):
    """`LeWinTransformerBlock.forward` `uformerWM/model.py:937-1019`."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    if min(H, W) <= WIN:          # `:892-894`
        shift = 0
    mask = _shift_mask(H, W, shift, x.dtype) if shift > 0 else None
    shortcut = x
    y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = _window_partition(y).view(-1, WIN * WIN, C)
    if (p + "modulator.weight") in sd:                      # `:996-999`
        win = win + sd[p + "modulator.weight"]
    a = window_attention(sd, p + "attn.", win, heads, mask)
    y = _window_reverse(a.view(-1, WIN, WIN, C), H, W)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = shortcut + _drop(p, 0, y.view(B, L, C))
    z = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    return x + _drop(p, 1, leff(sd, p + "mlp.", z))


def basic_layer(sd, p, x, depth, heads):
    """`BasicUformerLayer` `uformerWM/model.py:1057-1090` (shift on odd blocks)."""
    for i in range(depth):
        x = lewin_block(sd, "%sblocks.%d." % (p, i), x, heads, 0 if i % 2 == 0 else WIN // 2)
    return x


def _tok2img(x):
    B, L, C = x.shape
    H = int(math.sqrt(L))
    return x.transpose(1, 2).contiguous().view(B, C, H, H)


def downsample(sd, p, x):
    """`Downsample.forward` `uformerWM/model.py:768-775`."""
    o = F.conv2d(_tok2img(x), sd[p + "conv.0.weight"], sd[p + "conv.0.bias"], stride=2, padding=1)
    return o.flatten(2).transpose(1, 2).contiguous()


def upsample(sd, p, x):
    """`Upsample.forward` `uformerWM/model.py:794-800`."""
    o = F.conv_transpose2d(_tok2img(x), sd[p + "deconv.0.weight"], sd[p + "deconv.0.bias"], stride=2)
    return o.flatten(2).transpose(1, 2).contiguous()


def input_proj(sd, p, x):
    """`InputProj.forward` `uformerWM/model.py:824-829` (LeakyReLU slope 0.01)."""
    o = F.leaky_relu(F.conv2d(x, sd[p + "proj.0.weight"], sd[p + "proj.0.bias"], padding=1), 0.01)
    return o.flatten(2).transpose(1, 2).contiguous()


def encoder_stages(sd, p, y, taps=None, tag="enc"):
    """`Encoder.forward` `uformerWM/model.py:1381-1394` (== `EncoderTransformerWM` `:1570-1579`)."""
    convs = []
    for s in range(4):
        y = basic_layer(sd, "%sencoderlayer_%d." % (p, s), y, DEPTHS[s], HEADS[s])
        convs.append(y)
        if taps is not None:
            taps["%s.conv%d" % (tag, s)] = y
        y = downsample(sd, "%sdowsample_%d." % (p, s), y)
        if taps is not None:
            taps["%s.pool%d" % (tag, s)] = y
    y = basic_layer(sd, p + "conv.", y, DEPTHS[4], HEADS[4])
    convs.append(y)
    if taps is not None:
        taps["%s.conv4" % tag] = y
    return convs


def decoder_stages(sd, p, convs, taps=None):
    """`Decoder.forward` `uformerWM/model.py:1221-1240`."""
    y = convs[4]
    for s in range(4):
        up = upsample(sd, "%supsample_%d." % (p, s), y)
        y = torch.cat([up, convs[3 - s]], -1)
        y = basic_layer(sd, "%sdecoderlayer_%d." % (p, s), y, DEPTHS[5 + s], HEADS[5 + s])
        if taps is not None:
            taps["dec.deconv%d" % s] = y
    return y


def wm_encode(sd, message):
    """`ConvAutoencoder.encode` `uformerWM/model.py:1720-1726`."""
    p = "encoder_wm."
    x = F.relu(F.conv2d(message, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1))
    x = F.max_pool2d(x, 2, 2)
    x = F.relu(F.conv2d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1))
    return F.max_pool2d(x, 2, 2)


def wm_decode_logits(sd, feat):
    """`ConvAutoencoder.decode` `uformerWM/model.py:1711-1718` without the sigmoid."""
    p = "encoder_wm."
    x = F.relu(F.conv_transpose2d(feat, sd[p + "t_conv1.weight"], sd[p + "t_conv1.bias"], stride=2))
    return F.conv_transpose2d(x, sd[p + "t_conv2.weight"], sd[p + "t_conv2.bias"], stride=2)


def extractor_features(sd, y, taps=None):
    """`EncoderTransformerWM.forward` `uformerWM/model.py:1568-1583`."""
    p = "decoder_wm."
    t = input_proj(sd, p + "input_proj.", y)
    conv4 = encoder_stages(sd, p, t, taps, "ext")[4]
    c5 = F.conv2d(conv4.unsqueeze(1), sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=(16, 8))
    return c5.squeeze(1).reshape(c5.shape[0], 4, 8, 8)


def wm_decode(sd, y, taps=None, return_logits=False):
    """`UformerAudio.wm_decode` `uformerWM/model.py:2379-2382`."""
    feat = extractor_features(sd, y, taps)
    if taps is not None:
        taps["ext.feat"] = feat
    logits = wm_decode_logits(sd, feat)
    wm = torch.sigmoid(logits)
    return (wm, logits) if return_logits else wm


def istft255(spec_b2ft, length=None):
    """`torch.istft(y.permute(0,2,3,1), n_fft=255)` `uformerWM/model.py:2458`."""
    z = torch.view_as_complex(spec_b2ft.permute(0, 2, 3, 1).contiguous())
    return torch.istft(z, n_fft=N_FFT, length=length, return_complex=False)


def stft255(wave):
    """`torch.stft(istft, n_fft=255).permute(0,3,1,2)` `uformerWM/model.py:2463`."""
    z = torch.stft(wave, n_fft=N_FFT, return_complex=True)
    return torch.view_as_real(z).permute(0, 3, 1, 2).contiguous()


def embed(sd, x, message, taps=None):
    """`UformerAudio.forward` `uformerWM/model.py:2386-2421`: returns (y, noise, wm_pred, wm_pred_logits)."""
    feat_wm = wm_encode(sd, message)                                  # (B,4,8,8)
    feat = feat_wm.reshape(feat_wm.shape[0], feat_wm.shape[1], -1)      # (B,4,64)
    feat_expand = feat.repeat((1, 16, 8))                               # (B,64,512)
    t = input_proj(sd, "input_proj.", x)
    if taps is not None:
        taps["emb.inproj"] = t
    convs = encoder_stages(sd, "encoder.", t, taps, "enc")
    conv4 = convs[4]
    # MaxPool2d((16,8)) on a 3-D (B,64,512) tensor: batch acts as channel (`:2398-2400`)
    c4ds = F.max_pool2d(conv4, kernel_size=(16, 8), stride=(16, 8)).reshape(conv4.shape[0], 4, 8, 8)
    wm_pred_logits = wm_decode_logits(sd, feat_wm + c4ds)
    concat = torch.cat([feat_expand, conv4], dim=2)                     # (B,64,1024)
    d3 = decoder_stages(sd, "decoder.", convs[:4] + [concat], taps)
    B, L, C = d3.shape
    noise = F.conv2d(d3.transpose(1, 2).view(B, C, IMG, IMG), sd["output_proj.proj.0.weight"],
                     sd["output_proj.proj.0.bias"], padding=1)         # `:857-865`
    y = x + noise
    return y, noise, torch.sigmoid(wm_pred_logits), wm_pred_logits


def feature_extract(sd, x, message):
    """`UformerAudio.feature_extract` `uformerWM/model.py:2345-2377`: (y, wm_pred) with wm_pred the image codec's own
    reconstruction `ConvAutoencoder.forward` `:1733-1748` (no bottleneck term, unlike `forward`)."""
    y, _, _, _ = embed(sd, x, message)
    return y, torch.sigmoid(wm_decode_logits(sd, wm_encode(sd, message)))


def forward(sd, x, message, taps=None, return_logits=False):
    """`UformerAudio.forward` `uformerWM/model.py:2384-2511` -> (stft_new, noise, wm_pred, wm)."""
    y, noise, wm_pred, _ = embed(sd, x, message, taps)
    if taps is not None:
        taps["emb.y"] = y
    wave = istft255(y)                                                  # `:2458`
    if taps is not None:
        taps["emb.wave"] = wave
    s = stft255(wave)                                                   # `:2463`
    if taps is not None:
        taps["emb.roundtrip"] = s
    s = F.conv2d(s, sd["stft_layer.0.weight"], sd["stft_layer.0.bias"], padding=1)
    s = F.conv2d(F.relu(s), sd["stft_layer.2.weight"], sd["stft_layer.2.bias"], padding=1)  # `:2305-2309,2465`
    wm, logits = wm_decode(sd, y, taps, return_logits=True)             # `:2508-2509` (reads y, not stft_new)
    if return_logits:
        return s, noise, wm_pred, wm, logits
    return s, noise, wm_pred, wm


# --------------------------------------------------------------------------- state_dict schema
def state_dict_schema():
    """Names and shapes of every tensor in the `Uformer_audio` state_dict,
    in the order the reference registers them.  Values: (shape, kind)."""
    out = {}

    def conv(name, co, ci, kh, kw):
        out[name + ".weight"] = ((co, ci, kh, kw), "conv_w")
        out[name + ".bias"] = ((co,), "bias")

    def block(p, C, heads, mod):
        if mod:
            out[p + "modulator.weight"] = ((64, C), "embed")
        out[p + "norm1.weight"] = ((C,), "ln_w")
        out[p + "norm1.bias"] = ((C,), "ln_b")
        out[p + "attn.relative_position_bias_table"] = ((225, heads), "table")
        out[p + "attn.relative_position_index"] = ((64, 64), "index")
        out[p + "attn.qkv.to_q.weight"] = ((C, C), "lin_w")
        out[p + "attn.qkv.to_q.bias"] = ((C,), "bias")
        out[p + "attn.qkv.to_kv.weight"] = ((2 * C, C), "lin_w")
        out[p + "attn.qkv.to_kv.bias"] = ((2 * C,), "bias")
        out[p + "attn.proj.weight"] = ((C, C), "lin_w")
        out[p + "attn.proj.bias"] = ((C,), "bias")
        out[p + "norm2.weight"] = ((C,), "ln_w")
        out[p + "norm2.bias"] = ((C,), "ln_b")
        out[p + "mlp.linear1.0.weight"] = ((4 * C, C), "lin_w")
        out[p + "mlp.linear1.0.bias"] = ((4 * C,), "bias")
        out[p + "mlp.dwconv.0.weight"] = ((4 * C, 1, 3, 3), "conv_w")
        out[p + "mlp.dwconv.0.bias"] = ((4 * C,), "bias")
        out[p + "mlp.linear2.0.weight"] = ((C, 4 * C), "lin_w")
        out[p + "mlp.linear2.0.bias"] = ((C,), "bias")

    def enc(p):
        conv(p + "input_proj.proj.0", EMBED, 2, 3, 3)
        for s in range(4):
            C = EMBED << s
            for i in range(DEPTHS[s]):
                block("%sencoderlayer_%d.blocks.%d." % (p, s, i), C, HEADS[s], False)
            conv("%sdowsample_%d.conv.0" % (p, s), 2 * C, C, 4, 4)
        for i in range(DEPTHS[4]):
            block("%sconv.blocks.%d." % (p, i), EMBED * 16, HEADS[4], False)

    conv("input_proj.proj.0", EMBED, 2, 3, 3)
    conv("output_proj.proj.0", 2, 2 * EMBED, 3, 3)
    enc("encoder.")
    ups = [(1024, 256), (512, 128), (256, 64), (128, 32)]
    for s in range(4):
        out["decoder.upsample_%d.deconv.0.weight" % s] = ((ups[s][0], ups[s][1], 2, 2), "conv_w")
        out["decoder.upsample_%d.deconv.0.bias" % s] = ((ups[s][1],), "bias")
        C = 2 * ups[s][1]
        for i in range(DEPTHS[5 + s]):
            block("decoder.decoderlayer_%d.blocks.%d." % (s, i), C, HEADS[5 + s], True)
    conv("encoder_wm.conv1", 16, 1, 3, 3)
    conv("encoder_wm.conv2", 4, 16, 3, 3)
    out["encoder_wm.t_conv1.weight"] = ((4, 16, 2, 2), "conv_w")
    out["encoder_wm.t_conv1.bias"] = ((16,), "bias")
    out["encoder_wm.t_conv2.weight"] = ((16, 1, 2, 2), "conv_w")
    out["encoder_wm.t_conv2.bias"] = ((1,), "bias")
    enc("decoder_wm.")
    conv("decoder_wm.conv2", 1, 1, 8, 8)
    conv("stft_layer.0", 4, 2, 3, 3)
    conv("stft_layer.2", 2, 4, 3, 3)
    return out
