"""`utils.get_arch(opt)` factory of the reference (`uformerWM/utils/model_utils.py:60-108`) for
the architectures on the hot path."""
from ..model import UformerAudio


def get_arch(opt):
    arch = opt.arch
    print('You choose ' + arch + '...')
    if arch == 'Uformer_audio':
        return UformerAudio(img_size=opt.train_ps, embed_dim=32, win_size=8, token_projection='linear',
                            token_mlp='leff', depths=[1, 2, 8, 8, 2, 8, 8, 2, 1], modulator=True, dd_in=2,
                            in_chans=2, audio_scale=getattr(opt, 'audio_scale', '0'),
                            precision=getattr(opt, 'precision', 'mixed'))
    raise Exception("Arch error!")


def load_checkpoint(model, weights):
    """`uformerWM/utils/model_utils.py:27-47`: load a checkpoint, stripping DataParallel's 'module.'."""
    import torch
    ck = torch.load(weights, map_location='cpu')
    sd = ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    sd = {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}
    model.load_state_dict(sd)
