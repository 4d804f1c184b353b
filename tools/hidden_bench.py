"""BASELINE config 3: HiDDeN decoder + noise layer on STFT magnitudes, 128 x 2 s utterances (512 clips).

    python tools/hidden_bench.py [--utterances 128] [--iters 5]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    from image_in_speech_watermarking_b200 import synthetic as SY
    from image_in_speech_watermarking_b200.hidden import noise_layers as NL, audio_test as HT, noise_argparser as NA
    from image_in_speech_watermarking_b200.hidden.model.decoder import Decoder
    from image_in_speech_watermarking_b200.hidden.options import HiDDenConfiguration
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args()
    cfg = HiDDenConfiguration(H=128, W=128, message_length=30, encoder_blocks=4, encoder_channels=64, decoder_blocks=7,
                              decoder_channels=64, use_discriminator=True, use_vgg=False, discriminator_blocks=3,
                              discriminator_channels=64, decoder_loss=1, encoder_loss=0.7, adversarial_loss=1e-3)
    torch.manual_seed(0)
    d = Decoder(cfg, precision=a.precision).cuda().eval()
    B = a.utterances
    waves = SY.synth_speech_batch(0, 4, 2.0).repeat((B + 3) // 4, 1)[:B].cuda()
    msgs = torch.stack([SY.synth_image_binary(i) for i in range(B)]).cuda()
    np.random.seed(0)
    noiser = NL.Noiser(NA.parse_noise("cropout((0.25,0.35),(0.25,0.35))+dropout(0.25,0.35)+quant()"), torch.device("cuda"))
    for _ in range(2):
        HT.attack_and_decode(waves, msgs, d, noiser)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        dec, st = HT.attack_and_decode(waves, msgs, d, noiser)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    clips = dec.shape[0]
    print(json.dumps({"workload": "BASELINE configs[2]: HiDDeN decoder + noise layer on STFT magnitudes, %d x 2 s" % B,
                      "precision": a.precision, "clips": clips, "ms_per_batch": ms, "clips_per_s": clips / (ms * 1e-3), "audio_s_per_s": B * 2.0 / (ms * 1e-3),
                      "decoder_tflops": 7.84e9 * clips / (ms * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
