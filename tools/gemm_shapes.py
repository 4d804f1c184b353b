"""Times the bf16 tcgen05 dense layer on every GEMM shape of one Uformer pass (per `clips` clips)
and prints TFLOP/s, effective GB/s and the binding roofline.  GPU tool, not a test."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image_in_speech_watermarking_b200 import _lib

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib = _lib.load()
shapes = []   # (name, M, N, K, count per forward, bytes_in_out)
def stage(name, tokens, C, blocks):
    M = clips * tokens
    shapes.append((name + ".qkv", M, 3 * C, C, blocks))
    shapes.append((name + ".proj", M, C, C, blocks))
    shapes.append((name + ".lin1", M, 4 * C, C, blocks))
    shapes.append((name + ".lin2", M, C, 4 * C, blocks))
for s, (tok, C, nb) in enumerate([(16384, 32, 1), (4096, 64, 2), (1024, 128, 8), (256, 256, 8), (64, 512, 2)]):
    stage("enc%d" % s, tok, C, nb * 2)          # encoder + extractor
    if s < 4:
        shapes.append(("down%d" % s, clips * tok // 4, 2 * C, 16 * C, 2))
for s, (tok, C, nb) in enumerate([(256, 512, 8), (1024, 256, 8), (4096, 128, 2), (16384, 64, 1)]):
    stage("dec%d" % s, tok, C, nb)
    shapes.append(("up%d" % s, clips * tok // 4, 2 * C, [1024, 512, 256, 128][s], 1))
tot_ms = 0; tot_fl = 0
print("%-12s %9s %5s %5s %3s %9s %9s %9s" % ("layer", "M", "N", "K", "cnt", "us", "TFLOP/s", "GB/s"))
for name, M, N, K, cnt in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    # call the bf16 kernel through wmk_linear_f32 would add casts; use the fp32 entry on small shapes only
    Af = A.float(); Wf = W.float(); b = torch.zeros(N, device="cuda"); C_ = torch.empty(M, N, device="cuda")
    _lib.profile_enable(True); _lib.profile_collect()
    for _ in range(3):
        _lib.check(lib.wmk_linear_f32(_lib.ptr(Af), _lib.ptr(Wf), _lib.ptr(b), _lib.ptr(C_), M, N, K, 1, 0, _lib.stream_ptr()))
    _lib.profile_collect()
    for _ in range(5):
        _lib.check(lib.wmk_linear_f32(_lib.ptr(Af), _lib.ptr(Wf), _lib.ptr(b), _lib.ptr(C_), M, N, K, 1, 0, _lib.stream_ptr()))
    r = _lib.profile_collect()["gemm"]
    us = r["ms"] / r["launches"] * 1e3
    fl = 2.0 * M * N * K
    by = 2.0 * M * K + 2.0 * N * K + 4.0 * M * N      # bf16 in, fp32 out in this harness
    print("%-12s %9d %5d %5d %3d %9.1f %9.1f %9.1f" % (name, M, N, K, cnt, us, fl / us / 1e6, by / us / 1e3))
    tot_ms += us * cnt / 1e3; tot_fl += fl * cnt
    del A, W, Af, Wf, C_
print("total GEMM time per forward+extract of %d clips: %.2f ms, %.1f TFLOP/s" % (clips, tot_ms, tot_fl / tot_ms / 1e9))
