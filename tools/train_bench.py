"""BASELINE config 5: ModelA training step (embed + Gaussian attack + extract, forward/backward, fused Adam),
per-GPU batch = 32 utterances x 3 s = 192 clips (256 x 3 s over 8 GPUs), data parallel.

    python tools/train_bench.py [--clips 192] [--steps 10]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py

Prints one JSON line (rank 0): clips/s and audio-s/s over all ranks, ms/step (max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from image_in_speech_watermarking_b200.model import ModelA
    from image_in_speech_watermarking_b200 import cnn_train as CT, train_modelA as TM, _lib
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=192)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    m = ModelA().cuda().train()
    m.attack = TM.gaussian_attack(0.05)
    opt = CT.FlatAdam(m.parameters(), lr=2e-4, weight_decay=0.02)
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.rand(a.clips, 2, 128, 128, device="cuda", generator=g)
    wm = (torch.rand(a.clips, 1, 32, 32, device="cuda", generator=g) > 0.5).float()
    step = lambda: TM.train_step(m, opt, x, wm)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.load().wmk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, l1, l2 = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        clips = a.clips * world
        print(json.dumps({"workload": "BASELINE configs[4]: ModelA train step, %d clips/GPU (3 s utterances), Gaussian attack, Adam" % a.clips,
                          "n_gpus": world, "ms_per_step": float(ms), "clips_per_s": clips / (float(ms) * 1e-3),
                          "audio_s_per_s": clips * 0.5 / (float(ms) * 1e-3), "loss": float(loss),
                          "gpu_launches_per_step": (_lib.load().wmk_launch_count() - l0) // a.steps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
