import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from image_in_speech_watermarking_b200 import sharding as SH
r = int(os.environ['RANK']); torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
g = torch.ones(69_000_000, device='cuda')
for name, fn in (("wmk", lambda: SH.allreduce_grads(g)), ("torch", lambda: dist.all_reduce(g))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if r == 0: print(name, "all-reduce of 276 MB: %.2f ms" % ((time.time() - t0) / 5 * 1e3), flush=True)
dist.destroy_process_group()
