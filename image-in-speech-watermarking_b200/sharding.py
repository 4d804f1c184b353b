"""Multi-GPU layout of the hot path: utterances are independent, so every rank owns a contiguous
block of the batch (all clips of an utterance stay on one GPU - the ISTFT overlap-add needs its
neighbours) and the only exchange is ONE all-reduce of the statistics vector per evaluation
(the reference is single-process, `uformerWM/evaluate.py:372-374`)."""
import torch
import torch.distributed as dist

# layout of the reduced vector (float64)
STAT_KEYS = ("bit_err_clean", "bits_clean", "bit_err_att", "bits_att", "sum_snr_db", "sum_audio_mse",
             "sum_wm_mse_att", "utterances")


def shard_range(n_utt, rank, world):
    """Contiguous block of utterance indices owned by `rank`."""
    per = (n_utt + world - 1) // world
    return range(min(n_utt, rank * per), min(n_utt, (rank + 1) * per))


def stats_vector(stats):
    """Per-utterance stats of `audio_test.embed_attack_extract` (B, 7) -> the additive 8-vector."""
    B = stats.shape[0]
    dev = stats.device
    one = lambda v: torch.tensor(float(v), device=dev, dtype=torch.float64)
    return torch.stack([stats[:, 4].sum(), one(1024.0 * B), stats[:, 5].sum(), stats[:, 6].sum(), stats[:, 0].sum(),
                        stats[:, 1].sum(), stats[:, 3].sum(), one(B)])


_comm = None          # (ncclComm_t handle, world size) of wmk_comm_create, made on first use


def wmk_comm():
    """The library's own NCCL communicator over the ranks of the default process group (`include/wmk.h`
    wmk_comm_create): rank 0 draws the ncclUniqueId, torch.distributed only carries its 128 bytes.  None when
    torch.distributed is not initialised or has a single rank."""
    global _comm
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if _comm is None:
        import ctypes
        from . import _lib
        lib = _lib.load()
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.wmk_comm_unique_id(buf))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        buf = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
        h = ctypes.c_void_p()
        _lib.check(lib.wmk_comm_create(buf, world, rank, ctypes.byref(h)))
        _comm = (h, world)
    return _comm[0]


def allreduce_stats(vec):
    """Sum the statistics vector over all ranks: `wmk_stats_allreduce_f64` (NCCL through the C ABI, on the current
    stream) for device vectors, gloo in the CPU tests; identity when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if vec.is_cuda and vec.dtype == torch.float64 and vec.is_contiguous():
            from . import _lib
            _lib.check(_lib.load().wmk_stats_allreduce_f64(_lib.ptr(vec), vec.numel(), wmk_comm(), _lib.stream_ptr()))
        else:
            dist.all_reduce(vec)
    return vec


def allreduce_grads(flat_grad):
    """Sum the flat fp32 gradient buffer of the data-parallel training step over all ranks (`wmk_grad_allreduce_f32`
    on the device, gloo on the CPU); returns the world size."""
    world = 1
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size()
        if world > 1:
            if flat_grad.is_cuda and flat_grad.dtype == torch.float32 and flat_grad.is_contiguous():
                from . import _lib
                _lib.check(_lib.load().wmk_grad_allreduce_f32(_lib.ptr(flat_grad), flat_grad.numel(), wmk_comm(),
                                                              _lib.stream_ptr()))
            else:
                dist.all_reduce(flat_grad)
    return world


def summarize(vec):
    v = [float(x) for x in vec]
    return {"ber_clean": v[0] / v[1], "ber_attacked": v[2] / v[3], "mean_snr_db": v[4] / v[7],
            "mean_audio_mse": v[5] / v[7], "mean_wm_mse_attacked": v[6] / v[7], "utterances": int(v[7])}
