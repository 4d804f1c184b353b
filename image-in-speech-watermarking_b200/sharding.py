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


def allreduce_stats(vec):
    """Sum the statistics vector over all ranks (NCCL on GPUs, gloo in the CPU tests); identity when
    torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec)
    return vec


def summarize(vec):
    v = [float(x) for x in vec]
    return {"ber_clean": v[0] / v[1], "ber_attacked": v[2] / v[3], "mean_snr_db": v[4] / v[7],
            "mean_audio_mse": v[5] / v[7], "mean_wm_mse_attacked": v[6] / v[7], "utterances": int(v[7])}
