#!/usr/bin/env python
"""Benchmark of the embed -> attack -> extract hot path (BASELINE.json metric: audio-seconds/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One "step" = one pass of the hot path over one batch of synthetic utterances:
STFT -> UformerAudio.forward (embed + in-model ISTFT/STFT projection + clean extract) -> ISTFT ->
attack chain -> STFT -> UformerAudio.wm_decode -> SNR / MSE / BER statistics.
Workload at every N (weak scaling, per-GPU work fixed): BASELINE.json configs[1] - 64 utterances
x 3 s, 16 kHz, 32x32 binary images, 'awgn-20+low_pass' attack, random-init weights.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_UTT, SECONDS, SR = 64, 3.0, 16000
ATTACK = "awgn-20+low_pass"
GFLOP_PER_CLIP_FWD, GFLOP_PER_CLIP_EXT = 53.76, 10.43
DTYPE = {"mixed": "bf16 operands / fp32 accumulate (embedder); split-bf16 'bf16x3' = hi*hi+lo*hi+hi*lo / fp32 accumulate "
                  "(extractor: bit parity at |logit| >= 1e-4)",
         "bf16": "bf16", "fp32": "f32"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="mixed", choices=["mixed", "bf16", "fp32"])
    ap.add_argument("--utterances", type=int, default=B_UTT)
    ap.add_argument("--chunk", type=int, default=0, help="clips per internal pass (0 = the whole batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config(args):
    return {"workload": "BASELINE configs[1]: uformerWM embed+attack+extract, %d x %.0f s utterances per GPU, "
                        "32x32 binary image, attacks %s, random-init Uformer_audio" % (args.utterances, SECONDS, ATTACK),
            "utterances_per_gpu": args.utterances, "seconds_per_utterance": SECONDS, "sample_rate": SR,
            "clips_per_utterance": 6, "attack": ATTACK, "precision": args.precision,
            "l2": "per-step working set (%d clips x ~40 MB activations) exceeds the 126 MB L2; no flush needed"
                  % (6 * args.utterances)}


# ----------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_pipeline_step(sd, first_index, n_utt):
    """The oracle port of the reference's CPU path on `n_utt` utterances; returns audio seconds done."""
    import numpy as np
    import torch
    from oracle import pipeline as P
    from image_in_speech_watermarking_b200 import synthetic as SY
    for i in range(n_utt):
        wave = SY.synth_speech(first_index + i, SECONDS)[None]
        msg = SY.synth_image_binary(first_index + i)[None]
        data = P.prepare_data(wave)
        rng = np.random.default_rng(first_index + i)
        P.reconstruct_audio(data, msg, sd, attack=ATTACK, draws={"awgn": rng.standard_normal(wave.shape[-1])})
    return n_utt * SECONDS


def cpu_state_dict():
    from oracle import uformer as O
    from image_in_speech_watermarking_b200 import synthetic as SY
    return SY.init_state_dict(O.state_dict_schema(), "reference", 0)


def cpu_baseline(budget_s=12.0, max_utt=4):
    import torch
    torch.set_num_threads(os.cpu_count())
    sd = cpu_state_dict()
    cpu_pipeline_step(sd, 1000, 1)                       # warm-up (allocator, MKL threads)
    t0 = time.time()
    done = 0
    while done < max_utt and (time.time() - t0 < budget_s or done == 0):
        cpu_pipeline_step(sd, done, 1)
        done += 1
    dt = time.time() - t0
    return {"value": done * SECONDS / dt, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d utterance(s) of %.0f s (6 clips each) of the same workload, oracle port of the reference's "
                      "PyTorch-CPU path, fp32, torch threads=%d, %.1f s wall" % (done, SECONDS, os.cpu_count(), dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count())
    sd = cpu_state_dict()
    n_utt = 1                                            # bounded sample per step
    for w in range(max(1, min(args.warmup, 1))):
        cpu_pipeline_step(sd, 1000 + w, n_utt)
    t0 = time.time()
    for k in range(args.steps):
        cpu_pipeline_step(sd, k, n_utt)
    dt = time.time() - t0
    val = args.steps * n_utt * SECONDS / dt
    cb = {"value": val, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
          "sample": "each step = %d utterance of %.0f s of the workload (of %d per GPU), oracle port of the "
                    "reference's PyTorch-CPU path, torch threads=%d" % (n_utt, SECONDS, args.utterances, os.cpu_count())}
    print(json.dumps({"impl": "reference", "metric": "embed+attack+extract audio-seconds per second", "value": val,
                      "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(args),
                      "cpu_baseline": cb, "gpu_launches": 0,
                      "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML from a Python thread DURING the timed
    region (an `nvidia-smi -lms` subprocess was measured to slow the timed steps by ~30%)."""

    def __init__(self, index):
        self.index, self.rows, self.on, self.h, self.nv = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while self.on:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.nv is None:
            return
        self.on = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.on = False
        self.t.join()
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = [("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        mx = None
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            pass
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": [n for n, b in names if bits & b], "samples": len(sm)}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from image_in_speech_watermarking_b200 import _lib, synthetic as SY
    from image_in_speech_watermarking_b200.model import UformerAudio
    from image_in_speech_watermarking_b200 import audio_test as PT, sharding as SH

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B = args.utterances
    # all clips of the rank's batch in one pass (12 GEMM-sized passes of 32 clips would be launch bound)
    chunk = args.chunk or 6 * B
    model = UformerAudio(precision=args.precision, clips_per_pass=chunk).cuda().eval()   # reference-style random init
    host_w = SY.synth_speech_batch(rank * B, B, SECONDS).pin_memory()
    host_m = torch.stack([SY.synth_image_binary(rank * B + i) for i in range(B)]).pin_memory()
    waves, msgs = host_w.to(dev), host_m.to(dev)
    stats_sum = torch.zeros(8, device=dev, dtype=torch.float64)

    def step(w, m):
        r = PT.embed_attack_extract(w, m, model, ATTACK, seed=1, want_outputs=False)
        # the path's only collective: one all-reduce of the BER / SNR statistics vector (NCCL)
        return SH.allreduce_stats(SH.stats_vector(r["stats"]))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        step(waves, msgs)
    sync()
    if rank == 0:
        sampler.start()
    # ---- device-resident timing (value): K steps, CUDA events on the launching stream
    l0 = lib.wmk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        torch.cuda.nvtx.range_push("wmk_timed_step")     # ncu --nvtx --nvtx-include "wmk_timed_step/" (tools/profile_round.sh)
        vec = step(waves, msgs)
        torch.cuda.nvtx.range_pop()
    e1.record()
    sync()
    launches = (lib.wmk_launch_count() - l0) // args.steps
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # ---- per-kernel-family durations: the same steps again with a CUDA-event pair around every
    # launch (kept out of the timed region above: ~8k extra event records per step cost host time)
    _lib.profile_enable(True)
    step(waves, msgs)
    _lib.profile_collect()                       # first profiled step only fills the event pool
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step(waves, msgs)
    p1.record()
    sync()
    ms_profiled = p0.elapsed_time(p1) / args.steps
    fam = _lib.profile_collect()
    _lib.profile_enable(False)
    # ---- end to end through the public API with host buffers
    sync()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        w = host_w.to(dev, non_blocking=True)
        m = host_m.to(dev, non_blocking=True)
        out = step(w, m).cpu()
    t1.record()
    sync()
    ms_e2e = t0.elapsed_time(t1) / args.steps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_audio = B * SECONDS * world
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_bw = peaks.get("hbm_gbs", 6550.0)
    src = "MEASURED_PEAKS.json (sustained figures: kernels timed inside a long step)" if peaks \
        else "fallback 1.4 PFLOP/s / 6.55 TB/s (B200_PROFILING.md)"
    step_ms_families = {k: round(v["ms"] / args.steps, 3) for k, v in fam.items()}
    kname = "gemm_tcgen05_persistent_kernel" if args.precision != "fp32" else "gemm_fp32_kernel"

    def fam_roof(name, bound):
        f = fam[name]
        sec = f["ms"] * 1e-3
        if sec <= 0:
            return None
        if bound == "tensor":
            ach, peak, unit = f["work"] / sec / 1e12, peak_tf, "TFLOP/s"
        else:
            ach, peak, unit = f["work"] / sec / 1e9, peak_bw, "GB/s"
        n = max(1, f["launches"])
        return {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "launches_per_step": f["launches"] // args.steps, "avg_launch_ms": f["ms"] / n,
                "algorithmic_work_per_launch": f["work"] / n, "share_of_step": f["ms"] / args.steps / ms_profiled}

    # The dense layers are ONE kernel template launched on ~45 shapes: launches whose arithmetic intensity is
    # below the B200 ridge (214 FLOP/B: the C<=128 stages, K = 32..128) are HBM bound, the rest tensor bound
    # (classified per launch in csrc/wmk_common.cuh gemm_work()).  `roofline` is the class with the larger
    # share of the step; the other class is `roofline_other`.
    r_hbm, r_tc = fam_roof("gemm_hbm", "hbm"), fam_roof("gemm", "tensor")
    both = [r for r in (r_hbm, r_tc) if r]
    both.sort(key=lambda r: -r["share_of_step"])
    roofline = dict(both[0]) if both else {"bound": "tensor", "achieved": 0.0, "peak": peak_tf, "unit": "TFLOP/s", "frac": 0.0}
    traffic, traffic_src = None, None
    try:                # dram__bytes_read + dram__bytes_write per launch of the dense-layer kernel, from the committed ncu pass
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))      # regenerated by tools/make_profiles.py
        traffic = tj["families"]["gemm"]["dram_bytes_per_launch"]
        traffic_src = "profiles/r01_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d dense-layer " \
                      "launches of one 64 x 3 s step, both roofline classes)" % tj["families"]["gemm"]["launches"]
    except Exception:
        pass
    roofline.update({"kernel": kname, "peak_source": src, "traffic": traffic, "traffic_source": traffic_src,
                     "family_ms_per_step": step_ms_families, "profiled_step_ms": ms_profiled})
    if len(both) > 1:
        roofline["roofline_other"] = both[1]
    gm = fam["gemm"]["ms"] + fam["gemm_hbm"]["ms"]
    if gm > 0:
        roofline["all_dense_tflops"] = (fam["gemm"]["work"] + fam["gemm_hbm"]["work2"]) / (gm * 1e-3) / 1e12
    for nm, key in (("stft", "stft"), ("istft", "istft")):
        st = fam[nm]
        if st["ms"] > 0:
            roofline[key + "_gbs"] = st["work"] / (st["ms"] * 1e-3) / 1e9
            roofline[key + "_frac_of_hbm"] = roofline[key + "_gbs"] / peak_bw
    line = {"metric": "embed+attack+extract audio-seconds per second", "value": total_audio / (ms * 1e-3),
            "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE[args.precision], "data": "synthetic", "config": config(args),
            "e2e": {"value": total_audio / (ms_e2e * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": host_w.numel() * 4 + host_m.numel() * 4, "d2h_bytes_per_step": 8 * 8,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "stats": {"ber_clean": float(vec[0] / vec[1]), "ber_attacked": float(vec[2] / vec[3]),
                      "mean_snr_db": float(vec[4] / vec[7])}}
    if world == 1 and not args.no_cpu_baseline:
        # BASELINE metric "STFT GB/s vs HBM peak": the front-end kernels on their own at a size that fills the GPU
        # (2048 x 3 s; the launches inside the step above cover 64 utterances in ~20 us and are latency dominated)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import stft_bench
            torch.cuda.synchronize()
            time.sleep(2.0)        # a kernel timed alone: let the board leave the power-capped state of the step loop
            sb = stft_bench.measure(2048, SECONDS, 20)
            line["stft_standalone"] = {k: {"gbs": v["gbs"], "frac_of_hbm": v["gbs"] / peak_bw, "ms": v["ms"]} for k, v in sb.items()}
            line["stft_standalone"]["workload"] = "2048 x 3 s; algorithmic bytes 1276 B/frame (n_fft 255, hop 63), 1536 B/frame (stft256)"
        except Exception as e:                       # never lose the bench line over the side measurement
            line["stft_standalone"] = {"error": str(e)}
        line["cpu_baseline"] = cpu_baseline()
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
