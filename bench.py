#!/usr/bin/env python
"""Benchmark of the VITS inference hot path (BASELINE.json metric: audio-seconds/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one `SynthesizerTrn.infer` over one batch of synthetic utterances (weak scaling:
every rank processes its own batch; no data-path collective -- utterances are independent,
SURVEY.md §8e).  Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement".

`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU ops =
the same ATen/oneDNN kernels the reference module calls) on the host cores.  The real
reference cannot travel to the GPU box (/root/reference is absent there).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# NCCL prints its version banner on stdout; the contract is ONE JSON line there, so send NCCL's log to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload -> (config, n_vocab, n_speakers, batch per GPU, phonemes, length_scale, cpu sample utterances)
WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (batch 256 x 128 phonemes, full infer)
    "multilingual_v3_b256x128": ("multilingual_v3", 256, 2, 256, 128, 2.8, 48),
    # BASELINE.json configs[1]-like full path on the heavy HiFi-GAN V1 generator
    "baker_v1_b64x128": ("baker_v1", 256, 1, 64, 128, 3.4, 4),
    # small smoke-sized workload
    "multilingual_v3_b8x32": ("multilingual_v3", 256, 2, 8, 32, 2.8, 8),
}
DEFAULT_WORKLOAD = "multilingual_v3_b256x128"
SCALES = (0.667, None, 0.8)  # noise_scale, length_scale (per workload), noise_scale_w -- every reference caller

# per-frame work of the HiFi-GAN generator (SURVEY.md §8d / BASELINE.md §4)
GEN_FLOP_PER_FRAME = {"multilingual_v3": 45.36e6, "baker_v1": 614.9e6}
GEN_LAYER_BYTES_PER_FRAME = {"multilingual_v3": 0.749e6, "baker_v1": 4.05e6}


def make_batch(n_vocab, n_speakers, B, Tx, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randint(0, n_vocab, (B, Tx), generator=gen)
    lens = torch.full((B,), Tx, dtype=torch.long)
    sid = torch.randint(0, n_speakers, (B,), generator=gen)
    return x, lens, sid


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload):
    """DRAM bytes per generator call from the committed ncu capture (profiles/generator_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "generator_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("workload") == workload:
            return float(d["generator_dram_bytes_per_step"]), d.get("source")
    except Exception:
        pass
    return None, None


def pick_cpu_threads(workload):
    """oneDNN convs on tiny channel counts degrade when oversubscribed: try a few thread counts on one
    utterance and keep the fastest (reported as `cores`)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    cands = sorted({c for c in (8, 16, 32, 64, n) if c <= n})
    best, best_t = cands[0], float("inf")
    for c in cands:
        _, dt = cpu_reference_leg(workload, 1, c)
        if dt < best_t:
            best, best_t = c, dt
    return best


def cpu_reference_leg(workload, n_utts, threads):
    """The CPU arm: oracle port of SynthesizerTrn.infer on `n_utts` utterances of the workload."""
    from oracle import vits_oracle as O
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config
    cfg_name, n_vocab, n_spk, B, Tx, ls, _ = WORKLOADS[workload]
    hps = builtin_config(cfg_name)
    sd = synth.make_state_dict(hps.model, n_vocab, n_spk, seed=hps.train.seed)
    w = O.fold_weight_norm(sd)
    torch.set_num_threads(threads)
    x, lens, sid = make_batch(n_vocab, n_spk, B, Tx, 5678)
    x, lens, sid = x[:n_utts], lens[:n_utts], sid[:n_utts]
    t0 = time.perf_counter()
    with torch.no_grad():
        r = O.infer(w, hps.model, x, lens, sid, SCALES[0], ls, SCALES[2], folded=True)
    dt = time.perf_counter() - t0
    audio_s = float(r["y_lengths"].sum()) * hps.data.hop_length / hps.data.sampling_rate
    return audio_s, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name, n_vocab, n_spk, B, Tx, ls, n_utts = WORKLOADS[args.workload]
    threads = pick_cpu_threads(args.workload)
    for _ in range(1 if args.warmup else 0):
        cpu_reference_leg(args.workload, min(2, n_utts), threads)
    tot_audio, tot_t = 0.0, 0.0
    for _ in range(args.steps):
        a, t = cpu_reference_leg(args.workload, n_utts, threads)
        tot_audio += a
        tot_t += t
    v = tot_audio / tot_t
    line = {
        "impl": "reference", "metric": "audio-seconds/sec (VITS infer)", "value": v, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "config": cfg_name, "phonemes": Tx,
                   "sample": f"{n_utts} utterances of the workload per step"},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{n_utts} utterances x {Tx} phonemes per step, oracle port of SynthesizerTrn.infer"},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import wetts_b200
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (wetts_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    cfg_name, n_vocab, n_spk, B, Tx, ls, n_cpu = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    from wetts_b200 import _lib
    _lib.check(_lib.load().wetts_set_option(b"tensor_cores", int(args.tensor_cores)))
    _lib.check(_lib.load().wetts_set_option(b"fused_resblock", int(args.fused_resblock)))
    hps = builtin_config(cfg_name)
    sd = synth.make_state_dict(hps.model, n_vocab, n_spk, seed=hps.train.seed)
    net = wetts_b200.build_model(hps, n_vocab, n_spk, sd, dev)
    hop, sr = hps.data.hop_length, hps.data.sampling_rate
    ns, nsw = SCALES[0], SCALES[2]

    x, lens, sid = make_batch(n_vocab, n_spk, B, Tx, 5678 + rank)
    xh, lh, sh = x.pin_memory(), lens.pin_memory(), sid.pin_memory()
    xd, ld, sdv = x.to(dev), lens.to(dev), sid.to(dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    noise_w = torch.randn(B, 2, Tx, device=dev, generator=gen) if net.use_sdp else None

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- probe once to size noise_z and the outputs
    o, _, y_mask, _ = net.infer(xd, ld, sdv, ns, ls, nsw, noise_w=noise_w, return_attn=False)
    Ty = y_mask.shape[2]
    frames = int(net.last_y_lengths.sum())
    noise_z = torch.randn(B, hps.model.inter_channels, Ty, device=dev, generator=gen)
    audio_s_rank = frames * hop / sr
    # e2e draws its own noise (as the reference does), so Ty varies for SDP models: size for the worst case
    out_cap = int(o.numel() * 1.5) + 4096
    out_host = torch.empty(out_cap, dtype=torch.float32).pin_memory()   # flat: the D2H copy stays contiguous
    del o

    e2e_bytes = [0]

    def step_resident():
        return net.infer(xd, ld, sdv, ns, ls, nsw, noise_w=noise_w, noise_z=noise_z)

    def step_e2e():
        a = xh.to(dev, non_blocking=True)
        b_ = lh.to(dev, non_blocking=True)
        c = sh.to(dev, non_blocking=True)
        o_, *_ = net.infer(a, b_, c, ns, ls, nsw, return_attn=False)   # noise drawn on device, as the reference does
        n = min(o_.numel(), out_cap)
        out_host[:n].copy_(o_.reshape(-1)[:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        e2e_bytes[0] = n * 4
        return o_

    def timed(fn, steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    launches0 = net.launch_count()
    sampler.start()
    if args.profile_range:
        torch.cuda.profiler.start()
    total_ms = timed(step_resident, args.steps)
    if args.profile_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = net.launch_count() - launches0

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    # ---- dominant block: the HiFi-GAN generator, timed alone with CUDA events on the launch stream
    z_in = torch.randn(B, hps.model.inter_channels, Ty, device=dev, generator=gen)
    g_in = net.emb_g(sdv)[:, :, None] if n_spk > 0 else None
    for _ in range(2):
        net.dec(z_in, g=g_in)
    gsteps = max(2, min(args.steps, 5))
    gen_ms = timed(lambda: net.dec(z_in, g=g_in), gsteps) / gsteps

    tot_audio = torch.tensor([audio_s_rank], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(tot_audio)
    audio_job = float(tot_audio.item())
    ms_per_step = total_ms / args.steps
    value = audio_job / (ms_per_step / 1e3)
    e2e_value = audio_job / (e2e_ms / args.steps / 1e3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = pick_cpu_threads(args.workload)
        a_s, dt = cpu_reference_leg(args.workload, n_cpu, threads)
        cpu = {"value": a_s / dt, "unit": "audio-s/s", "cores": threads, "kind": "port",
               "sample": f"{n_cpu} utterances x {Tx} phonemes, oracle port of SynthesizerTrn.infer, {dt:.1f} s"}

    if rank == 0:
        hbm_peak, which = peaks()
        all_frames = B * Ty                                    # the generator runs the padded tail too (finding 9)
        gen_bytes = GEN_LAYER_BYTES_PER_FRAME[cfg_name] * all_frames
        gen_flop = GEN_FLOP_PER_FRAME[cfg_name] * all_frames
        traffic, traffic_src = measured_traffic(args.workload) if (B, args.fused_resblock, args.tensor_cores) == (WORKLOADS[args.workload][3], 1, 1) else (None, None)
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12        # TFLOP/s at the clock seen under load
        line = {
            "metric": "audio-seconds/sec (VITS infer)", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "config": cfg_name, "tensor_cores": bool(args.tensor_cores),
                       "fused_resblock": bool(args.fused_resblock), "batch_per_gpu": B, "phonemes": Tx,
                       "frames_max": Ty, "valid_frames_per_gpu": frames, "length_scale": ls,
                       "sampling_rate": sr, "scales": [ns, ls, nsw], "parallelism": f"batch-sharded x{world}",
                       "l2": "working set >> L2 (multi-GB activations per step); no explicit flush"},
            "rtf": 1.0 / value,
            "e2e": {"value": e2e_value, "unit": "audio-s/s",
                    "h2d_bytes_per_step": int(x.numel() * 8 + lens.numel() * 8 + sid.numel() * 8),
                    "d2h_bytes_per_step": int(e2e_bytes[0] + 8), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "HiFi-GAN generator conv stack (wetts_generator_forward)",
                         "achieved": gen_bytes / (gen_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gen_bytes / (gen_ms / 1e3) / 1e9 / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": gen_bytes,
                         "bytes": "layer-boundary algorithmic bytes (BASELINE.md §4)", "peak_source": which,
                         "ms": gen_ms,
                         "fp32_fma": {"achieved": gen_flop / (gen_ms / 1e3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                      "frac": gen_flop / (gen_ms / 1e3) / 1e12 / fp32_peak,
                                      "note": "binding roof of this stack (SURVEY.md §0 finding 6); peak = 148 SM x 128 FMA x 2 x sm clock under load"}},
            "generator_share_of_step": gen_ms / ms_per_step,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override utterances per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile-range", action="store_true",
                    help="wrap the timed region in cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--tensor-cores", type=int, default=1, help="0: force the fp32 SIMT kernels")
    ap.add_argument("--fused-resblock", type=int, default=1, help="0: one launch per generator conv (no fused MRF stage kernel)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
