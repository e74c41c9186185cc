#!/usr/bin/env python
"""Headline benchmark: images/sec of the vit-ad scoring hot path (DeiT-B encoder + GMM/MDN head with 100
Gaussians + image scores + 224x224 anomaly maps), 224x224 synthetic images, batch 32 per GPU.

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)

A step = one batch of 32 images through encoder -> head -> scores/maps.  `value` is timed with the inputs
resident in HBM; `e2e` runs the same step through the validator API from pinned host memory (H2D of the
images and D2H of scores+maps inside the timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "vit-ad_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "images/sec (DeiT enc+GMM score, 224^2, bs32)"
WORKLOAD = "enc_deit + MDN/GMM head (100 Gaussians) validation scoring, synthetic 224x224 MVTecAD-shaped images, batch 32 per GPU"
ENC_GFLOP_PER_IMG = 35.31
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the GPU is busy, from a background thread: an NVML
    query can take milliseconds, and from the launching thread it would let the GPU's queue run dry mid-measurement."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz, self.h = [], set(), None, None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def sample(self):
        if self.h is None:
            return
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            try:
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self, period_s: float = 0.02):
        import threading

        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                self.sample()
                self._stop.wait(period_s)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        if getattr(self, "_thread", None) is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_images_per_sec(sample_images: int, steps: int, warmup: int, gaussians: int):
    """The oracle port of the reference algorithm (oracle/vitad_oracle.py) on all host cores, fp32, no_grad:
    encoder -> MDN head -> scores/maps as numpy (the span of ValidatorMDN.py:123-168)."""
    import torch

    from oracle import vitad_oracle as O
    from oracle import weights as W

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    enc_sd = W.make_deit_state_dict(seed=11)
    mdn_sd = W.make_mdn_state_dict(seed=21, num_gaussians=gaussians)
    imgs = W.synthetic_images(seed=1, batch=sample_images)
    g = O.gumbel_noise((sample_images, 196, gaussians), torch.Generator().manual_seed(7))

    def step():
        with torch.no_grad():
            tok, _ = O.deit_forward(enc_sd, imgs)
            L = O.mdn_patch_loglik(tok, mdn_sd, g)
            s, m = O.mdn_scores(O.mdn_probability_map(L), 224, 16)
        return s.numpy(), m.numpy()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_images * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    ips, ms, cores = cpu_reference_images_per_sec(sample, args.steps, min(args.warmup, 1), args.gaussians)
    sample_desc = f"{sample} images per step (of the 32-image batch), {args.steps} steps, oracle port, torch fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample_desc},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from vitad import synth_weights as W  # seeded synthetic weights / inputs (product package, not oracle/)
    from vitad import _lib, ops
    from vitad.encoders import EncoderDeit
    from vitad.mdn import GaussianMixtureDensityNetwork
    from vitad.parallel import init_from_env
    from vitad.validators import ValidatorMdn

    rank, world, local = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, K = args.batch, args.gaussians
    NBUF = 4

    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11))
    head = GaussianMixtureDensityNetwork(768, 768, K)
    head.load_state_dict(W.make_mdn_state_dict(seed=21, num_gaussians=K))
    enc, head = enc.to(dev).eval(), head.to(dev).eval()

    host_imgs = [W.synthetic_images(seed=100 + rank * NBUF + i, batch=B).pin_memory() for i in range(NBUF)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    gum = [-torch.empty(B, 196, K, device=dev).exponential_().log() for _ in range(NBUF)]
    props = {"dataset": "synthetic", "dataclass": "bench", "num_gaussians": K, "fp_thres": 0.3}
    validator = ValidatorMdn([head], enc, None, props, gumbel=lambda bi, shape: gum[bi % NBUF])
    stream = torch.cuda.current_stream()
    mdn_events = []

    def step_device(i, timed=False):
        f = enc(dev_imgs[i % NBUF])
        x = f.patch_embedding
        if timed:  # bracket the dominant kernel (fused projection + logsumexp) on its launch stream
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            head.timing_hook = (e0, e1)
            mdn_events.append((e0, e1))
        prob, scores = head.score(x, gum[i % NBUF])
        head.timing_hook = None
        maps, _ = ops.bilinear_up(prob.view(-1, 14, 14), 224, align_corners=True, post_one_minus=True)
        return scores, maps

    e2e_wall = []  # host wall time between consecutive results of the streaming loop

    def run_e2e(n):
        """n batches through the validator's public streaming loop: every batch is copied from pinned host memory
        (H2D, copy-in stream), scored, and its scores + maps land in host numpy arrays (D2H, copy-out stream); the
        copies of neighbouring batches overlap the kernels of the current one."""
        loader = ((host_imgs[i % NBUF], None, None) for i in range(n))
        got, acc = 0, 0.0
        t_prev = time.perf_counter()
        for _bi, s_np, m_np in validator.iter_scores(loader):
            t_now = time.perf_counter()
            e2e_wall.append((t_now - t_prev) * 1e3)
            t_prev = t_now
            assert m_np.shape[0] == s_np.shape[0]  # the caller holds both arrays on the host (pinned staging views)
            acc += float(s_np[0]) + float(m_np[-1, 0, -1, -1])  # touch first and last bytes of what arrived
            got += s_np.shape[0]
        assert got == n * B and acc == acc

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step_device(i)
        torch.cuda.synchronize()
        clocks = ClockSampler(local)
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        clocks.start()
        e0.record()
        for i in range(args.steps):
            step_device(i, timed=True)
        e1.record()
        torch.cuda.synchronize()
        clocks.stop()
        barrier()
        launches = _lib.launch_count() - launches0
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        mdn_ms = statistics.mean(a.elapsed_time(b) for a, b in mdn_events)

        # end to end through the validator API from pinned host memory
        run_e2e(3)
        barrier()
        torch.cuda.synchronize()
        e2e_wall.clear()
        e0.record()
        run_e2e(args.steps)
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))

        # p50 latency of one image (batch 1) through the same device-resident step, synchronising every iteration
        lat_ms = []
        if world == 1:
            one = dev_imgs[0][:1].contiguous()
            g1 = gum[0][:1].contiguous()
            for it in range(10 + 100):
                t0 = time.perf_counter()
                f1 = enc(one)
                prob1, _s1 = head.score(f1.patch_embedding, g1)
                ops.bilinear_up(prob1.view(-1, 14, 14), 224, align_corners=True, post_one_minus=True)
                torch.cuda.synchronize()
                if it >= 10:
                    lat_ms.append((time.perf_counter() - t0) * 1e3)
            # the same step captured once in a CUDA graph (vitad.graphed.GraphedStep) and replayed
            lat_graph_ms, graph_err = [], None
            try:
                from vitad.graphed import GraphedStep

                def one_step(img):
                    f1 = enc(img)
                    prob1, s1 = head.score(f1.patch_embedding, g1)
                    m1, _ = ops.bilinear_up(prob1.view(-1, 14, 14), 224, align_corners=True, post_one_minus=True)
                    return s1, m1

                gstep = GraphedStep(one_step, one)
                for it in range(10 + 100):
                    t0 = time.perf_counter()
                    gstep(one)
                    torch.cuda.synchronize()
                    if it >= 10:
                        lat_graph_ms.append((time.perf_counter() - t0) * 1e3)
            except Exception as e:  # report, do not hide
                graph_err = repr(e)[:200]

        # per-launch-site device times of 3 more steps through the library's event profiler (CUDA events on the launch
        # stream around every launch site; they break programmatic-dependent-launch overlap, so these times are an
        # upper bound of what the kernels cost inside the timed region above)
        import ctypes as C
        _lib.lib.vitad_profile_enable.argtypes = [C.c_int]
        _lib.lib.vitad_profile_report.argtypes = [C.c_char_p, C.c_int]
        _lib.lib.vitad_profile_report.restype = C.c_int
        PSTEPS = 3
        _lib.lib.vitad_profile_enable(1)
        for i in range(PSTEPS):
            step_device(i)
        buf = C.create_string_buffer(1 << 16)
        _lib.lib.vitad_profile_report(buf, len(buf))
        _lib.lib.vitad_profile_enable(0)
        sites = [ln.split() for ln in buf.value.decode().strip().split("\n") if ln and not ln.startswith("__span__")]

    if rank != 0:
        return
    peaks, peak_src = load_peaks()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    mdn_flops = 2.0 * (B * 196) * (2 * K * 768) * 768  # unpadded logical dims
    achieved = mdn_flops / (mdn_ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("gmm_fused_bytes_per_launch")
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    rows_tok = B * 198

    def site_work(name):
        """(bound, algorithmic work per launch, unit) of a launch site; unpadded logical sizes."""
        if name.startswith("gemm_epi"):
            n, k = int(name.split("_n")[1].split("_")[0]), int(name.split("_k")[1].split("_")[0])
            m = B * 196 if name.startswith("gemm_epi4") else rows_tok
            return "tensor", 2.0 * m * n * k / 1e12, "TFLOP/s"
        if name.startswith("attention"):
            return "tensor", 4.0 * B * 12 * 198 * 198 * 64 / 1e12, "TFLOP/s"
        if name == "gmm_fused":
            return "tensor", mdn_flops / 1e12, "TFLOP/s"
        if name == "layernorm":
            return "hbm", rows_tok * 768 * (4 + 2) / 1e9, "GB/s"
        if name == "gmm_mean":
            return "hbm", (768 * B * 196 + B * 196) * 4 / 1e9, "GB/s"
        if name == "gmm_logpi":
            return "fp32", 2.0 * B * 196 * 768 * K / 1e12, "TFLOP/s"
        return None, 0.0, ""

    kernels = []
    for name, cnt, us_total in sites:
        cnt, us_total = int(cnt), float(us_total)
        bound, work, unit = site_work(name)
        us = us_total / cnt
        ent = {"site": name, "launches_per_step": cnt / PSTEPS, "us_per_launch": round(us, 2)}
        if bound:
            ach = work / (us * 1e-6)
            pk = peak if bound == "tensor" else (hbm_peak if bound == "hbm" else None)
            ent.update({"bound": bound, "achieved": round(ach, 1), "unit": unit, "frac": round(ach / pk, 3) if pk else None})
        kernels.append(ent)
    kernels.sort(key=lambda e: -e["us_per_launch"] * e["launches_per_step"])
    out = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "gaussians": K,
                   "weights": "random init of the reference constructors (seeded)",
                   "l2": f"{NBUF} rotating input batches; per-step working set (~0.45 GB of weights) exceeds the 126 MB L2, no flush",
                   "parallelism": f"batch-sharded x{world}, weight replica per rank, no data-path collective"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": B * 3 * 224 * 224 * 4, "d2h_bytes_per_step": B * 4 + B * 224 * 224 * 4,
                "result_interval_ms": {"p50": round(statistics.median(e2e_wall), 3), "max": round(max(e2e_wall), 3)}},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "gmm fused sigma/mu projection + logsumexp (gemm4_tc_kernel<208,1,EpiMdn<104>> on 33 clusters of four + the CTA-pair kernel on the 16 SMs they leave idle) + feature mean",
                     "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src + ", sustained bf16/fp16 dense",
                     "flops_per_launch": mdn_flops, "ms_per_launch": mdn_ms,
                     "encoder_tflops": ENC_GFLOP_PER_IMG * B * 1e9 / ((ms_step - mdn_ms) * 1e-3) / 1e12},
        "kernels": kernels,
    }
    if lat_ms:
        out["latency_bs1_ms"] = {"p50": round(statistics.median(lat_ms), 4), "p90": round(sorted(lat_ms)[89], 4),
                                 "iters": len(lat_ms), "how": "host wall clock around one batch-1 step + device synchronize"}
        if lat_graph_ms:
            out["latency_bs1_ms"]["cuda_graph_p50"] = round(statistics.median(lat_graph_ms), 4)
            out["latency_bs1_ms"]["cuda_graph_p90"] = round(sorted(lat_graph_ms)[89], 4)
        if graph_err:
            out["latency_bs1_ms"]["cuda_graph_error"] = graph_err
    if world == 1 and not args.no_cpu_baseline:
        sample = 8
        ips, _, cores = cpu_reference_images_per_sec(sample, 2, 1, K)
        out["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": f"{sample} images x 2 steps (1 warm-up) of the same workload, oracle port, torch fp32"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--gaussians", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        try:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
