#!/usr/bin/env python
"""bench.py -- scene-flow frames/sec of the SeFlow++ hot path on synthetic 100k-point frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

A step = one pass of the hot path over one frame triple (t-1, t, t+1; BASELINE.json configs[1]:
"SeFlow++ inference, synthetic Scania-shaped 100k-pt frame-pairs").  Prints ONE JSON line (rank 0).
  value   frames/s, whole job, inputs resident in HBM (device-timed, max over ranks)
  e2e     frames/s through the public API (himo_b200.engine.SeFlowPPEngine.infer) with pinned HOST
          buffers: H2D of the three clouds and D2H of the per-point flow inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"
  knn / voxelize / fastnsf / sustained / pipeline: the other rows of the hot path (bench_extras.py), every rank its own copy
--impl reference times the CPU restatement of the reference's path (oracle/deflowpp_ref.py, torch CPU,
all host threads) on the same config; /root/reference itself cannot run here (no hydra/lightning/h5py
and CUDA-only native ops) and does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_POINTS = 100_000
METRIC = "scene-flow frames/sec on 100k-pt pairs"
UNIT = "frames/s"
# one config dict for BOTH arms (the driver compares them textually): BASELINE.json configs[1]
CONFIG = {"workload": "SeFlow++ (DeFlowPP) inference, synthetic Scania-shaped 100k-pt frame triples, 1 frame/step/GPU",
          "n_points": N_POINTS, "grid": "512x512 pillars, 3 frames", "precision": "fp32",
          "l2": "per-step working set ~1.2 GB of activations > 126 MB L2 (no flush needed)"}


def backbone_flops(composed: bool = False) -> float:
    """Algorithmic FLOPs (2*MAC) of UNetThreeFrame on 3x[32,512,512] (OSF/src/models/basic/unet.py:101-166).
    composed=True: the FLOPs the kernels EXECUTE when u3 (1x1 on the skip) is folded into u4 on the host."""
    from himo_b200.weights import DECODER_BLOCKS, ENCODER_LAYERS
    fl = 0.0
    res = 512
    for _, cin, cout, stride in ENCODER_LAYERS:
        res = res // stride
        fl += 3 * 2.0 * res * res * cout * cin * 9
    low = 64
    for _, skip, latent, out in DECODER_BLOCKS:
        hi = low * 2
        fl += 2.0 * low * low * latent * skip            # u1 1x1 @ low res
        if not composed:
            fl += 2.0 * hi * hi * latent * latent        # u3 1x1 on the skip (skip channels == latent)
        fl += 2.0 * hi * hi * out * (2 * latent) * 9     # u4
        fl += 2.0 * hi * hi * out * out * 9              # u5
        low = hi
    fl += 2.0 * 512 * 512 * 96 * 96 * 9                  # decoder_step4
    return fl


def decoder_flops(n_points: int, iters: int = 2) -> float:
    per_pt = iters * 3 * 2.0 * 288 * 192 + 2.0 * 288 * 48 + 2.0 * 48 * 3 + 2.0 * 3 * 96
    return per_pt * n_points


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Index of the next sample: brackets the timed region inside an already running sampler."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        window = self.rows[first:last] if last is not None else self.rows[first:]
        if len(window) < 2:          # very short timed region: widen to the neighbouring samples (still under load)
            window = self.rows[max(0, first - 2):(last + 2 if last is not None else None)]
        self.rows = window
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def make_frames(rank: int, count: int = 2):
    from himo_b200 import frames
    out = []
    for k in range(count):
        tr = frames.lidar_triple(N_POINTS, seed=1000 * 2 + 17 * rank + k, t=1.0 + 0.3 * k)
        fr = {"pc0": tr["pc0"], "pc1": tr["pc1"], "pch1": tr["pch1"], "pose0": tr["pose0"], "pose1": tr["pose1"],
              "poseh1": tr["poseh1"], "frames": tr["frames"]}
        out.append(fr)
    return out


def run_reference(args):
    """CPU arm: the oracle restatement of DeFlowPP.forward + final-flow packing, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from himo_b200 import weights
    from oracle import deflowpp_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = weights.synth_deflowpp_state_dict(0)
    frames_ = make_frames(0, 1)
    fr = frames_[0]

    def step():
        res = deflowpp_ref.deflowpp_forward(sd, fr["pch1"], fr["pc0"], fr["pc1"], fr["poseh1"], fr["pose0"], fr["pose1"])
        gm = torch.zeros(fr["pc0"].shape[0], dtype=torch.bool)
        return deflowpp_ref.final_flow(torch.from_numpy(fr["pc0"]), gm, fr["pose0"], fr["pose1"], res)

    t0 = time.perf_counter()
    step()
    t_first = time.perf_counter() - t0
    warm = max(0, min(args.warmup, 1 if t_first > 2 else args.warmup) - 1)
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, int(150.0 / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    fps = steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "reference_impl": "CPU restatement of the reference path (oracle/deflowpp_ref.py, torch CPU fp32, all host threads); "
                              "/root/reference itself needs hydra / lightning / h5py and CUDA-only native ops",
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} frame triple(s) of the bench workload"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the knn / voxelize / fastnsf / sustained / pipeline keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from himo_b200 import _lib, weights
    from himo_b200.deflowpp import cal_pose0to1
    from himo_b200.engine import SeFlowPPEngine

    sd = weights.synth_deflowpp_state_dict(0)
    eng = SeFlowPPEngine(sd, device=dev, precision=args.precision, max_points=N_POINTS)
    net = eng.net
    host_frames = make_frames(rank, 2)
    L = _lib.lib()
    if os.environ.get("HIMO_PDL") is not None:          # A/B knob (profiles/): programmatic dependent launch on / off
        L.himo_conv_set_pdl(int(os.environ["HIMO_PDL"]))

    # ---- device-resident copies for the kernel-only arm
    dev_frames = []
    for fr in host_frames:
        d = {k: torch.from_numpy(fr[k]).to(dev) for k in ("pc0", "pc1", "pch1")}
        d["T0"] = cal_pose0to1(torch.from_numpy(fr["pose0"]), torch.from_numpy(fr["pose1"]))
        d["Th"] = cal_pose0to1(torch.from_numpy(fr["poseh1"]), torch.from_numpy(fr["pose1"]))
        d["T12"] = d["T0"][:3, :4].contiguous().float().flatten().to(dev)
        dev_frames.append(d)
    # the engine's slots: one network replica (shared weights, own workspace) + compute stream each.  Step i runs on slot
    # i % n, so that n frame triples are in flight and the kernels of one fill the tail waves of the others -- the same
    # arrangement `infer_stream` (the e2e arm, save.py) uses.
    lanes = [(s.net, s.stream, torch.empty((N_POINTS + 16, 3), dtype=torch.float32, device=dev),
              torch.empty((N_POINTS + 16, 3), dtype=torch.float32, device=dev)) for s in eng._slots]
    in_flight = [len(lanes)]

    def step_resident(i, events=None):
        d = dev_frames[i % len(dev_frames)]
        lane_net, lane_stream, flow_buf, final_buf = lanes[0 if events is not None else i % in_flight[0]]
        with torch.cuda.stream(lane_stream):
            out = lane_net.forward_triple(d["pch1"], d["pc0"], d["pc1"], d["Th"], d["T0"], compact=False, stage_events=events,
                                          flow_all_out=flow_buf)
            st = L.himo_final_flow(_lib.ptr(d["pc0"]), d["pc0"].shape[0], _lib.ptr(d["T12"]), _lib.ptr(out["flow_all"]),
                                   None, _lib.ptr(final_buf), _lib.stream_ptr(dev))
        _lib.check(st, "himo_final_flow")

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for lane in lanes:                        # the lanes' streams start after e0 ...
            lane[1].wait_event(e0)
        for i in range(steps):
            fn(i)
        for lane in lanes:                        # ... and e1 is recorded once all of them have drained
            torch.cuda.current_stream().wait_stream(lane[1])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- warm-up, then the timed kernel-only region
    sampler = ClockSampler(local_rank)
    sampler.start()                      # started before the warm-up so that it is already streaming samples
    for i in range(args.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    for _ in range(200):                 # nvidia-smi takes a moment to produce its first row
        if sampler.mark() > 0:
            break
        step_resident(0)
        torch.cuda.synchronize()
        time.sleep(0.01)
    mark0 = sampler.mark()
    launches0 = L.himo_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = (L.himo_launch_count() - launches0)
    value = world * args.steps / (ms_total / 1e3)
    in_flight[0] = 1                     # the same K steps one at a time on one stream (what round 1 and the ncu lists time),
    pdl_env = os.environ.get("HIMO_PDL")  # with programmatic dependent launch on, as a single-stream user would run it
    if pdl_env is None:
        L.himo_conv_set_pdl(1)
    for i in range(3):
        step_resident(i)
    ms_single = timed(step_resident, args.steps)
    if pdl_env is None:
        L.himo_conv_set_pdl(0 if len(lanes) > 1 else 1)
    in_flight[0] = len(lanes)

    # ---- end-to-end arm: host buffers -> public API -> host result
    def run_e2e(steps):
        # the streaming form of the public API: host frames in, host flow arrays out, copies of the
        # neighbouring frames overlapped with the network on separate CUDA streams
        n = 0
        for _ in eng.infer_stream(host_frames[i % len(host_frames)] for i in range(steps)):
            n += 1
        assert n == steps
    run_e2e(args.warmup)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record(); torch.cuda.synchronize()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # device span, never below host wall
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    barrier()
    e2e_value = world * args.steps / (ms_e2e / 1e3)
    mark1 = sampler.mark()
    h2d, d2h = eng.h2d_bytes, eng.d2h_bytes

    # ---- per-stage split (one profiled pass per frame; CUDA events on the launching stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for e in ev:
        e.record()
    torch.cuda.synchronize()
    stage = np.zeros(3)
    reps = max(3, min(10, args.steps))
    if pdl_env is None:
        L.himo_conv_set_pdl(1)           # one network alone on its stream, as in the single_stream figure
    for i in range(reps):
        step_resident(i, ev)
        torch.cuda.synchronize()
        stage += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
    stage /= reps
    if pdl_env is None:
        L.himo_conv_set_pdl(0 if len(lanes) > 1 else 1)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1590.0 * 0.88))

    # ---- the other rows of the hot path (every rank runs its own copy: weak scaling)
    extras = {}
    if not args.no_extras:
        import bench_extras as X
        fr0 = host_frames[0]
        for name, fn in (("sustained", lambda: X.sustained(step_resident, sampler, 3.0)),
                         ("knn", lambda: X.knn(dev, (fr0["pc0"], fr0["pc1"]), hbm_gbs, rank)),
                         ("voxelize", lambda: X.voxelize(dev, hbm_gbs, rank)),
                         ("fastnsf", lambda: X.fastnsf(dev, fr0, peak_burst)),
                         ("pipeline", lambda: X.pipeline(eng, host_frames, rank, local_rank))):
            barrier()
            try:
                extras[name] = fn()
            except Exception as e:                # a failing extra must not cost the headline line
                extras[name] = {"error": repr(e)[:300]}
        barrier()
    clocks = sampler.stop(mark0, mark1)

    def reduce_extras(ex):
        """Whole-job view over the ranks: throughputs summed, times / fractions maxed (slowest rank)."""
        if world == 1 or not ex:
            return ex
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, ex)
        if rank != 0:
            return ex
        SUM = ("frames_per_s", "save_frames_per_s", "pairs_per_s", "algorithmic_gbs", "algorithmic_tflops", "frames", "steps")

        def walk(vals, key=None):
            v0 = vals[0]
            if isinstance(v0, dict):
                return {k: walk([v[k] for v in vals if isinstance(v, dict) and k in v], k) for k in v0}
            if isinstance(v0, bool) or not isinstance(v0, (int, float)):
                return v0
            nums = [v for v in vals if isinstance(v, (int, float))]
            return sum(nums) if key in SUM else max(nums)
        out = walk(gathered)
        out["reduction"] = f"{world} ranks: " + ", ".join(SUM) + " summed over ranks, every other number is the max over ranks"
        return out
    extras = reduce_extras(extras)

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    composed = bool(getattr(net, "compose_skip", False))
    fl_alg, fl_exec = backbone_flops(False), backbone_flops(composed)
    # With several networks in flight the launches of different frames overlap, so a kernel's own duration is not separable
    # inside the timed region: the backbone's time per step is taken as (timed ms per step) x (its share of a step that runs
    # alone on one stream, CUDA events at the stage boundaries).  `single_stream` repeats the direct form: FLOPs / stage time.
    share = stage[1] / max(stage.sum(), 1e-9)
    t_back = (ms_total / args.steps) * share / 1e3
    achieved = fl_exec / t_back / 1e12
    achieved_single = fl_exec / (stage[1] / 1e3) / 1e12
    # the timed region is tens of milliseconds at the boost clock: the burst cuBLAS figure is the denominator
    # (VERDICT r01); the sustained one applies to the >= 3 s loop reported under "sustained"
    peak_tf = peak_burst
    peak_src = ("MEASURED_PEAKS.json bf16_tflops (measured, burst: the backbone stage is timed over a sub-second region)"
                if peaks else "B200_PROFILING.md fallback 1590 TFLOP/s")
    traffic, traffic_note = None, "no ncu capture for this build of csrc/conv.cu (profiles/r02_backbone_traffic.json absent or stale)"
    try:          # dram__bytes_read.sum + dram__bytes_write.sum of the conv launches of one step (ncu capture), valid only
        import hashlib   # for the kernel source it was captured with
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_backbone_traffic.json")))
        sha = hashlib.sha256(open(os.path.join(ROOT, "himo_b200", "csrc", "conv.cu"), "rb").read()).hexdigest()[:16]
        if tj.get("conv_cu_sha16") == sha:
            traffic = float(tj["dram_bytes_read_per_step"]) + float(tj["dram_bytes_write_per_step"])
            traffic_note = "DRAM bytes per step over the backbone launches (ncu --set full, profiles/r02_backbone_traffic.json, same conv.cu)"
    except (OSError, KeyError, ValueError):
        pass
    mma_mult = 3 if args.precision == "fp32" else 1
    cfg = dict(CONFIG)
    cfg["precision"] = args.precision
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (split-fp16 operands x3 MMAs, fp32 accumulate)" if args.precision == "fp32" else "bf16",
        "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "in_flight": {"networks": len(lanes),
                      "note": "value and e2e run step i on slot i % n (own workspace + stream, shared weights); "
                              "single_stream = the same K steps back to back on one stream",
                      "single_stream": {"value": world * args.steps / (ms_single / 1e3), "ms_per_step": ms_single / args.steps}},
        "clocks": clocks,
        "stages_ms": {"embedder": stage[0], "backbone": stage[1], "decoder": stage[2]},
        "roofline": {"bound": "tensor", "kernel": "k_conv_umma / k_conv_rows2 / k_conv_wide (the convolution launches of a step; the "
                                                   "stage time also holds the 3 bilinear upsamples, ~4 % of it)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": traffic, "traffic_unit": traffic_note,
                     "peak_source": peak_src,
                     "executed_gflop_per_step": fl_exec / 1e9, "algorithmic_gflop_per_step": fl_alg / 1e9,
                     "flops_note": "achieved = EXECUTED FLOPs / stage time; the reference formulation has %.1f GFLOP more "
                                   "(the three 1x1 u3 convolutions, folded into u4 on the host)" % ((fl_alg - fl_exec) / 1e9),
                     "time_basis": "timed ms_per_step x backbone share of a stand-alone step (%.3f)" % share,
                     "single_stream": {"achieved": achieved_single, "frac": achieved_single / peak_tf,
                                       "note": "executed FLOPs / backbone stage time of a step alone on one stream"},
                     "tensor_issue_multiplier": mma_mult,
                     "tensor_issue_frac": achieved * mma_mult / peak_tf,
                     "frac_of_sustained_peak": achieved / peak_sust},
    }
    if extras:
        if isinstance(extras.get("sustained"), dict) and "ms_per_step" in extras["sustained"]:
            sus = extras["sustained"]
            sus["backbone_tflops_at_stage_share"] = fl_exec / (sus["ms_per_step"] * 1e-3 * stage[1] / max(stage.sum(), 1e-9)) / 1e12
            sus["frac_of_sustained_peak"] = sus["backbone_tflops_at_stage_share"] / peak_sust
        line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        from oracle import deflowpp_ref
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        fr = host_frames[0]
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = deflowpp_ref.deflowpp_forward(sd, fr["pch1"], fr["pc0"], fr["pc1"], fr["poseh1"], fr["pose0"], fr["pose1"])
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": 1.0 / statistics.median(ts), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "3 frame triples of the bench workload (median), torch CPU fp32"}
        # parity of this very run against the CPU path (EPE vs reference, north_star)
        got = eng.infer(fr)
        gm = torch.zeros(fr["pc0"].shape[0], dtype=torch.bool)
        ref = deflowpp_ref.final_flow(torch.from_numpy(fr["pc0"]), gm, fr["pose0"], fr["pose1"], res).numpy()
        line["epe_vs_cpu_reference"] = {"max_abs": float(np.abs(got - ref).max()),
                                        "mean_epe": float(np.linalg.norm(got - ref, axis=1).mean())}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
