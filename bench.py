#!/usr/bin/env python3
"""Benchmark of the receive DSP chain (fs/4 mix -> FIR pair -> SSB/AM demod -> biquad cascade), BASELINE.json metric:
demodulated Msamples/s summed over all channels, and % of the HBM roofline.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's own C chain on the host cores)

Workload at every N: BASELINE config 3 per GPU — 4096 channels x 10 s at 44.1 kHz (3446 blocks of 128 samples),
mode = {AM, USB, LSB, CW}[c mod 4], the sketch's FIR tables and live biquad cascade, state carried across updates.
One "step" is one pass over that whole batch, fed in updates of --blocks-per-update blocks.  N GPUs = N independent
channel shards (weak scaling, no collective on the data path).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

FS = 44100.0
CHANNELS = 4096
SECONDS = 10.0
BLOCK = 128
ALGO_BYTES_PER_SAMPLE = 4  # 2 B int16 IF in + 2 B int16 audio out


def ncu_traffic():
    """DRAM bytes per launch of the chain kernel from the committed `ncu --set full` capture of this command
    (profiles/rNN_chain_kernel_ncu_metrics.json, newest round), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_chain_kernel_ncu_metrics.json")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch"]), os.path.basename(files[-1])
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled under the bench load (B200_PROFILING.md recipe).  The sampler runs from before
    the warm-up; samples are time-stamped and the ones inside the timed region are reported.  A default run's timed region is
    ~0.1 s (a few sampler periods), so when fewer than 3 samples fall inside it the samples of the identical load right before
    and after it (warm-up steps, untimed cool-down steps) are used as well and `window` says so."""

    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None
        self.t_load0 = self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": None}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[5:9]]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if len(inside) < 3:
            inside = [r for r in rows if self.t_load0 is not None and r[0] >= self.t_load0]
            window = "timed region plus the identical untimed load around it (warm-up, cool-down): the region is shorter than 3 sampler periods"
        if inside:
            reasons = sorted({n for r in inside for n, a in zip(names, r[3]) if a})
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)), reasons=reasons,
                       samples=len(inside), window=window)
        return out


def configure_chain(g, m, n_channels, ch0):
    """mode = {AM,USB,LSB,CW}[c mod 4] with the sketch's tables; biquad1 low-pass + biquad2 notch (the live cascade)."""
    modes = m.synth.mixed_modes(n_channels, ch0)
    g.setup_like_sketch(m.capi.MODE_AM)
    for c, md in enumerate(modes):
        if md != m.capi.MODE_AM:
            g.tune(md, c, 1)
    return modes


def cpu_chain(lib, m, K, modes):
    from chain_helpers import tables_for
    o = lib.chain(len(modes))
    for c, md in enumerate(modes):
        o.set_mode(c, 1, md)
        o.fir_init(c, 1, *tables_for(K, md))
    o.biquad_set_coefficients(0, 0, len(modes), 0, K["biquad1_lowpass_coef"])
    o.biquad_set_coefficients(1, 0, len(modes), 0, K["biquad2_notch_coef"])
    return o


def cpu_checker():
    """oracle/_ref (the reference's own compiled sources) when it travelled, else the oracle port."""
    import oracle_lib as ol
    if ol.have_ref():
        return ol.CheckerLib("ref"), "reference"
    return ol.CheckerLib("orc"), "port"


def time_cpu(m, target_s=8.0, n_threads=0):
    """Reference C chain on the host cores over a bounded sample of the workload. Returns dict for cpu_baseline."""
    lib, kind = cpu_checker()
    K = m.load_ref_constants()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    n_threads = n_threads or cores  # explicit: torchrun exports OMP_NUM_THREADS=1, which would silently serialise the baseline
    n_ch = max(4, 4 * n_threads)
    modes = m.synth.mixed_modes(n_ch)
    x = np.stack([m.synth.channel_stream(c % 16, modes[c], 64 * BLOCK, FS) for c in range(n_ch)])
    o = cpu_chain(lib, m, K, modes)
    t0 = time.perf_counter()
    _, used = o.run(x, n_threads)
    cal = time.perf_counter() - t0  # calibration pass (also warms the threads)
    reps = int(max(1, min(target_s / max(cal, 1e-4), (1 << 28) / x.size)))  # <= 512 MB of input
    xx = np.ascontiguousarray(np.tile(x, (1, reps)))
    runs, t0 = 0, time.perf_counter()
    while True:
        o.run(xx, n_threads)  # state carries on: one long stream per channel
        runs += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or runs >= 64:
            break
    o.close()
    samples = xx.size * runs
    return {"value": samples / dt / 1e6, "unit": "Msamples/s", "cores": int(used), "kind": kind,
            "sample": f"{n_ch} channels x {reps * 64 * runs} blocks ({samples / 1e6:.1f} Msamples, {dt:.1f} s) of the C3 mode mix, "
                      f"{'oracle/_ref: reference CMSIS/Teensy sources' if kind == 'reference' else 'oracle/msdr_oracle.c port'}, gcc -O2, OpenMP",
            "seconds": dt, "samples": samples}


def run_reference(args, rank, world):
    import minimal_sdr_b200 as m
    if rank != 0:
        return
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = time_cpu(m, target_s=max(1.0, min(10.0, 60.0 / max(1, args.steps + args.warmup))))
        if i >= args.warmup:
            vals.append(info)
    tot_s = sum(v["seconds"] for v in vals)
    tot_n = sum(v["samples"] for v in vals)
    value = tot_n / tot_s / 1e6
    cb = {k: info[k] for k in ("unit", "cores", "kind", "sample")}
    cb["value"] = value
    line = {"impl": "reference", "metric": "demodulated Msamples/s (all channels)", "value": value, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "q15/q31 fixed point (int16 data, int32 accumulate)",
            "data": "synthetic", "config": workload_config(args), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    nb = int(round(SECONDS * FS)) // BLOCK + (1 if int(round(SECONDS * FS)) % BLOCK else 0)
    return {"workload": "C3: batched 4096 channels x 10 s mixed AM/SSB/CW per GPU, state carried across 128-sample blocks",
            "channels_per_gpu": args.channels, "blocks_per_channel": nb, "blocks_per_update": args.blocks_per_update,
            "fs_hz": FS, "modes": "{AM,USB,LSB,CW}[c mod 4]", "fir": "sketch tables: AM 102 taps (bw 2800 @ 24 kHz design, used as-is), SSB/CW 86 taps",
            "biquads": "biquad1 low-pass (0.9*IF, Q 0.54) + biquad2 notch (fs/8, Q 15), integer Q2.30",
            "l2": "inputs (3.6 GB per step) and outputs far exceed the 126 MB L2; no explicit flush", "sharding": "independent channel shards per GPU",
            "layout": getattr(args, "layout", "updates"),
            **({"with_frontend": "raw 12-bit ADC codes through msdr_frontend_update_device (DC-block, amplifier, AGC) before every chain update"}
               if getattr(args, "with_frontend", False) else {})}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=CHANNELS)
    ap.add_argument("--blocks-per-update", type=int, default=1024,
                    help="128-sample blocks per msdr_chain_update_device call (state is carried from call to call); 1024 = 2.97 s of signal")
    ap.add_argument("--seconds", type=float, default=SECONDS)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--with-frontend", action="store_true",
                    help="study: feed raw 12-bit ADC codes through the front-end conditioning kernel (DC-block, amplifier, AGC; SURVEY 8f rank 1) "
                         "in front of every chain update; the default line measures the north-star path only")
    ap.add_argument("--layout", choices=["updates", "rows"], default="updates",
                    help="device-resident input layout: 'updates' = one contiguous [channels][samples] batch per update (how a streaming "
                         "receiver holds its block batches), 'rows' = one 10 s row per channel, updates are column windows of it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import minimal_sdr_b200 as m

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    C = args.channels
    n_samples = int(round(args.seconds * FS))
    nb_total = (n_samples + BLOCK - 1) // BLOCK
    L = nb_total * BLOCK
    bpu = max(1, min(args.blocks_per_update, nb_total))
    shard = m.shard.weak_scaling_shard(C, world, rank)  # this rank's slice of the global channel space
    ch0 = shard.ch0

    g = m.ReceiveChain(C, device=local_rank)
    g.set_option("variant", args.variant)
    configure_chain(g, m, C, ch0)
    x = m.synth.torch_batch(C, L, dev, FS, ch0=ch0)
    y = torch.empty_like(x) if args.layout == "rows" else None
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream: kernels and timing events share it
    assert stream.cuda_stream != 0
    g.set_stream(stream.cuda_stream)
    updates = [(b0, min(bpu, nb_total - b0)) for b0 in range(0, nb_total, bpu)]
    if args.layout == "rows":
        stride = x.stride(0)
        calls = [(x.data_ptr() + 2 * b0 * BLOCK, y.data_ptr() + 2 * b0 * BLOCK, nb, stride) for b0, nb in updates]
    else:  # one contiguous batch per update; same samples, same order of processing
        xs = [x[:, b0 * BLOCK:(b0 + nb) * BLOCK].contiguous() for b0, nb in updates]
        ys = [torch.empty_like(t) for t in xs]
        calls = [(xi.data_ptr(), yi.data_ptr(), nb, xi.stride(0)) for (b0, nb), xi, yi in zip(updates, xs, ys)]
        if args.e2e_steps <= 0:
            del x
        torch.cuda.empty_cache()

    fe = None
    if args.with_frontend:
        fe = m.Frontend(C, device=local_rank)
        fe.set_stream(stream.cuda_stream)
        adc = [(xi.to(torch.int32) // 16 + 2048).clamp_(0, 4095).to(torch.int16) for xi in (xs if args.layout != "rows" else [x])]
        if args.layout == "rows":
            adc_calls = [adc[0].data_ptr() + 2 * b0 * BLOCK for b0, nb in updates]
        else:
            adc_calls = [a.data_ptr() for a in adc]
        torch.cuda.synchronize()

    def step():
        for i, (d_in, d_out, nb, stride) in enumerate(calls):
            if fe is not None:  # raw codes -> conditioned IF samples, written over the chain's input buffer
                fe.update_device(adc_calls[i], d_in, nb, stride)
            g.update_device(d_in, d_out, nb, stride)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    step()  # first touch (plan upload, allocations) before the sampler's load window
    barrier()
    clocks.t_load0 = time.time()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = g.launch_count()
    fe_l0 = fe.launch_count() if fe is not None else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()  # cudaProfilerStart: `ncu --profile-from-start off` sees exactly the timed region
    clocks.t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks.t1 = time.time()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = g.launch_count() - l0 + ((fe.launch_count() - fe_l0) if fe is not None else 0)
    if clocks.t1 - clocks.t0 < 0.5:  # untimed cool-down under the same load so that the clock sampler sees it
        t_end = time.time() + 0.5
        while time.time() < t_end:
            step()
            torch.cuda.synchronize()
    clk = clocks.stop()
    ms = m.shard.max_over_ranks(ms, dev)  # multi-GPU timing rule: slowest rank

    samples_per_step = C * L
    value = world * samples_per_step * args.steps / (ms * 1e-3) / 1e6  # Msamples/s, whole job

    # ---- e2e: the user-facing call with HOST (pinned) buffers; H2D and D2H inside the timed region
    e2e = None
    if args.e2e_steps > 0:
        hin = m.capi.PinnedBuffer((C, L))
        hout = m.capi.PinnedBuffer((C, L))
        hin.array[:] = x.cpu().numpy()
        g.update(hin.array, out=hout.array)  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            g.update(hin.array, out=hout.array)
        barrier()
        dt = time.perf_counter() - t0
        dt = m.shard.max_over_ranks(dt, dev)
        e2e = {"value": world * samples_per_step * args.e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(samples_per_step * 2), "d2h_bytes_per_step": int(samples_per_step * 2), "steps": args.e2e_steps,
               "api": "msdr_chain_update (C ABI, pinned host buffers, whole batch per call)"}
        checksum = int(hout.array[:: max(1, C // 64), ::4096].astype(np.int64).sum())
        e2e["result_checksum"] = checksum
        hin.free(); hout.free()

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_s = (ms * 1e-3) / max(1, launches)
        samples_per_launch = samples_per_step * args.steps / max(1, launches)
        achieved = ALGO_BYTES_PER_SAMPLE * samples_per_launch / per_launch_s / 1e9
        # second ceiling (SURVEY 8d asks for the issue-side fraction next to the HBM one): a channel group of 32 advances one
        # sample per biquad sample-step of its chain warps, 43.5 cycles isolated (profiles/r01_microbench_lat.txt), and an SM
        # hosts one group (two when there are more groups than SMs)
        import torch
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        groups = (C + 31) // 32
        concurrent = groups if groups <= sms else min(groups, 2 * sms)
        f_hz = float((clk or {}).get("sm_mhz") or 1965.0) * 1e6
        rec_ceiling = min(C, concurrent * 32) * f_hz / 43.5 / 1e6
        recurrence = {"bound": "serial biquad recurrence, issue-bound inside its warp", "cycles_per_sample_step": 43.5,
                      "concurrent_channel_groups": int(concurrent), "ceiling": rec_ceiling, "unit": "Msamples/s per GPU",
                      "frac": value / world / rec_ceiling, "source": "tools/microbench/bqstep2.cu, lat.cu; DESIGN.md section 6"}
        line = {
            "metric": "demodulated Msamples/s (all channels)", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "q15/q31 fixed point (int16 data, int32 wrapping accumulate, Q2.30 biquad)", "data": "synthetic",
            "config": workload_config(args), "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic()[0],
                         "traffic_source": ncu_traffic()[1], "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * samples_per_launch,
                         "peak_source": peak_src, "kernel": "msdr::v4::chain_kernel (fused mix + tensor-core FIR pair + demod + biquad cascade)",
                         "algorithmic_bytes_per_sample": ALGO_BYTES_PER_SAMPLE, "samples_per_launch": samples_per_launch,
                         "avg_launch_ms": per_launch_s * 1e3, "recurrence": recurrence,
                         "note": "not HBM-bound: 4096 channels are 128 biquad chains, each an exact-arithmetic serial recurrence (34 cycles per sample in the recurrence warps, feed-forward products in helper warps); the FIR pair runs beside them on the tensor cores (tcgen05 kind::i8); see DESIGN.md section 6"},
            "e2e": e2e,
        }
        if world == 1 and not args.no_cpu:
            cb = time_cpu(m)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
