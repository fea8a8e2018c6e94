#!/usr/bin/env python
"""bench.py -- headline benchmark of the partitioned-convolution hot path.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): S independent mono 48 kHz
streams sharing one 2 s synthetic IR (96 000 taps -> P = 188 partitions), block B = 512.  One STEP = one UPOLA
block step for all S streams of a rank: forward FFT of the new block into the frequency-domain delay line (FDL),
multiply-accumulate over all P partitions, inverse FFT, overlap-add.

  value   real-time 48 kHz channels sustained = S * B / (step time) / 48000, inputs resident in HBM
  e2e     the same through irb_engine_process(): HOST (pinned) buffers, H2D and D2H inside the timed region
  roofline  the FDL-MAC kernel: algorithmic bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline / --impl reference: the reference's own fp::convolution::convolvePeriodic (oracle/_ref, compiled
          unmodified) on the box's host cores, one whole stream per thread

Launch: `python bench.py --gpus 1` or, for N > 1, under torchrun (one rank per GPU, streams sharded by rank,
no data-path collective; torch.distributed is used only for the barrier and the max-over-ranks time).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000.0
METRIC = "48kHz RT channels per B200 (block 512, 2s IR)"
UNIT = "channels"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=65536, help="streams per GPU (configs[2] names 1024; >= 10k is the north-star target; "
                                                                "65536 resident streams keep p99 well under the block period)")
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--ir-seconds", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the post-run comparison of the measured block step with its two-launch form")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: min(steps, 32))")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the bounded CPU sample")
    ap.add_argument("--two-launch", action="store_true", help="block step as k_fwd + k_mac instead of the fused single launch (A/B measurement)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c4"],
                    help="c3 (default, the headline line): streams sharing one IR.  c4: BASELINE configs[3], --streams-total streams with PER-STREAM "
                         "10 s IRs, block 1024, sharded by stream over the ranks (strong scaling; not the headline metric)")
    ap.add_argument("--streams-total", type=int, default=8192, help="c4: streams of the whole job")
    a = ap.parse_args()
    if a.workload == "c4":
        a.block, a.ir_seconds = 1024, 10.0
    return a


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload(args):
    B = args.block
    Lh = int(round(args.ir_seconds * SR))
    P = int(np.ceil(np.float32(Lh) / np.float32(B)))
    return B, Lh, P


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_sample(args, target_seconds, threads=None, probe=None):
    """The reference's convolvePeriodic (oracle/_ref; the C port if _ref did not travel) on the host cores:
    `threads` workers, one whole mono stream each, same IR / block size as the GPU workload."""
    import oracle
    from irbaboon_b200 import synth
    B, Lh, P = workload(args)
    h = synth.decaying_ir(2000, Lh)
    if oracle.have_reference():
        ref = oracle.Reference()
        T = threads or max(1, ref.hardware_threads())
        if probe is None:
            secs, _ = ref.bench_convolve_periodic(T, T, int(SR // 4), h, B, 1003)          # probe: 0.25 s of audio per stream
            probe = max(secs, 1e-3) / 0.25
        per_audio_second = probe                                                            # wall seconds per audio second with T streams on T threads
        # T..8T streams of up to 10 s each: about target_seconds of wall time on all host threads
        audio_s = float(np.clip(target_seconds / per_audio_second, 0.5, 10.0))
        rounds = int(np.clip(target_seconds / (audio_s * per_audio_second), 1, 8))
        Lx = int(audio_s * SR) // B * B
        streams = T * rounds
        secs, chk = ref.bench_convolve_periodic(T, streams, Lx, h, B, 1003)
        kind = "reference"
    else:                                                                                   # scalar C port, 1 thread
        orc = oracle.Oracle()
        T, streams = 1, 1
        Lx = int(2 * SR) // B * B
        x = synth.white_noise(1003, 0, Lx)
        t0 = time.perf_counter()
        orc.convolve_periodic(x, h, B)
        secs, chk, kind = time.perf_counter() - t0, 0.0, "port"
    rt = streams * (Lx / SR) / secs
    return {"value": rt, "unit": UNIT, "cores": T, "kind": kind, "seconds": secs, "probe": probe,
            "sample": "%d streams x %.2f s white noise each through fp::convolution::convolvePeriodic(B=%d, %d-tap IR), %d threads, one stream per thread"
                      % (streams, Lx / SR, B, Lh, T)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every step is a bounded sample of the workload; the whole run (warm-up included) is sized for about two minutes
    budget = 120.0 / max(1, args.steps + args.warmup)
    per_step = float(np.clip(budget, 0.25, args.cpu_seconds))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(args, per_step, probe=last["probe"] if last else None)
        if i >= args.warmup:
            vals.append(last)
    secs = sum(v["seconds"] for v in vals)
    rt = float(np.mean([v["value"] for v in vals]))
    B, Lh, P = workload(args)
    line = {"metric": METRIC, "value": rt, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(1, len(vals)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "configs[2]: streams sharing one %.1f s IR, block %d (CPU: %s)" % (args.ir_seconds, B, last["sample"]),
                                            "block": B, "ir_taps": Lh, "partitions": P},
            "cpu_baseline": {"value": rt, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": rt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from irbaboon_b200 import engine as eng
    from irbaboon_b200 import sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator is created; stdout must carry one JSON line, so the
        # file descriptor points at stderr while the process group comes up (first barrier included)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, Lh, P = workload(args)
    per_stream_ir = args.workload == "c4"
    S = args.streams
    if per_stream_ir:                                  # strong scaling: the job's streams are cut into contiguous ranges
        lo, hi = sharding.stream_range(rank, world, args.streams_total)
        S = hi - lo
    bins = B + 1
    if per_stream_ir:
        e = eng.Engine(B, P, S, S, device=local)
        irs = [synth.decaying_ir(2000 + j, Lh, j) for j in range(8)]          # 8 distinct IRs cycled: every stream still owns its spectra
        for s_ in range(S):
            e.set_ir(s_, irs[(lo + s_) % 8])
            e.bind(s_, s_ + 1, s_)
    else:
        h = synth.decaying_ir(2000, Lh)
        e = eng.Engine(B, P, S, 1, device=local)
        e.set_ir(0, h)
    if args.two_launch:
        e.set_fused_step(False)
    stream = torch.cuda.Stream()                       # the kernels and the timing events share this stream
    torch.cuda.set_stream(stream)
    e.set_stream(stream.cuda_stream)

    # synthetic input: NBUF distinct white-noise blocks per stream, resident in HBM, cycled step by step
    NBUF = 4
    gen = torch.Generator(device="cuda").manual_seed(1003 + rank)
    d_in = (torch.rand((NBUF, S, B), device="cuda", generator=gen, dtype=torch.float32) * 2 - 1).contiguous()
    d_out = torch.empty((S, B), device="cuda", dtype=torch.float32)

    def step(i):
        e.process_device(d_in[i % NBUF].data_ptr(), d_out.data_ptr(), 1)

    # fill the FDL (steady state needs P blocks of history) + the requested warm-up
    for i in range(max(args.warmup, 3) + P):
        step(i)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count()
    e.set_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    step_ms, mac_ms = e.timings()
    e.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None

    total_ms = sharding.max_over_ranks(total_ms, device="cuda")
    ms_per_step = total_ms / args.steps
    # final host-side gather of the last output block in global stream order (outside the timed region; the only
    # cross-rank data movement of the whole job)
    total_streams = args.streams_total if per_stream_ir else world * S
    gathered = sharding.gather_streams(d_out.cpu().numpy(), total_streams, device="cuda")
    value = total_streams * B / (ms_per_step * 1e-3) / SR

    # roofline of the dominant kernel (FDL MAC, fused with the inverse FFT + overlap-add epilogue)
    # FDL read per stream + the shared IR once per GPU (c3) or every stream's own IR spectra (c4)
    # With the fused step (default for c3) the kernel is the WHOLE block step: add the new spectrum's write and the audio in/out
    # (SURVEY 8d bytes_blk = bytes_mac + (B+1)*8 + 2*B*4 per stream).
    fused = (not per_stream_ir) and (not args.two_launch) and not e.mac_plan()[0]
    alg_bytes = 2 * S * P * bins * 8 if per_stream_ir else (S + 1) * P * bins * 8
    if fused:
        alg_bytes += S * (bins * 8 + 2 * B * 4)
    mac_avg_ms = float(np.mean(mac_ms)) if len(mac_ms) else float("nan")
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (mac_avg_ms * 1e-3) / 1e9
    tma = fused and e.fft_size // 2 >= 256 and os.environ.get("IRB_MAC_TMA", "1") != "0"      # the dispatch rule of launch_mac_t (irb_engine.cu)
    kname = "k_mac_slots<%d,INV> (per-stream-IR FDL multiply-accumulate + inverse FFT + overlap-add)" % (e.fft_size // 2) if per_stream_ir else \
            ("k_mac_tma<%d> (one launch per block step: forward FFT + TMA-streamed FDL multiply-accumulate + inverse FFT + overlap-add)" if tma else
             "k_mac<%d,INV,FUSE> (one launch per block step: forward FFT + FDL multiply-accumulate + inverse FFT + overlap-add)" if fused else
             "k_mac<%d,INV> (FDL multiply-accumulate + inverse FFT + overlap-add)") % (e.fft_size // 2)
    roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": mac_avg_ms, "kernel_share_of_step": mac_avg_ms / float(np.mean(step_ms)),
            "traffic": None}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            if tj.get("streams") == S and tj.get("block") == B and tj.get("partitions") == P and bool(tj.get("fused", False)) == fused \
                    and ("k_mac_tma" in tj.get("kernel", "")) == tma:
                roof["traffic"] = tj["dram_bytes_per_launch"]
        except Exception:
            pass
    period_ms = 1e3 * B / SR
    lat = {"block_period_ms": period_ms, "p50_ms": float(np.percentile(step_ms, 50)), "p99_ms": float(np.percentile(step_ms, 99)),
           "max_ms": float(step_ms.max()), "realtime": bool(np.percentile(step_ms, 99) < period_ms)}

    # ---- e2e: host (pinned) buffers through irb_engine_process, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        K2 = args.e2e_steps or min(args.steps, 32)
        K2 = max(4, min(K2, (2 << 30) // (S * B * 4)))          # at most ~2 GB of pinned host memory per direction
        if per_stream_ir:
            K2 = min(K2, 8)
        hin = eng.pinned_empty((K2, S, B))
        hout = eng.pinned_empty((K2, S, B))
        rng = np.random.default_rng(1003 + rank)
        hin[:] = rng.random((K2, S, B), dtype=np.float32) * 2 - 1
        e.set_stream(None)
        e.process(hin[:4], hout[:4])                      # warm-up of the copy path
        barrier()
        REP = 4                                           # the feed is continuous: REP back-to-back submissions of K2 blocks, one wait
        t0 = time.perf_counter()
        for _ in range(REP):
            e.submit(hin, hout)                           # copies and kernels enqueued; the pipeline stays full across submissions
        e.wait()                                          # every output block is back on the host
        dt = (time.perf_counter() - t0) / REP
        dt = sharding.max_over_ranks(dt, device="cuda")
        # strict block-by-block round trip (what a live host callback sees)
        t0 = time.perf_counter()
        for i in range(min(K2, 16)):
            e.process(hin[i], hout[i])
        dt1 = (time.perf_counter() - t0) / min(K2, 16)
        e2e = {"value": total_streams * B * K2 / dt / SR, "unit": UNIT, "h2d_bytes_per_step": S * B * 4, "d2h_bytes_per_step": S * B * 4,
               "steps": K2, "ms_per_step": 1e3 * dt / K2, "api": "4 x irb_engine_submit(host in, host out, n_blocks=%d) + irb_engine_wait, pinned buffers, wall clock" % K2,
               "blockwise_roundtrip_ms": 1e3 * dt1, "checksum": float(np.abs(hout[-1]).sum())}
        eng.pinned_free(hin)
        eng.pinned_free(hout)

    # ---- self-check after the timed region, at the bench's own size (the GPU full, every SM holding its resident CTAs): the
    # block step as measured against its two-launch form on EVERY channel, bit for bit (the oracle comparisons live in tests/) ----
    selfcheck = None
    if not args.no_selfcheck:
        KC = 4
        rng = np.random.default_rng(77 + rank)
        xin = (rng.random((KC, S, B), dtype=np.float32) * 2 - 1).astype(np.float32)
        e.set_stream(None)
        e.reset()
        ya = e.process(xin).copy()
        e.reset()
        if not per_stream_ir:
            e.set_fused_step(bool(args.two_launch))     # the other form of the block step
        yb = e.process(xin).copy()
        selfcheck = {"blocks": KC, "channels_compared": int(S),
                     "against": "same step repeated" if per_stream_ir else ("fused step" if args.two_launch else "two-launch step"),
                     "bit_identical": bool(np.array_equal(ya, yb))}
        selfcheck["ok"] = sharding.max_over_ranks(0.0 if selfcheck["bit_identical"] else 1.0, device="cuda") == 0.0      # every rank
        del xin, ya, yb

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(args, args.cpu_seconds)
        cpu.pop("seconds", None)
        cpu.pop("probe", None)

    if rank == 0:
        wl = ("configs[3]: %d streams with per-stream %.0f s IRs (%d taps, %d partitions), block %d, sharded by stream over %d GPU(s)"
              % (total_streams, args.ir_seconds, Lh, P, B, world)) if per_stream_ir else \
             ("configs[2] shape: %d independent mono streams per GPU sharing one %.1f s IR (%d taps, %d partitions), block %d" % (S, args.ir_seconds, Lh, P, B))
        line = {"metric": METRIC if not per_stream_ir else "48kHz RT channels (block 1024, per-stream 10s IRs)", "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong" if per_stream_ir else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl,
                           "streams_per_gpu": S, "block": B, "ir_taps": Lh, "partitions": P, "fft_size": e.fft_size,
                           "state_bytes_per_gpu": int(e.state_bytes), "l2_policy": "inputs larger than L2 (FDL %.2f GB per GPU)" % (S * P * B * 8 / 1e9), "mac_plan": list(e.mac_plan()),
                           "sharding": "contiguous stream ranges by rank, IR replicated, no data-path collective; host gather of outputs",
                           "gathered_output_shape": list(gathered.shape), "channel_samples_per_s": value * SR},
                "latency": lat, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "selfcheck": selfcheck, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
