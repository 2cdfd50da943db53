"""bench.py -- features refined per second on BASELINE config 2 (dense-cluster 2D video).

    python bench.py --gpus N --steps K --warmup W            # this repo, N GPUs (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One step = one pass of the hot path (``refine_leastsq``: cluster-level least-squares refinement) over
one video of ``--gpus x --frames`` frames of 1024x1024 uint8 with ~2100 features per frame in
clusters of 2-6.  Weak scaling: rank r owns frames
[r * frames, (r + 1) * frames) of that ONE video; frames are independent, so there is no collective
on the data path.

  value   whole-job features/s with all inputs resident in HBM (frames, packed parameters, bounds):
          the C-ABI launches only (ctk_frame_max + ctk_refine_batch per size bin), CUDA-event timed;
  e2e     the same metric through the public API with HOST frames and a pandas DataFrame: clustering,
          packing, pinned staging, H2D, kernels, D2H and DataFrame write-back inside the timed
          region.  N = 1: ``clustertracking_b200.refine_leastsq(f, reader, 11)``; N > 1: the sharded
          product path ``parallel.refine_leastsq_sharded`` on the one video (every rank passes only
          its own rows, the merged table is gathered on rank 0 -- inside the timed region).

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = "1"

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (1024, 1024)
PITCH, SIZE, DIAMETER, NOISE = 44, 2.75, 11, 8
N_LAYOUTS = 8
WORKLOAD = ("config2: 2D synthetic video, %d frames/GPU of 1024x1024 uint8, ~2100 gaussian features "
            "per frame in overlapping clusters of 2-6, refine_leastsq fit_function='gauss', "
            "diameter 11")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU (config 2: 1000)")
    ap.add_argument("--cpu-frames", type=int, default=2, help="frames of the cpu_baseline sample")
    ap.add_argument("--precision", default="float32", choices=["float32", "float64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic video
# --------------------------------------------------------------------------------------------------
def video_geometry(n_frames, seed):
    """Positions, signals and frame index of every feature of the video (host, numpy)."""
    from clustertracking_b200 import artificial
    rng = np.random.default_rng(seed)
    layouts = [artificial.clustered_positions(SHAPE, PITCH, SIZE, np.random.default_rng(seed * 1000 + g))[0]
               for g in range(min(N_LAYOUTS, n_frames))]
    pos, frame = [], []
    for t in range(n_frames):
        base = layouts[t % len(layouts)]
        pos.append(base + rng.normal(0., 0.5, base.shape))     # features diffuse 0.5 px rms per frame
        frame.append(np.full(len(base), t, dtype=np.int64))
    pos = np.concatenate(pos)
    frame = np.concatenate(frame)
    signal = rng.uniform(80., 160., len(pos))
    start = pos + rng.uniform(-0.5, 0.5, pos.shape)
    return pos, frame, signal, start


def start_dataframe(start, frame):
    import pandas as pd
    return pd.DataFrame(dict(y=start[:, 0], x=start[:, 1], signal=120., size=SIZE,
                             background=NOISE / 2., frame=frame))


def render_video_torch(pos, frame, signal, n_frames, device, seed, chunk=20):
    """Render the video on the GPU with torch (synthetic INPUT generation only, not the hot path):
    truncated gaussian spots as clustertracking/artificial.py:81-141 draws them, plus Poisson noise."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    h = int(np.ceil(4 * SIZE)) + 1
    oy, ox = torch.meshgrid(torch.arange(-h, h + 1, device=device),
                            torch.arange(-h, h + 1, device=device), indexing='ij')
    offs = torch.stack([oy.reshape(-1), ox.reshape(-1)], dim=1)              # (K, 2)
    stack = torch.empty((n_frames,) + SHAPE, dtype=torch.uint8, device=device)
    pos_t = torch.from_numpy(pos).to(device)
    frame_t = torch.from_numpy(frame).to(device)
    sig_t = torch.from_numpy(signal).to(device)
    bounds = np.searchsorted(frame, np.arange(0, n_frames + chunk, chunk))
    for ci, f0 in enumerate(range(0, n_frames, chunk)):
        f1 = min(n_frames, f0 + chunk)
        a, b = int(bounds[ci]), int(np.searchsorted(frame, f1))
        p = pos_t[a:b]
        pix = torch.floor(p).long()[:, None, :] + offs[None, :, :]          # (F, K, 2)
        rel = pix.double() - p[:, None, :]
        inside = (rel.abs() <= 4 * SIZE + 1).all(dim=2)
        inside &= (pix[..., 0] >= 0) & (pix[..., 0] < SHAPE[0]) & (pix[..., 1] >= 0) & (pix[..., 1] < SHAPE[1])
        r2 = ((rel / SIZE) ** 2).sum(dim=2)
        spot = torch.floor(sig_t[a:b, None] * torch.exp(-r2)).to(torch.int32)   # ndim/2 = 1
        flat = ((frame_t[a:b, None] - f0) * SHAPE[0] + pix[..., 0]) * SHAPE[1] + pix[..., 1]
        acc = torch.zeros((f1 - f0) * SHAPE[0] * SHAPE[1], dtype=torch.int32, device=device)
        acc.index_add_(0, flat[inside], spot[inside])
        acc = acc.view((f1 - f0,) + SHAPE)
        noise = torch.poisson(torch.full(acc.shape, float(NOISE), device=device), generator=gen)
        stack[f0:f1] = (acc + noise.to(torch.int32)).clamp_(0, 255).to(torch.uint8)
    return stack


# --------------------------------------------------------------------------------------------------
# clock sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(max(smax)) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------
# accounting (DESIGN.md "Measurement"): algorithmic bytes and flops of the refine kernel
# --------------------------------------------------------------------------------------------------
def refine_accounting(plan, stats, ids=None):
    """-> (HBM bytes, FP32 flops) one pass of the refine launches has to move / execute, from the
    per-cluster device counters.  Bytes: every masked pixel once per outer iteration, the parameter
    and bounds rows, the outputs.  Flops: per objective evaluation E*F_VAL + 2M, per
    normal-equation accumulation E*F_DER + 2*(v^2 (E+2Q)/2 + 2.5 v E + 2 M), with E pixel-feature
    pairs, Q shared pixel pairs, M union pixels, v free parameters per feature."""
    n = plan.cluster_sizes().astype(np.float64)
    if ids is not None:                       # accounting of one launch: its clusters only
        n, stats = n[ids], stats[ids]
    P = plan.problem.n_params
    px = np.dtype(plan.pixel_dtype).itemsize
    evals, accums, outer = stats[:, 0], stats[:, 1], np.maximum(stats[:, 2], 1)
    M, E, Q = stats[:, 3].astype(np.float64), stats[:, 4].astype(np.float64), stats[:, 5].astype(np.float64)
    # per cluster: masked pixels, parameter rows in + out (bounds come from the problem's tables,
    # they are never materialised), offsets / frame index / counters / cost / status
    nbytes = (M * px * outer + n * P * 8 * 2 + 8 + 4 + 32 + 8 + 4).sum()
    v = sum(1 for m in list(plan.problem.modes)[1:P] if m != 0)       # free parameters per feature
    f_val, f_der = 10., 6.
    flops = (evals * (E * f_val + 2 * M)
             + accums * (E * f_der + 2 * (v * v * (E + 2 * Q) / 2. + 2.5 * v * E + 2 * M))).sum()
    return float(nbytes), float(flops)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.)),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650., sm_max_mhz=1965., source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port, scipy SLSQP) on all host cores
# --------------------------------------------------------------------------------------------------
_WORKER = {}


def _ref_init(seed):
    import warnings
    warnings.simplefilter("ignore")
    from clustertracking_b200 import artificial
    _WORKER["artificial"] = artificial
    _WORKER["seed"] = seed
    _WORKER["frames"] = {}


def _ref_frame(k):
    art = _WORKER["artificial"]
    if k not in _WORKER["frames"]:
        frame, f0, _ = art.clustered_frame(SHAPE, PITCH, SIZE, NOISE, seed=_WORKER["seed"] + k)
        _WORKER["frames"][k] = (frame, f0)
    return len(_WORKER["frames"][k][1])


def _ref_refine(k):
    import warnings
    from oracle import cluster_oracle
    _ref_frame(k)
    frame, f0 = _WORKER["frames"][k]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cluster_oracle.refine_leastsq(f0.copy(), frame, DIAMETER)
    return len(out)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t_gen = time.time()
    with ctx.Pool(cores, initializer=_ref_init, initargs=(12345,)) as pool:
        # one frame of the config-2 workload per worker and step; frames are rendered untimed
        # (chunksize 1 with as many tasks as workers: every worker renders and keeps its own frame)
        n_tasks = cores
        pool.map(_ref_frame, range(n_tasks), chunksize=1)
        t_gen = time.time() - t_gen
        times, feats = [], 0
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            counts = pool.map(_ref_refine, range(n_tasks), chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
                feats += sum(counts)
            if step == 0 and dt * (args.warmup + args.steps) > 600 and n_tasks > 1:
                pass
    total = sum(times)
    value = feats / total
    sample = ("%d frames (one per host thread) of the config-2 workload per step, %d features/step; "
              "frames rendered by clustertracking_b200.artificial with per-frame seeds" %
              (n_tasks, feats // max(1, args.steps)))
    line = dict(impl="reference", metric="features_refined_per_sec", value=value, unit="features/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / max(1, args.steps), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD % args.frames, sample=sample),
                cpu_baseline=dict(value=value, unit="features/s", cores=cores, kind="port",
                                  sample=sample),
                e2e=dict(value=value, unit="features/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import warnings
    import torch
    import torch.distributed as dist
    import clustertracking_b200 as ctb
    from clustertracking_b200 import artificial, parallel, refine as ctb_refine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- inputs: ONE video of world * frames frames; rank r owns frames [r * frames, (r + 1) * frames)
    # (rendered on its device; no rank ever holds the whole video or the whole table) ----------------
    n_frames = args.frames
    first = rank * n_frames
    pos, frame, signal, start = video_geometry(n_frames, seed=7 + rank)
    d_stack = render_video_torch(pos, frame, signal, n_frames, device, seed=100 + rank)
    torch.cuda.synchronize()
    host_stack = torch.empty(d_stack.shape, dtype=torch.uint8, pin_memory=True)
    host_stack.copy_(d_stack)
    torch.cuda.synchronize()
    reader = artificial.FrameStack(host_stack.numpy(), first_frame=first)
    f0 = start_dataframe(start, frame + first)
    f0.index = f0.index + int(sum_before(len(f0), rank, world, dist, device))
    n_features = len(f0)

    # ---- value: everything resident in HBM, C-ABI launches only ----------------------------------
    def resident(precision, steps, warmup):
        plan = ctb_refine.prepare(f0.copy(), reader, DIAMETER, precision=precision)
        session = ctb_refine.DeviceSession(plan, device)
        session.frames.register(d_stack, 0)                 # frames are already resident
        slices = session.schedule()                         # every size class once

        def one_step(events=None):
            session.frames.launch_frame_max(0, n_frames, events)
            session.run(slices, events)

        for _ in range(max(3, warmup)):
            one_step()
        torch.cuda.synchronize()
        stats = session.d_stats.cpu().numpy()
        status = session.d_status.cpu().numpy()
        launches0 = session.launches + session.frames.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        events = []
        barrier()
        ev0.record()
        for _ in range(steps):
            one_step(events)
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        launches = session.launches + session.frames.launches - launches0
        return dict(plan=plan, session=session, slices=slices, stats=stats, status=status, ms=ms,
                    events=events, launches=launches, steps=steps)

    sampler = ClockSampler(local_rank)
    sampler.start()
    run = resident(args.precision, args.steps, args.warmup)
    clocks = sampler.stop()
    plan, session, slices, stats, status = (run[k] for k in ("plan", "session", "slices", "stats", "status"))
    resident_ms, kernel_events, gpu_launches = run["ms"], run["events"], run["launches"]
    total_features = sum_over_ranks(n_features)
    value = total_features * args.steps / (resident_ms * 1e-3)

    # the reference's own arithmetic (everything float64) beside the float32 headline
    other = "float64" if args.precision == "float32" else "float32"
    run2 = resident(other, max(1, min(args.steps, 3)), 3)
    value_other = total_features * run2["steps"] / (run2["ms"] * 1e-3)
    del run2["session"]

    # kernel-level durations of the refine launches (same timed region, same stream)
    refine_ms = sum(a.elapsed_time(b) for kind, a, b, _ in kernel_events if kind == "refine") / args.steps
    fmax_ms = sum(a.elapsed_time(b) for kind, a, b, _ in kernel_events if kind == "frame_max") / args.steps
    nbytes, flops = refine_accounting(plan, stats)
    peaks = measured_peaks()
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    fp32_peak = sms * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    # the dominant launch: the (size class, kind) whose launches take the longest per step
    by_label = {}
    for kind, a, b, label in kernel_events:
        if kind == "refine":
            by_label[label] = by_label.get(label, 0.) + a.elapsed_time(b) / args.steps
    dom_label = max(by_label, key=by_label.get)
    dominant = None
    if dom_label[1] == "main":
        cap, start_, count = next(s for s in slices if s[0] == dom_label[0])
        ids = session.d_work[start_:start_ + count].cpu().numpy()
        sub_bytes, sub_flops = refine_accounting(plan, stats, ids)
        dominant = dict(cap=int(cap), clusters=int(count), ms=float(by_label[dom_label]),
                        nbytes=sub_bytes, flops=sub_flops)
    traffic, traffic_source = None, None
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
        if dominant is None or not (name.startswith("r0") and name.endswith("_traffic.json")):
            continue
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            t = json.load(fh)
        if (t.get("frames_per_gpu") == n_frames and t.get("max_cluster_features") == dominant["cap"]
                and t.get("precision") == args.precision):
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
            traffic_source = "profiles/%s (ncu --set full capture of this launch)" % name
            break
    peak_fp32_source = ("%d SMs x 128 lanes x 2 x %.0f MHz (sm_max_mhz, %s; MEASURED_PEAKS.json has no "
                        "FP32 figure)" % (sms, peaks["sm_max_mhz"], peaks["source"]))
    if dominant is not None:
        share = 100. * dominant["ms"] / (refine_ms + fmax_ms)
        kernel = ("refine_kernel, launch of the size class up to %d features (%d clusters, %.0f %% of "
                  "the step)" % (dominant["cap"], dominant["clusters"], share))
        # the BINDING bound of this kernel is FP32 issue, not HBM (arithmetic intensity ~1e2 flop/B)
        roofline = dict(bound="fp32", kernel=kernel,
                        achieved=dominant["flops"] / (dominant["ms"] * 1e-3) / 1e12, peak=fp32_peak,
                        unit="TFLOP/s", peak_source=peak_fp32_source, ms_per_launch=dominant["ms"],
                        algorithmic_flops_per_launch=dominant["flops"],
                        traffic=traffic, traffic_source=traffic_source,
                        all_refine_launches=dict(ms_per_step=refine_ms,
                                                 achieved=flops / (refine_ms * 1e-3) / 1e12,
                                                 algorithmic_flops_per_feature=flops / n_features))
        roofline_hbm = dict(bound="hbm", kernel=kernel,
                            achieved=dominant["nbytes"] / (dominant["ms"] * 1e-3) / 1e9,
                            peak=peaks["hbm_gbs"], unit="GB/s", traffic=traffic,
                            traffic_source=traffic_source, peak_source=peaks["source"],
                            ms_per_launch=dominant["ms"],
                            algorithmic_bytes_per_launch=dominant["nbytes"],
                            all_refine_launches=dict(ms_per_step=refine_ms,
                                                     achieved=nbytes / (refine_ms * 1e-3) / 1e9,
                                                     algorithmic_bytes_per_feature=nbytes / n_features))
    else:
        kernel = "refine_kernel (all size classes of one step)"
        roofline = dict(bound="fp32", kernel=kernel, achieved=flops / (refine_ms * 1e-3) / 1e12,
                        peak=fp32_peak, unit="TFLOP/s", peak_source=peak_fp32_source,
                        ms_per_step=refine_ms, traffic=None,
                        algorithmic_flops_per_feature=flops / n_features)
        roofline_hbm = dict(bound="hbm", kernel=kernel, achieved=nbytes / (refine_ms * 1e-3) / 1e9,
                            peak=peaks["hbm_gbs"], unit="GB/s", traffic=None,
                            peak_source=peaks["source"], ms_per_step=refine_ms,
                            algorithmic_bytes_per_feature=nbytes / n_features)
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    roofline_hbm["frac"] = roofline_hbm["achieved"] / roofline_hbm["peak"]
    frame_bytes = float(d_stack.numel())
    roofline_fmax = dict(bound="hbm", kernel="frame_max_kernel", achieved=frame_bytes / (fmax_ms * 1e-3) / 1e9,
                         peak=peaks["hbm_gbs"], unit="GB/s", ms_per_step=fmax_ms)
    roofline_fmax["frac"] = roofline_fmax["achieved"] / roofline_fmax["peak"]
    del run["session"], session

    # ---- e2e: public API, host frames and DataFrame, copies inside the timed region -----------------
    def call_api(table, rd):
        """One end-to-end pass.  N = 1: refine_leastsq.  N > 1: the product's sharded path on ONE
        video -- every rank passes only ITS frames, the merged table is gathered on rank 0."""
        if world == 1:
            return ctb.refine_leastsq(table, rd, DIAMETER, precision=args.precision)
        return parallel.refine_leastsq_sharded(table, rd, DIAMETER, presharded=True, gather='root',
                                               precision=args.precision)

    step_log = []

    def time_api(table, rd, steps, warmup):
        times, out = [], None
        for step in range(warmup + steps):
            out = None
            barrier()
            t0 = time.perf_counter()
            out = call_api(table, rd)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            if step >= warmup:
                times.append(dt)
                step_log.append((round(1e3 * dt, 1), dict(ctb_refine.LAST_CALL).get("setup_ms"), dict(ctb_refine.LAST_CALL).get("launched_at_ms"),
                                 dict(ctb_refine.LAST_CALL).get("returned_at_ms"), dict(ctb_refine.LAST_CALL).get("labelling"),
                                 dict(ctb_refine.LAST_CALL).get("phases_ms")))
        return sum(times), out

    e2e_s, out = time_api(f0, reader, args.steps, args.warmup)
    if rank == 0 and os.environ.get("CTK_BENCH_STEPS"):
        for entry in step_log:
            print("e2e step", entry, file=sys.stderr, flush=True)
    info = dict(ctb_refine.LAST_CALL)
    h2d, d2h = sum_over_ranks(info["h2d_bytes"]), sum_over_ranks(info["d2h_bytes"])
    api = ("clustertracking_b200.refine_leastsq(DataFrame, FrameStack(pinned host uint8), 11)" if world == 1 else
           "clustertracking_b200.parallel.refine_leastsq_sharded(this rank's rows of ONE %d-frame video, "
           "FrameStack(pinned host uint8), 11, presharded=True, gather='root'): frames sharded over %d "
           "ranks, merged DataFrame on rank 0, gather inside the timed region"
           % (n_frames * world, world))
    e2e = dict(value=total_features * args.steps / e2e_s, unit="features/s",
               h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
               ms_per_step=1e3 * e2e_s / args.steps, api=api, host_ms=info.get("phases_ms"),
               labelling=info.get("labelling"))
    if world > 1:
        e2e["sharded_ms_rank0"] = dict(parallel.LAST_GATHER)
    if world > 1 and rank == 0:
        assert out is not None and len(out) == int(total_features)
        assert np.all(np.diff(out['frame'].values) >= 0)
    out_first = out if world == 1 else None
    if world > 1:
        # strong scaling: ONE video of `frames` frames over all ranks (frames / N each)
        share = max(1, n_frames // world)
        sel = f0['frame'].values < first + share
        f_strong = f0[sel]
        strong_s, _ = time_api(f_strong, reader, max(2, args.steps // 2), 2)
        strong_features = sum_over_ranks(int(sel.sum()))
        e2e["strong"] = dict(value=strong_features * max(2, args.steps // 2) / strong_s, unit="features/s",
                             ms_per_step=1e3 * strong_s / max(2, args.steps // 2),
                             frames_total=share * world,
                             what="the same call on ONE video of %d frames, %d per rank"
                                  % (share * world, share))
    else:
        # the same call from a PAGEABLE numpy stack (what a user who never heard of pinned memory has)
        pageable = artificial.FrameStack(np.array(host_stack.numpy(), copy=True), first_frame=first)
        page_s, _ = time_api(f0, pageable, max(2, args.steps // 2), 2)
        e2e["pageable"] = dict(value=total_features * max(2, args.steps // 2) / page_s, unit="features/s",
                               ms_per_step=1e3 * page_s / max(2, args.steps // 2),
                               what="the same call with the frames in pageable host memory")
        del pageable

    # ---- cpu_baseline (rank 0, N=1 only) and parity on the same sample ------------------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cluster_oracle
        k = min(args.cpu_frames, n_frames)
        sel = f0['frame'].values < k
        sub = f0[sel].copy()
        sub_reader = artificial.FrameStack(reader.stack[:k])
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = cluster_oracle.refine_leastsq(sub.copy(), sub_reader, DIAMETER)
        dt = time.perf_counter() - t0
        cpu_baseline = dict(value=len(sub) / dt, unit="features/s", cores=1, kind="port",
                            sample="first %d frames of the same video (%d features), oracle port "
                                   "(scipy SLSQP) on 1 thread, %.1f s" % (k, len(sub), dt))

        def compare(got):
            assert np.array_equal(got.index.values, want.index.values)
            both = ~np.isnan(got['cost'].values) & ~np.isnan(want['cost'].values)
            return dict(
                sample_features=int(len(sub)),
                cluster_membership_identical=bool(np.array_equal(got['cluster'].values, want['cluster'].values)),
                failures_ours=int(np.isnan(got['cost'].values).sum()),
                failures_oracle=int(np.isnan(want['cost'].values).sum()),
                max_abs_dpos_px=float(np.abs(got[['y', 'x']].values[both] - want[['y', 'x']].values[both]).max()),
                max_rel_dsignal=float(np.abs(got['signal'].values[both] / want['signal'].values[both] - 1).max()),
                max_abs_dcost=float(np.abs(got['cost'].values[both] - want['cost'].values[both]).max()))

        parity = {args.precision: compare(out_first[out_first['frame'].values < k])}
        parity[other] = compare(ctb.refine_leastsq(sub.copy(), sub_reader, DIAMETER, precision=other))

    if rank == 0:
        ok = status == 0
        f32 = args.precision == "float32"
        line = dict(
            metric="features_refined_per_sec", value=value, unit="features/s", n_gpus=args.gpus,
            steps=args.steps, warmup=max(3, args.warmup), ms_per_step=resident_ms / args.steps,
            higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32" if f32 else "f64", data="synthetic",
            config=dict(workload=WORKLOAD % n_frames, frames_per_gpu=n_frames,
                        video="one video of %d frames, %d per GPU" % (n_frames * world, n_frames),
                        features_per_gpu=int(n_features), clusters_per_gpu=int(plan.n_clusters),
                        l2="inputs larger than L2: %.2f GB of frames per GPU vs 126 MB"
                           % (frame_bytes / 1e9) if frame_bytes > 2.6e8 else
                           "WARNING: inputs (%.0f MB) not larger than L2" % (frame_bytes / 1e6),
                        pixel_arithmetic=args.precision,
                        normal_equations=("J^T J, its Cholesky factor and the model values float32; "
                                          "J^T r, parameters, bounds and the objective float64" if f32
                                          else "float64 throughout"),
                        converged_clusters=int(ok.sum()), failed_clusters=int((~ok).sum()),
                        mean_evaluations_per_cluster=float(stats[:, 0].mean()),
                        mean_outer_iterations=float(stats[:, 2].mean()),
                        mean_union_pixels=float(stats[:, 3].mean())),
            clocks=clocks, e2e=e2e, gpu_launches=int(gpu_launches),
            roofline=roofline, roofline_hbm=roofline_hbm, roofline_frame_max=roofline_fmax,
            cpu_baseline=cpu_baseline, parity_vs_oracle=parity)
        line["value_" + ("f64" if f32 else "f32")] = dict(
            value=value_other, unit="features/s", ms_per_step=run2["ms"] / run2["steps"],
            steps=run2["steps"], pixel_arithmetic=other,
            failed_clusters=int((run2["status"] != 0).sum()),
            mean_evaluations_per_cluster=float(run2["stats"][:, 0].mean()),
            what="the same resident measurement with precision=%r" % other)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sum_before(n, rank, world, dist, device):
    """Number of rows held by the ranks before this one (so that index labels run on)."""
    if world == 1:
        return 0
    import torch
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    counts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(counts, mine)
    return int(sum(int(c.item()) for c in counts[:rank]))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
