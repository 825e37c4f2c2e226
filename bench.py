#!/usr/bin/env python
"""Benchmark of the MD_RDM depth-map fusion path (BASELINE.json metric: fused depth maps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

ours      : BASELINE.json configs[1] - standalone fusion path on B200, batch 16, scales 8/16/32:
            inputs (ordinary map + raw pair matrices: 1 f32 64x64 + 5 f64 256x64 per image)
            resident in HBM, one step = quantize + ALS + decompose + weighted reconstruction of one
            batch (5 kernel launches replayed from a CUDA graph).  Steps rotate over a ring of
            resident batches larger than L2 and over 32 streams (batches in flight).
            e2e = the public host API (FusionPlan.run_pinned: decoder maps in pinned host memory
            -> fused 128x128 log-depth maps in pinned host memory), H2D and D2H copies inside
            the timed region, pair build fused in front.
reference : the reference's own CPU algorithm for the same path (oracle/literal.py: the
            loop-for-loop restatement of the reference, which is Python and cannot travel to
            the GPU box), one image per step, all host threads.
A step lasts ~11 us, so the 32-lane pipeline needs a few hundred steps to fill: K = 2000 by default (22 ms);
with K = 37 the same code reports ~0.8 M maps/s, with K = 5 ~0.3 M (start-up and host launch time dominate).
One JSON line on stdout (rank 0).  Weak scaling: every rank runs K steps on its own batches;
no collective on the data path (torch.distributed is used for the barrier and the max only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "fused_depth_maps_per_sec"
UNIT = "maps/s"
BATCH = 16
SCALES = (8, 16, 32)
LAUNCHES_PER_STEP = 5   # sparsify, compact-page ALS, dense ALS (8x8 maps), select, fused tail


# ----------------------------------------------------------------------------- workload arithmetic
def algorithmic_bytes(scales=SCALES):
    """Per-image stage contract bytes (SURVEY.md 8d / DESIGN.md 'Algorithmic bytes')."""
    import math
    E = {s: (64 * 64 if s == 8 else (s // 16) ** 2 * 256 * 64) for s in scales}
    raw = {s: E[s] * (4 if s == 8 else 8) for s in scales}
    Rq = {s: 4 * E[s] for s in scales}
    bins = {s: E[s] for s in scales}
    mp = {s: 4 * s * s for s in scales}
    comp = {s: 8 * sum(4 ** k for k in range(1, int(math.log2(s)) + 1)) for s in scales}
    comp_d1 = 8 * 85
    kmax = max([3] + [int(math.log2(s)) for s in scales])
    yhat = 4 * sum(4 ** k for k in range(kmax + 1))
    out = {
        "pair": sum(mp[s] + (8 * (s // 2) ** 2 if s > 8 else 0) + raw[s] for s in scales),
        "quantize": sum(raw[s] + Rq[s] + bins[s] for s in scales),
        "als": sum(Rq[s] + mp[s] for s in scales),
        "decompose": sum(mp[s] for s in scales) + 512 + sum(comp[s] for s in scales) + comp_d1,
        "reconstruct": sum(comp[s] for s in scales) + comp_d1 + yhat + 131072,
    }
    out["path"] = out["quantize"] + out["als"] + out["decompose"] + out["reconstruct"]
    # per-kernel split of the same contract bytes: the sparsify kernel does the quantize stage of the page
    # scales (raw read, Rq and bins as the stage's logical outputs), the compact-page ALS kernel the ALS stage
    # of those scales (Rq in, maps out), the dense kernel both stages of the 8x8 map
    pg = [s for s in scales if s > 8]
    out["sparsify_kernel"] = sum(raw[s] + Rq[s] + bins[s] for s in pg)
    out["als_sparse_kernel"] = sum(Rq[s] + mp[s] for s in pg)
    out["als_dense_kernel"] = sum(raw[s] + 2 * Rq[s] + bins[s] + mp[s] for s in scales if s == 8)
    out["als_iterate_kernel"] = out["quantize"] + sum(Rq[s] for s in scales)   # the three iterate-phase kernels together
    out["als_select_kernel"] = sum(mp[s] for s in scales)
    out["tail_kernel"] = out["decompose"] + out["reconstruct"]
    return out


def synthetic_batch(B: int, scales, seed: int):
    """SURVEY 8d synthetic decoder outputs: x_d1 = randint(1,90) DORN counts, relative maps
    exp(0.3 randn), weights abs(randn(K,1)) as network/RDM_Net.py:449-465 initialises them."""
    import math
    g = torch.Generator().manual_seed(seed)
    x_d1 = torch.randint(1, 90, (B, 1, 8, 8), generator=g, dtype=torch.int64)
    rel = [torch.exp(0.3 * torch.randn(B, 1, s, s, generator=g)) for s in scales]
    K = [1, 1, 1, 1, 0, 0, 0, 0]
    for s in scales:
        for k in range(1, int(math.log2(s)) + 1):
            K[k] += 1
    weights = [torch.abs(torch.randn(k, 1, generator=g)) for k in K if k > 0]
    return x_d1, rel, weights


def batch_seed(rank: int, batch_idx: int) -> int:
    """seed of the synthetic batch `batch_idx` of GPU `rank`."""
    return 1234 + 1000 * rank + batch_idx


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def dist_max(value: float, device=None) -> float:
    """Max over ranks (identity when not distributed)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the GPU is busy."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, cuda_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            try:   # CUDA and NVML orderings differ under CUDA_VISIBLE_DEVICES: go through the UUID
                self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)).encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover - NVML missing
            self._nv, self.error = None, repr(e)

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- ours
def build_ring(dev, rank, n_plans, source):
    """`n_plans` resident batches (each its own buffers + CUDA graph), inputs generated per SURVEY 8d."""
    import md_rdm_b200.ops  # noqa: F401
    from md_rdm_b200.fusion import FusionPlan
    R = torch.ops.rdm
    ring = []
    for b in range(n_plans):
        x_d1, rel, weights = synthetic_batch(BATCH, SCALES, seed=batch_seed(rank, b))
        plan = FusionPlan(BATCH, SCALES, source, device=dev, want_bins=True)
        rel_d = [r.to(dev) for r in rel]
        if source == "raw":   # raw pair matrices derived from the maps with the pair-build kernels (not timed)
            srcs = [R.pair_v1(r) if r.shape[2] == 8 else R.pair_id(r)[0] for r in rel_d]
        else:
            srcs = rel_d
        plan.load_inputs(x_d1.to(dev), srcs, torch.cat([w.reshape(-1) for w in weights]).to(dev))
        plan.host_inputs = (x_d1, rel)
        plan.capture()
        ring.append(plan)
    torch.cuda.synchronize()
    return ring


def timed_steps(ring, streams, steps, fn):
    """Run `steps` steps, step i on stream i % S with plan i % len(ring); device time by CUDA events
    recorded on the current stream around fork/join of the side streams."""
    cur = torch.cuda.current_stream()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    start.record(cur)
    for s in streams:
        s.wait_event(start)
    for i in range(steps):
        with torch.cuda.stream(streams[i % len(streams)]):
            fn(ring[i % len(ring)])
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        cur.wait_event(ev)
    end.record(cur)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    return start.elapsed_time(end), wall_ms


def time_serial(fns, reps):
    """Average duration of the launches issued by `fns` (one per ring entry) back to back on ONE stream.
    The launches are captured into a CUDA graph first, so the host launch rate (several microseconds per
    Python + ctypes call) is not what gets measured for the short kernels."""
    n = len(fns)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for f in fns[:4]:
            f()
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for f in fns:
            f()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rounds = max(reps // n, 1)
    with torch.cuda.stream(stream):
        g.replay()
        start.record(stream)
        for _ in range(rounds):
            g.replay()
        end.record(stream)
    stream.synchronize()
    return start.elapsed_time(end) / (rounds * n) * 1e-3   # seconds per launch


def stage_kernel_table(dev, batch):
    """Stand-alone streaming kernels of the path (the drop-in ops), graph-timed (8 calls per graph, no host
    dispatch in the timed region): algorithmic GB/s per kernel (SURVEY 8d byte formulas)."""
    import md_rdm_b200.ops  # noqa: F401
    from md_rdm_b200.codebooks import default_quantization
    R = torch.ops.rdm
    g = torch.Generator().manual_seed(7)
    q = default_quantization()
    x8 = torch.exp(0.3 * torch.randn(batch, 1, 8, 8, generator=g)).to(dev)
    x32 = torch.exp(0.3 * torch.randn(batch, 1, 32, 32, generator=g)).to(dev)
    raw32, _ = R.pair_id(x32)
    thr, lvl = q.device_tables(32, dev)
    y = (0.5 + 9.5 * torch.rand(batch, 1, 128, 128, generator=g, dtype=torch.float64)).to(dev)
    comps = [torch.randn(batch, 1, 2 ** k, 2 ** k, generator=g).to(dev) for k in range(8)]
    cases = {
        "pair_v1": (lambda: R.pair_v1(x8), batch * (256 + 16384)),
        "pair_id_32": (lambda: R.pair_id(x32), batch * (4096 + 2048 + 4 * 131072)),
        "lloyd_quantize_f64_32": (lambda: R.lloyd_quantize(raw32, thr, lvl), raw32.numel() * 17),
        "gm_normalize+decompose_gt128": (lambda: R.decompose(R.gm_normalize(y), False), batch * (3 * 131072 + 174760)),
        "recombination_128": (lambda: R.recombination(comps, 7), batch * (131072 + 4 * 21845)),
    }
    out = {}
    for name, (fn, nbytes) in cases.items():
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            fn()
        stream.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            for _ in range(8):
                fn()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            gr.replay()
            s0.record(stream)
            for _ in range(4):
                gr.replay()
            s1.record(stream)
        stream.synchronize()
        t = s0.elapsed_time(s1) / 32 * 1e-3
        out[name] = {"us": round(t * 1e6, 2), "gbs": round(nbytes / t / 1e9, 1)}
        del gr
    return out


def pcie_d2h_peak(dev):
    """Measured device->host copy rate of this box (pinned memory): large copies, and copies of the e2e result
    size (2 MB) issued back to back - the ceiling of any path that returns 128 KB of f64 log-depth per image."""
    out = {}
    for name, nbytes, reps in (("d2h_64MB_gbs", 64 << 20, 10), ("d2h_2MB_gbs", 2 << 20, 200)):
        src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        s1.record()
        torch.cuda.synchronize()
        out[name] = round(nbytes * reps / (s0.elapsed_time(s1) * 1e-3) / 1e9, 2)
    return out


def cpu_baseline_port(images: int):
    """The reference's CPU algorithm (literal port) on a bounded sample of the same workload."""
    from oracle import fusion_ref as fr
    from oracle import literal as lit
    books = fr.load_codebooks()
    x_d1, rel, weights = fr.synthetic_batch(images, SCALES, seed=batch_seed(0, 0))
    t0 = time.perf_counter()
    lit.fusion_forward_literal(x_d1, rel, weights, books)
    dt = time.perf_counter() - t0
    # vectorised restatement (same arithmetic, Python loops removed): the "best-effort CPU" figure
    x16 = fr.synthetic_batch(BATCH, SCALES, seed=batch_seed(0, 0))
    fr.fusion_forward(*x16, books)
    t1 = time.perf_counter()
    nb = 5
    for _ in range(nb):
        fr.fusion_forward(*x16, books)
    dv = time.perf_counter() - t1
    return images / dt, (nb * BATCH) / dv, dt


def run_ours(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to STDOUT when the first communicator is created: keep stdout for the
        # one JSON line by pointing fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from md_rdm_b200 import _cabi
    _cabi.load()   # fail loudly before anything is timed

    from md_rdm_b200.fusion import capture_lane, capture_ring
    ab = algorithmic_bytes(SCALES)
    n_plans = max(args.ring // args.streams, 1) * args.streams   # whole plans per lane
    ring = build_ring(dev, rank, n_plans, "raw")
    ring_in_bytes = sum(p.h2d_bytes() for p in ring)
    streams = [torch.cuda.Stream() for _ in range(args.streams)]
    K, W = args.steps, args.warmup
    replay = lambda p: p.replay()   # noqa: E731
    # `streams` independent lanes: lane j owns plans j, j+S, ... of the ring, one stream and one CUDA graph
    # that runs its plans back to back (3 kernels each).  Lanes are replayed round-robin and never join, so
    # several batches stay in flight and the host launches K / (ring/S) graphs instead of K.
    S = args.streams
    lanes = [capture_lane(ring[j::S]) for j in range(S)]
    per_lane = n_plans // S

    def run_steps(k):
        """Exactly k steps: lane graphs round-robin, then single-plan graphs for the remainder."""
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        for _, st in lanes:
            st.wait_event(fork)
        if k < 3 * per_lane:
            # a handful of steps: one single-plan graph per step, each on its own lane, instead of serialising them
            # inside one or two lane graphs (beyond ~12 steps the host cost of the extra launches outweighs that)
            for i in range(k):
                with torch.cuda.stream(lanes[i % S][1]):
                    ring[i % n_plans].replay()
        else:
            n_graphs = k // per_lane
            for i in range(n_graphs):
                g, st = lanes[i % S]
                with torch.cuda.stream(st):
                    g.replay()
            rem = k - n_graphs * per_lane
            for i in range(rem):
                with torch.cuda.stream(lanes[(n_graphs + i) % S][1]):
                    ring[((n_graphs + i) % S)].replay()
        for _, st in lanes:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    def timed(k, fn):
        cur = torch.cuda.current_stream()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        start.record(cur)
        fn(k)
        end.record(cur)
        torch.cuda.synchronize()
        return start.elapsed_time(end), (time.perf_counter() - t0) * 1e3

    with ClockSampler(local_rank) as clk:
        # warm-up: at least W steps, and long enough for NVML to see the clocks under load
        timed(max(W, 3), run_steps)
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            timed(10 * n_plans, run_steps)
        dist_barrier()
        dev_ms, wall_ms = timed(K, run_steps)
        dist_barrier()
        dev_ms = dist_max(max(dev_ms, wall_ms), dev)

        # single-stream latency of one step, and per-kernel launch durations (one stream, back to back)
        lat_ms, _ = timed_steps(ring, streams[:1], max(K, 50), replay)
        lat_ms /= max(K, 50)
        reps = max(K, 100)
        t_iter = time_serial([(lambda p=p: p.run_als_phase(1)) for p in ring], reps)
        t_spf = time_serial([(lambda p=p: p.run_als_phase(4)) for p in ring], reps)
        t_sps = time_serial([(lambda p=p: p.run_als_phase(8)) for p in ring], reps)
        t_dns = time_serial([(lambda p=p: p.run_als_phase(16)) for p in ring], reps)
        t_sel = time_serial([(lambda p=p: p.run_als_phase(2)) for p in ring], reps)
        t_tail = time_serial([(lambda p=p: p.run_tail()) for p in ring], reps)

        # end-to-end through the public host API: pinned host maps -> pinned host log-depth
        # One step = one user call FusionPlan.submit_pinned(): a graph of [H2D copy of the packed inputs,
        # the three kernels, D2H copy of the log-depth maps], calls issued round-robin on the streams
        # (asynchronous API), one stream synchronisation at the end.
        # 8 calls in flight are enough to keep the PCIe link busy; a deeper ring only enlarges the set of pinned
        # result buffers the host has to absorb (2 MB each)
        e2e_lanes = max(min(args.streams, args.e2e_lanes), 1)
        e2e_ring = build_ring(dev, rank, e2e_lanes, "map")
        for p in e2e_ring:
            hb = p._host_buffers()
            hb["x_d1"].copy_(p.host_inputs[0])
            for s, t in zip(p.scales, p.host_inputs[1]):
                hb["src"][s].copy_(t)
            p.capture_e2e()
        e2e_step = lambda p: p.submit_pinned()   # noqa: E731
        e2e_streams = streams[:len(e2e_ring)]
        timed_steps(e2e_ring, e2e_streams, max(W, 3), e2e_step)
        dist_barrier()
        e_dev_ms, e_wall_ms = timed_steps(e2e_ring, e2e_streams, K, e2e_step)
        dist_barrier()
        e_ms = dist_max(max(e_dev_ms, e_wall_ms), dev)
        # the same calls batched: one graph launch per pass over the e2e ring (copies included)
        e2e_graph = capture_ring(e2e_ring, len(e2e_streams), e2e=True)
        nb = len(e2e_ring)
        timed(nb, lambda k: e2e_graph.replay())
        reps = max(K // nb, 1)
        eb_dev_ms, eb_wall_ms = timed(K, lambda k: [e2e_graph.replay() for _ in range(reps)])
        eb_ms = dist_max(max(eb_dev_ms, eb_wall_ms), dev) / (reps * nb) * K
        pcie = pcie_d2h_peak(dev) if rank == 0 else {}
        stage = {}
        if rank == 0 and not args.no_stage_table:
            stage = {"batch_16": stage_kernel_table(dev, 16), "batch_256": stage_kernel_table(dev, 256)}
    clocks = clk.summary()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic, traffic_spf = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_spf = tj.get("als_sparse_kernel_dram_bytes_per_launch"), tj.get("als_sparsify_raw_kernel_dram_bytes_per_launch")
    # dominant kernel by time: the compact-page ALS iterations
    achieved = ab["als_sparse_kernel"] * BATCH / t_sps / 1e9
    kernel_s = {"als_sparsify_raw": t_spf, "als_sparse": t_sps, "als_dense": t_dns, "als_select": t_sel, "fuse_tail": t_tail}
    kernel_b = {"als_sparsify_raw": ab["sparsify_kernel"], "als_sparse": ab["als_sparse_kernel"], "als_dense": ab["als_dense_kernel"],
                "als_select": ab["als_select_kernel"], "fuse_tail": ab["tail_kernel"]}

    out = {
        "metric": METRIC, "value": world * K * BATCH / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {
            "workload": f"BASELINE configs[1]: standalone fusion path, batch 16, scales {'/'.join(map(str, SCALES))}; inputs = ordinary 8x8 "
                        f"map + raw pair matrices (1 f32 64x64 + {sum((s // 16) ** 2 for s in SCALES if s > 8)} f64 256x64 per image) resident "
                        "in HBM; quantize + ALS + decompose + weighted reconstruction -> bins, relative maps, y_hat, 128x128 f64 log-depth",
            "batch": BATCH, "scales": list(SCALES), "images_per_step_per_gpu": BATCH,
            "dtype_detail": "f32: 8x8 pair ratios + Lloyd compare, ALS, y_hat; f64: page pair ratios + Lloyd compare, decomposition, logs, recombination",
            "l2_policy": f"inputs larger than L2: ring of {n_plans} resident batches = {ring_in_bytes / 1e6:.0f} MB of inputs (L2 126 MB)",
            "batches_in_flight": args.streams,
            "cuda_graph": f"{args.streams} lanes (streams), one graph launch per {n_plans // args.streams} steps of a lane, {LAUNCHES_PER_STEP} kernels per step",
            "launches_per_step": LAUNCHES_PER_STEP,
            "single_stream_ms_per_step": lat_ms,
            "algorithmic_bytes_per_image": ab,
            "kernel_ms": {**{k: v * 1e3 for k, v in kernel_s.items()}, "als_iterate_phase": t_iter * 1e3},
            "kernel_gbs": {k: kernel_b[k] * BATCH / v / 1e9 for k, v in kernel_s.items()},
            "kernel_note": "serial launch durations on one stream (CUDA-graph timed) and algorithmic GB/s by the SURVEY 8d contract bytes; "
                           "als_iterate_phase = sparsify + compact-page ALS + dense ALS back to back",
            "path_gbs_at_value": ab["path"] * BATCH * K / (dev_ms * 1e-3) / 1e9,
            "stage_kernels": stage,
            "stage_kernels_note": "stand-alone drop-in ops, CUDA-graph timed, algorithmic GB/s; at batch 16 the inputs are L2-resident "
                                  "and the kernels are launch-latency bound, at batch 256 they stream from HBM",
        },
        "clocks": clocks,
        "e2e": {"value": world * K * BATCH / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e_ring[0].h2d_bytes(),
                "d2h_bytes_per_step": e2e_ring[0].d2h_bytes(),
                "calls_in_flight": len(e2e_ring),
                "api": "FusionPlan.submit_pinned (source='map'): one CUDA-graph launch per call = H2D copy of the pinned decoder maps, "
                       "pair build + Lloyd + ALS + decompose + reconstruction, D2H copy of the log-depth maps",
                "d2h_gbs_at_value": world * K * e2e_ring[0].d2h_bytes() / (e_ms * 1e-3) / 1e9 / world,
                "pcie_measured": pcie,
                "batched_value": world * K * BATCH / (eb_ms * 1e-3),
                "batched_note": "same calls with one graph launch per pass over the ring (host launch rate removed)"},
        "gpu_launches": K * LAUNCHES_PER_STEP,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "als_sparse_kernel (100 ALS iterations per page on the compact form, one warp per page; issue/latency "
                               "bound by construction - its inputs are 16 KB per page - see DESIGN.md 4.1)",
                     "algorithmic_bytes_per_launch": ab["als_sparse_kernel"] * BATCH, "launch_seconds": t_sps, "peak_source": peak_src,
                     "path_at_value": {"achieved": ab["path"] * BATCH * K / (dev_ms * 1e-3) / 1e9,
                                       "frac": ab["path"] * BATCH * K / (dev_ms * 1e-3) / 1e9 / peak,
                                       "note": "whole path (quantize + ALS + decompose + reconstruct contract bytes) over the measured step "
                                               "time: the figure north_star's 60 % target is stated on"},
                     "streaming_kernel": {"kernel": "als_sparsify_raw_kernel (reads every raw pair matrix once: structure check + Lloyd)",
                                          "achieved": ab["sparsify_kernel"] * BATCH / t_spf / 1e9,
                                          "frac": ab["sparsify_kernel"] * BATCH / t_spf / 1e9 / peak,
                                          "algorithmic_bytes_per_launch": ab["sparsify_kernel"] * BATCH, "launch_seconds": t_spf,
                                          "traffic": traffic_spf}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lit_rate, vec_rate, dt = cpu_baseline_port(args.cpu_images)
        out["cpu_baseline"] = {"value": lit_rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{args.cpu_images} images (one call) of the batch-16 workload through oracle/literal.py, "
                                         f"the loop-for-loop restatement of the reference ({dt:.1f} s)",
                               "vectorised_port_value": vec_rate, "host_cpus": os.cpu_count()}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import fusion_ref as fr
    from oracle import literal as lit
    torch.set_num_threads(os.cpu_count() or 1)
    books = fr.load_codebooks()
    K, W = args.steps, args.warmup
    per_step = 1   # images per step: a bounded sample of the batch-16 workload
    batches = [fr.synthetic_batch(per_step, SCALES, seed=batch_seed(0, i)) for i in range(4)]
    # The literal port (the reference's own loop structure) costs ~0.3-1 s per image.  If K steps of it would
    # not finish within a few minutes, the vectorised restatement of the same arithmetic is timed instead
    # and the line says so: the bound on the run time takes precedence.
    t0 = time.perf_counter()
    lit.fusion_forward_literal(*batches[0], books)
    t_lit = time.perf_counter() - t0
    use_literal = (K + W) * t_lit <= args.ref_budget_s
    if use_literal:
        step = lambda b: lit.fusion_forward_literal(*b, books)   # noqa: E731
        what = "oracle/literal.py (loop-for-loop port of the reference)"
    else:
        step = lambda b: fr.fusion_forward(*b, books)            # noqa: E731
        what = (f"oracle/fusion_ref.py (vectorised port: {K + W} steps of the literal port at {t_lit:.2f} s each "
                f"would exceed the {args.ref_budget_s:.0f} s budget)")
    for i in range(W):
        step(batches[i % 4])
    t0 = time.perf_counter()
    for i in range(K):
        step(batches[i % 4])
    dt = time.perf_counter() - t0
    val = K * per_step / dt
    sample = f"{per_step} image per step of the batch-16 scales-8/16/32 workload, {what}"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] on the host CPU: decoder maps -> pair build + Lloyd + ALS + decompose + weighted "
                               "reconstruction, reference algorithm", "batch": per_step, "scales": list(SCALES),
                   "literal_port_s_per_image": t_lit},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="batches in flight (CUDA stream branches)")
    ap.add_argument("--e2e-lanes", type=int, default=8, help="host calls in flight in the e2e measurement")
    ap.add_argument("--ring", type=int, default=128, help="resident input batches (ring > L2)")
    ap.add_argument("--cpu-images", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-table", action="store_true")
    ap.add_argument("--scales", default="8,16,32", help="relative decoder scales (default: BASELINE configs[1]); "
                    "8,16,32,64 is the configuration network/RDM_Net.py:96-97 names")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="--impl reference: wall-clock budget for the literal port")
    args = ap.parse_args()
    global SCALES
    SCALES = tuple(int(v) for v in args.scales.split(","))
    if args.impl == "reference":
        args.steps = 20 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else max(args.warmup, 1)
        run_reference(args)
    else:
        args.steps = 2000 if args.steps is None else args.steps
        args.warmup = 20 if args.warmup is None else max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
