"""Batched fast entry point of the fusion path: all relative decoder scales of one batch in a fixed
number of launches (`FusionPlan.launches_per_run`: compact page form, ALS on the compact pages, dense ALS
of the 8x8 maps, arg-min/normalise/re-tile, fused decompose+combine+recombine), with every buffer
preallocated, optional CUDA-graph replay and a pinned-host end-to-end call.

This sits beside the drop-in names (md_rdm_b200.computations / rdm_net) and computes exactly
what RN:103-133 + network/module.py:132 compute for decoder 1 plus relative decoders at
`scales`: it is what `bench.py` times.

source = "map": inputs are the decoder maps; pair build + Lloyd + ALS are fused and the pair
                matrices never exist in HBM (the call a user of RDM_Net makes).
source = "raw": inputs are materialised raw pair matrices (f32 64x64 for scale 8, f64
                P x 256 x 64 for scales >= 16), as BASELINE.json's standalone fusion config
                states it: quantize + ALS + decompose + reconstruct.
"""
from __future__ import annotations

import contextlib
import ctypes
import weakref
from collections import OrderedDict
from ctypes import c_void_p
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi
from ._cabi import AlsScale, check, i32_array, load, ptr_array
from .codebooks import Quantization, default_quantization
from .ops import split_yhat, tail_layout

LIMIT_8 = 30
LIMIT_PAGE = 100


class FusionPlan:
    def __init__(self, n_images: int, scales: Sequence[int] = (8, 16, 32), source: str = "map", group: Optional[int] = None,
                 device="cuda", quant: Optional[Quantization] = None, want_bins: bool = True, want_values: bool = False,
                 want_A: bool = False, limit_8: int = LIMIT_8, limit_page: int = LIMIT_PAGE, overlap: bool = False, flags: int = 0,
                 compact_result: bool = False):
        if source not in ("map", "raw"):
            raise ValueError("source must be 'map' or 'raw'")
        self.lib = load()                      # raises if librdm_b200.so is missing: no fallback
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FusionPlan needs a CUDA device (md_rdm_b200 has no CPU path)")
        self.N = int(n_images)
        self.group = int(group) if group else self.N
        if self.N % self.group:
            raise ValueError("n_images must be a multiple of group")
        self.scales = tuple(int(s) for s in scales)
        if any(s not in (8, 16, 32, 64, 128) for s in self.scales):
            raise ValueError("relative decoder scales must be in {8,16,32,64,128}")
        # decoder 10 (128x128, RN:61): pair build + Lloyd + ALS run in the same launches as the other scales (64 pages
        # per image); its 175 KB f64 pyramid does not fit beside the others in one CTA's shared memory, so the tail of
        # such a plan is COMPOSED from the stand-alone kernels (decompose, log_stack, make_pred, recombination)
        # instead of the single fuse_tail launch
        self.composed_tail = 128 in self.scales
        self.source = source
        self.quant = quant or default_quantization()
        dev, N = self.device, self.N
        f32, f64 = torch.float32, torch.float64
        K, w_off, kmax, n_w = tail_layout(self.scales)
        self.K, self.w_off, self.kmax, self.n_weights = K, w_off, kmax, n_w

        # ---- inputs (static buffers: copy into them, then run()).  x_d1 and the per-scale sources are
        # views of ONE device byte buffer so that a host call needs a single H2D copy.
        shapes = [("x_d1", (N, 1, 8, 8), torch.int64)]
        for s in self.scales:
            if source == "map":
                shapes.append((s, (N, 1, s, s), f32))
            elif s == 8:
                shapes.append((s, (N, 64, 64), f32))
            else:
                shapes.append((s, (N, (s // 16) ** 2, 256, 64), f64))
        self._layout, off = [], 0
        for key, shape, dt in shapes:
            nbytes = int(torch.empty((), dtype=dt).element_size()) * int(torch.Size(shape).numel())
            self._layout.append((key, shape, dt, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self._in_bytes = off
        self._in_dev = torch.zeros((off,), dtype=torch.uint8, device=dev)
        views = self._views(self._in_dev)
        self.x_d1 = views["x_d1"]
        self.x_d1.fill_(1)
        self.weights = torch.ones((n_w,), dtype=f32, device=dev)
        self.src: Dict[int, torch.Tensor] = {s: views[s].fill_(1) for s in self.scales}
        # ---- outputs
        self.rel: Dict[int, torch.Tensor] = {}
        self.pages: Dict[int, torch.Tensor] = {}
        self.bins: Dict[int, torch.Tensor] = {}
        self.values: Dict[int, torch.Tensor] = {}
        self.record: Dict[int, torch.Tensor] = {}
        self.kstar: Dict[int, torch.Tensor] = {}
        self._ws: Dict[int, torch.Tensor] = {}
        self._side: Optional[torch.cuda.Stream] = None   # fork/join branch of run()
        self.overlap = bool(overlap)
        self._tables = {}
        G = N // self.group
        descs = (AlsScale * len(self.scales))()
        for i, s in enumerate(self.scales):
            rows = 64 if s == 8 else 256
            P = 1 if s == 8 else (s // 16) ** 2
            limit = limit_8 if s == 8 else limit_page
            if source == "map":
                kind = _cabi.SRC_MAP_F32
            elif s == 8:
                kind = _cabi.SRC_RAW_F32
            else:
                kind = _cabi.SRC_RAW_F64
            thr, lvl = self.quant.device_tables(s, dev)
            self._tables[s] = (thr, lvl)
            self.rel[s] = torch.empty((N, 1, s, s), dtype=f32, device=dev)
            self.pages[s] = torch.empty((N, P, rows), dtype=f32, device=dev)
            self.record[s] = torch.empty((G, P, limit + 1), dtype=f32, device=dev)
            self.kstar[s] = torch.empty((G, P), dtype=torch.int32, device=dev)
            self._ws[s] = torch.empty((N * self.lib.rdm_als_ws_floats(rows, P, limit),), dtype=f32, device=dev)
            if want_bins:
                self.bins[s] = torch.empty((N, P, rows, 64), dtype=torch.uint8, device=dev)
            if want_values:
                self.values[s] = torch.empty((N, P, rows, 64), dtype=f32, device=dev)
            d = descs[i]
            d.src, d.src_kind, d.rows, d.pages, d.side, d.limit, d.flags = self.src[s].data_ptr(), kind, rows, P, s, limit, int(flags)
            d.thresholds, d.levels = thr.data_ptr(), lvl.data_ptr()
            d.bins_out = self.bins[s].data_ptr() if want_bins else None
            d.values_out = self.values[s].data_ptr() if want_values else None
            d.pages_out, d.map_out, d.ws = self.pages[s].data_ptr(), self.rel[s].data_ptr(), self._ws[s].data_ptr()
            d.record_out, d.kstar_out = self.record[s].data_ptr(), self.kstar[s].data_ptr()
        self._descs = descs
        self.yhat = torch.empty((N, (4 ** (kmax + 1) - 1) // 3), dtype=f32, device=dev)
        self.depth = torch.empty((N, 1, 128, 128), dtype=f64, device=dev)
        # Opt-in compact result: no decoder is finer than 2^kmax, so the log-depth map is constant on blocks of
        # 2^(7-kmax) pixels; `depth_compact` (N,1,2^kmax,2^kmax) holds every distinct value (depth == its nearest-
        # neighbour upsampling, bit for bit) in 1/4^(7-kmax) of the bytes.  The pinned end-to-end call then returns it
        # INSTEAD of the full map (expand_compact() rebuilds the full map on the host).
        self.compact_result = bool(compact_result)
        self.depth_compact = torch.empty((N, 1, 1 << kmax, 1 << kmax), dtype=f64, device=dev) if compact_result else None
        self.A: List[torch.Tensor] = [torch.empty((N, K[k], 4 ** k), dtype=f64, device=dev) for k in range(kmax + 1)] if want_A else []
        self._rel_ptrs = ptr_array([self.rel[s].data_ptr() for s in self.scales])
        self._sides = i32_array(list(self.scales))
        self._a_ptrs = ptr_array([self.A[k].data_ptr() if (want_A and k <= kmax) else None for k in range(8)])
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_e2e: Optional[torch.cuda.CUDAGraph] = None
        self._pinned = None
        has_pages, has_8 = any(s > 8 for s in self.scales), 8 in self.scales
        # sparsify + compact-page ALS (page scales), dense ALS (8x8 maps and the page fallback), fused tail
        self.launches_per_run = (2 if has_pages else 0) + (1 if self.scales else 0) + 1
        if self.composed_tail:   # gm + decompose per decoder, log_stack + make_pred per slot, recombination
            self.launches_per_run += 1 + len(self.scales) + 2 * (kmax + 1)
            if compact_result:
                raise ValueError("compact_result needs kmax < 7: with a 128x128 decoder the map has no constant blocks")

    @staticmethod
    def phase_masks() -> Dict[str, int]:
        """The ALS launches by `rdm_als_fused_phases` mask bit (bench.py times them one by one)."""
        return {"als_sparsify": _cabi.PHASE_SPARSIFY, "als_sparse": _cabi.PHASE_PAGES, "als_dense": _cabi.PHASE_DENSE}

    @staticmethod
    def kernel_names() -> Dict[str, str]:
        return {"als_sparsify": "als_sparsify_raw_kernel / als_sparsify_map_kernel (structure check + Lloyd: reads every raw pair matrix once)",
                "als_sparse": "als_pages_kernel (100 ALS iterations per page on the compact form, one CTA per (batch, page), one warp per "
                              "image, batch-wide arg-min + normalise + re-tile inside; small launches take its cluster form "
                              "als_pages_cluster_kernel: 4 CTAs of 4 warps per (batch, page))",
                "als_dense": "als_kernel (dense ALS of the 8x8 maps, one cluster per batch, arg-min inside; fallback for pages without pair structure)",
                "fuse_tail": "fuse_tail_kernel (decompose + combine + recombine)"}

    def _views(self, buf: torch.Tensor):
        """Typed views (x_d1 and one source per scale) of a packed input byte buffer."""
        return {key: buf[off:off + nbytes].view(dt).view(shape) for key, shape, dt, off, nbytes in self._layout}

    # ------------------------------------------------------------------ device path
    def run(self, overlap: Optional[bool] = None) -> torch.Tensor:
        """Enqueue the whole path on the current stream; returns the (N,1,128,128) f64 log-depth buffer.

        After the compact page form is built (PHASE_SPARSIFY) the ALS on the compact pages (PHASE_PAGES) and the
        dense ALS of the 8x8 maps plus the fallback for pages without pair structure (PHASE_DENSE, reads the flags
        the sparsify launch wrote) are independent launches - so with `overlap` the dense launch goes to a side
        stream that forks from and joins the current one (plain events: capturable into a CUDA graph, where it
        becomes two parallel branches).  It shortens one call alone but costs throughput with many calls in flight
        (the join costs more than the idle SMs it fills), so it is off unless the plan was built with
        `overlap=True` (latency-bound callers)."""
        with self._guard():
            ov = self.overlap if overlap is None else overlap
            self._run_als(ov)
            self._run_tail(ov)
        return self.depth

    def _guard(self):
        """Make the plan's device current for the launches (a plan built for cuda:1 may be run while cuda:0 is
        current); free when it already is."""
        if torch.cuda.current_device() == (self.device.index if self.device.index is not None else torch.cuda.current_device()):
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)

    def run_als(self, overlap: Optional[bool] = None) -> Dict[int, torch.Tensor]:
        """Only the ALS launches (pair build / Lloyd / ALS / arg-min / re-tile): fills and returns `self.rel`."""
        with self._guard():
            self._run_als(self.overlap if overlap is None else overlap)
        return self.rel

    def _run_als(self, overlap: bool) -> None:
        cur = torch.cuda.current_stream(self.device)
        st = c_void_p(cur.cuda_stream)
        lib = self.lib
        if self.scales:
            n, descs = len(self.scales), self._descs
            if overlap and any(s > 8 for s in self.scales) and 8 in self.scales:
                if self._side is None:
                    self._side = torch.cuda.Stream(self.device)
                side = self._side
                check(lib.rdm_als_fused_phases(descs, n, self.N, self.group, _cabi.PHASE_SPARSIFY, st), "rdm_als_fused_phases")
                side.wait_stream(cur)
                check(lib.rdm_als_fused_phases(descs, n, self.N, self.group, _cabi.PHASE_DENSE, c_void_p(side.cuda_stream)), "rdm_als_fused_phases")
                check(lib.rdm_als_fused_phases(descs, n, self.N, self.group, _cabi.PHASE_PAGES, st), "rdm_als_fused_phases")
                cur.wait_stream(side)
            else:
                check(lib.rdm_als_fused(descs, n, self.N, self.group, st), "rdm_als_fused")

    def run_from_features(self, feats: Dict[int, torch.Tensor], conv_w: Dict[int, torch.Tensor], conv_b: Dict[int, Optional[torch.Tensor]],
                          write_maps: bool = False) -> torch.Tensor:
        """start one step earlier, from the feature blocks the relative decoders' 1x1 conv heads
        consume (RN:146, RN:157): feats[s] (N,C_s,s,s) f32, conv_w[s] = conv1.weight (1,C_s,1,1) or (C_s,), conv_b[s] =
        conv1.bias (1,) or None.  One launch per scale does the conv AND builds the compact pair-matrix form, so
        for s >= 16 the decoder map never reaches HBM (unless `write_maps`); then the page ALS, the dense ALS (8x8 map)
        and the tail run as in run().  Needs source == "map".  `x_d1` and `weights` are taken from the plan's buffers."""
        if self.source != "map":
            raise RuntimeError("run_from_features needs a plan built with source='map'")
        lib = self.lib
        with self._guard():
            st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self.enqueue_conv_heads(feats, conv_w, conv_b, write_maps)
            if self.scales:
                check(lib.rdm_als_fused_phases(self._descs, len(self.scales), self.N, self.group, _cabi.PHASE_PAGES | _cabi.PHASE_DENSE, st),
                      "rdm_als_fused_phases")
            self._run_tail()
        return self.depth

    def enqueue_conv_heads(self, feats, conv_w, conv_b, write_maps: bool = False) -> None:
        """The conv-head launches of run_from_features alone (one per scale)."""
        st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        for i, s in enumerate(self.scales):
            f = feats[s]
            if f.dtype != torch.float32 or f.dim() != 4 or f.shape[0] != self.N or f.shape[2] != s or f.shape[3] != s or not f.is_contiguous():
                raise RuntimeError(f"run_from_features: feats[{s}] must be a contiguous ({self.N},C,{s},{s}) f32 tensor")
            w = conv_w[s].reshape(-1).float().contiguous()
            b = conv_b.get(s)
            b = None if b is None else b.reshape(-1).float().contiguous()
            if w.numel() != f.shape[1]:
                raise RuntimeError(f"run_from_features: conv_w[{s}] has {w.numel()} weights for {f.shape[1]} channels")
            map_out = c_void_p(self.src[s].data_ptr()) if (s == 8 or write_maps) else c_void_p(0)
            desc = ctypes.pointer(self._descs[i]) if s >= 16 else None
            check(self.lib.rdm_conv_head_f32(c_void_p(f.data_ptr()), c_void_p(w.data_ptr()), c_void_p(b.data_ptr()) if b is not None else c_void_p(0),
                                             self.N, int(f.shape[1]), s, map_out, desc, st), "rdm_conv_head_f32")

    def run_als_phase(self, phase_mask: int) -> None:
        """Only the ALS launches selected by phase_mask (`_cabi.PHASE_*`): used by bench.py to time the kernels one by one."""
        with self._guard():
            st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            check(self.lib.rdm_als_fused_phases(self._descs, len(self.scales), self.N, self.group, phase_mask, st), "rdm_als_fused_phases")

    def _run_tail_composed(self) -> None:
        """RN:117-133 + MOD:132 from the stand-alone kernels (plans with the 128x128 decoder)."""
        from . import computations as cp
        R = torch.ops.rdm
        B = self.N
        rows = [cp.decompose_depth_map([], R.gm_normalize(self.x_d1), 3)[::-1]]
        for s in self.scales:
            rows.append(cp.decompose_depth_map([], self.rel[s], s.bit_length() - 1, relative_map=True)[::-1])
        A = cp.relative_fine_detail_matrix(rows, True)
        ws, off = [], 0
        for k in range(self.kmax + 1):
            ws.append(self.weights[off:off + self.K[k]].view(self.K[k], 1))
            off += self.K[k]
        for k, a in enumerate(A):
            if self.A:
                self.A[k].copy_(a)
        y_hat = cp.make_pred(ws, A, True, False)
        self.yhat.copy_(torch.cat([y.reshape(B, -1) for y in y_hat], 1))
        self.depth.copy_(cp.recombination(list(y_hat)))

    def _run_tail(self, latency_bound: Optional[bool] = None) -> None:
        if latency_bound is None:
            latency_bound = self.overlap
        if self.composed_tail:
            return self._run_tail_composed()
        st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        # a latency-bound caller (overlap=True) also takes four CTAs per image in the tail; otherwise the
        # launch picks the count from the batch (one per image from 16 images up: best with many calls in flight)
        check(self.lib.rdm_fuse_tail_bands(c_void_p(self.x_d1.data_ptr()), self._rel_ptrs, self._sides, len(self.scales),
                                           c_void_p(self.weights.data_ptr()), self.N, c_void_p(self.yhat.data_ptr()),
                                           c_void_p(self.depth.data_ptr()),
                                           c_void_p(self.depth_compact.data_ptr()) if self.depth_compact is not None else c_void_p(0),
                                           self._a_ptrs, 4 if latency_bound else 0, st), "rdm_fuse_tail_bands")

    def expand_compact(self, compact: torch.Tensor) -> torch.Tensor:
        """(N,1,2^kmax,2^kmax) compact result -> the (N,1,128,128) map it stands for (nearest-neighbour, exact)."""
        r = 1 << (7 - self.kmax)
        return compact.repeat_interleave(r, 2).repeat_interleave(r, 3)

    def run_tail(self) -> None:
        with self._guard():
            self._run_tail()

    def capture(self) -> None:
        """Record run() into a CUDA graph (the library does no host sync and no allocation)."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.run()                      # warm-up outside capture (function attributes, module load)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run()
            self._graph = g

    def replay(self) -> torch.Tensor:
        if self._graph is None:
            self.capture()
        self._graph.replay()
        return self.depth

    # ------------------------------------------------------------------ inputs
    def load_inputs(self, x_d1: torch.Tensor, srcs: Sequence[torch.Tensor], weights: Optional[torch.Tensor] = None) -> None:
        self.x_d1.copy_(x_d1.reshape(self.x_d1.shape), non_blocking=True)
        for s, t in zip(self.scales, srcs):
            self.src[s].copy_(t.reshape(self.src[s].shape), non_blocking=True)
        if weights is not None:
            self.weights.copy_(weights.reshape(-1), non_blocking=True)

    def yhat_list(self) -> List[torch.Tensor]:
        """The reference's y_hat: list of (N,1,2^k,2^k) f32 views, k = 0..kmax."""
        return split_yhat(self.yhat, self.kmax)

    # ------------------------------------------------------------------ host end-to-end
    def _host_buffers(self):
        """Pinned host staging: one packed input buffer (typed views `x_d1`, `src[s]`) and the output."""
        if self._pinned is None:
            h_in = torch.zeros((self._in_bytes,), dtype=torch.uint8).pin_memory()
            v = self._views(h_in)
            self._pinned = dict(packed=h_in, x_d1=v["x_d1"], src={s: v[s] for s in self.scales},
                                depth=torch.empty((self.depth_compact if self.compact_result else self.depth).shape, dtype=self.depth.dtype).pin_memory())
        return self._pinned

    def h2d_bytes(self) -> int:
        return self.x_d1.numel() * 8 + sum(t.numel() * t.element_size() for t in self.src.values())

    def d2h_bytes(self) -> int:
        return (self.depth_compact if self.compact_result else self.depth).numel() * 8

    def _enqueue_e2e(self) -> None:
        hb = self._host_buffers()
        with self._guard():
            self._in_dev.copy_(hb["packed"], non_blocking=True)      # one H2D copy for all inputs
            self.run()
            hb["depth"].copy_(self.depth_compact if self.compact_result else self.depth, non_blocking=True)   # D2H of the fused log-depth maps

    def capture_e2e(self) -> None:
        """CUDA graph of the whole host call: H2D copy node, the kernels of run(), D2H copy node."""
        self._host_buffers()
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue_e2e()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_e2e()
            self._graph_e2e = g

    def submit_pinned(self, use_graph: bool = True) -> None:
        """Asynchronous host call on the current stream: inputs are taken from this plan's pinned
        staging buffers (`_host_buffers()`), the result lands in its pinned `depth` buffer."""
        if use_graph:
            if self._graph_e2e is None:
                self.capture_e2e()
            self._graph_e2e.replay()
        else:
            self._enqueue_e2e()

    def run_pinned(self, use_graph: bool = True) -> torch.Tensor:
        """Synchronous host call with the inputs already staged in the pinned buffers."""
        self.submit_pinned(use_graph)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pinned["depth"]

    def run_host(self, x_d1: torch.Tensor, srcs: Sequence[torch.Tensor], use_graph: bool = True) -> torch.Tensor:
        """Host tensors in -> host (pinned) log-depth out, synchronous: stage, H2D copy, the three
        launches, D2H copy, stream sync."""
        hb = self._host_buffers()
        hb["x_d1"].copy_(x_d1.reshape(hb["x_d1"].shape))
        for s, t in zip(self.scales, srcs):
            hb["src"][s].copy_(t.reshape(hb["src"][s].shape))
        return self.run_pinned(use_graph)


def capture_lane(plans: Sequence[FusionPlan], e2e: bool = False):
    """A CUDA graph that runs `plans` one after the other on ONE stream, plus that stream.  Several lanes
    replayed round-robin keep several batches in flight without any join between them (a fork/join graph
    over the whole ring drains the GPU at every replay)."""
    dev = plans[0].device
    with torch.cuda.device(dev):
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for p in plans:                 # warm-up outside capture
                p._enqueue_e2e() if e2e else p.run()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for p in plans:
                p._enqueue_e2e() if e2e else p.run()
    return g, stream


def capture_ring(plans: Sequence[FusionPlan], n_streams: int, e2e: bool = False) -> torch.cuda.CUDAGraph:
    """ONE CUDA graph that runs every plan of `plans` once, plan i on branch i % n_streams (fork/join
    inside the capture): a whole ring of batches per host launch, so the host launch rate
    (~25-50 us per Python graph replay) does not bound a multi-batch pipeline."""
    dev = plans[0].device
    with torch.cuda.device(dev):
        side = [torch.cuda.Stream() for _ in range(n_streams)]
        for p in plans:                 # warm-up outside capture
            p._enqueue_e2e() if e2e else p.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream()
            for s in side:
                s.wait_stream(cur)
            for i, p in enumerate(plans):
                with torch.cuda.stream(side[i % n_streams]):
                    p._enqueue_e2e() if e2e else p.run()
            for s in side:
                cur.wait_stream(s)
    return g


_PLAN_CACHE_MAX = 8
_plans: "OrderedDict[tuple, tuple]" = OrderedDict()   # key -> (plan, weakref to the caller's Quantization or None)


def clear_plans() -> None:
    """Drop the cached plans of `fuse_maps` (frees their device workspaces)."""
    _plans.clear()


def fuse_maps(x_d1: torch.Tensor, rel_maps: Sequence[torch.Tensor], weights: Sequence[torch.Tensor] | torch.Tensor,
              quant: Optional[Quantization] = None):
    """Functional form: decoder outputs in, (depth (B,1,128,128) f64, y_hat list, filled relative maps) out.
    x_d1 (B,1,8,8) int64 DORN counts, rel_maps[i] (B,1,s_i,s_i) f32; weights: the `Weights.weight_list`
    (list of (K,1)) or a flat tensor.  One arg-min group = the whole call, like one reference forward
    (RN:103-133 + network/module.py:132).

    Gradients: when autograd is recording and any weight requires grad, the tail goes through
    `ops.fuse_tail_autograd`, so `depth` and `y_hat` carry gradients to the `Weights` parameters (the only
    gradients the training loss needs); the relative maps get none (Lloyd severs them).  Otherwise the
    whole call is the inference plan.  Calls are single-stream: a cached plan owns static input / output buffers
    (results are cloned before returning), so two threads or streams must not share one (B, scales, device, quant)."""
    B = x_d1.shape[0]
    scales = tuple(int(r.shape[2]) for r in rel_maps)
    key = (B, scales, str(x_d1.device), id(quant) if quant is not None else None)
    hit = _plans.get(key)
    if hit is not None and quant is not None and hit[1]() is not quant:
        hit = None                                   # id() reused by another object after collection: stale codebooks
    if hit is None:
        plan = FusionPlan(B, scales, "map", device=x_d1.device, quant=quant, want_bins=False)
        _plans[key] = (plan, weakref.ref(quant) if quant is not None else None)
        while len(_plans) > _PLAN_CACHE_MAX:
            _plans.popitem(last=False)
    else:
        plan = hit[0]
        _plans.move_to_end(key)
    w_list = [weights] if torch.is_tensor(weights) else [w for w in weights if w.numel()]
    needs_grad = torch.is_grad_enabled() and any(w.requires_grad for w in w_list)
    flat = torch.cat([w.reshape(-1) for w in w_list]).float()
    if needs_grad:
        from .ops import fuse_tail_autograd, split_yhat as _split
        plan.load_inputs(x_d1, rel_maps, None)
        rel = [t.clone() for t in plan.run_als().values()]
        depth, yhat = fuse_tail_autograd(x_d1.contiguous(), rel, flat)
        return depth, _split(yhat, plan.kmax), rel
    plan.load_inputs(x_d1, rel_maps, flat.detach())
    plan.run()
    return plan.depth.clone(), [y.clone() for y in plan.yhat_list()], [plan.rel[s].clone() for s in scales]
