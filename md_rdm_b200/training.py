"""The reference's training step on the fast path (BASELINE config 3).

Mirrors `RelativeDephModule.training_step` + `compute_final_depth` + `compute_ordinal_target` + `normalize`
(network/module.py:64-97, 119-149, "MOD") for the configuration RN:96-97 names (decoder 1 + relative decoders):
the CNN is out of scope, so the step starts from what the decoders emit - the DORN logits of decoder 1 and the
relative decoder maps - and ends with the gradients of `Weights` (and of the DORN logits, through Ordinal_Loss).

    DORN head (RN:313-345)                       rdm::dorn_regression        -> x_d1 (int64 counts), ord (f64)
    pair build + Lloyd + ALS (RN:358-396)        FusionPlan.run_als          -> filled relative maps
    decompose + combine + recombine (RN:117-133, MOD:132)   ops.fuse_tail_autograd -> y_hat, final depth (autograd to Weights)
    GT: resize 226->128 (MOD:68), mask (MOD:74-78), normalise + decompose n=7 (MOD:123), ordinal d0 (MOD:126)
                                                 rdm::gt_prepare (one launch) or the stand-alone ops
    losses: MSE(final, y) (MOD:89) + detached per-scale component loss (CP:499-510) + Ordinal_Loss (loss.py:17-59)

Gradients: only `Weights` and the DORN logits receive any; Lloyd severs everything upstream of ALS.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import computations as cp
from . import ops  # noqa: F401
from .fusion import FusionPlan
from .loss import depth2label_sid
from .ops import fuse_tail_autograd, split_yhat

R = torch.ops.rdm


class TrainingStep:
    def __init__(self, batch: int, scales: Sequence[int] = (8, 16, 32), device="cuda", gt_side: int = 226, K: int = 90,
                 fused_gt: bool = True):
        self.device = torch.device(device)
        self.B, self.scales = int(batch), tuple(int(s) for s in scales)
        # overlap: the dense ALS of the 8x8 maps runs beside the page ALS (a lone batch is latency-bound)
        self.plan = FusionPlan(self.B, self.scales, "map", device=self.device, want_bins=False, overlap=True)
        self._gt_stream: Optional[torch.cuda.Stream] = None
        dev, f32, f64 = self.device, torch.float32, torch.float64
        self.weights = torch.ones((self.plan.n_weights,), dtype=f32, device=dev, requires_grad=True)
        self.logits = torch.zeros((self.B, 2 * K, 8, 8), dtype=f32, device=dev, requires_grad=True)
        self.y_raw = torch.ones((self.B, 1, gt_side, gt_side), dtype=f64, device=dev)
        self.fused_gt = bool(fused_gt) and hasattr(R, "gt_prepare")
        self._host: Optional[Dict[str, torch.Tensor]] = None
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_out: Optional[Dict[str, torch.Tensor]] = None
        self._loss_host: Optional[torch.Tensor] = None
        self._grad_host: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ inputs
    def load(self, rel: Sequence[torch.Tensor], y_raw: torch.Tensor, logits: torch.Tensor, weights: torch.Tensor) -> None:
        with torch.no_grad():
            for s, t in zip(self.scales, rel):
                self.plan.src[s].copy_(t.reshape(self.plan.src[s].shape), non_blocking=True)
            self.y_raw.copy_(y_raw, non_blocking=True)
            self.logits.copy_(logits, non_blocking=True)
            self.weights.copy_(weights.reshape(-1), non_blocking=True)
        self._host_src = dict(rel=[t.clone() for t in rel], y_raw=y_raw.clone(), logits=logits.clone(), weights=weights.reshape(-1).clone())

    def host_copy(self) -> Dict[str, torch.Tensor]:
        return self._host_src

    # ------------------------------------------------------------------ ground truth (MOD:68, 74-78, 119-127)
    def targets(self):
        """(y masked 128x128 f64, packed level-major target pyramid or None, component targets [d0, f1..f7] f64,
        ordinal target (B,1,8,8) int32)."""
        if self.fused_gt:
            y, pyr, ord_t = R.gt_prepare(self.y_raw)
            return y, pyr, ops.unpack_pyramid(pyr, self.B, 128, False), ord_t
        y = cp.resize(self.y_raw, 128)                                         # MOD:68
        y = (y * (y > 0)) + ((y <= 0) + 1e-4)                                  # MOD:74-78
        comps = cp.decompose_depth_map([], R.gm_normalize(y), 7)[::-1]        # MOD:123, MOD:145-149
        ord_t = depth2label_sid(cp.resize(y, 8))                               # MOD:126 / MOD:134-143 (same tensor)
        comps[0] = cp.decompose_depth_map([], R.gm_normalize(ord_t.long()), 3)[::-1][0]
        return y, None, comps, ord_t

    # ------------------------------------------------------------------ one step
    def step(self) -> Dict[str, torch.Tensor]:
        self.weights.grad = None
        self.logits.grad = None
        cur = torch.cuda.current_stream(self.device)
        # The ground truth does not depend on the network: it is prepared on a second stream that forks here and
        # joins before the losses (plain stream waits: capturable, two parallel branches of the CUDA graph).  The fork
        # also orders this step's allocations on that stream behind the previous step's work.
        if self._gt_stream is None:
            self._gt_stream = torch.cuda.Stream(self.device)
        self._gt_stream.wait_stream(cur)
        with torch.cuda.stream(self._gt_stream):
            # the second branch: everything that does not wait for the ALS - ground truth, the DORN head (RN:313-345) and
            # its Ordinal_Loss (loss.py:17-59); autograd runs their backward kernels on this stream as well
            y, pyr, comps, ord_t = self.targets()
            x_d1, ord_ = R.dorn_regression(self.logits)
            ord_loss = R.ordinal_loss(ord_, ord_t)
        rel = self.plan.run_als()                                              # RN:358-396 for every relative decoder
        cur.wait_stream(self._gt_stream)                                       # join: the tail needs decoder 1's map
        # (a training step is alone on the GPU: four CTAs per image make the tail launch shorter)
        final, yhat = fuse_tail_autograd(x_d1, [rel[s] for s in self.scales], self.weights, bands=4)   # RN:117-133 + MOD:132
        # CP:499-510: per-scale MSE, summed through torch.as_tensor => detached: it only enters the VALUE of the loss, so it
        # is taken on the second stream while this one goes on to the MSE and the backward pass (no host sync anywhere)
        self._gt_stream.wait_stream(cur)
        with torch.cuda.stream(self._gt_stream), torch.no_grad():
            if pyr is not None:
                fine = R.component_loss(yhat, pyr, self.plan.kmax)
            else:
                fine = torch.stack([torch.nn.functional.mse_loss(a.double(), b) for a, b in zip(split_yhat(yhat, self.plan.kmax), comps)]).sum()
        mse = torch.nn.functional.mse_loss(final, y)                           # MOD:89
        (mse + ord_loss).backward()                                            # MOD:90-95: `fine` carries no gradient
        cur.wait_stream(self._gt_stream)
        loss = mse.detach() + fine + ord_loss.detach()                         # MOD:90-92, in the reference's order
        return {"loss": loss.detach(), "mse": mse.detach(), "fine": fine, "ord": ord_loss.detach(), "final": final.detach()}

    # ------------------------------------------------------------------ CUDA graph of the whole step
    def capture(self) -> None:
        """Record one whole step (forward, losses, backward) into a CUDA graph: the step is ~60 small launches, so
        replaying it removes the host dispatch that otherwise dominates.  Inputs are the static buffers filled by
        `load()` / `step_from_host()`; the loss terms and the gradients land in static tensors."""
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):                      # warm-up outside capture (lazy initialisations, allocator)
                    self.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.step()
            self._graph, self._graph_out = g, out

    def replay(self) -> Dict[str, torch.Tensor]:
        """One step from the captured graph; `weights.grad` / `logits.grad` and the returned loss terms are the
        graph's static tensors (overwritten by the next replay)."""
        if self._graph is None:
            self.capture()
        self._graph.replay()
        return self._graph_out

    def launches_per_step(self) -> int:
        """Kernels of librdm_b200 per step (torch's own elementwise kernels for the losses come on top)."""
        gt = 2 if self.fused_gt else 7      # gt_prepare + component_loss
        # dorn_regression (2) + ALS launches + fuse_tail + GT + ordinal loss (2) + backward: ordinal, dorn, tail (2)
        return 2 + (self.plan.launches_per_run - 1) + 1 + gt + 2 + 2 + 2

    # ------------------------------------------------------------------ host end to end
    def pin_host(self) -> None:
        h = self._host_src
        self._host = dict(rel=[t.pin_memory() for t in h["rel"]], y_raw=h["y_raw"].pin_memory(), logits=h["logits"].pin_memory())
        self._loss_host = torch.zeros((), dtype=torch.float64).pin_memory()
        self._grad_host = torch.zeros((self.plan.n_weights,), dtype=torch.float32).pin_memory()

    def h2d_bytes(self) -> int:
        h = self._host_src
        return sum(t.numel() * t.element_size() for t in h["rel"]) + h["y_raw"].numel() * 8 + h["logits"].numel() * 4

    def d2h_bytes(self) -> int:
        return 8 + 4 * self.plan.n_weights

    def step_from_host(self) -> float:
        """Pinned host inputs -> device, one step, loss and Weights gradient back on the host (synchronous)."""
        h = self._host
        with torch.no_grad():
            for s, t in zip(self.scales, h["rel"]):
                self.plan.src[s].copy_(t, non_blocking=True)
            self.y_raw.copy_(h["y_raw"], non_blocking=True)
            self.logits.copy_(h["logits"], non_blocking=True)
        out = self.replay() if self._graph is not None else self.step()
        self._loss_host.copy_(out["loss"], non_blocking=True)
        self._grad_host.copy_(self.weights.grad, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return float(self._loss_host)
