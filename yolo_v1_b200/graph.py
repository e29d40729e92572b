"""CUDA-graph capture of the loss call and its collective (BASELINE config 5).

What `train.py:166-171` runs around the loss per iteration is `loss = lossLayer(pred, target)` followed by the
logging of the four components; under data parallelism the job-wide numbers need one all-reduce of the 5-float terms
vector (yolo_v1_b200/dist.py).  At the reference's batch size (12 x 14 x 14 cells, train.py:38-41) the loss kernel
runs for a few microseconds and everything else is launch plumbing -- so the whole sequence
    fused loss kernel  ->  terms copy  ->  NCCL all_reduce(terms)  [-> / world_size]
is captured ONCE into a CUDA graph over static buffers and replayed per step: one graph launch instead of four
eager launches and their Python.  The small-call kernel (csrc/loss_small.cu) needs no workspace reset, so the graph
holds no memset node either.

PyTorch is plumbing here (graph capture, streams, the process group); the arithmetic is libyolo1_b200.so.
"""
import torch
import torch.distributed as dist

from .loss import yolo_loss_fused

__all__ = ["GraphedLoss"]


class _GraphedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, runner):
        loss, terms = runner.run(pred.detach(), target)
        ctx.runner, ctx.serial, ctx.need = runner, runner.serial, pred.requires_grad
        ctx.mark_non_differentiable(terms)
        return loss, terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        if not ctx.need:
            return None, None, None
        r = ctx.runner
        if r.serial != ctx.serial:
            raise RuntimeError("GraphedLoss: the static gradient buffer was overwritten by a later call before this "
                               "backward() ran; call backward() before the next forward()")
        # a fresh tensor: the static buffer is rewritten by the next replay
        return r.static_grad * grad_loss.to(r.static_grad.dtype), None, None


class GraphedLoss:
    """loss + terms all-reduce as one CUDA graph.

        g = GraphedLoss(batch_size, S, B, C, from_logits=..., group=None, average=True)
        loss = g(pred, target)          # autograd-connected 0-dim tensor; g.global_terms = all-reduced float32[5]
        loss.backward()

    The graph is captured on the first call from that call's shapes / strides / dtypes (later calls must match) on
    a side stream, after one eager warm-up that also initialises the NCCL communicator.  `group=None` uses the
    default process group when torch.distributed is initialised and no collective otherwise.  Call `release()`
    before `destroy_process_group()`."""

    def __init__(self, batch_size, S, B=2, C=20, l_coord=5.0, l_noobj=0.5, coord_mode="reference",
                 from_logits=False, group=None, average=True):
        self.cfg = dict(batch_size=batch_size, S=S, B=B, C=C, l_coord=l_coord, l_noobj=l_noobj,
                        coord_mode=coord_mode, from_logits=from_logits)
        self.group, self.average = group, average
        self.graph = None
        self.serial = 0
        self.static_pred = self.static_target = self.static_grad = None
        self.local_terms = self.global_terms = None

    def _collective(self):
        return dist.is_available() and dist.is_initialized()

    def _body(self):
        yolo_loss_fused(self.static_pred, self.static_target, want_grad=True, out_grad=self.static_grad,
                        out_terms=self.local_terms, workspace=self._ws, **self.cfg)
        self.global_terms.copy_(self.local_terms)
        if self._collective():
            dist.all_reduce(self.global_terms, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                self.global_terms.div_(dist.get_world_size(self.group))

    def _capture(self, pred, target):
        dev = pred.device
        self.static_pred = torch.empty_strided(pred.shape, pred.stride(), dtype=pred.dtype, device=dev)
        self.static_target = torch.empty_strided(target.shape, target.stride(), dtype=torch.float32, device=dev)
        self.static_grad = torch.empty_strided(pred.shape, pred.stride(), dtype=pred.dtype, device=dev)
        self.local_terms = torch.zeros(5, dtype=torch.float32, device=dev)
        self.global_terms = torch.zeros(5, dtype=torch.float32, device=dev)
        self._ws = torch.zeros(1 << 17, dtype=torch.uint8, device=dev)
        self.static_pred.copy_(pred)
        self.static_target.copy_(target)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._body()                      # warm-up: lazy initialisation (NCCL communicator, kernel attributes)
            side.synchronize()
            if self._collective():
                dist.barrier(group=self.group)    # every rank has finished its warm-up collective before any captures
                torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            # thread_local: ProcessGroupNCCL's watchdog thread polls CUDA events while this thread captures; under the
            # default (global) capture mode such a call from another thread invalidates the capture
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = g

    def release(self):
        """Drop the captured graph and its static buffers.  A graph that captured an NCCL collective must be gone
        before `torch.distributed.destroy_process_group()` -- the communicator's teardown otherwise waits for it
        (measured: a 2-rank job hangs in destroy_process_group with the graph alive)."""
        self.graph = None
        self.static_pred = self.static_target = self.static_grad = None
        self.serial += 1          # a backward() still pending on the old buffers raises instead of reading freed memory
        torch.cuda.synchronize()

    def run(self, pred, target):
        """pred / target -> (loss 0-dim clone, global_terms); gradient in self.static_grad until the next call."""
        if self.graph is None:
            self._capture(pred, target)
        if (pred.shape != self.static_pred.shape or pred.stride() != self.static_pred.stride() or
                pred.dtype != self.static_pred.dtype):
            raise ValueError("GraphedLoss was captured for pred %s %s %s" % (
                tuple(self.static_pred.shape), self.static_pred.stride(), self.static_pred.dtype))
        self.static_pred.copy_(pred)
        self.static_target.copy_(target)
        self.graph.replay()
        self.serial += 1
        return self.local_terms[4].clone(), self.global_terms

    def __call__(self, pred, target):
        loss, _ = _GraphedLossFn.apply(pred, target, self)
        return loss
