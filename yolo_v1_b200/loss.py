"""Host-side mirror of the reference's loss interface (/root/reference/v1Loss.py:9-118) on top of the
C ABI (include/yolo1_b200.h).  PyTorch supplies device memory, streams and autograd plumbing; the
arithmetic runs in libyolo1_b200.so (yolo_v1_b200/csrc/loss.cu).

`YOLOLossV1` keeps the reference constructor and forward signature, so `train.py:101,167` work as they
are: `lossLayer = YOLOLossV1(batch_size, S, B, clsN, lambda_coord, lambda_noobj, _logger=..., _vis=...)`,
`loss = lossLayer(pred, target); loss.backward()`.
"""
import contextlib
import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import host as _host

__all__ = ["YOLOLossV1", "yolo_loss_fused", "yolo_loss_from_objects", "scale_grad_"]

_COORD_MODES = {"reference": _lib.COORD_REFERENCE, "paper": _lib.COORD_PAPER}
TERM_NAMES = ("location", "contain", "not_contain", "classify", "total")
_WS_BYTES = None


def _dtype_code(t):
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    if t.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    raise TypeError("pred must be float32 or bfloat16, got %s" % t.dtype)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def yolo_loss_fused(pred, target, batch_size, S=None, B=2, C=20, l_coord=5.0, l_noobj=0.5,
                    coord_mode="reference", want_grad=True, variant=0, out_grad=None, out_terms=None,
                    workspace=None, from_logits=False):
    """One fused pass: loss terms AND d total / d pred (v1Loss.py:22-118 + its autograd backward).

    pred  : CUDA tensor [N,S,S,5B+C], float32 or bfloat16, ANY strides (the backbone's permuted NCHW view,
            backbones/OriginResNet.py:189, is read in place -- no .contiguous()).
    target: CUDA float32 tensor of the same shape.
    Returns (loss, grad, terms): loss = terms[4] (0-dim view), grad laid out like pred (or None),
    terms = float32[5] on the device: location, contain, not_contain, classify (each / batch_size, the four
    numbers v1Loss.py:108 logs) and the total.  Stream ordered on the current stream; no host sync.
    from_logits=True: `pred` holds the head's PRE-sigmoid outputs (OriginResNet.py:186-188); the sigmoid is applied
    inside the kernel and `grad` is d total / d logit (yolo1_loss_fwd_bwd_logits).
    """
    if pred.dim() != 4 or pred.shape != target.shape:
        raise ValueError("pred and target must both be [N,S,S,5B+C]; got %s and %s" %
                         (tuple(pred.shape), tuple(target.shape)))
    if not pred.is_cuda:
        raise RuntimeError("yolo_loss_fused needs CUDA tensors (host tensors go through "
                           "yolo_v1_b200.host.HostContext / YOLOLossV1, which stage them to the GPU)")
    N, S_, S2, D = pred.shape
    if S is None:
        S = S_
    if S_ != S or S2 != S or D != 5 * B + C:
        raise ValueError("shape %s does not match S=%d B=%d C=%d" % (tuple(pred.shape), S, B, C))
    dev = pred.device
    if target.device != dev:
        target = target.to(dev, non_blocking=True)
    if target.dtype != torch.float32:
        target = target.float()
    L = _lib.lib()
    # (a small call is shorter than this function: the device guard is entered only when the tensors live on another
    #  device than the current one, and the constant workspace size is asked for once)
    guard = torch.cuda.device(dev) if torch.cuda.current_device() != dev.index else contextlib.nullcontext()
    with guard:
        grad = None
        if want_grad:
            grad = out_grad if out_grad is not None else torch.empty_like(pred)
            if grad.shape != pred.shape or grad.dtype != pred.dtype or grad.device != dev:
                raise ValueError("out_grad must match pred in shape, dtype and device")
        terms = out_terms if out_terms is not None else torch.empty(5, dtype=torch.float32, device=dev)
        global _WS_BYTES
        if _WS_BYTES is None:
            _WS_BYTES = int(L.yolo1_loss_workspace_bytes(N, S, B, C))    # independent of the call (include/yolo1_b200.h)
        ws_bytes = _WS_BYTES
        ws = workspace if workspace is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if ws.numel() * ws.element_size() < ws_bytes:
            raise ValueError("workspace too small: need %d bytes" % ws_bytes)
        gstr = _lib.strides4(grad) if grad is not None else None
        common = (pred.data_ptr(), _lib.strides4(pred), _dtype_code(pred),
                  target.data_ptr(), _lib.strides4(target),
                  grad.data_ptr() if grad is not None else None, gstr,
                  terms.data_ptr(), N, S, B, C, float(l_coord), float(l_noobj), 1.0 / float(batch_size),
                  _COORD_MODES[coord_mode], ws.data_ptr(), ws.numel() * ws.element_size())
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if from_logits:
            rc = L.yolo1_loss_fwd_bwd_logits(*common, stream)
        else:
            rc = L.yolo1_loss_fwd_bwd_ex(*common, int(variant), stream)
        if rc != 0:
            _lib.check(rc, "yolo1_loss_fwd_bwd")
    return terms[4], grad, terms


def yolo_loss_from_objects(pred, boxes, labels, offsets, batch_size, S=None, B=2, C=20, l_coord=5.0, l_noobj=0.5,
                           coord_mode="reference", want_grad=True, from_logits=False, out_grad=None, check=False,
                           variant=0, workspace=None):
    """The fused loss fed with object lists instead of a dense target (yolo1_loss_fwd_bwd_objects):
    identical to `yolo_loss_fused(pred, encode_targets(boxes, labels, offsets), ...)` without ever writing or
    reading the [N,S,S,5B+C] target -- boxes [n,4] (cx,cy,w,h), labels [n], offsets [N+1] CUDA tensors in the CSR
    form of `pack_objects`.  Returns (loss, grad, terms).  check=True costs one host sync and raises IndexError
    when a box centre / label lies outside the grid / class range (as the reference encoder would)."""
    if not pred.is_cuda:
        raise RuntimeError("yolo_loss_from_objects needs CUDA tensors")
    N, S_, S2, D = pred.shape
    if S is None:
        S = S_
    if S_ != S or S2 != S or D != 5 * B + C or offsets.shape[0] != N + 1:
        raise ValueError("shapes do not match: pred %s, offsets %s, S=%d B=%d C=%d" %
                         (tuple(pred.shape), tuple(offsets.shape), S, B, C))
    dev = pred.device
    boxes = boxes.to(device=dev, dtype=torch.float32).contiguous().reshape(-1, 4)
    labels = labels.to(device=dev, dtype=torch.int32).contiguous().reshape(-1)
    offsets = offsets.to(device=dev, dtype=torch.int64).contiguous()
    L = _lib.lib()
    with torch.cuda.device(dev):
        grad = None
        if want_grad:
            grad = out_grad if out_grad is not None else torch.empty_like(pred)
        terms = torch.empty(5, dtype=torch.float32, device=dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        ws_bytes = int(L.yolo1_loss_objects_workspace_bytes(N, S, B, C))
        ws = workspace if workspace is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if ws.numel() * ws.element_size() < ws_bytes:
            raise ValueError("workspace too small: need %d bytes" % ws_bytes)
        if grad is not None and (grad.shape != pred.shape or grad.dtype != pred.dtype or grad.device != dev):
            raise ValueError("out_grad must match pred in shape, dtype and device")
        rc = L.yolo1_loss_fwd_bwd_objects_ex(
            pred.data_ptr(), _lib.strides4(pred), _dtype_code(pred), int(bool(from_logits)),
            boxes.data_ptr() if boxes.numel() else None, labels.data_ptr() if labels.numel() else None,
            offsets.data_ptr(), grad.data_ptr() if grad is not None else None,
            _lib.strides4(grad) if grad is not None else None, terms.data_ptr(), N, S, B, C,
            float(l_coord), float(l_noobj), 1.0 / float(batch_size), _COORD_MODES[coord_mode],
            ws.data_ptr(), ws.numel() * ws.element_size(), status.data_ptr(), int(variant), _stream_ptr(dev))
        _lib.check(rc, "yolo1_loss_fwd_bwd_objects")
    if check and int(status.item()) != 0:
        raise IndexError("yolo_loss_from_objects: a box centre or label lies outside the grid / class range")
    return terms[4], grad, terms


def scale_grad_(grad, grad_out):
    """grad *= grad_out (a 0-dim device tensor) in place; a no-op launch when grad_out == 1."""
    L = _lib.lib()
    # the dense storage behind `grad` (any stride order): scale it linearly
    flat = grad.as_strided((grad.numel(),), (1,)) if grad.is_contiguous() or _is_dense(grad) else None
    if flat is None:
        grad.mul_(grad_out.to(grad.dtype))
        return grad
    go = grad_out.detach().to(device=grad.device, dtype=torch.float32).reshape(1)
    with torch.cuda.device(grad.device):
        rc = L.yolo1_scale_grad(flat.data_ptr(), _dtype_code(grad), flat.numel(), go.data_ptr(),
                                _stream_ptr(grad.device))
        _lib.check(rc, "yolo1_scale_grad")
    return grad


def _is_dense(t):
    """True when t's elements tile one contiguous block in some dimension order."""
    dims = sorted(range(t.dim()), key=lambda d: t.stride(d))
    expect = 1
    for d in dims:
        if t.shape[d] == 1:
            continue
        if t.stride(d) != expect:
            return False
        expect *= t.shape[d]
    return True


_TWICE = ("the fused YOLO loss computed d loss / d pred in forward() and handed that buffer to autograd in the first "
          "backward(); it cannot be back-propagated a second time.  Construct YOLOLossV1(..., retain_graph=True) to "
          "keep the buffer (each backward() then returns a fresh, scaled copy), as the reference's autograd graph "
          "would under loss.backward(retain_graph=True).")


def _hand_over(ctx, grad_loss):
    """backward() of the fused functions.  Default: the stashed gradient is scaled in place (a no-op launch when
    grad_output == 1, the `loss.backward()` case) and handed to autograd -- zero copies, once.  A second backward()
    over the same graph raises, as PyTorch does for freed saved tensors.  `retain` mode keeps the stash untouched
    and returns `stash * grad_output` in a fresh tensor every time (correct for retain_graph=True and for repeated
    torch.autograd.grad with different grad_outputs; costs one extra pass over the gradient)."""
    if not ctx.had_grad:
        return None
    grad = ctx.grad
    if grad is None:
        raise RuntimeError(_TWICE)
    if ctx.retain:
        return grad * grad_loss.to(device=grad.device, dtype=grad.dtype)
    ctx.grad = None
    if grad.is_cuda:
        return scale_grad_(grad, grad_loss)
    # host tensors: the scalar is already on the host, so the no-op test for grad_output == 1 is free
    if float(grad_loss) != 1.0:
        grad.mul_(float(grad_loss))
    return grad


class _FusedYoloLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, cfg, retain):
        need = pred.requires_grad
        loss, grad, terms = yolo_loss_fused(pred.detach(), target, want_grad=need, **cfg)
        ctx.grad, ctx.had_grad, ctx.retain = grad, need, retain
        ctx.mark_non_differentiable(terms)
        return loss.clone(), terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        return _hand_over(ctx, grad_loss), None, None, None


class _FusedYoloLossObjects(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, boxes, labels, offsets, cfg, retain):
        need = pred.requires_grad
        loss, grad, terms = yolo_loss_from_objects(pred.detach(), boxes, labels, offsets, want_grad=need, **cfg)
        ctx.grad, ctx.had_grad, ctx.retain = grad, need, retain
        ctx.mark_non_differentiable(terms)
        return loss.clone(), terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        return _hand_over(ctx, grad_loss), None, None, None, None, None


class YOLOLossV1(nn.Module):
    """Drop-in for the reference `YOLOLossV1` (v1Loss.py:9-118): same constructor, same forward.

    Differences, all opt-in or silent improvements:
      * the loss AND its gradient are computed in one CUDA pass (`loss.backward()` only scales);
      * nothing is printed per call unless `_logger`/`_vis` are given or `verbose=True`
        (the reference prints four numbers on every call, v1Loss.py:110, forcing host syncs);
      * keyword-only extras: `coord_mode` ('reference' = the row-slice behaviour of v1Loss.py:101,
        'paper' = xy plain / wh sqrt), `verbose`, `from_logits` (feed the head's pre-sigmoid output; the
        sigmoid of OriginResNet.py:188 and its backward are fused into the loss kernel), `retain_graph`
        (allow several backward passes over one forward; see `_hand_over`).
    `_device` is accepted for signature compatibility; tensors are used where they live.  Host (CPU)
    tensors are staged through the pipelined host-buffer path of the library (no CPU arithmetic).
    The module has no parameters or buffers (state_dict() of an enclosing model is unchanged).
    """

    def __init__(self, _batch_size, _S, _B, _clsN, _l_coord=5., _l_noobj=0.5, _device='cuda:0', _logger=None,
                 _vis=None, *, coord_mode="reference", verbose=False, from_logits=False, retain_graph=False):
        super().__init__()
        if coord_mode not in _COORD_MODES:
            raise ValueError("coord_mode must be 'reference' or 'paper'")
        self.S = _S
        self.B = _B
        self.device = _device
        self.C = _clsN
        self.lambda_coord = _l_coord
        self.lambda_noobj = _l_noobj
        self.batch_size = _batch_size
        self.logger = _logger
        self.vis = _vis
        self.coord_mode = coord_mode
        self.verbose = verbose
        self.from_logits = from_logits   # inputs are pre-sigmoid head outputs: sigmoid fused into the kernel
        # retain_graph=True: backward() may run several times over one forward() (loss.backward(retain_graph=True),
        # repeated torch.autograd.grad): each call returns stash * grad_output in a fresh tensor.  Default: the stash
        # is handed over once with no copy and a second backward() raises.
        self.retain_graph = retain_graph
        self.last_terms = None   # device float32[5] of the latest call (read lazily: no sync unless asked)
        self._host_ctx = None

    def _cfg(self):
        return dict(batch_size=self.batch_size, S=self.S, B=self.B, C=self.C, l_coord=self.lambda_coord,
                    l_noobj=self.lambda_noobj, coord_mode=self.coord_mode, from_logits=self.from_logits)

    def forward(self, pred_tensor, target_tensor):
        if pred_tensor.is_cuda:
            loss, terms = _FusedYoloLoss.apply(pred_tensor, target_tensor, self._cfg(), self.retain_graph)
        else:
            loss, terms = self._forward_host(pred_tensor, target_tensor)
        self.last_terms = terms
        if self.logger or self.vis or self.verbose:
            self._report(terms)
        return loss

    def forward_objects(self, pred_tensor, boxes, labels, offsets):
        """forward(pred, encoder(boxes, labels)) without the dense target: the ragged object lists (CSR, see
        `yolo_v1_b200.pack_objects`) go straight into the kernel.  Additive API; CUDA tensors only."""
        cfg = self._cfg()
        loss, terms = _FusedYoloLossObjects.apply(pred_tensor, boxes, labels, offsets, cfg, self.retain_graph)
        self.last_terms = terms
        if self.logger or self.vis or self.verbose:
            self._report(terms)
        return loss

    def _forward_host(self, pred, target):
        if self.from_logits:
            raise RuntimeError("from_logits=True needs CUDA tensors")
        if self._host_ctx is None:
            self._host_ctx = _host.HostContext(self.S, self.B, self.C)
        return _HostYoloLoss.apply(pred, target, self._host_ctx, self._cfg(), self.retain_graph)

    def _report(self, terms):
        t = terms.detach().float().cpu().tolist()   # one sync, only when someone listens
        if self.logger:
            self.logger.info('location loss : %.5f contain loss : %.5f not contain loss: %.5f classify loss : %.5f'
                             % (t[0], t[1], t[2], t[3]))   # v1Loss.py:108
        elif self.verbose:
            print('location loss : %.5f' % t[0], 'contain loss : %.5f' % t[1], 'not contain loss: %.5f' % t[2],
                  'classify loss : %.5f' % t[3])           # v1Loss.py:110
        if self.vis:                                       # v1Loss.py:113-116
            self.vis.plot('location loss', t[0])
            self.vis.plot('confidence loss', t[1])
            self.vis.plot('no object loss', t[2])
            self.vis.plot('classify loss', t[3])


class _HostYoloLoss(torch.autograd.Function):
    """CPU tensors in, CPU tensors out; the arithmetic still runs on the GPU (pipelined H2D/kernel/D2H)."""

    @staticmethod
    def forward(ctx, pred, target, hctx, cfg, retain):
        need = pred.requires_grad
        terms, grad = hctx.loss(pred.detach(), target, batch_size=cfg["batch_size"], l_coord=cfg["l_coord"],
                                l_noobj=cfg["l_noobj"], coord_mode=cfg["coord_mode"], want_grad=need)
        ctx.grad, ctx.had_grad, ctx.retain = grad, need, retain
        ctx.mark_non_differentiable(terms)
        return terms[4].clone(), terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        return _hand_over(ctx, grad_loss), None, None, None, None
