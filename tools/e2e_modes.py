"""Development aid: the host-buffer loss call (yolo1_loss_fwd_bwd_host) in every transfer mode, with and without the
gradient, config-3 size.  python tools/e2e_modes.py"""
import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_v1_b200 as y
from yolo_v1_b200 import synth
N, S = 65536, 14
pred, target = synth.make_loss_inputs(N, S, seed=1, device="cuda")
hp = torch.empty(pred.shape, pin_memory=True); ht = torch.empty(pred.shape, pin_memory=True); hg = torch.empty(pred.shape, pin_memory=True)
hp.copy_(pred); ht.copy_(target); torch.cuda.synchronize()
for chunk in (0, 8192):
    ctx = y.HostContext(S, chunk_images=chunk)
    for mode in (0, 1, 2, 3, 4):
        for grad in (True, False):
            ctx.set_zero_copy(mode)
            for _ in range(2): ctx.loss(hp, ht, batch_size=N, out_grad=hg if grad else None, want_grad=grad)
            t0 = time.perf_counter()
            for _ in range(5): ctx.loss(hp, ht, batch_size=N, out_grad=hg if grad else None, want_grad=grad)
            ms = (time.perf_counter() - t0) / 5 * 1e3
            print("chunk %5d mode %d grad=%d: %.1f ms  %.0f Mcells/s" % (chunk, mode, grad, ms, N * S * S / ms / 1e3), flush=True)
    ctx.close()
