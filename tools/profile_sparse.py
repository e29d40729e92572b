"""One launch each of the dense streaming kernel (variant 0) and the sector-read kernels (variants 40, 42, 100) at
config-3 size, for ncu: DRAM bytes and sectors per request show whether the memory system fetches less when the
kernel asks for less.
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,\
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum \
        --clock-control none -k regex:loss_ python tools/profile_sparse.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth

N, S = 65536, 14
pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
grad = torch.empty_like(pred)
terms = torch.empty(5, device="cuda")
ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
for v in (0, 40, 42, 100):
    for want_grad in (True, False):
        y.yolo_loss_fused(pred, target, batch_size=N, variant=v, want_grad=want_grad,
                          out_grad=grad if want_grad else None, out_terms=terms, workspace=ws)
torch.cuda.synchronize()
print("profile_sparse ok")
