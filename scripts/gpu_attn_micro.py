"""Micro-benchmark of the tensor-core attention core on the fv0 level-2 shape (for ncu / timing)."""
import sys, time
import torch
sys.path.insert(0, ".")
from fastervit_b200 import lib
lib.load()
S, heads, hd, groups = 53, 8, 32, 1024
qkv = torch.randn(groups * S, 3 * heads * hd, device="cuda").half()
bias = torch.randn(heads, S, S, device="cuda") + 8
out = torch.zeros(groups * S, heads * hd, device="cuda", dtype=torch.half)
def run():
    lib.call("fvit_attn_tc_fwd", qkv.data_ptr(), qkv.stride(0), groups, S, heads, hd, bias.data_ptr(), hd ** -0.5,
             out.data_ptr(), out.stride(0))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print("attn_tc fv0-L2 shape: %.1f us per launch" % (e0.elapsed_time(e1) * 100))
