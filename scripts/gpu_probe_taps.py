"""GPU probe: isolate which tap configuration of fvit_gemm misbehaves. Each case runs in its own
process with a timeout so a device-side trap or hang cannot take the others down."""
import subprocess
import sys

CASES = {
    "one_tap_pos": "taps=[(5,0)]",
    "one_tap_neg": "taps=[(-5,0)]",
    "two_taps_zero": "taps=[(0,0),(0,0)]",
    "nine_taps_zero": "taps=[(0,0)]*9",
    "nine_taps_real": "taps=[((dy-1)*18+(dx-1),0) for dy in range(3) for dx in range(3)]",
}

BODY = r'''
import torch, sys
sys.path.insert(0, ".")
from fastervit_b200 import lib
lib.load()
torch.manual_seed(0)
rows, cin, cout = 864, 64, 64
{taps}
nt = len(taps)
a = torch.randn(rows, cin, device="cuda").half()
w = (torch.randn(cout, nt, cin, device="cuda") * 0.05).half()
out = torch.zeros(rows, cout, device="cuda")
lib.gemm(a, w.view(cout, nt * cin), kc=cin, taps=taps, out_f32=out)
torch.cuda.synchronize()
ref = torch.zeros(rows, cout, device="cuda")
af = a.float()
for t, (sh, _) in enumerate(taps):
    sa = torch.zeros_like(af)
    if sh >= 0:
        sa[: rows - sh] = af[sh:]
    else:
        sa[-sh:] = af[: rows + sh]
    ref += sa @ w[:, t].float().t()
err = ((out - ref).abs().max() / ref.abs().max()).item()
print("rel err", err)
'''

for name, taps in CASES.items():
    try:
        r = subprocess.run([sys.executable, "-c", BODY.format(taps=taps)], capture_output=True, text=True,
                           timeout=60)
        print(name, "rc", r.returncode, r.stdout.strip()[-200:], r.stderr.strip()[-300:].replace("\n", " | "))
    except subprocess.TimeoutExpired:
        print(name, "TIMEOUT")
    sys.stdout.flush()
