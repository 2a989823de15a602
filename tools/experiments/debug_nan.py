import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import cp360_b200
from cp360_b200 import _lib
dev = torch.device("cuda", 0)
w, C, B = 24, 16, 2
torch.manual_seed(w * 1000 + C)
c2e = cp360_b200.Cube2Equi(w)
x = torch.randn((6 * B, C, w, w), device=dev)
x[0, C // 2, w // 2, w // 2] = float("nan")
full = c2e.to_equi_nn(x)
want_v, want_i = torch.max(full, 1)
sal = c2e.to_equi_max(x)
P = 8 * w * w
sal2 = torch.empty((B, 2 * w, 4 * w), device=dev)
arg = torch.empty((B, 2 * w, 4 * w), dtype=torch.int32, device=dev)
scratch = torch.zeros((B, 2 * w, 4 * w), dtype=torch.int64, device=dev)
taps, wts = c2e._plan_on(dev)
st = torch.cuda.current_stream().cuda_stream
_lib.check(_lib.lib().cp360_c2e_max_arg_fwd(x.data_ptr(), taps.data_ptr(), wts.data_ptr(), sal2.data_ptr(), arg.data_ptr(), scratch.data_ptr(), B, C, w, st))
torch.cuda.synchronize()
idx = torch.isnan(want_v).nonzero()
print("want NaN at", idx[:4].tolist(), "count", len(idx))
for i in idx[:4].tolist():
    b, y, xx = i
    k = int(scratch[b, y, xx].item()) & 0xffffffffffffffff
    print("pix", i, "full[:,ch]=", full[b, :, y, xx].tolist()[:10], "sal", sal[b, y, xx].item(), "sal2", sal2[b, y, xx].item(), "arg", arg[b, y, xx].item(), "key hi %08x lo %08x" % (k >> 32, k & 0xffffffff), "want_i", want_i[b, y, xx].item())
