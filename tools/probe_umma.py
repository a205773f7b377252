"""Cycles per 128x256x16 bf16 UMMA issued back to back from shared memory (see msacl_umma_probe)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import msacl_b200
from msacl_b200 import _lib
lib = _lib.load()
out = torch.zeros(4, dtype=torch.float64, device="cuda")
for mode in (0, 1, 2, 3, 4):
    for iters in (64, 512):
        out.zero_()
        _lib.check(lib.msacl_umma_probe(mode, iters, out.data_ptr(), _lib.current_stream()))
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        print("mode", mode, "iters", iters, "cycles/UMMA %.1f" % o[0], ("tcgen05.ld rate %.1f B/cycle" % o[1]) if mode == 2 else "", flush=True)
