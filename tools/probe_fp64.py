import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch, msacl_b200
from msacl_b200 import _lib
lib = msacl_b200.load_library()
sink = torch.rand(128, device='cuda')
for mode, iters in ((0, 20000), (3, 2000)):
    fl = C.c_double(0)
    for _ in range(2): _lib.check(lib.msacl_ffma_probe(mode, iters, sink.data_ptr(), C.byref(fl), _lib.current_stream()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _lib.check(lib.msacl_ffma_probe(mode, iters, sink.data_ptr(), C.byref(fl), _lib.current_stream())); b.record(); torch.cuda.synchronize()
    print('mode', mode, 'TFLOP/s', fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
