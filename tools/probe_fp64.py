"""FP64 rates on this GPU: DFMA throughput at full occupancy (mode 3) and the DMUL / DADD issue interval seen by ONE warp per
SM sub-partition (modes 4-6) -- the situation of the env phase of the fused rollout kernel."""
import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch, msacl_b200
from msacl_b200 import _lib
lib = msacl_b200.load_library()
sink = torch.rand(128, device='cuda')
clock_ghz = torch.cuda.clock_rate() / 1e6 if hasattr(torch.cuda, "clock_rate") else None
for mode, iters in ((0, 20000), (3, 2000), (4, 4000), (5, 4000), (6, 4000), (7, 4000)):
    fl = C.c_double(0)
    for _ in range(2): _lib.check(lib.msacl_ffma_probe(mode, iters, sink.data_ptr(), C.byref(fl), _lib.current_stream()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _lib.check(lib.msacl_ffma_probe(mode, iters, sink.data_ptr(), C.byref(fl), _lib.current_stream())); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if mode < 4:
        print('mode', mode, 'TFLOP/s', fl.value / (ms * 1e-3) / 1e12)
    elif mode == 7:
        print('mode 7: ns per float64 -> float32 -> float64 round trip + DMUL %.2f (cycles at 1.9 GHz: %.1f)' % (ms * 1e6 / fl.value, ms * 1e6 / fl.value * 1.9))
    else:
        print('mode', mode, 'chains', {4: 8, 5: 2, 6: 1}[mode], 'ns per FP64 warp-instruction %.2f' % (ms * 1e6 / fl.value),
              '(cycles at 1.9 GHz: %.1f)' % (ms * 1e6 / fl.value * 1.9))
