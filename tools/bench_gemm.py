"""msacl_gemm_tc throughput for the learner's layer shapes: algorithmic TFLOP/s (2 m n k / t) and the fraction of the measured
bf16 tensor peak incl. the 6 (or 3) bf16 products issued per algorithmic product."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msacl_b200
from msacl_b200 import _lib
from msacl_b200.learner import _desc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 1400.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    pass
lib = msacl_b200.load_library()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(rows, 256, device="cuda", generator=g)
W = torch.randn(256, 256, device="cuda", generator=g) / 16
b = torch.randn(256, device="cuda", generator=g)
Y = torch.empty(rows, 256, device="cuda")
S = 74
parts = torch.empty(S, 256, 256, device="cuda")
cases = {
    "forward  [rows,256]x[256,256]^T +bias relu": lambda p: _desc(a=X, a_rs=256, a_ks=1, b=W, b_rs=256, b_ks=1, m=rows, n=256, k=256, c=Y, ldc=256, bias=b, act=1, precision=p),
    "dgrad    [rows,256]x[256,256]   *relu'": lambda p: _desc(a=X, a_rs=256, a_ks=1, b=W, b_rs=1, b_ks=256, m=rows, n=256, k=256, c=Y, ldc=256, mask=X, mask_ld=256, mask_act=1, precision=p),
    "wgrad    [256,rows]x[rows,256]  split-K 74": lambda p: _desc(a=X, a_rs=1, a_ks=256, b=X, b_rs=1, b_ks=256, m=256, n=256, k=rows, c=parts, ldc=256, split_k=S, c_split_stride=65536, precision=p),
}
packed_cases = {k + "  [weights pre-packed]": v for k, v in cases.items() if "wgrad" not in k}
for name, mk in list(cases.items()) + list(packed_cases.items()):
    for prec in (6, 3):
        d = mk(prec)
        if "pre-packed" in name:
            if (rows + 127) // 128 < 148:
                continue
            pk = torch.empty(int(lib.msacl_gemm_packed_b_bytes(256, prec)), dtype=torch.uint8, device="cuda")
            _lib.check(lib.msacl_gemm_pack_b(C.byref(d), pk.data_ptr(), _lib.current_stream()))
            d.b_packed = pk.data_ptr()
        for _ in range(3):
            _lib.check(lib.msacl_gemm_tc(C.byref(d), _lib.current_stream()))
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        reps = 10
        for _ in range(reps):
            _lib.check(lib.msacl_gemm_tc(C.byref(d), _lib.current_stream()))
        e.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(e) / reps
        tf = 2.0 * rows * 256 * 256 / (ms * 1e-3) / 1e12
        print(json.dumps({"kernel": "gemm_tc", "case": name, "rows": rows, "precision": prec, "ms": round(ms, 4), "algorithmic_tflops": round(tf, 1),
                          "tensor_tflops_incl_split": round(tf * prec, 1), "bf16_peak_sustained": peak, "frac_incl_split": round(tf * prec / peak, 3),
                          "frac_algorithmic": round(tf / peak, 3)}), flush=True)
# cuBLAS fp32 (what the torch engine runs) for context
for _ in range(3):
    torch.mm(X, W.t(), out=Y)
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10):
    torch.mm(X, W.t(), out=Y)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
print(json.dumps({"kernel": "cuBLAS fp32 torch.mm (context)", "rows": rows, "ms": round(ms, 4), "algorithmic_tflops": round(2.0 * rows * 65536 / (ms * 1e-3) / 1e12, 1)}))
