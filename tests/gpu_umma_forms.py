"""GPU probe of the two tcgen05 operand forms the next kernels build on: MN-major operands read from the activation layout, and the
TS form (A operand in TMEM).  Usage: gpu_umma_forms.py  (prints max errors; used to pin tests/test_gpu_parity.py::test_umma_operand_forms)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from playableenvironments_b200 import _cabi  # noqa: E402

L = _cabi.lib()
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(5)
for mode in sys.argv[1:] or ["2", "1a", "1b"]:
    for n, k in [(128, 128), (256, 64), (64, 16)]:
        if mode == "2":
            a, b = torch.randn(128, k, generator=g).cuda(), torch.randn(n, k, generator=g).cuda()
            ref = a.half().float() @ b.half().float().t()
            lbo = sbo = 0
            m = 2
        else:
            a, b = torch.randn(k, 128, generator=g).cuda(), torch.randn(k, n, generator=g).cuda()
            ref = a.half().float().t() @ b.half().float()
            lbo, sbo = (128, 2048) if mode == "1a" else (2048, 128)
            m = 1
        d = torch.full((128, n), float("nan"), device="cuda")
        _cabi.check(L.pe_debug_umma_gemm2(m, a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, lbo, sbo, stream))
        torch.cuda.synchronize()
        print(f"mode {mode} n={n} k={k} lbo={lbo} sbo={sbo}: max err / scale = {float((d - ref).abs().max() / ref.abs().max()):.3e}", flush=True)
