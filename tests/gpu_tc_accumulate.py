"""Diagnostic: how does tcgen05.mma (kind::f16, fp32 accumulate) round its accumulation?  D = A B^T with fp16-exact operands (every
product is exact in fp32), compared with the exactly rounded result: error size, its growth with K and its BIAS (a truncating adder
leaves a systematic error that does not average out)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from playableenvironments_b200 import _cabi

L = _cabi.lib()
stream = torch.cuda.current_stream().cuda_stream
out = {}
for positive in (False, True):
    for k in (16, 64, 128, 256):
        g = torch.Generator().manual_seed(k + positive)
        a = torch.randn(128, k, generator=g).half().float()
        b = torch.randn(256, k, generator=g).half().float()
        if positive:
            a, b = a.abs(), b.abs()
        d = torch.empty(128, 256, device="cuda")
        a_d, b_d = a.cuda(), b.cuda()
        _cabi.check(L.pe_debug_umma_gemm(a_d.data_ptr(), b_d.data_ptr(), None, d.data_ptr(), 256, k, stream))
        torch.cuda.synchronize()
        exact = a.double() @ b.double().t()
        err = (d.cpu().double() - exact)
        ulp = torch.abs(exact).clamp_min(1e-30).log2().floor().exp2() * 2.0 ** -23
        rn = exact.float().double() - exact                      # what one correctly rounded result would leave
        sgemm = (a_d @ b_d.t()).cpu().double() - exact  # cuBLAS fp32 (FFMA chain) on the same data
        out[f"{'positive' if positive else 'signed'} K={k}"] = {
            "rms_err_ulp": float((err / ulp).pow(2).mean().sqrt()), "mean_err_ulp": float((err / ulp).mean()),
            "mean_signed_by_result_ulp": float((err / ulp * torch.sign(exact)).mean()), "max_err_ulp": float((err / ulp).abs().max()),
            "rn_rms_ulp": float((rn / ulp).pow(2).mean().sqrt()), "sgemm_rms_ulp": float((sgemm / ulp).pow(2).mean().sqrt()),
            "sgemm_mean_ulp": float((sgemm / ulp).mean())}
print(json.dumps(out, indent=1))
