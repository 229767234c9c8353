"""Times the d2s tcgen05 GEMMs of Block.forward against cuBLAS (+ the separate d2s elementwise kernels) at the bench shapes."""
import sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
D = 384
for T in (197, 138, 97, 68):
    M = 1024 * T
    x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16()
    fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16()
    proj = torch.nn.Linear(D, D).cuda().bfloat16()
    ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
    u = torch.randn(M, 4 * D, device="cuda", dtype=torch.bfloat16)
    def ref():
        v = fc1(x); ops.bias_act_(v, None, ops.ACT_GELU); return v
    fl = 2.0 * M * 4 * D * D
    a = t(ref); c = t(lambda: fc1(x))
    b1 = t(lambda: ops.linear_act(x, fc1.weight, fc1.bias, ops.ACT_GELU))
    b2 = t(lambda: ops.linear_act(x, fc1.weight, fc1.bias, ops.ACT_GELU))
    b3 = t(lambda: ops.linear_act(x, fc1.weight, fc1.bias, ops.ACT_NONE))
    print(f"T={T} fc1: cuBLAS+GELU {a:.1f} us | cuBLAS alone {c:.1f} us ({fl / c / 1e6:.0f} TF/s) | 1-CTA fused {b1:.1f} us ({fl / b1 / 1e6:.0f} TF/s)"
          f" | pair fused {b2:.1f} us ({fl / b2 / 1e6:.0f} TF/s) | pair no-act {b3:.1f} us")
    x3 = x.view(1024, T, D)
    def ref2():
        y = fc2(u); return ops.add_layernorm(x3, y.view(1024, T, D), ln.weight, ln.bias, 1e-6)
    a = t(ref2); c = t(lambda: fc2(u))
    b = t(lambda: ops.linear_residual_ln(u, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6))
    b0 = t(lambda: ops.linear_residual_ln(u, fc2.weight, fc2.bias, x, want_norm=False))
    print(f"T={T} fc2+add+LN: cuBLAS+addLN {a:.1f} us (cuBLAS alone {c:.1f} us, {fl / c / 1e6:.0f} TF/s) | fused {b:.1f} us ({fl / b / 1e6:.0f} TF/s) | fused, no LN {b0:.1f} us")
    def ref3():
        y = proj(x); return ops.add_layernorm(x3, y.view(1024, T, D), ln.weight, ln.bias, 1e-6)
    fl = 2.0 * M * D * D
    a = t(ref3); c = t(lambda: proj(x))
    b = t(lambda: ops.linear_residual_ln(x, proj.weight, proj.bias, x, ln.weight, ln.bias, 1e-6))
    bytes_ = M * D * 2 * 4
    print(f"T={T} proj+add+LN: cuBLAS+addLN {a:.1f} us (cuBLAS alone {c:.1f} us) | fused {b:.1f} us ({bytes_ / b / 1e3:.0f} GB/s algorithmic)")
