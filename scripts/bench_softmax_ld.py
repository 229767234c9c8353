"""Times the padded-row policy softmax kernels of the training attention at the training shape (B=256, H=6, T=197)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, d2s
lib, ops = d2s.pkg._lib, d2s.pkg.ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for B, H, T in ((256, 6, 197), (256, 6, 138), (256, 6, 97)):
    Tp = (T + 7) // 8 * 8
    S = torch.randn(B * H, Tp, Tp, device="cuda").bfloat16()
    G = torch.randn(B * H, Tp, Tp, device="cuda").bfloat16()
    P = torch.empty_like(S); dS = torch.empty_like(S)
    pol = (torch.rand(B, T, device="cuda") > 0.3).float()
    stats = torch.empty(B, H, T, 2, device="cuda")
    gp = torch.zeros(B, T, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.call("d2s_softmax_policy_fwd_ld", S.data_ptr(), pol.data_ptr(), B, H, T, Tp, Tp, 1e-6, P.data_ptr(), stats.data_ptr(), st)
    b = lambda: lib.call("d2s_softmax_policy_bwd_ld", S.data_ptr(), pol.data_ptr(), G.data_ptr(), stats.data_ptr(), B, H, T, Tp, Tp, 1e-6, dS.data_ptr(), gp.data_ptr(), st)
    tf, tb = t(f), t(b)
    by = B * H * T * Tp * 2
    print(f"B={B} H={H} T={T}: fwd {tf:.1f} us ({2 * by / tf / 1e3:.0f} GB/s) | bwd {tb:.1f} us ({3 * by / tb / 1e3:.0f} GB/s)")
B, H, T = 256, 6, 197
Tp = 200
S = torch.randn(B * H, Tp, Tp, device="cuda").bfloat16(); G = torch.randn(B * H, Tp, Tp, device="cuda").bfloat16(); dS = torch.empty_like(S)
pol = (torch.rand(B, T, device="cuda") > 0.3).float(); stats = torch.rand(B, H, T, 2, device="cuda") + 1; gp = torch.zeros(B, T, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("bwd, no gpolicy:", t(lambda: lib.call("d2s_softmax_policy_bwd_ld", S.data_ptr(), pol.data_ptr(), G.data_ptr(), stats.data_ptr(), B, H, T, Tp, Tp, 1e-6, dS.data_ptr(), None, st)))
print("bwd, no policy :", t(lambda: lib.call("d2s_softmax_policy_bwd_ld", S.data_ptr(), None, G.data_ptr(), stats.data_ptr(), B, H, T, Tp, Tp, 1e-6, dS.data_ptr(), None, st)))
print("bwd, in place  :", t(lambda: lib.call("d2s_softmax_policy_bwd_ld", S.data_ptr(), pol.data_ptr(), G.data_ptr(), stats.data_ptr(), B, H, T, Tp, Tp, 1e-6, G.data_ptr(), gp.data_ptr(), st)))
