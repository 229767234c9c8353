"""One forward + one backward launch of the padded-row policy softmax kernels at the training shape, for ncu."""
import sys; sys.path.insert(0, "/root/repo")
import torch, d2s
lib = d2s.pkg._lib
B, H, T = 256, 6, 197
Tp = (T + 7) // 8 * 8
S = torch.randn(B * H, Tp, Tp, device="cuda").bfloat16(); G = torch.randn(B * H, Tp, Tp, device="cuda").bfloat16()
P = torch.empty_like(S); dS = torch.empty_like(S)
pol = (torch.rand(B, T, device="cuda") > 0.3).float()
stats = torch.empty(B, H, T, 2, device="cuda"); gp = torch.zeros(B, T, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    lib.call("d2s_softmax_policy_fwd_ld", S.data_ptr(), pol.data_ptr(), B, H, T, Tp, Tp, 1e-6, P.data_ptr(), stats.data_ptr(), st)
    lib.call("d2s_softmax_policy_bwd_ld", S.data_ptr(), pol.data_ptr(), G.data_ptr(), stats.data_ptr(), B, H, T, Tp, Tp, 1e-6, dS.data_ptr(), gp.data_ptr(), st)
torch.cuda.synchronize()
print("ok")
