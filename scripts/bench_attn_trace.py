"""Where does an attention-forward CTA spend its cycles?  Needs a profiling build (D2S_NVCC_EXTRA=-DD2S_ATTN_TRACE_BUILD python
dense2sparse-vit_b200/build.py): the kernel then accumulates clock64 totals per CTA -- control warp: waits on k_full / q_full /
tmem_free / p_full / v_full / o_full and everything else (issue); softmax warp 0: wait s_full, softmax pass, publish, wait
o_full, epilogue, rest.  Prints the mean over CTAs in cycles per (image, head) unit.  Diagnostic only.

    python scripts/bench_attn_trace.py [T ...]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402

ops = d2s.pkg.ops
lib = ctypes.CDLL(os.path.join(ROOT, "dense2sparse-vit_b200", "libd2s_b200.so"))
lib.d2s_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, H = 1024, 6
for T in [int(a) for a in sys.argv[1:]] or [197, 138, 97]:
    qkv = (torch.randn(B, T, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    for _ in range(3):
        out, _ = ops.attention_core(qkv, H)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out, _ = ops.attention_core(qkv, H)
    e.record()
    torch.cuda.synchronize()
    per_sm = 2 if T > 128 else 4
    ctas = min(B * H, per_sm * 148)
    buf = (ctypes.c_longlong * (ctas * 16))()
    rc = lib.d2s_debug_attn_trace(buf, ctas)
    assert rc == 0, rc
    t = torch.tensor(list(buf), dtype=torch.float64).view(ctas, 2, 8)
    units = B * H / ctas
    m = t.mean(0) / units
    print(f"T={T}: {s.elapsed_time(e) * 1e3:.1f} us, {ctas} CTAs, {units:.1f} units per CTA; cycles per unit (mean over CTAs)")
    print("  control : k_full %.0f  q_full %.0f  tmem_free %.0f  p_full %.0f  v_full %.0f  o_full(last) %.0f  issue/other %.0f  | total %.0f"
          % tuple(m[0].tolist()))
    print("  softmax0: s_full %.0f  LDTM+wait %.0f  pack/STTM/publish %.0f  o_full %.0f  epilogue %.0f  exps+sums %.0f  other %.0f  | total %.0f"
          % tuple(m[1].tolist()))
