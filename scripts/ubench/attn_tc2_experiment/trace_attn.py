import os, sys, ctypes
os.environ["D2S_ATTN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, d2s
import numpy as np
T = int(sys.argv[1]) if len(sys.argv) > 1 else 197
B = 1024
qkv = torch.randn(B, T, 1152, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    d2s.pkg.ops.attention_core(qkv, 6)
torch.cuda.synchronize()
lib = d2s._lib.load()
buf = np.zeros(4 * 16 * 8, dtype=np.int64)
lib.d2s_debug_read_trace.argtypes = [ctypes.c_void_p]
assert lib.d2s_debug_read_trace(buf.ctypes.data) == 0
tr = buf.reshape(4, 16, 8)
t0 = tr[tr > 0].min()
names = {0: ["wait qk_empty", "got", "wait v_empty", "got"], 1: ["wait qk_full", "got", "S issued", "v_full got", "p_full[t0] got", "p_full[t1] got", "PV issued"],
         2: ["wait s_full", "got", "pass1 done", "pass2 done", "o_full got", "epi done"], 3: ["wait s_full", "got", "pass1 done", "pass2 done", "o_full got", "epi done"]}
for role, nm in [(0, "producer"), (1, "mma"), (2, "wg0"), (3, "wg1")]:
    print("==", nm, names[role])
    for it in range(10):
        row = tr[role, it]
        print(it, [int(x - t0) if x > 0 else None for x in row[:len(names[role])]])
