"""Guard-zone tests: the stand-in for compute-sanitizer, which is closed on this GPU pool (profiles/r02_sanitizer_closed.txt).

Every kernel below is called through the C ABI with each OUTPUT embedded in a larger allocation whose guard zones hold a
sentinel bit pattern that must survive the launch (out-of-bounds writes), each INPUT followed by a NaN-poisoned guard zone
(an out-of-bounds read that reaches a result turns it into NaN), outputs pre-poisoned with NaN (every element must be
written), and twice in a row (run-to-run determinism of everything that is not an fp32 atomic sum).  Shapes are ragged on
purpose: row counts that are not multiples of the 128-row tiles, T with a nearly empty second tile, odd batch sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096          # elements on either side
SENT = 0x5A


class Guarded:
    """A tensor of `shape` inside a larger byte buffer; `ok()` is True when the guard zones still hold the sentinel."""

    def __init__(self, shape, dtype, fill=None, poison_after=False):
        n = 1
        for s in shape:
            n *= s
        es = torch.empty(0, dtype=dtype).element_size()
        gb = GUARD * es
        self.raw = torch.full((gb + n * es + gb,), SENT, dtype=torch.uint8, device="cuda")
        self.t = self.raw[gb:gb + n * es].view(dtype).view(shape)
        if fill is not None:
            self.t.copy_(fill)
        else:
            self.t.fill_(float("nan") if dtype.is_floating_point else -1)
        self.gb = gb
        if poison_after:                     # inputs: what lies behind the tensor is NaN, not something plausible
            self.raw[gb + n * es:].view(torch.int16).fill_(-1)       # 0xFFFF: NaN as bf16, NaN pairs as fp32
            self.tail = self.raw[gb + n * es:].clone()
        else:
            self.tail = None

    def ok(self):
        head = bool((self.raw[:self.gb] == SENT).all())
        tail = self.raw[self.raw.numel() - self.gb:]
        return head and (bool((tail == SENT).all()) if self.tail is None else bool(torch.equal(tail, self.tail)))

    @property
    def ptr(self):
        return self.t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,T,H,pol", [(3, 197, 6, True), (2, 138, 3, False), (5, 97, 2, True), (1, 129, 1, True), (2, 208, 2, False)])
def test_attention_forward_and_backward_stay_in_bounds(d2s, B, T, H, pol):
    lib = d2s._lib
    hd, D = 64, H * 64
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = Guarded((B, T, 3 * D), torch.bfloat16, (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.7).bfloat16(), poison_after=True)
    go = Guarded((B, T, D), torch.bfloat16, (torch.randn(B, T, D, device="cuda", generator=g) * 0.5).bfloat16(), poison_after=True)
    policy = None
    if pol:
        p = (torch.rand(B, T, device="cuda", generator=g) > 0.3).float()
        p[:, 0] = 1
        policy = Guarded((B, T), torch.float32, p, poison_after=True)
    results = []
    for _ in range(2):
        out, cls = Guarded((B, T, D), torch.bfloat16), Guarded((B, H, T), torch.float32)
        stats = Guarded((B, H, T, 4), torch.float32)
        dqkv = Guarded((B, T, 3 * D), torch.bfloat16)
        gpol = Guarded((B, T), torch.float32, torch.zeros(B, T, device="cuda")) if pol else None
        lib.call("d2s_attn_policy_fwd", qkv.ptr, policy.ptr if pol else None, 1, B, T, H, hd, hd ** -0.5, 1e-6, out.ptr, cls.ptr,
                 stats.ptr, _stream())
        lib.call("d2s_attn_policy_bwd", qkv.ptr, policy.ptr if pol else None, out.ptr, go.ptr, cls.ptr, None, stats.ptr, B, T, H, hd,
                 hd ** -0.5, dqkv.ptr, gpol.ptr if pol else None, _stream())
        torch.cuda.synchronize()
        for name, gd in (("out", out), ("cls_row", cls), ("stats", stats), ("dqkv", dqkv), ("gpolicy", gpol), ("qkv", qkv), ("gout", go)):
            if gd is not None:
                assert gd.ok(), f"{name}: guard zone overwritten"
        for name, gd in (("out", out), ("cls_row", cls), ("stats", stats), ("dqkv", dqkv), ("gpolicy", gpol)):
            if gd is not None:
                assert bool(torch.isfinite(gd.t.float()).all()), f"{name}: element left unwritten or fed by an out-of-bounds read"
        results.append((out.t.clone(), cls.t.clone(), dqkv.t.clone()))
    for a, b in zip(*results):
        assert torch.equal(a, b)


@pytest.mark.parametrize("M", [1, 127, 129, 300, 197 * 3])
def test_gemm_family_stays_in_bounds(d2s, M):
    lib = d2s._lib
    D, HID = 384, 1536
    g = torch.Generator(device="cuda").manual_seed(M)
    rn = lambda *s, k=1.0: (torch.randn(*s, device="cuda", generator=g) * k).bfloat16()   # noqa: E731
    h = Guarded((M, D), torch.bfloat16, rn(M, D), poison_after=True)
    x = Guarded((M, D), torch.bfloat16, rn(M, D), poison_after=True)
    a4 = Guarded((M, HID), torch.bfloat16, rn(M, HID, k=0.5), poison_after=True)
    w1, b1, w2, b2 = rn(HID, D, k=0.05), rn(HID, k=0.1), rn(D, HID, k=0.03), rn(D, k=0.1)
    wp = rn(D, D, k=0.05)
    gam, bet = (1 + 0.1 * torch.randn(D, device="cuda", generator=g)).bfloat16(), rn(D, k=0.1)
    for _ in range(2):
        o_sum, o_norm = Guarded((M, D), torch.bfloat16), Guarded((M, D), torch.bfloat16)
        lib.call("d2s_mlp_residual_ln_bf16", h.ptr, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), x.ptr, gam.data_ptr(),
                 bet.data_ptr(), 1e-6, M, D, HID, 1, 0, o_sum.ptr, o_norm.ptr, _stream())
        p_sum, p_norm = Guarded((M, D), torch.bfloat16), Guarded((M, D), torch.bfloat16)
        lib.call("d2s_linear_residual_ln_bf16", h.ptr, wp.data_ptr(), b2.data_ptr(), x.ptr, gam.data_ptr(), bet.data_ptr(), 1e-6, M, D, D,
                 p_sum.ptr, p_norm.ptr, _stream())
        f_sum, f_norm = Guarded((M, D), torch.bfloat16), Guarded((M, D), torch.bfloat16)
        lib.call("d2s_linear_residual_ln_bf16", a4.ptr, w2.data_ptr(), b2.data_ptr(), x.ptr, gam.data_ptr(), bet.data_ptr(), 1e-6, M, D, HID,
                 f_sum.ptr, f_norm.ptr, _stream())
        act, pre = Guarded((M, HID), torch.bfloat16), Guarded((M, HID), torch.bfloat16)
        lib.call("d2s_linear_act_pair_bf16", h.ptr, w1.data_ptr(), b1.data_ptr(), M, HID, D, 1, act.ptr, pre.ptr, _stream())
        # 192-column tiles with resident input rows (qkv: N = 1152; the predictors' Linear + GELU: N = 384)
        wq, bq = rn(3 * D, D, k=0.05), rn(3 * D, k=0.1)
        qkv, g384 = Guarded((M, 3 * D), torch.bfloat16), Guarded((M, D), torch.bfloat16)
        lib.call("d2s_linear_act_pair_bf16", h.ptr, wq.data_ptr(), bq.data_ptr(), M, 3 * D, D, 0, qkv.ptr, None, _stream())
        lib.call("d2s_linear_act_pair_bf16", h.ptr, wp.data_ptr(), b2.data_ptr(), M, D, D, 1, g384.ptr, None, _stream())
        # statistics instead of the normalised copy, and the MLP kernel that applies the LayerNorm to its input tile itself
        s_sum, s_st = Guarded((M, D), torch.bfloat16), Guarded((M, 2), torch.float32)
        lib.call("d2s_linear_residual_stats_bf16", h.ptr, wp.data_ptr(), b2.data_ptr(), x.ptr, 1e-6, M, D, D, s_sum.ptr, s_st.ptr, _stream())
        l_sum, l_norm, l_st = Guarded((M, D), torch.bfloat16), Guarded((M, D), torch.bfloat16), Guarded((M, 2), torch.float32)
        lib.call("d2s_mlp_lnin_residual_ln_bf16", s_sum.ptr, s_st.ptr, gam.data_ptr(), bet.data_ptr(), w1.data_ptr(), b1.data_ptr(),
                 w2.data_ptr(), b2.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-6, M, D, HID, 1, 0, l_sum.ptr, l_norm.ptr, l_st.ptr, _stream())
        q2 = Guarded((M, 3 * D), torch.bfloat16)      # the next block's qkv projection normalising its input rows from those statistics
        lib.call("d2s_linear_lnin_act_pair_bf16", l_sum.ptr, l_st.ptr, gam.data_ptr(), bet.data_ptr(), wq.data_ptr(), bq.data_ptr(), M, 3 * D, D,
                 0, q2.ptr, _stream())
        torch.cuda.synchronize()
        for name, gd in (("mlp sum", o_sum), ("mlp norm", o_norm), ("proj sum", p_sum), ("proj norm", p_norm), ("fc2 sum", f_sum),
                         ("fc2 norm", f_norm), ("act", act), ("pre", pre), ("qkv", qkv), ("linear + gelu 384", g384), ("stats sum", s_sum),
                         ("stats", s_st), ("lnin sum", l_sum), ("lnin norm", l_norm), ("lnin stats", l_st), ("qkv lnin", q2), ("h", h), ("x", x),
                         ("a4", a4)):
            assert gd.ok(), f"{name}: guard zone overwritten"
            if gd not in (h, x, a4):
                assert bool(torch.isfinite(gd.t.float()).all()), f"{name}: element left unwritten or fed by an out-of-bounds read"


@pytest.mark.parametrize("B,T,D,K", [(3, 197, 384, 137), (1, 138, 768, 96), (7, 97, 192, 1), (2, 5, 8, 4)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gather_scatter_layernorm_stay_in_bounds(d2s, B, T, D, K, dtype):
    lib = d2s._lib
    code = 1 if dtype == torch.bfloat16 else 0
    g = torch.Generator(device="cuda").manual_seed(T + D)
    x = Guarded((B, T, D), dtype, torch.randn(B, T, D, device="cuda", generator=g).to(dtype), poison_after=True)
    idx = torch.stack([torch.sort(torch.randperm(T - 1, device="cuda", generator=g)[:K]).values for _ in range(B)])
    gam, bet = torch.ones(D, device="cuda", dtype=dtype), torch.zeros(D, device="cuda", dtype=dtype)
    out = Guarded((B, K + 1, D), dtype)
    lib.call("d2s_gather_tokens", x.ptr, code, B, T, D, idx.data_ptr(), K, 1, out.ptr, _stream())
    gx = Guarded((B, T, D), dtype)
    lib.call("d2s_scatter_tokens_bwd", out.ptr, code, B, T, D, idx.data_ptr(), K, 1, gx.ptr, _stream())
    gs, gn = Guarded((B, K + 1, D), dtype), Guarded((B, K + 1, D), dtype)
    if D % 8 == 0 and D <= 768:
        lib.call("d2s_gather_layernorm", x.ptr, idx.data_ptr(), gam.data_ptr(), bet.data_ptr(), code, B, T, D, K, 1e-6, gs.ptr, gn.ptr, _stream())
    torch.cuda.synchronize()
    for name, gd in (("gather", out), ("scatter", gx), ("gather+LN sum", gs), ("gather+LN norm", gn), ("x", x)):
        assert gd.ok(), f"{name}: guard zone overwritten"
    assert bool(torch.isfinite(out.t.float()).all()) and bool(torch.isfinite(gx.t.float()).all())
    ref = torch.cat([x.t[:, :1], torch.gather(x.t[:, 1:], 1, idx[..., None].expand(-1, -1, D))], 1)
    assert torch.equal(out.t, ref)


@pytest.mark.parametrize("B,N,K,S", [(1, 196, 98, 500), (3, 196, 137, 64), (2, 50, 7, 33)])
def test_perturbed_topk_stays_in_bounds(d2s, B, N, K, S):
    lib = d2s._lib
    g = torch.Generator(device="cuda").manual_seed(N + K)
    x = Guarded((B, N), torch.float32, torch.softmax(torch.randn(B, N, device="cuda", generator=g), -1), poison_after=True)
    nz = Guarded((B, S, N), torch.float32, torch.randn(B, S, N, device="cuda", generator=g), poison_after=True)
    outs = []
    for _ in range(2):
        ind, eg, gx = Guarded((B, K, N), torch.float32), Guarded((B, K, N), torch.float32), Guarded((B, N), torch.float32)
        lib.call("d2s_ptopk_fwd", x.ptr, nz.ptr, B, N, K, S, 0.05, ind.ptr, eg.ptr, _stream())
        go = torch.randn(B, K, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        lib.call("d2s_ptopk_bwd", go.data_ptr(), eg.ptr, B, N, K, gx.ptr, _stream())
        torch.cuda.synchronize()
        for name, gd in (("indicators", ind), ("egrad", eg), ("gx", gx), ("x", x), ("noise", nz)):
            assert gd.ok(), f"{name}: guard zone overwritten"
        assert bool(torch.isfinite(ind.t).all()) and bool(torch.isfinite(eg.t).all()) and bool(torch.isfinite(gx.t).all())
        torch.testing.assert_close(ind.t.sum(-1), torch.ones(B, K, device="cuda"), rtol=0, atol=1e-5)
        outs.append((ind.t.clone(), eg.t.clone(), gx.t.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("rows,D", [(1, 384), (333, 384), (50, 768), (1000, 192)])
def test_layernorm_train_kernels_and_colsum_stay_in_bounds(d2s, rows, D):
    lib = d2s._lib
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = Guarded((rows, D), torch.bfloat16, torch.randn(rows, D, device="cuda", generator=g).bfloat16(), poison_after=True)
    y = Guarded((rows, D), torch.bfloat16, torch.randn(rows, D, device="cuda", generator=g).bfloat16(), poison_after=True)
    w, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    s, h, st = Guarded((rows, D), torch.bfloat16), Guarded((rows, D), torch.bfloat16), Guarded((rows, 2), torch.float32)
    lib.call("d2s_add_layernorm_fwd", x.ptr, y.ptr, 1, w.data_ptr(), b.data_ptr(), rows, D, 1e-6, s.ptr, h.ptr, 1, st.ptr, _stream())
    dx = Guarded((rows, D), torch.bfloat16)
    dg, db = Guarded((D,), torch.float32, torch.zeros(D, device="cuda")), Guarded((D,), torch.float32, torch.zeros(D, device="cuda"))
    lib.call("d2s_add_layernorm_bwd", y.ptr, 1, s.ptr, 1, st.ptr, w.data_ptr(), x.ptr, rows, D, dx.ptr, dg.ptr, db.ptr, _stream())
    cs = Guarded((D,), torch.float32)
    lib.call("d2s_colsum_bf16", x.ptr, rows, D, cs.ptr, _stream())
    du, dbb = Guarded((rows, D), torch.bfloat16), Guarded((D,), torch.float32, torch.zeros(D, device="cuda"))
    lib.call("d2s_gelu_bwd_colsum_bf16", x.ptr, y.ptr, rows, D, du.ptr, dbb.ptr, _stream())
    torch.cuda.synchronize()
    for name, gd in (("sum", s), ("h", h), ("stats", st), ("dx", dx), ("dgamma", dg), ("dbeta", db), ("colsum", cs), ("du", du),
                     ("db", dbb), ("x", x), ("y", y)):
        assert gd.ok(), f"{name}: guard zone overwritten"
        if gd not in (x, y):
            assert bool(torch.isfinite(gd.t.float()).all()), f"{name}: element left unwritten or fed by an out-of-bounds read"
    torch.testing.assert_close(cs.t, x.t.float().sum(0), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("B,N,C", [(3, 196, 384), (2, 137, 384), (1, 5, 16), (5, 67, 768)])
def test_predictor_split_and_token_kl_stay_in_bounds(d2s, B, N, C):
    """d2s_pool_concat_fwd/bwd, d2s_token_kl_fwd (with x[:, 1:]-style batch strides) and d2s_layernorm_seg_fwd/bwd."""
    lib = d2s._lib
    g = torch.Generator(device="cuda").manual_seed(N + C)
    bf = torch.bfloat16
    h = Guarded((B, N, C), bf, torch.randn(B, N, C, device="cuda", generator=g).bfloat16(), poison_after=True)
    go = Guarded((B, N, C), bf, torch.randn(B, N, C, device="cuda", generator=g).bfloat16(), poison_after=True)
    pol = Guarded((B, N), torch.float32, (torch.rand(B, N, device="cuda", generator=g) > 0.3).float() + 0.5, poison_after=True)
    # token KL inputs: (B, N+1, C) tensors read from row 1 on (batch stride (N+1)*C)
    s = Guarded((B, N + 1, C), torch.float32, torch.randn(B, N + 1, C, device="cuda", generator=g), poison_after=True)
    t = Guarded((B, N + 1, C), bf, torch.randn(B, N + 1, C, device="cuda", generator=g).bfloat16(), poison_after=True)
    w = torch.ones(C, device="cuda")
    bb = torch.zeros(C, device="cuda")
    for _ in range(2):
        out, pooled, wsum = Guarded((B, N, C), bf), Guarded((B, C // 2), torch.float32), Guarded((B,), torch.float32)
        dh, dpol = Guarded((B, N, C), bf), Guarded((B, N), torch.float32)
        lib.call("d2s_pool_concat_fwd", h.ptr, pol.ptr, 1, B, N, C, out.ptr, pooled.ptr, wsum.ptr, _stream())
        lib.call("d2s_pool_concat_bwd", go.ptr, h.ptr, pol.ptr, pooled.ptr, wsum.ptr, 1, B, N, C, dh.ptr, dpol.ptr, _stream())
        kl, diff = Guarded((B * N,), torch.float32), Guarded((B * N, C), torch.float32)
        lib.call("d2s_token_kl_fwd", s.t[:, 1:].data_ptr(), 0, (N + 1) * C, t.t[:, 1:].data_ptr(), 1, (N + 1) * C, B, N, C, kl.ptr, diff.ptr,
                 _stream())
        hn, stats = Guarded((B, N, C), bf), Guarded((B * N, 2), torch.float32)
        dx = Guarded((B, N + 1, C), torch.float32)
        dgb = Guarded((2, C), torch.float32, torch.zeros(2, C, device="cuda"))
        if C <= 768:
            lib.call("d2s_layernorm_seg_fwd", s.ptr, 0, w.data_ptr(), bb.data_ptr(), B * N, C, N, 1, 1e-6, hn.ptr, 1, stats.ptr, _stream())
            lib.call("d2s_layernorm_seg_bwd", go.ptr, 1, s.ptr, 0, stats.ptr, w.data_ptr(), B * N, C, N, 1, dx.ptr, dgb.t[0].data_ptr(),
                     dgb.t[1].data_ptr(), _stream())
        torch.cuda.synchronize()
        named = [("out", out), ("pooled", pooled), ("wsum", wsum), ("dh", dh), ("dpolicy", dpol), ("kl", kl), ("diff", diff), ("hn", hn),
                 ("stats", stats), ("dx", dx), ("dgamma/dbeta", dgb)]
        for name, gd in named + [("h", h), ("gout", go), ("policy", pol), ("s", s), ("t", t)]:
            assert gd.ok(), f"{name}: guard zone overwritten"
        for name, gd in named:
            assert bool(torch.isfinite(gd.t.float()).all()), f"{name}: element left unwritten or fed by an out-of-bounds read"
        assert float(dx.t[:, 0].abs().max()) == 0.0                      # the skipped rows' gradient is written as zeros


@pytest.mark.parametrize("B,N,K", [(3, 196, 137), (2, 137, 96), (5, 96, 67), (1, 129, 1), (2, 7, 7), (150, 131, 60)])
def test_predictor_tail_kernel_stays_in_bounds(d2s, B, N, K):
    """d2s_predictor_a_tail_bf16: TMA tiles whose second box is nearly empty, rows past N of the last tile, the in-place u tile,
    the double-buffered key / bias rows (B > 148: several images per CTA)."""
    lib = d2s._lib
    H = 192
    g = torch.Generator(device="cuda").manual_seed(N + K)
    bf = torch.bfloat16
    rnd = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
    local = Guarded((B, N, H), bf, rnd(B, N, H, sc=0.6).bfloat16(), poison_after=True)
    per_image = Guarded((B, H), bf, rnd(B, H, sc=0.3).bfloat16(), poison_after=True)
    w2 = Guarded((H, 2 * H), bf, rnd(H, 2 * H, sc=0.08).bfloat16(), poison_after=True)
    w3 = Guarded((H // 2, H), bf, rnd(H // 2, H, sc=0.1).bfloat16(), poison_after=True)
    b3 = Guarded((H // 2,), bf, rnd(H // 2, sc=0.1).bfloat16(), poison_after=True)
    w4 = Guarded((2, H // 2), torch.float32, rnd(2, H // 2, sc=0.3), poison_after=True)
    b4 = Guarded((2,), torch.float32, rnd(2, sc=0.1), poison_after=True)
    prev = Guarded((B, N), torch.float32, (torch.rand(B, N, device="cuda", generator=g) > 0.3).float(), poison_after=True)
    outs = []
    for _ in range(2):
        logp, kept, pk = Guarded((B, N, 2), torch.float32), Guarded((B, K), torch.int64), Guarded((B, K), torch.float32)
        lib.call("d2s_predictor_a_tail_bf16", local.ptr, H, N * H, per_image.ptr, w2.ptr, w3.ptr, b3.ptr, w4.ptr, b4.ptr, prev.ptr, B, N, H, K,
                 logp.ptr, kept.ptr, pk.ptr, _stream())
        torch.cuda.synchronize()
        for name, gd in [("logp", logp), ("kept", kept), ("prev_kept", pk), ("local", local), ("per_image", per_image), ("w2", w2),
                         ("w3", w3), ("b3", b3), ("w4", w4), ("b4", b4), ("prev", prev)]:
            assert gd.ok(), f"{name}: guard zone overwritten"
        assert bool(torch.isfinite(logp.t).all()) and bool(torch.isfinite(pk.t).all()), "element left unwritten or fed by an out-of-bounds read"
        assert int(kept.t.min()) >= 0 and int(kept.t.max()) < N
        assert all(len(set(row.tolist())) == K for row in kept.t.cpu()[:8])
        outs.append((logp.t.clone(), kept.t.clone(), pk.t.clone()))
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1])), "run-to-run difference"


@pytest.mark.parametrize("n,begin,end", [(1000, 0, 1000), (1003, 13, 771), (4096, 8, 4088), (17, 1, 16)])
def test_adamw_flat_touches_only_its_range(d2s, n, begin, end):
    lib = d2s._lib
    g = torch.Generator(device="cuda").manual_seed(n)
    p0 = torch.randn(n, device="cuda", generator=g)
    p, m, v = Guarded((n,), torch.float32, p0), Guarded((n,), torch.float32, torch.zeros(n, device="cuda")), Guarded((n,), torch.float32, torch.zeros(n, device="cuda"))
    gr = Guarded((n,), torch.float32, torch.randn(n, device="cuda", generator=g), poison_after=True)
    sh = Guarded((n,), torch.bfloat16, torch.zeros(n, device="cuda").bfloat16())
    lr, st = torch.full((1,), 1e-3, device="cuda"), torch.ones(1, device="cuda")
    lib.call("d2s_adamw_flat_f32", p.ptr, gr.ptr, m.ptr, v.ptr, sh.ptr, begin, end, lr.data_ptr(), st.data_ptr(), 0.9, 0.999, 1e-8, 0.05, 1.0,
             _stream())
    torch.cuda.synchronize()
    for name, gd in (("p", p), ("m", m), ("v", v), ("shadow", sh), ("g", gr)):
        assert gd.ok(), f"{name}: guard zone overwritten"
    inside = torch.zeros(n, dtype=torch.bool, device="cuda")
    inside[begin:end] = True
    assert torch.equal(p.t[~inside], p0[~inside]) and float(m.t[~inside].abs().max() if (~inside).any() else 0.0) == 0.0
    assert bool((p.t[inside] != p0[inside]).all()) and bool(torch.isfinite(p.t).all())
    assert torch.equal(sh.t[inside], p.t[inside].bfloat16()) and float(sh.t[~inside].float().abs().max() if (~inside).any() else 0.0) == 0.0


@pytest.mark.parametrize("M,K", [(300, 768), (77, 3072), (1000, 64)])
def test_linear_residual_ln_at_768_columns_stays_in_bounds(d2s, M, K):
    lib = d2s._lib
    N = 768
    g = torch.Generator(device="cuda").manual_seed(M + K)
    bf = torch.bfloat16
    a = Guarded((M, K), bf, torch.randn(M, K, device="cuda", generator=g).bfloat16(), poison_after=True)
    w = Guarded((N, K), bf, (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16(), poison_after=True)
    x = Guarded((M, N), bf, torch.randn(M, N, device="cuda", generator=g).bfloat16(), poison_after=True)
    bias, gam, bet = (torch.randn(N, device="cuda", generator=g).bfloat16() for _ in range(3))
    outs = []
    for _ in range(2):
        s_, h_ = Guarded((M, N), bf), Guarded((M, N), bf)
        lib.call("d2s_linear_residual_ln_bf16", a.ptr, w.ptr, bias.data_ptr(), x.ptr, gam.data_ptr(), bet.data_ptr(), 1e-6, M, N, K, s_.ptr,
                 h_.ptr, _stream())
        torch.cuda.synchronize()
        for name, gd in (("out_sum", s_), ("out_norm", h_), ("a", a), ("w", w), ("x", x)):
            assert gd.ok(), f"{name}: guard zone overwritten"
        assert bool(torch.isfinite(s_.t.float()).all()) and bool(torch.isfinite(h_.t.float()).all())
        outs.append((s_.t.clone(), h_.t.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])      # deterministic
