"""GPU parity tests, op level: every C-ABI kernel family (called through ctypes via tacorl_b200.ops)
against the CPU oracle (oracle/tacorl_oracle.py) or the defining formula in fp64.
Tolerance for the fp32 path: 1e-4 relative L2 per tensor (BASELINE.json north_star)."""
import math

import pytest
import torch

from oracle import synth as S
from oracle import tacorl_oracle as O
from tests.gpu_util import DEV, assert_close, load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _ops():
    from tacorl_b200 import ops
    ops.set_precision("fp32")
    return ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 2048, 2048), (960, 182, 2048), (1024, 32, 192),
                                   (300, 64, 512), (257, 130, 33), (64, 64, 1000), (2000, 256, 128)])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False)])
def test_gemm_variants(M, N, K, tA, tB):
    ops = _ops()
    g = _g(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g)
    B = torch.randn((N, K) if tB else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    want = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    want = 0.5 * want + 0.25 * C0.double() + bias.double()
    C = C0.clone().to(DEV)
    ops.gemm(A.to(DEV), B.to(DEV), C, transA=tA, transB=tB, alpha=0.5, beta=0.25, bias=bias.to(DEV))
    assert_close("gemm", C, want, 2e-5)


def test_gemm_activation_pre_and_strided_rows():
    ops = _ops()
    g = _g(5)
    A = torch.randn(50, 96, generator=g)      # use columns 16:80 of a wider buffer
    W = torch.randn(40, 64, generator=g)
    b = torch.randn(40, generator=g)
    Ad = A.to(DEV)
    out_wide = torch.zeros(50, 100, device=DEV)
    pre = torch.empty(50, 40, device=DEV)
    ops.gemm(Ad[:, 16:80], W.to(DEV), out_wide[:, 10:50], transB=True, bias=b.to(DEV), act=2, Cpre=pre)
    z = A[:, 16:80].double() @ W.double().t() + b.double()
    assert_close("pre", pre, z, 2e-5)
    assert_close("silu", out_wide[:, 10:50], z * torch.sigmoid(z), 2e-5)
    assert float(out_wide[:, :10].abs().sum()) == 0 and float(out_wide[:, 50:].abs().sum()) == 0


@pytest.mark.parametrize("act", [None, "relu", "silu"])
def test_linear_fn_grads(act):
    ops = _ops()
    g = _g(11)
    x = torch.randn(6, 9, 48, generator=g)
    W = torch.randn(70, 48, generator=g) * 0.2
    b = torch.randn(70, generator=g)
    cot = torch.randn(6, 9, 70, generator=g)
    xd, Wd, bd = (t.clone().to(DEV).requires_grad_(True) for t in (x, W, b))
    y = ops.linear(xd, Wd, bd, act)
    (y * cot.to(DEV)).sum().backward()
    x64, W64, b64 = (t.double().requires_grad_(True) for t in (x, W, b))
    z = torch.nn.functional.linear(x64, W64, b64)
    yr = {None: z, "relu": torch.relu(z), "silu": torch.nn.functional.silu(z)}[act]
    (yr * cot.double()).sum().backward()
    assert_close("y", y, yr, RTOL)
    assert_close("dx", xd.grad, x64.grad, RTOL)
    assert_close("dW", Wd.grad, W64.grad, RTOL)
    assert_close("db", bd.grad, b64.grad, RTOL)


@pytest.mark.parametrize("n,h,w", [(3, 84, 84), (2, 128, 128), (2, 150, 200), (2, 200, 200), (1, 200, 200)])
def test_encoder_fwd_bwd_vs_oracle(n, h, w):
    ops = _ops()
    rec = load_golden("encoder_shapes")
    sd = S.synth_state_dict(rec["shapes"], rec["seed"])
    sd["model.6.temperature"] = torch.tensor([0.7])
    x = S.synth_images((n, 3, h, w), rec["seed"], f"enc{h}x{w}")
    P = O.params_from({k: v.double() for k, v in sd.items()})
    y = O.lmp_encoder(P, "", x.double())
    cot = torch.rand(y.shape, generator=S._gen(rec["seed"], f"cot{h}x{w}")) * 2 - 1
    (y * cot.double()).sum().backward()
    names = list(rec["shapes"].keys())
    params = [sd[k].clone().to(DEV).requires_grad_(True) for k in names]
    emb = ops.lmp_encoder(x.to(DEV), params)
    (emb * cot.to(DEV)).sum().backward()
    assert_close("emb", emb, y, RTOL)
    for k, p in zip(names, params):
        assert_close(f"grad {k}", p.grad, P[k].grad, RTOL)
    # the golden numbers come straight from the reference (fingerprints incl. out_head)
    for case in rec["cases"]:
        if (case["n"], case["h"], case["w"]) == (n, h, w):
            assert torch.allclose(emb[0, :8].cpu(), torch.tensor(case["out_head"]), rtol=1e-4, atol=1e-5)
            assert S.fingerprint_close(S.fingerprint(emb), case["out"], 2e-4)
            for k, p in zip(names, params):
                assert S.fingerprint_close(S.fingerprint(p.grad), case["grads"][k], 2e-4), k


def test_encoder_inference_chunked_matches_training_path():
    ops = _ops()
    rec = load_golden("encoder_shapes")
    sd = S.synth_state_dict(rec["shapes"], rec["seed"])
    params = [sd[k].to(DEV) for k in rec["shapes"]]
    x = S.synth_images((5, 3, 84, 84), 3, "inf").to(DEV)
    with torch.no_grad():
        a = ops.lmp_encoder(x, params)
    pg = [p.clone().requires_grad_(True) for p in params]
    b = ops.lmp_encoder(x, pg)
    assert_close("inference vs training forward", a, b, 1e-6)


@pytest.mark.parametrize("bidir,last_only,use_h0", [(False, False, False), (False, False, True), (True, False, False),
                                                    (True, True, False)])
@pytest.mark.parametrize("B,T,I,H", [(4, 7, 12, 40), (64, 16, 32, 256)])
def test_relu_rnn_vs_oracle(bidir, last_only, use_h0, B, T, I, H):
    ops = _ops()
    g = _g(B + T + I + H + bidir)
    # weights from the test's own generator, not torch's global one: the data must not depend on which tests ran before.
    # (With ~10^6 pre-activations per case, about one weight draw in 25 puts one of them within fp32 rounding of the ReLU
    # kink: measured with scripts/debug_rnn_flake.py, exactly ONE gate of the top layer then differs between the fp32 run
    # and the fp64 reference, the forward still agrees to 2e-7 and every gradient moves by ~1e-3, reproducibly for that
    # draw.  That is a property of comparing a ReLU network across precisions, not of the kernels.)
    rnn = torch.nn.RNN(I, H, num_layers=2, nonlinearity="relu", bidirectional=bidir, batch_first=True)
    bound = 1.0 / H ** 0.5
    sd = {k: (torch.rand(v.shape, generator=g) * 2 - 1) * bound for k, v in rnn.state_dict().items()}
    names = list(sd.keys())
    x = torch.randn(B, T, I, generator=g)
    D = 2 if bidir else 1
    h0 = torch.randn(2 * D, B, H, generator=g) * 0.3 if use_h0 else None
    P = {"r." + k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    out, hn = O.rnn_stack(P, "r.", x64, 2, bidir, None if h0 is None else h0.double())
    want = out[:, -1] if last_only else out
    cot = torch.randn(want.shape, generator=g)
    (want * cot.double()).sum().backward()
    ws = [sd[k].clone().to(DEV).requires_grad_(True) for k in names]
    xd = x.to(DEV).requires_grad_(True)
    got, hn_d = ops.relu_rnn(xd.transpose(0, 1), ws, 2, bidir, last_only, None if h0 is None else h0.to(DEV))
    if not last_only:
        got = got.transpose(0, 1)
        assert_close("h_n", hn_d, hn, RTOL)
    (got * cot.to(DEV)).sum().backward()
    assert_close("out", got, want, RTOL)
    assert_close("dx", xd.grad, x64.grad, RTOL)
    for k, w in zip(names, ws):
        assert_close(f"grad {k}", w.grad, P["r." + k].grad, RTOL, atol=1e-6)


def _dlm_inputs(rows, seed):
    g = _g(seed)
    logits = torch.randn(rows, 182, generator=g)
    logits[:, 120:180] = logits[:, 120:180] * 2 - 2
    logits[0, 120:130] = -7.0
    act = torch.rand(rows, 7, generator=g) * 2 - 1
    act[0, 0], act[0, 1], act[1, 2] = -1.0, 1.0, 0.9995
    logits[2, 60:70] = 5.0
    logits[2, 120:130] = -4.0
    act[:, 6] = torch.where(act[:, 6] > 0, 1.0, -1.0)
    return logits, act


def _oracle_dlm(logits, act):
    R = logits.shape[0]
    lp = logits[:, :60].reshape(R, 1, 6, 10)
    mu = logits[:, 60:120].reshape(R, 1, 6, 10)
    ls = logits[:, 120:180].reshape(R, 1, 6, 10)
    grip = logits[:, 180:].reshape(R, 1, 2)
    return O.dlm_loss(lp, torch.clamp(ls, min=-5.0), mu, grip, act.reshape(R, 1, 7))


def test_dlm_loss_and_grad_vs_oracle():
    ops = _ops()
    logits, act = _dlm_inputs(37, 3)
    l64 = logits.double().requires_grad_(True)
    want = _oracle_dlm(l64, act.double())
    want.backward()
    ld = logits.to(DEV).requires_grad_(True)
    got = ops.dlm_loss(ld, act.to(DEV))
    (got * 1.7).backward()
    assert abs(float(got) - float(want)) <= RTOL * abs(float(want))
    assert_close("dlogits", ld.grad, 1.7 * l64.grad, RTOL)


def test_dlm_known_answers_from_reference():
    ops = _ops()
    rec = load_golden("ops_kat")
    t = lambda k: torch.tensor(rec[k])
    lp, ls, mu, grip, act = t("lp"), t("ls"), t("mu"), t("grip"), t("act")
    B, T = act.shape[:2]
    logits = torch.cat([lp.reshape(B * T, -1), mu.reshape(B * T, -1), ls.reshape(B * T, -1), grip.reshape(B * T, -1)], 1)
    got = ops.dlm_loss(logits.to(DEV), act.reshape(B * T, 7).to(DEV))
    assert abs(float(got) - rec["dlm_loss"]) <= 1e-4 * abs(rec["dlm_loss"])
    pred, acc = ops.dlm_sample(logits.to(DEV), t("u1").reshape(B * T, 6, 10).to(DEV), t("u2").reshape(B * T, 6).to(DEV),
                               act.reshape(B * T, 7).to(DEV))
    assert torch.allclose(pred.cpu().view(B, T, 7), t("sample"), rtol=1e-4, atol=1e-5)
    want_acc = O.gripper_accuracy(t("sample"), act)
    assert abs(float(acc) - float(want_acc)) < 1e-6
    # TanhNormal log-probs
    mean, std, z, val = t("mean"), t("std"), t("z"), t("val")
    a = ops.tanh_logprob(mean.to(DEV), std.to(DEV), z.to(DEV), False)
    b = ops.tanh_logprob(mean.to(DEV), std.to(DEV), val.to(DEV), True)
    assert torch.allclose(a.cpu(), t("logp_pre"), rtol=1e-4, atol=1e-4)
    assert torch.allclose(b.cpu(), t("logp_val"), rtol=1e-4, atol=1e-4)


def test_heads_kl_tanh_grads_vs_oracle():
    ops = _ops()
    g = _g(9)
    B, Ld = 13, 16
    raw = torch.randn(B, 2 * Ld, generator=g) * 4
    raw[0, 0], raw[0, 1], raw[0, Ld], raw[0, Ld + 1] = 12.0, -11.0, 3.0, -6.0     # clamp branches
    raw2 = torch.randn(B, 2 * Ld, generator=g)
    eps = torch.randn(B, Ld, generator=g)
    cz = torch.randn(B, Ld, generator=g)
    # oracle (fp64)
    r64, r264 = raw.double().requires_grad_(True), raw2.double().requires_grad_(True)
    mu_p = torch.clamp(r64[:, :Ld], -9, 9)
    sd_p = torch.clamp(r64[:, Ld:], -5, 2).exp()
    mu_q = r264[:, :Ld]
    sd_q = torch.nn.functional.softplus(r264[:, Ld:]) + 1e-4
    kl = O.balanced_kl(mu_q, sd_q, mu_p, sd_p, 0.8)
    z = mu_q + sd_q * eps.double()
    plan = torch.tanh(z)
    lp = O.tanh_normal_log_prob(mu_q, sd_q, pre_tanh=z)
    lv = O.tanh_normal_log_prob(mu_p, sd_p, value=plan.detach())
    total = 3.0 * kl + (plan * cz.double()).sum() + 0.3 * lp.sum() - 0.2 * lv.sum()
    total.backward()
    # CUDA
    rd, r2d = raw.to(DEV).requires_grad_(True), raw2.to(DEV).requires_grad_(True)
    mp, sp = ops.gauss_head(rd)
    mq, sq = ops.softplus_head(r2d, 1e-4)
    kld = ops.kl_balanced(mq, sq, mp, sp, 0.8, True)
    pl, zd = ops.tanh_rsample(mq, sq, eps.to(DEV), True)
    lpd = ops.tanh_logprob(mq, sq, zd, False)
    lvd = ops.tanh_logprob(mp, sp, pl.detach(), True)
    tot = 3.0 * kld + (pl * cz.to(DEV)).sum() + 0.3 * lpd.sum() - 0.2 * lvd.sum()
    tot.backward()
    assert abs(float(kld) - float(kl)) <= RTOL * abs(float(kl))
    assert_close("plan", pl, plan, RTOL)
    assert_close("logp", lpd, lp, RTOL)
    assert_close("logp(value)", lvd, lv, RTOL)
    assert_close("d raw (policy head)", rd.grad, r64.grad, RTOL)
    assert_close("d raw (softplus head)", r2d.grad, r264.grad, RTOL)
    # un-balanced KL value
    kl_u = ops.kl_balanced(mq.detach(), sq.detach(), mp.detach(), sp.detach(), 0.8, False)
    want_u = O.balanced_kl(mu_q, sd_q, mu_p, sd_p, 0.8, kl_balancing=False)
    assert abs(float(kl_u) - float(want_u)) <= RTOL * abs(float(want_u))


def test_cql_loss_kernels_vs_formula():
    ops = _ops()
    g = _g(21)
    B, n = 9, 4
    q1 = torch.randn(13 * B, generator=g)
    q2 = torch.randn(13 * B, generator=g)
    lpc, lpn = torch.randn(n, B, generator=g) * 3, torch.randn(n, B, generator=g) * 3
    tq1, tq2 = torch.randn(B, generator=g), torch.randn(B, generator=g)
    rew = (torch.rand(B, generator=g) > 0.5).float()
    lap = torch.tensor([0.3])
    dens = math.log(0.5 ** 16)

    def formula(q, qd64):
        qd = q[:B]
        qr, qc, qn = (q[B + i * n * B: B + (i + 1) * n * B].view(n, B).t() for i in range(3))
        y = 10.0 * rew.double() + (1 - rew.double()) * 0.95 * torch.min(tq1, tq2).double()
        bell = ((qd - y) ** 2).mean()
        cat = torch.cat([qr - dens, qc - lpc.double().t(), qn - lpn.double().t()], dim=1)
        raw = torch.logsumexp(cat, dim=1).mean() - qd.mean()
        return bell, raw, qr.mean(), qc.mean()

    q1_64, q2_64, lap64 = q1.double().requires_grad_(True), q2.double().requires_grad_(True), lap.double().requires_grad_(True)
    b1, r1, qr1, qc1 = formula(q1_64, None)
    b2, r2, qr2, qc2 = formula(q2_64, None)
    ap = torch.clamp(lap64[0].exp(), 0.0, 1e6)
    c1, c2 = ap * (r1 - 5.0), ap * (r2 - 5.0)
    apl = (-c1 - c2) * 0.5
    l1, l2 = b1 + c1, b2 + c2
    g_lap = torch.autograd.grad(apl, lap64, retain_graph=True)[0]
    (1.3 * l1 + 0.7 * l2).backward()
    q1d, q2d = q1.to(DEV).requires_grad_(True), q2.to(DEV).requires_grad_(True)
    o1, o2, scal, dlap = ops.CqlCriticLossFn.apply(q1d, q2d, lpc.to(DEV), lpn.to(DEV), tq1.to(DEV), tq2.to(DEV),
                                                   rew.to(DEV), rew.to(DEV), lap.to(DEV), n, dens, 0.95, 10.0, 5.0,
                                                   1.0, 1.0, True)
    (1.3 * o1 + 0.7 * o2).backward()
    want = [b1, b2, c1, c2, ap, apl, l1, l2, q1_64[:B].mean(), qr1, qc1, q2_64[:B].mean(), qr2, qc2]
    for name, gv, wv in zip(ops.CQL_SCALARS, scal.cpu().tolist(), want):
        assert abs(gv - float(wv)) <= RTOL * max(1.0, abs(float(wv))), (name, gv, float(wv))
    assert_close("dq1", q1d.grad, q1_64.grad, RTOL)
    assert_close("dq2", q2d.grad, q2_64.grad, RTOL)
    assert abs(float(dlap) - float(g_lap)) <= RTOL * abs(float(g_lap))
    # actor / alpha losses
    logpi = torch.randn(B, 1, generator=g) * 2
    plp = torch.randn(B, 1, generator=g)
    la = torch.tensor([0.2])
    val, dla = ops.cql_alpha_loss(logpi.to(DEV), la.to(DEV), -7.0)
    want_a = -(la.double()[0] * (logpi.double() - 7.0)).mean()
    assert abs(float(val) - float(want_a)) <= RTOL * abs(float(want_a))
    assert abs(float(dla) - float(-(logpi.double() - 7.0).mean())) <= 1e-5
    for mode in (1, 2):
        lp64, a64, b64 = logpi.double().requires_grad_(True), plp.double().requires_grad_(True), tq1.double().view(B, 1).requires_grad_(True)
        alpha = math.exp(0.2)
        want_l = (alpha * lp64 - (a64 if mode == 1 else torch.min(a64, b64))).mean()
        want_l.backward()
        lpd, ad, bd = logpi.to(DEV).requires_grad_(True), plp.to(DEV).requires_grad_(True), tq1.view(B, 1).to(DEV).requires_grad_(True)
        got_l, out = ops.CqlActorLossFn.apply(mode, lpd, ad, bd if mode == 2 else None, la.to(DEV))
        got_l.backward()
        assert abs(float(got_l) - float(want_l)) <= RTOL * max(1.0, abs(float(want_l)))
        assert_close("d log_pi", lpd.grad, lp64.grad, RTOL)
        assert_close("d a", ad.grad, a64.grad, RTOL)
        if mode == 2:
            assert_close("d b", bd.grad, b64.grad, RTOL)


def test_adam_clip_polyak_vs_oracle():
    ops = _ops()
    from tacorl_b200.optim import FlatAdam, FlatBuffer, polyak_update
    g = _g(33)
    shapes = [(5, 3), (7,), (129, 65), (1,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    ref = [p.clone().double() for p in ps]
    st = O.new_adam_state(ref)
    opt = FlatAdam(mine, lr=3e-3, max_grad_norm=1.0)
    for it in range(4):
        gs = [torch.randn(s, generator=g) * (3 if it % 2 == 0 else 0.01) for s in shapes]
        for p, gr in zip(mine, gs):
            p.grad = gr.clone().to(DEV)
        opt.step()
        g64 = [x.double().clone() for x in gs]
        O.clip_grad_norm(g64, 1.0)
        O.adam_step(ref, g64, st, 3e-3)
    for a, b in zip(mine, ref):
        assert_close("adam param", a, b, 1e-5)
    tgt = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    tref = [t.detach().cpu().double().clone() for t in tgt]
    tb = FlatBuffer(tgt)
    polyak_update(tb, opt.flat_params, 0.005)
    O.polyak_update(tref, ref, 0.005)
    for a, b in zip(tgt, tref):
        assert_close("polyak", a, b, 1e-6)
    x = torch.randn(100003, generator=g)
    assert abs(float(ops.sqnorm(x.to(DEV))) - float((x.double() ** 2).sum())) <= 1e-5 * float((x.double() ** 2).sum())


def test_library_fails_loudly_on_cpu_tensors():
    ops = _ops()
    from tacorl_b200._lib import TacorlLibraryError
    with pytest.raises(TacorlLibraryError):
        ops.linear(torch.randn(2, 3), torch.randn(4, 3), torch.randn(4))


def test_transformer_plan_recogniser_with_dropout_masks_vs_oracle():
    """PlanRecognitionTransformersNetwork (training mode, dropout 0.1) with an explicit mask tape against the
    oracle restatement fed the same masks: output distribution parameters and every gradient."""
    _ops()
    from tacorl_b200.networks.plan_encoders.plan_recognition_transformer import PlanRecognitionTransformersNetwork
    from tacorl_b200.utils.rng import noise_tape
    g = _g(41)
    B, T, D, Hh, FF, p = 5, 8, 32, 8, 64, 0.1
    net = PlanRecognitionTransformersNetwork(state_dim=D, latent_plan_dim=16, num_heads=Hh, num_layers=2,
                                             encoder_hidden_size=FF, fc_hidden_size=96, max_position_embeddings=T,
                                             dropout_p=p)
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    sd = S.synth_state_dict(shapes, 23)
    net.load_state_dict(sd)
    net.to(DEV).train()
    emb = torch.randn(B, T, D, generator=g)

    def mk(shape):
        return (torch.rand(shape, generator=g) > p).float() / (1 - p)

    masks = {"input": mk((T, B, D))}
    tape = [masks["input"]]
    for l in range(2):
        masks[f"attn{l}"] = mk((B * Hh, T, T)); masks[f"drop1_{l}"] = mk((T, B, D))
        masks[f"ff{l}"] = mk((T, B, FF)); masks[f"drop2_{l}"] = mk((T, B, D))
        tape += [masks[f"attn{l}"], masks[f"drop1_{l}"], masks[f"ff{l}"], masks[f"drop2_{l}"]]
    P = {"pr." + k: v.double().requires_grad_(True) for k, v in sd.items()}
    e64 = emb.double().requires_grad_(True)
    mu, sdv = O.plan_recognition_transformer(P, "pr.", e64, num_heads=Hh, num_layers=2,
                                             masks={k: v.double() for k, v in masks.items()})
    cm, cs = torch.randn(B, 16, generator=g), torch.randn(B, 16, generator=g)
    ((mu * cm.double()).sum() + (sdv * cs.double()).sum()).backward()
    ed = emb.to(DEV).requires_grad_(True)
    with noise_tape(tape) as tp:
        dist = net(ed)
        assert len(tp) == 0
    ((dist.normal_mean * cm.to(DEV)).sum() + (dist.normal_std * cs.to(DEV)).sum()).backward()
    assert_close("mean", dist.normal_mean, mu, RTOL)
    assert_close("std", dist.normal_std, sdv, RTOL)
    assert_close("d emb", ed.grad, e64.grad, RTOL)
    for k, prm in net.named_parameters():
        want = P["pr." + k].grad
        if want is None:
            assert prm.grad is None, k
            continue
        assert_close(f"grad {k}", prm.grad, want, RTOL, atol=1e-6)
    net.eval()
    with torch.no_grad():
        d2 = net(ed)
    mu_e, _ = O.plan_recognition_transformer(P, "pr.", e64, num_heads=Hh, num_layers=2, masks=None)
    assert_close("eval mean", d2.normal_mean, mu_e, RTOL)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_uint8_frames_equal_prenormalised_float_frames(prec):
    """raw uint8 frames with ScaleImageTensor+Normalize fused into the first kernel give the same
    embeddings and gradients as feeding the reference's pre-normalised float images."""
    from tacorl_b200 import ops
    ops.set_precision(prec)
    try:
        rec = load_golden("encoder_shapes")
        sd = S.synth_state_dict(rec["shapes"], rec["seed"])
        names = list(rec["shapes"].keys())
        u8 = torch.randint(0, 256, (3, 3, 84, 84), generator=_g(77), dtype=torch.uint8)
        xf = (u8.float() / 255.0 - 0.5) / 0.5
        outs = []
        for x in (xf, u8):
            params = [sd[k].clone().to(DEV).requires_grad_(True) for k in names]
            emb = ops.lmp_encoder(x.to(DEV), params, (0.5, 0.5))
            emb.sum().backward()
            outs.append((emb.detach(), [p.grad for p in params]))
        tol = 1e-5 if prec == "fp32" else 2e-3
        assert_close("emb", outs[1][0], outs[0][0], tol)
        for k, a, b in zip(names, outs[1][1], outs[0][1]):
            assert_close(f"grad {k}", a, b, tol * 10, atol=1e-6)
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("rows,ins,widths,acts,two_seg,detach", [
    (64, (32,), (256, 256, 32), ("relu", "relu"), False, False),          # goal encoder
    (64, (32, 32), (256, 256, 256, 32), ("silu",) * 3, True, False),      # MLPPolicy on (state, goal): fc_mean | fc_log_std
    (832, (64, 16), (256, 256, 256, 1), ("silu",) * 3, False, False),     # MLPQNetwork over 13 * B rows: layer-by-layer route
    (256, (64, 16), (256, 256, 256, 1), ("silu",) * 3, False, False),     # 16 row blocks: partial slabs + reduction
    (64, (64, 16), (256, 256, 256, 1), ("silu",) * 3, False, True),       # actor loss through Q: input gradient only
    (3, (32,), (64, 64, 32), ("relu", "relu"), False, False),
    (100, (128, 32), (256, 64), ("silu",), True, False),                  # 2 row blocks, ragged
    (64, (32, 32), (256, 256, 256, 14), ("silu",) * 3, 3, False),         # discrete-gripper MLPPolicy: mean | log_std | gripper
    (40, (64,), (256, 14), ("silu",), 3, False),                          # three segments through the partial-slab reduction
])
def test_fused_mlp_chain_vs_fp64(rows, ins, widths, acts, two_seg, detach):
    """ops.mlp_chain (one launch forward, one backward) against a plain fp64 evaluation: values, input gradients, every
    weight / bias gradient (incl. the per-row-block partial reduction and the two-segment last layer)."""
    from tacorl_b200 import ops
    ops.set_precision("fp32")
    g = torch.Generator().manual_seed(rows + sum(widths))
    xs = [torch.randn(rows, i, generator=g) for i in ins]
    dims = [sum(ins)] + list(widths)
    layers = []
    for l in range(len(widths)):
        bound = 1.0 / dims[l] ** 0.5
        if two_seg and l == len(widths) - 1:
            parts = [6, 6, widths[l] - 12] if two_seg == 3 else [widths[l] // 2, widths[l] - widths[l] // 2]
            layers.append([((torch.rand(h, dims[l], generator=g) * 2 - 1) * bound, torch.randn(h, generator=g) * 0.1)
                           for h in parts])
        else:
            layers.append([((torch.rand(widths[l], dims[l], generator=g) * 2 - 1) * bound, torch.randn(widths[l], generator=g) * 0.1)])
    cot = torch.randn(rows, widths[-1], generator=g)
    # fp64 reference
    x64 = [x.double().requires_grad_(True) for x in xs]
    l64 = [[(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in sg] for sg in layers]
    h = torch.cat(x64, dim=-1)
    for l, sg in enumerate(l64):
        W = torch.cat([w for w, _ in sg], 0)
        b = torch.cat([bb for _, bb in sg], 0)
        h = torch.nn.functional.linear(h, W, b)
        if l < len(acts):
            h = torch.relu(h) if acts[l] == "relu" else torch.nn.functional.silu(h)
    (h * cot.double()).sum().backward()
    # CUDA
    xd = [x.to(DEV).requires_grad_(True) for x in xs]
    ld = [[(W.to(DEV).requires_grad_(not detach), b.to(DEV).requires_grad_(not detach)) for W, b in sg] for sg in layers]
    out = ops.mlp_chain(xd if len(xd) > 1 else xd[0], [sg if len(sg) > 1 else sg[0] for sg in ld], acts)
    (out * cot.to(DEV)).sum().backward()
    assert_close("out", out, h, 2e-6)
    for a, b in zip(xd, x64):
        assert_close("dx", a.grad, b.grad, 1e-5)
    for sgd, sg64 in zip(ld, l64):
        for (W, b), (W64, b64) in zip(sgd, sg64):
            if detach:
                assert W.grad is None and b.grad is None
            else:
                assert_close("dW", W.grad, W64.grad, 1e-5)
                assert_close("db", b.grad, b64.grad, 1e-5)
