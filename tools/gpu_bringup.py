"""GPU bring-up: kernel-by-kernel checks with diagnostics (run under gpurun)."""
import sys, time, traceback
import numpy as np, torch
sys.path.insert(0, ".")
from mapf_gpt_b200 import engine as E, maps, weights as W
import oracle
from oracle import gpt_oracle as G

torch.manual_seed(0)
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)

def gemm_case(M, N, K, variant, tag):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    try:
        Cc = E.test_gemm(A, B, variant)
        torch.cuda.synchronize()
    except Exception as ex:
        print(f"[gemm {tag}] M{M} N{N} K{K} v{variant:#x}: EXC {ex}", flush=True); return False
    ref = A.float() @ B.float().t()
    err = (Cc - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    ok = rel < 2e-3
    print(f"[gemm {tag}] M{M} N{N} K{K} v{variant:#x}: max_abs_err {err:.4g} rel {rel:.3g} {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        bad = ((Cc - ref).abs() > 1e-2 * ref.abs().max()).float()
        print("   bad frac", bad.mean().item(), "rows bad(first 16 of 128):", bad[:128].mean(1)[:16].tolist(),
              "cols bad(first 16):", bad.mean(0)[:16].tolist(), flush=True)
        print("   C[0,:8]", Cc[0,:8].tolist(), "ref", ref[0,:8].tolist(), flush=True)
    return ok

ok_all = True
for (cfg, N, K) in [(0, 160, 32), (0, 160, 160), (0, 480, 160), (0, 160, 640), (1, 256, 64), (1, 768, 256), (1, 256, 1024), (2, 128, 64), (2, 384, 192)]:
    ok = gemm_case(256, N, K, cfg, "base")
    if not ok:
        gemm_case(256, N, K, cfg | 0x100, "swap-lbo-sbo")
    ok_all &= ok
gemm_case(128 * 300, 480, 160, 0, "big")

def attn_case(n_seq, n_head, hs, scale=1.0):
    q = (torch.randn(n_seq, n_head, 256, hs, device=dev) * scale).bfloat16()
    k = (torch.randn(n_seq, n_head, 256, hs, device=dev) * scale).bfloat16()
    v = torch.randn(n_seq, n_head, 256, hs, device=dev).bfloat16()
    try:
        o = E.test_attention(q, k, v); torch.cuda.synchronize()
    except Exception as ex:
        print(f"[attn] hs{hs}: EXC {ex}", flush=True); return False
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    err = (o.float() - ref).abs().max().item()
    ok = err < 3e-2
    print(f"[attn] seq{n_seq} head{n_head} hs{hs} scale{scale}: max_abs_err {err:.4g} {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        d = (o.float() - ref).abs()
        print("   err by q-tile:", d[:, :, :128].max().item(), d[:, :, 128:].max().item(), " by d(first 8):", d.amax((0,1,2))[:8].tolist(), flush=True)
        print("   o[0,0,0,:8]", o[0,0,0,:8].float().tolist(), "ref", ref[0,0,0,:8].tolist(), flush=True)
    return ok

for hs in (32, 64):
    ok_all &= attn_case(2, 3, hs, 1.0)
    attn_case(3, 5, hs, 0.3)

# ---- env kernels vs oracle
def env_case(name, n, E_envs, steps=12, seed=3):
    m = maps.load_map(name)
    grid = m["grid"]; H, Wd = grid.shape
    eng = E.RolloutEngine(E_envs, n, H, Wd)
    st = []; gl = []
    for e in range(E_envs):
        s, g = maps.sample_instance(m, n, seed, e); st.append(s); gl.append(g)
    st = np.stack(st); gl = np.stack(gl)
    eng.reset(0, grid, st, gl)
    orc = []
    for e in range(E_envs):
        o = oracle.ObsOracle(grid); o.create_agents(st[e], gl[e]); orc.append(o)
    # BFS parity
    bad = 0
    for e in range(min(E_envs, 2)):
        for a in range(0, n, max(1, n // 4)):
            _, f = orc[e].partial(a)
            bad += int((eng.cost2go(e, a) != f).sum())
    rng = np.random.default_rng(seed)
    pos = st.copy(); act = np.full((E_envs, n), -1, np.int32)
    tok_bad = 0; pos_bad = 0
    for t in range(steps):
        eng.update_agents(pos if t == 0 else None, None, act)
        toks = eng.generate_observations()
        for e in range(E_envs):
            orc[e].update_agents(pos[e], gl[e], act[e])
            tok_bad += int((toks[e, :n].astype(np.int32) != orc[e].generate_observations()).sum())
        act = rng.integers(0, 5, (E_envs, n)).astype(np.int32)
        mv = np.where(rng.random((E_envs, n)) < 0.85, act, rng.integers(0, 5, (E_envs, n))).astype(np.int32)
        newpos = eng.env_step(mv)
        for e in range(E_envs):
            p2, _ = oracle.pogema_step_soft(grid, pos[e], mv[e])
            pos_bad += int((p2 != newpos[e, :n]).sum())
            pos[e] = p2
    print(f"[env] {name} n{n} E{E_envs}: bfs_bad {bad} tok_bad {tok_bad} pos_bad {pos_bad}", flush=True)
    met = eng.metrics()
    print("   metrics[0]", met[0].tolist(), flush=True)
    eng.close()
    return bad == 0 and tok_bad == 0 and pos_bad == 0

try:
    ok_all &= env_case("validation-random-seed-000", 32, 2)
    ok_all &= env_case("validation-mazes-seed-000", 64, 3)
    ok_all &= env_case("wfi_warehouse", 192, 2)
    ok_all &= env_case("Berlin_1_256_03", 256, 2)
    ok_all &= env_case("puzzle-00", 4, 5)
except Exception:
    traceback.print_exc(); ok_all = False

# ---- full forward vs torch fp32 oracle
def fwd_case(model, n_rows=6, scale=3.0):
    cfg = W.model_config(model)
    sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), scale)
    eng = E.RolloutEngine(1, 8, 31, 31)
    t0 = time.time(); eng.load_model(sd, cfg); t1 = time.time()
    rng = np.random.default_rng(5)
    toks = rng.integers(0, 67, (n_rows, 256)).astype(np.int8)
    lg = eng.forward_tokens(toks)
    sdd = {k: v.to(dev) for k, v in sd.items()}
    ref = G.forward_logits(sdd, cfg.n_layer, cfg.n_head, torch.from_numpy(toks.astype(np.int64)).to(dev))[:, :5].cpu().numpy()
    err = np.abs(lg - ref).max()
    print(f"[fwd] {model} rows{n_rows}: load {t1-t0:.1f}s max_abs_err {err:.4g} (ref absmax {np.abs(ref).max():.3g})", flush=True)
    print("   mine", lg[0].tolist(), "\n   ref ", ref[0].tolist(), flush=True)
    eng.close()
    return err < 5e-2

for mname in ("2M", "6M", "85M"):
    try:
        ok_all &= fwd_case(mname)
    except Exception:
        traceback.print_exc(); ok_all = False
print("BRINGUP", "ALL OK" if ok_all else "HAS FAILURES", flush=True)
