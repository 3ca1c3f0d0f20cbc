"""GPU parity tests, API level: the drop-in PPO / AsyncTools classes against the golden fixtures produced by the REAL
reference (tests/golden/gen_golden.py ran /root/reference's PPO.learn / AsyncPPO.worker / utils on the same inputs).

Tolerances: rollouts, masks, buffers and GAE returns bit-exact; advantages, losses and post-update weights 1e-5
relative (float32), as BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch as t

pytestmark = pytest.mark.gpu

from oracle import cref, ppo as oppo  # noqa: E402

ENVS = {"cartpole": "CartPole-v1", "pendulum": "Pendulum-v1", "acrobot": "Acrobot-v1", "mountaincar": "MountainCar-v0",
        "mountaincarcont": "MountainCarContinuous-v0"}


@pytest.fixture(scope="module")
def api():
    if not t.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import AsyncTools
    import PPO as ppo_pkg
    import prl_b200

    return dict(AsyncTools=AsyncTools, PPO=ppo_pkg, prl=prl_b200)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize])


def close_weights(got, want, frac=1.0, dmax=1e-6):
    """Post-update weights: 1e-5 relative (+2e-6 absolute) - the north_star bar - on a fraction `frac` of the components and
    `dmax` absolute on all of them.  The defaults (every component, 1e-6) are what the tensor-core path achieves on every
    fixture (measured: max |diff| 1.2e-7 .. 3.6e-7, printed by the tests).  The continuous fixture (12 AdamW steps on the
    fp32-FMA path) is checked at frac = 0.9995, dmax = 1e-5 (achieved: 99.99 %, 6.9e-6): that escape hatch is AdamW's
    conditioning, not slack - a component whose gradient is near zero moves by lr * m / sqrt(v) with m / sqrt(v) decided by
    rounding noise, and the REFERENCE's own float32 result sits 8e-6 away from the float64 evaluation of the same update
    there (tests/test_oracle.py::test_reference_float32_update_is_conditioned_at_1e5)."""
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    ok = d <= 1e-5 * np.abs(want) + 2e-6
    assert ok.mean() >= frac, (ok.mean(), d.max())
    assert d.max() <= dmax, d.max()
    return float(ok.mean()), float(d.max()), float((d / (np.abs(want) + 1e-3)).max())


def make_ppo(api, g, **kw):
    cont = bool(g["is_continuous"])
    p = api["PPO"].PPO(is_continuous=cont, observ_dim=int(g["O"]), action_dim=int(g["A"]), action_scaling=2.0 if cont else None,
                       lr=float(g["lr"]), k_epochs=int(g["k_epochs"]), batch_size=int(g["batch_size"]),
                       mini_batch_size=int(g["mini_batch_size"]), use_RND=bool(g["use_rnd"]), beta=float(g["beta"]),
                       gamma=float(g["gamma"]), GAE_lambda=float(g["GAE_lambda"]), policy_clip=float(g["policy_clip"]), **kw)
    p.show_progress = False
    sd = {k[len("init."):]: t.from_numpy(g[k]) for k in g.files if k.startswith("init.")}
    p.policy.load_state_dict(sd)
    p.policy_old.load_state_dict(p.policy.state_dict())
    assert np.array_equal(p.policy.flat.cpu().numpy(), g["init_flat"])  # flat layout == parameters() order
    if bool(g["use_rnd"]):
        p.rnd.load_state_dict({k[len("rnd_init."):]: t.from_numpy(g[k]) for k in g.files if k.startswith("rnd_init.")})
    return p


@pytest.mark.parametrize("name,roll,path", [("discrete", "cartpole", "tensor"), ("continuous", "pendulum", "tensor"), ("rnd", "acrobot", "tensor"),
                                            ("discrete_1step", "cartpole", "tensor"), ("continuous_1step", "pendulum", "tensor"),
                                            ("continuous", "pendulum", "fp32"), ("continuous_1step", "pendulum", "fp32"), ("discrete", "cartpole", "fp32")])
def test_learn_matches_reference_post_update_weights(api, golden, name, roll, path, capsys):
    """PPO.learn() on the reference's memory contents -> the reference's post-update weights and AdamW moments, through the
    tensor-core update (the default for every BASELINE policy) and through the fp32-FMA one."""
    g, r = golden("learn_" + name), golden("rollout_" + roll)
    ppo = make_ppo(api, g)
    assert ppo.update_path == "tensor"
    ppo.update_path = path
    for i in range(len(r["states"])):  # fill PPO.memory the way buffer_to_target_buffer_transfer does (float32 items)
        ppo.memory.states.append(r["states"][i].copy())
        ppo.memory.actions.append(r["actions"][i].copy())
        ppo.memory.rewards.append(np.float32(r["rewards"][i]))
        ppo.memory.dones.append(np.float32(r["dones"][i]))
    ppo.learn()
    assert len(ppo.memory.states) == 0
    got = ppo.policy.flat.cpu().numpy()
    tol = dict(frac=0.9995, dmax=1e-5) if name == "continuous" else {}
    fr, dmax, rmax = close_weights(got, g["post_flat"], **tol)
    with capsys.disabled():
        print(f"\n[parity] learn_{name} ({ppo.update_path} path, {ppo.optimizer.step_count} optimiser steps) post-update weights vs the REAL "
              f"reference: {100 * fr:.2f} % of components within 1e-5*|w| + 2e-6, max |diff| {dmax:.2e}, max |diff| / (|w| + 1e-3) {rmax:.2e}")
    assert np.array_equal(ppo.policy_old.flat.cpu().numpy(), got)
    for k in g.files:  # through the state_dict keys as well (checkpoint layout)
        if k.startswith("post."):
            close_weights(ppo.policy.state_dict()[k[5:]].cpu().numpy().ravel(), g[k].ravel(), frac=0.99 if name == "continuous" else 1.0, dmax=tol.get("dmax", 1e-6))
    # AdamW moments: sums of signed gradients, so components near zero carry cancellation noise - 1e-5 of the largest
    for got_m, want_m in ((ppo.optimizer.exp_avg, g["post_exp_avg"]), (ppo.optimizer.exp_avg_sq, g["post_exp_avg_sq"])):
        np.testing.assert_allclose(got_m.cpu().numpy(), want_m, rtol=1e-4, atol=1e-5 * np.abs(want_m).max())
    if bool(g["use_rnd"]):
        for k in g.files:
            if k.startswith("rnd_post."):
                close_weights(ppo.rnd.state_dict()[k[9:]].cpu().numpy().ravel(), g[k].ravel())
    # loss of the first minibatch step vs the oracle's restatement of the reference loss
    l = ppo.last_losses[0].cpu().numpy()
    mb = int(g["mini_batch_size"])
    p0 = oppo.unflatten(g["init_flat"], bool(g["is_continuous"]), int(g["O"]), int(g["A"]))
    sl = slice(0, mb)
    want = float(oppo.ppo_loss(p0, bool(g["is_continuous"]), t.from_numpy(r["states"][sl]), t.from_numpy(r["actions"][sl]),
                               t.from_numpy(g["eval_logp"][sl]), t.from_numpy(g["advantages"][sl]), t.from_numpy(g["gae_returns"][sl]),
                               float(g["policy_clip"])))
    if not bool(g["use_rnd"]):  # with RND the golden advantages were recorded after reward mixing; loss checked above via weights
        assert (l[0] + 0.5 * l[1] - 0.01 * l[2]) / l[3] == pytest.approx(want, rel=1e-5, abs=1e-6)


def test_learn_large_minibatch_matches_oracle_learn(api, capsys):
    """PPO.learn() at BASELINE's minibatch size: >= 2^17 rows from a sampled CartPole rollout, mini_batch_size = 65 536
    (two full minibatches of 3.5 tiles per CTA + a partial one), 2 epochs = 6+ optimiser steps through the fused tcgen05
    step kernel - against oracle.ppo.learn on the same memory contents, in float32 (what the reference runs) and in
    float64 (the truth both float32 runs are judged by)."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    t.manual_seed(11)
    ppo = P.PPO(is_continuous=False, observ_dim=4, action_dim=2, k_epochs=2, batch_size=1024, mini_batch_size=65536, lr=1e-3)
    ppo.show_progress = False
    assert ppo.update_path == "tensor"
    ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1", max_episode_steps=64), ppo=ppo, num_envs=8192, steps=1)
    ap.worker()
    s, a, r, dn = ppo.memory.device_view(4, 1, ppo.device)
    N = s.shape[0]
    assert N >= 1 << 17, N
    mem = dict(states=s.cpu().numpy(), actions=a.cpu().numpy()[:, 0], rewards=r.cpu().numpy(), dones=dn.cpu().numpy())
    init = ppo.policy.flat.cpu().numpy().copy()
    ppo.learn()
    got = ppo.policy.flat.cpu().numpy()
    kw = dict(lr=1e-3, k_epochs=2, policy_clip=0.2, gae_lambda=0.95, gamma=0.995, mini_batch_size=65536)
    p32 = oppo.unflatten(init, False, 4, 2)
    oppo.learn(p32, False, mem, **kw)
    p64 = {k: v.double() for k, v in oppo.unflatten(init, False, 4, 2).items()}
    oppo.learn(p64, False, mem, dtype=t.float64, **kw)
    w32 = oppo.flatten(p32, False).numpy(); w64 = oppo.flatten(p64, False).numpy()
    fr, dmax, rmax = close_weights(got, w32)
    d64 = np.abs(got - w64).max(); o64 = np.abs(w32.astype(np.float64) - w64).max()
    with capsys.disabled():
        print(f"\n[parity] learn() N={N}, mini_batch 65 536, {ppo.optimizer.step_count} optimiser steps: vs oracle float32 {100 * fr:.2f} % within "
              f"1e-5*|w| + 2e-6, max |diff| {dmax:.2e}; vs oracle float64 max |diff| {d64:.2e} (oracle float32 itself: {o64:.2e})")
    assert d64 <= 1e-6, d64   # achieved 2.8e-7 - the same distance the float32 oracle keeps from the float64 one
    assert not np.array_equal(got, init)


def test_learn_with_cuda_graph_epochs_is_bit_identical(api):
    """use_cuda_graph replays one captured epoch k_epochs times (device-resident AdamW step counter): same launches,
    same order, so the weights must be bit-identical to the launch-by-launch path."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    out = []
    for graph in (False, True):
        t.manual_seed(3)
        ppo = P.PPO(is_continuous=False, observ_dim=4, action_dim=2, k_epochs=5, batch_size=256, mini_batch_size=512)
        ppo.show_progress = False
        ppo.use_cuda_graph = graph
        ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1", max_episode_steps=64), ppo=ppo, num_envs=256, steps=1)
        for _ in range(4):   # later rounds reuse the captured epoch when the full minibatches are the same launches (the last,
            ap.worker()      # partial one is launched eagerly) or capture a new one when their number changed
            assert len(ppo.memory.states) >= 4 * 512   # >= 4 minibatches x 5 epochs = 20 optimiser steps
            ppo.learn()
        if graph:
            assert 1 <= len(ppo._graph_cache) <= 4
        out.append((ppo.policy.flat.cpu().numpy(), ppo.optimizer.exp_avg_sq.cpu().numpy(), ppo.optimizer.step_count,
                    int(ppo.optimizer.step_dev[0].item())))
    assert np.array_equal(bits(out[0][0]), bits(out[1][0])) and np.array_equal(bits(out[0][1]), bits(out[1][1]))
    assert out[0][2] == out[1][2] == out[0][3] == out[1][3] and out[0][2] >= 80


def test_fused_optimizer_step_matches_separate_kernels(api):
    """prl_ppo_step_tc (gradient + fixed-order reduction + clip + AdamW in one cooperative launch) against the three-kernel
    sequence: same reduction order, so the gradient is bit-identical; the weights agree to float32 rounding of the norm."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    out = []
    for fused in (False, True):
        t.manual_seed(5)
        ppo = P.PPO(is_continuous=False, observ_dim=4, action_dim=2, k_epochs=3, batch_size=256, mini_batch_size=1000)
        ppo.show_progress = False
        ppo.fused_optimizer = fused
        ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1", max_episode_steps=64), ppo=ppo, num_envs=256, steps=1)
        ap.worker()
        ppo.learn()
        out.append((ppo.policy.flat.cpu().numpy(), ppo._grad.cpu().numpy(), ppo.optimizer.exp_avg.cpu().numpy(), ppo.optimizer.step_count,
                    int(ppo.optimizer.step_dev[0].item()), ppo.last_losses.cpu().numpy()))
    a, b = out
    assert a[3] == b[3] == a[4] == b[4] and a[3] >= 6
    np.testing.assert_allclose(b[0], a[0], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(b[2], a[2], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(b[5], a[5], rtol=1e-6)


def test_learn_returns_early_below_batch_size(api, golden):
    g = golden("learn_discrete")
    ppo = make_ppo(api, g)
    ppo.batch_size = 1000
    for i in range(10):
        ppo.memory.push(np.zeros(4), np.zeros(1), np.ones(1), np.bool_(False))
    before = ppo.policy.flat.clone()
    ppo.learn()
    assert len(ppo.memory.states) == 10 and t.equal(before, ppo.policy.flat)  # PPO.py:123-124: nothing happens, memory kept


def test_compute_gae_api_matches_reference(api, golden):
    g = golden("learn_discrete")
    ppo = make_ppo(api, g)
    out = ppo.compute_gae(g["gae_rewards"], g["gae_dones"], g["gae_values"], g["gae_next_value"])
    assert isinstance(out, list) and len(out) == len(g["gae_returns"])
    assert np.array_equal(bits(np.array(out, np.float32)), bits(g["gae_returns"]))
    # a buffer that does NOT end with done=1 uses the bootstrap value (PPO.py:113)
    r = np.ones(5, np.float32); d = np.zeros(5, np.float32); v = np.linspace(0, 1, 5).astype(np.float32)
    want = cref.gae(r, d, v, 0.7, ppo.gamma, ppo.GAE_lambda)
    assert np.array_equal(bits(np.array(ppo.compute_gae(r, d, v, np.float32(0.7)), np.float32)), bits(want))


@pytest.mark.parametrize("key", list(ENVS))
def test_stepwise_api_reproduces_reference_worker_trace(api, golden, key):
    """EnvVectorizer.step + utils.* driven by the reference's taped actions: every per-step array the reference saw
    (compact obs, fp64 rewards, dones, truncates, the mask) and the final flat buffer, bit for bit."""
    A, prl = api["AsyncTools"], api["prl"]
    utils = A.utils
    g = golden("rollout_" + key)
    E, T = int(g["E"]), int(g["T"])
    env = A.AsyncPPO.EnvVectorizer(prl.make(ENVS[key], max_episode_steps=T), E)
    buf = A.AsyncPPO.VecMemory(E)
    mem = api["PPO"].Memory()
    states = env.reset_to(g["init_state"])[0]
    off = 0
    for step in range(int(g["nsteps"])):
        mask = env.envs_active
        assert np.array_equal(mask, g["mask_before"][step])
        n = int(g["step_counts"][step])
        assert utils.number_of_active_environments(mask) == n
        assert np.array_equal(utils.indexes_of_active_environments(E, mask), np.arange(E)[~mask])
        assert np.array_equal(bits(states), bits(g["seen_states"][off:off + n]))
        actions = g["tape"][step][~mask]
        nxt, rew, dones, truncs, infos = env.step(actions)
        assert rew.dtype == np.float64 and dones.dtype == np.bool_ and nxt.dtype == np.float32 and len(infos) == n
        assert np.array_equal(bits(nxt), bits(g["step_obs"][off:off + n]))
        assert np.array_equal(bits(rew), bits(g["step_rewards"][off:off + n]))
        assert np.array_equal(dones, g["step_dones"][off:off + n]) and np.array_equal(truncs, g["step_truncs"][off:off + n])
        fin = dones | truncs
        utils.buffer_append(buf, states, actions, rew, fin, mask, E)
        states = utils.inactive_states_dropout(nxt, fin)
        env.envs_active = utils.update_active_environments_list(mask, fin)
        off += n
    assert np.all(env.envs_active)
    # the list-of-lists view of the device VecMemory, then the env-major transfer
    lens = [len(x) for x in buf.rewards]
    assert sum(lens) == len(g["rewards"])
    utils.buffer_to_target_buffer_transfer(buf, mem)
    assert all(len(x) == 0 for x in buf.rewards)
    assert len(mem.states) == len(g["states"])
    for name in ("states", "actions", "rewards", "dones"):
        got = np.array(list(getattr(mem, name)), np.float32)
        assert np.array_equal(bits(got), bits(g[name])), name
    assert np.array_equal(bits(env.sim.get_state().cpu().numpy()[:, : g["final_state"].shape[1]]), bits(g["final_state"]))


@pytest.mark.parametrize("key,cont", [("cartpole", False), ("pendulum", True), ("acrobot", False), ("mountaincar", False), ("mountaincarcont", True)])
def test_fused_worker_equals_stepwise_worker(api, key, cont):
    """AsyncPPO.worker(): the one-launch fused rollout and the step-by-step loop (PPO.get_action -> EnvVectorizer.step ->
    utils.*) draw the same Philox numbers and must fill ppo.memory identically (bit-exact), sampled actions included."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    E, T = 96, 48
    d = prl.make(ENVS[key], max_episode_steps=T)
    out = []
    for fused in (True, False):
        t.manual_seed(7)
        ppo = P.PPO(is_continuous=cont, observ_dim=d.observ_dim, action_dim=d.action_dim, action_scaling=2.0 if cont else None)
        ap = A.AsyncPPO.AsyncPPO(env=d, ppo=ppo, num_envs=E, steps=1)
        ap.fused = fused
        ap.worker()
        ap.worker()  # second episode appends behind the first (memory is only cleared by learn())
        s, a, r, dn = ppo.memory.device_view(d.observ_dim, d.action_dim if cont else 1, ppo.device)
        out.append((s.cpu().numpy(), a.cpu().numpy(), r.cpu().numpy(), dn.cpu().numpy(), int(ap.step_score), float(ap.reward_score)))
    f, s = out
    assert f[4] == s[4] == len(f[2]) and f[4] >= 2 * E
    for i in range(4):
        assert np.array_equal(bits(f[i]), bits(s[i])), i
    assert f[5] == pytest.approx(s[5], rel=1e-5)
    assert (f[3].sum() == 2 * E)  # every episode ends with done = 1 (termination or truncation)


@pytest.mark.parametrize("key,rnd,envs", [("cartpole", False, 300), ("acrobot", False, 300), ("acrobot", True, 300), ("mountaincar", False, 300),
                                          ("pendulum", False, 300), ("cartpole", False, 37)])   # 37 envs: the unaligned transfer / GAE paths
def test_fused_worker_by_products_equal_the_separate_passes(api, key, rnd, envs):
    """The fused worker records, per transition, the acting policy's log-prob and state value (prl_rollout_eval) and - without
    RND - the GAE returns (prl_gae_columns on the time-major planes).  They must be the bits PPO.learn's own passes produce on
    the transferred rows (PPO.py:134-154 old-policy evaluation, :107-120 compute_gae), and learn() must end with the same
    weights whether it consumes them or recomputes them.  Also covers rows accumulated over two worker() calls."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    from prl_b200 import ops

    def run(fuse):
        t.manual_seed(11)
        env = prl.make(ENVS[key], max_episode_steps=48)
        ppo = P.PPO(is_continuous=env.is_continuous, observ_dim=env.observ_dim, action_dim=env.action_dim, k_epochs=2, batch_size=64,
                    mini_batch_size=512, use_RND=rnd, beta=0.01, action_scaling=2.0 if env.is_continuous else None)
        ppo.show_progress = False
        ppo.fuse_evaluation = fuse
        ap = A.AsyncPPO.AsyncPPO(env=env, ppo=ppo, num_envs=envs, steps=1)
        ap.worker()
        ap.worker()   # rows of a second call land behind the first one's
        return ppo, ap

    ppo, ap = run(True)
    m = ppo.memory
    N = m._dev_count
    pre = m.evaluated(N, ppo._eval_tag())
    assert pre is not None and N >= 2 * envs
    states, actions, rewards, dones = m.device_view(ppo.observ_dim, 1, ppo.device)
    logp, value, _ = ops.policy_evaluate(ppo.policy_old.flat, ppo.is_continuous, ppo.observ_dim, ppo.action_dim, states, actions)
    assert np.array_equal(bits(pre[0].cpu().numpy()), bits(logp.cpu().numpy()))
    assert np.array_equal(bits(pre[1].cpu().numpy()), bits(value.cpu().numpy()))
    if rnd:
        assert pre[2] is None     # the intrinsic reward is only known in learn(): GAE stays there
    else:
        ret = ops.gae(rewards, dones, value, ppo.gamma, ppo.GAE_lambda)
        assert np.array_equal(bits(pre[2].cpu().numpy()), bits(ret.cpu().numpy()))
        want = cref.gae(rewards.cpu().numpy(), dones.cpu().numpy(), value.cpu().numpy(), value.cpu().numpy()[-1], ppo.gamma, ppo.GAE_lambda)
        assert np.array_equal(bits(pre[2].cpu().numpy()), bits(want))
    # a policy_old that changed after the rollout invalidates them (learn() then evaluates again, like the reference)
    ppo.policy_old.flat.mul_(1.0)
    assert m.evaluated(N, ppo._eval_tag()) is None
    # same final weights with and without the fused by-products
    a, _ = run(True)
    b, _ = run(False)
    assert a.memory.evaluated(a.memory._dev_count, a._eval_tag()) is not None and b.memory.evaluated(b.memory._dev_count, b._eval_tag()) is None
    calls0 = dict(ops._lib.CALL_COUNTS)
    a.learn()
    assert ops._lib.CALL_COUNTS.get("prl_policy_evaluate", 0) == calls0.get("prl_policy_evaluate", 0)   # the pass is really skipped
    b.learn()
    assert np.array_equal(bits(a.policy.flat.cpu().numpy()), bits(b.policy.flat.cpu().numpy()))
    if rnd:
        assert np.array_equal(bits(a.rnd.pred_flat.cpu().numpy()), bits(b.rnd.pred_flat.cpu().numpy()))


def test_auto_reset_worker_opt_in(api):
    """AsyncPPO.auto_reset (opt-in; the default stays the reference's drop-out worker): exactly num_envs x rollout_steps transitions
    per worker(), several episodes per env column, by-products (log-prob, value, GAE returns over columns with dones in the
    middle) still the bits of the separate passes, learn() consumes them."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    from prl_b200 import ops

    t.manual_seed(5)
    ppo = P.PPO(is_continuous=False, observ_dim=4, action_dim=2, k_epochs=2, batch_size=64, mini_batch_size=4096)
    ppo.show_progress = False
    ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1", max_episode_steps=50), ppo=ppo, num_envs=200, steps=1)
    ap.worker()
    n_ref = len(ppo.memory.states)           # reference semantics: one episode per env, ragged
    assert 200 <= n_ref < 200 * 50
    ppo.memory.clear()
    ap.auto_reset, ap.rollout_steps = True, 80
    ap.step_score = 0
    ap.worker()
    m = ppo.memory
    N = m._dev_count
    assert N == 200 * 80 == int(ap.step_score)
    states, actions, rewards, dones = m.device_view(4, 1, ppo.device)
    d = dones.cpu().numpy().reshape(200, 80)
    assert d[:, -1].all() and d.sum() > 2 * 200          # every column closed; untrained CartPole: several episodes per env
    runs = np.diff(np.concatenate([[-1], np.flatnonzero(d[0])]))   # episode lengths of env 0
    assert runs.max() <= 50
    pre = m.evaluated(N, ppo._eval_tag())
    assert pre is not None
    logp, value, _ = ops.policy_evaluate(ppo.policy_old.flat, False, 4, 2, states, actions)
    ret = ops.gae(rewards, dones, value, ppo.gamma, ppo.GAE_lambda)
    for got, want in zip(pre, (logp, value, ret)):
        assert np.array_equal(bits(got.cpu().numpy()), bits(want.cpu().numpy()))
    before = ppo.policy.flat.clone()
    ppo.learn()
    assert len(m.states) == 0 and not t.equal(before, ppo.policy.flat) and bool(t.isfinite(ppo.policy.flat).all())


@pytest.mark.parametrize("key", ["cartpole", "pendulum", "acrobot", "mountaincar", "mountaincarcont"])
def test_batch1_playback_loop_like_test_py(api, key):
    """The reference's Test.py loop (Test.py:19-33): `state, _ = env.reset()`, batch-1 `ppo.get_action`, `env.step(action)` until
    done | truncate - on the single-env surface of the descriptor gym.make hands out.  Seeded like gymnasium (numpy stream), the
    trace must be the oracle env's, bit for bit, under the same actions."""
    P, prl = api["PPO"], api["prl"]
    from oracle import envs as oenvs

    env = prl.make(ENVS[key], max_episode_steps=60)
    cont = env.is_continuous
    t.manual_seed(2)
    ppo = P.PPO(is_continuous=cont, observ_dim=env.observ_dim, action_dim=env.action_dim, action_scaling=2.0 if cont else None)
    ref = oenvs.make(ENVS[key], max_episode_steps=60)
    state, _ = env.reset(seed=123)
    want, _ = ref.reset(seed=123)
    assert state.dtype == np.float32 and np.array_equal(bits(state), bits(np.asarray(want, np.float32)))
    steps = 0
    while True:
        action = ppo.get_action(t.from_numpy(state).unsqueeze(0))
        assert action.shape == ((1, env.action_dim) if cont else (1,))
        state, reward, done, truncate, _ = env.step(action.squeeze(0))
        w_state, w_reward, w_done, w_trunc, _ = ref.step(action.squeeze(0))
        steps += 1
        assert np.array_equal(bits(state), bits(np.asarray(w_state, np.float32))) and reward == float(w_reward), steps
        assert done == bool(w_done) and truncate == bool(w_trunc), steps
        if done or truncate:
            break
    assert 1 <= steps <= 60
    totals = prl.play(ppo, env, episodes=2, seed=5)
    assert len(totals) == 2 and all(np.isfinite(totals))


def test_reward_score_is_bit_reproducible(api):
    """AsyncPPO.reward_score (AsyncPPO.py:131) of the fused worker: the per-CTA reward sums are added in CTA order by the last CTA to
    finish, not by atomics - the same rollout gives the same bits, and they are the float64 sum of the stored rewards."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    got = []
    for _ in range(3):
        t.manual_seed(21)
        ppo = P.PPO(is_continuous=True, observ_dim=3, action_dim=1, action_scaling=2.0)
        ap = A.AsyncPPO.AsyncPPO(env=prl.make("Pendulum-v1", max_episode_steps=40), ppo=ppo, num_envs=5000, steps=1)
        ap.step_score = 0
        ap.reward_score = 0
        ap.worker()
        rewards = ppo.memory.device_view(3, 1, ppo.device)[2]
        got.append((np.float64(ap.reward_score).tobytes(), int(ap.step_score), float(rewards.double().sum().item())))
    assert got[0][0] == got[1][0] == got[2][0] and got[0][1] == 5000 * 40
    assert np.frombuffer(got[0][0], np.float64)[0] == pytest.approx(got[0][2], rel=1e-6)   # (the stored rewards are float32 copies)


def test_rollout_forms_store_the_same_bits(api, monkeypatch):
    """The fused worker picks how many rows of the forward a thread owns from the env count (csrc/rollout.cu: 1 = 32 envs per CTA
    when there are few envs and the step is latency-bound, 7 = 224-env CTAs where they fill the GPU's CTA slots more evenly, 8 = 256)
    - a row's arithmetic does not depend on that: the same transitions, old log-probs, values and returns, bit for bit."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    got = {}
    for form in (None, "8", "7", "1"):
        if form is None:
            monkeypatch.delenv("PRL_ROLLOUT_RPT", raising=False)
        else:
            monkeypatch.setenv("PRL_ROLLOUT_RPT", form)
        for env_id, cont, O, AD in (("CartPole-v1", False, 4, 2), ("Pendulum-v1", True, 3, 1), ("Acrobot-v1", False, 6, 3)):
            t.manual_seed(33)
            ppo = P.PPO(is_continuous=cont, observ_dim=O, action_dim=AD, action_scaling=2.0 if cont else None)
            ap = A.AsyncPPO.AsyncPPO(env=prl.make(env_id, max_episode_steps=60), ppo=ppo, num_envs=1237, steps=1)
            ap.step_score = 0
            ap.reward_score = 0
            ap.worker()
            n = len(ppo.memory.states)
            rows = [x.clone() for x in ppo.memory.device_view(O, 1, ppo.device)]
            pre = ppo.memory.evaluated(n, ppo._eval_tag())
            assert pre is not None
            got.setdefault(env_id, []).append((n, int(ap.step_score), float(ap.reward_score), rows + [x.clone() for x in pre if x is not None]))
    monkeypatch.delenv("PRL_ROLLOUT_RPT", raising=False)
    for env_id, runs in got.items():
        for other in runs[1:]:
            assert other[:2] == runs[0][:2] and other[2] == pytest.approx(runs[0][2], rel=1e-12)
            assert len(other[3]) == len(runs[0][3])
            for x, y in zip(runs[0][3], other[3]):
                assert t.equal(x, y), env_id


def test_reference_unittest_call_patterns(api):
    """The duck-typing the reference's own unittests rely on."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    ppo = P.PPO(is_continuous=False, observ_dim=4, action_dim=2)
    for _ in range(100):
        ppo.memory.push(state=np.random.randn(4), action=np.random.randint(0, 2, size=(1,)), reward=np.random.rand(1),
                        done=np.random.choice([True, False]))
    assert ppo.get_action(t.from_numpy(np.random.randn(1, 4))).shape == (1,)   # float64 input tensor
    assert len(ppo.batch_packer([t.randn(224, 4), t.randint(0, 2, (224,)), t.rand(224)], batch_size=32)) == 3
    assert len(ppo.batch_packer(t.randn(224, 4), batch_size=32)) == 7
    pc = P.PPO(is_continuous=True, observ_dim=4, action_dim=2, action_scaling=1.0)
    a = pc.get_action(t.randn(1, 4))
    assert a.shape == (1, 2) and a.dtype == np.float32
    for cont in (False, True):
        ac = P.ActorCritic(is_continuous=cont, observ_dim=4, action_dim=2)
        ac.get_dist(state=t.randn(1, 4).cuda()).sample()
        assert ac.get_state_value(state=t.randn(1, 4).cuda()).shape == (1,)
        acts = t.randn(1, 2).cuda() if cont else t.randint(0, 2, size=(1, 2)).cuda()
        lp, v, ent = ac.get_evaluate(states=t.randn(1, 4).cuda(), actions=acts)
        assert lp.shape == (1,) and v.shape == (1,) and ent.dim() == 0
    rnd = P.RND(4, 4, beta=0.001)
    loader = t.utils.data.DataLoader(t.randn(32, 4).cuda(), batch_size=16)
    assert rnd.compute_intrinsic_reward(loader).shape == (32,)
    rnd.update_pred(loader)
    utils = A.utils
    utils.indexes_of_active_environments(4, np.random.choice([False, True], 4))
    utils.range_of_active_environments(np.random.choice([False, True], 4))
    x = np.random.randn(4, 4); dn = np.random.choice([False, True], 4)
    assert np.array_equal(utils.inactive_states_dropout(x, dn), x[~dn])        # float64 rows
    utils.buffer_append(buffer=A.AsyncPPO.VecMemory(num_envs=4), states=np.random.randn(4, 4), actions=np.random.randint(0, 2, size=4),
                        rewards=np.random.randn(4), dones=np.random.choice([False, True], size=4),
                        is_env_terminal=np.random.choice([False, True], size=4), num_envs=4)
    act = np.random.choice([False, True], size=4)
    utils.update_active_environments_list(act, np.random.choice([False, True], size=4 - np.sum(act)))
    buffer = A.AsyncPPO.VecMemory(num_envs=4)
    total = 0
    for i in range(4):
        n = np.random.randint(0, 100); total += n
        buffer.states[i] = [np.random.randn(4) for _ in range(n)]
        buffer.actions[i] = [np.random.randint(0, 2) for _ in range(n)]
        buffer.rewards[i] = [np.random.randn() for _ in range(n)]
        buffer.dones[i] = [np.random.choice([False, True]) for _ in range(n)]
    target = P.Memory()
    utils.buffer_to_target_buffer_transfer(buffer, target_buffer=target)
    assert len(target.states) == total and all(len(x) == 0 for x in buffer.states)
    vm = A.AsyncPPO.VecMemory(num_envs=4)
    vm.push(idx=0, state=np.random.randn(1), action=np.random.randint(0, 2, size=(1,)), reward=np.random.rand(1), done=np.bool_(True))
    assert vm.states[0][0].dtype == np.float32
    vm.clear()
    ev = A.AsyncPPO.EnvVectorizer(env=prl.make("CartPole-v1"), num_envs=4)
    obs, infos = ev.reset()
    assert obs.shape == (4, 4) and obs.dtype == np.float32 and len(infos) == 4
    ev.step(actions=np.random.randint(0, 2, size=4))


def test_reference_unittests_run_unmodified(api, capsys):
    """SURVEY.md section 2, component 12: the reference's own unittests/ (test_PPO.py, test_utils.py, test_AsyncPPO.py),
    byte-identical copies staged at build() time by oracle/stage_reference.py under oracle/_ref/dropin/unittests/, executed as
    they are: each file puts its parent directory (empty) first on sys.path and imports `PPO` / `AsyncTools`, which PYTHONPATH
    resolves to the drop-in packages of the B200 build.  `import gymnasium` is served by tests/stubs/gymnasium (the image has no
    gymnasium wheel) whose make() is prl_b200.make."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.join(root, "oracle", "_ref", "dropin", "unittests")
    pkg = os.path.join(root, "parallel-reinforcement-learning_b200")
    files = sorted(f for f in (os.listdir(staged) if os.path.isdir(staged) else []) if f.startswith("test_") and f.endswith(".py"))
    if not files:
        pytest.skip("the reference's unittests are not staged (build() stages them where /root/reference exists)")
    assert files == ["test_AsyncPPO.py", "test_PPO.py", "test_utils.py"]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([pkg, os.path.join(root, "tests", "stubs")]), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *files], cwd=staged, env=env, capture_output=True,
                       text=True, timeout=900)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    with capsys.disabled():
        print(f"\n[reference unittests, unmodified, against the B200 build] {tail}")
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail and "error" not in tail


def test_memory_del_slice_clears_device_rows(api):
    """`del memory.states[:]` (the reference's Memory.clear idiom, Memory.py:26-30) on rows that live on the device."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    ppo = P.PPO(is_continuous=False, observ_dim=4, action_dim=2)
    ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1", max_episode_steps=16), ppo=ppo, num_envs=8, steps=1)
    ap.worker()
    m = ppo.memory
    n = len(m.states)
    assert n >= 8 and len(m.dones) == n and m.states[0].shape == (4,) and float(m.dones[n - 1]) == 1.0 and m.actions[-1].shape == ()
    del m.states[:]
    assert len(m.states) == 0 and len(m.actions) == n
    del m.actions[:], m.rewards[:], m.dones[:]
    assert len(m.actions) == len(m.rewards) == len(m.dones) == 0 and m._dev_count == 0
    ap.worker()
    assert len(m.states) >= 8
    m.verify_transfers()


def test_async_ppo_run_trains_cartpole(api):
    """train.py's flow (reference train.py:8-36) at the unittest's size: 4 envs, 1000 steps - then a larger batch to see
    the mean episode length move (CartPole reward = episode length)."""
    A, P, prl = api["AsyncTools"], api["PPO"], api["prl"]
    t.manual_seed(0)
    model = P.PPO(is_continuous=False, action_dim=2, observ_dim=4)
    model.show_progress = False
    ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1"), ppo=model, num_envs=4, steps=1000)
    ap.show_progress = False
    ap.run()
    t.manual_seed(1)
    model = P.PPO(is_continuous=False, action_dim=2, observ_dim=4, k_epochs=4, batch_size=1024, mini_batch_size=2048, lr=3e-3)
    model.show_progress = False
    ap = A.AsyncPPO.AsyncPPO(env=prl.make("CartPole-v1"), ppo=model, num_envs=512, steps=1)
    ap.show_progress = False
    first = None
    for it in range(12):
        ap.step_score = 0; ap.reward_score = 0
        ap.worker()
        mean_len = float(ap.step_score) / 512
        first = mean_len if first is None else first
        model.learn()
    assert np.isfinite(model.policy.flat.cpu().numpy()).all()
    assert mean_len > 1.5 * first, (first, mean_len)  # the policy improves


def test_env_vectorizer_seeded_reset_matches_seeded_numpy_envs(api):
    """EnvVectorizer.reset(seed=s): env i starts where a gymnasium CartPole seeded with reset(seed=s+i) starts
    (Generator(PCG64(SeedSequence(s+i))).uniform(-0.05, 0.05, 4), observation = float32(state)); the next unseeded
    reset() - the reference's own call, AsyncPPO.py:48-62 - continues those generators."""
    AsyncTools, prl = api["AsyncTools"], api["prl"]
    E, s = 37, 2024
    vec = AsyncTools.AsyncPPO.EnvVectorizer(prl.make("CartPole-v1"), num_envs=E)
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(s + i))) for i in range(E)]
    obs, infos = vec.reset(seed=s)
    assert len(infos) == E
    assert np.array_equal(bits(obs), bits(np.stack([g.uniform(-0.05, 0.05, size=(4,)) for g in gens]).astype(np.float32)))
    obs, _ = vec.reset()
    assert np.array_equal(bits(obs), bits(np.stack([g.uniform(-0.05, 0.05, size=(4,)) for g in gens]).astype(np.float32)))
    with pytest.raises(ValueError):
        vec.reset(seed=[1, 2, 3])


def test_checkpoint_round_trip_uses_reference_keys(api, tmp_path):
    P = api["PPO"]
    a = P.PPO(is_continuous=False, observ_dim=6, action_dim=3, use_RND=True)
    a.save_weights(str(tmp_path))
    sd = t.load(str(tmp_path / "Policy_weights.pth"), weights_only=True)
    assert list(sd) == oppo.param_keys(False) and tuple(sd["actor.3.weight"].shape) == (3, 64)
    b = P.PPO(is_continuous=False, observ_dim=6, action_dim=3, use_RND=True)
    b.load_weights(str(tmp_path))
    assert t.equal(a.policy.flat, b.policy.flat) and t.equal(b.policy_old.flat, b.policy.flat)
    assert t.equal(a.rnd.pred_flat, b.rnd.pred_flat) and t.equal(a.rnd.target_flat, b.rnd.target_flat)
    b.load_weights(str(tmp_path / "missing"))  # FileNotFoundError is swallowed like the reference (PPO.py:276-277)
