"""GPU parity: K6 DQN TD kernel and DQN / PER-DQN learner updates vs the torch-CPU oracle."""
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle.learners import DQNLearnerOracle
from oracle.nets import DeepQNetworkOracle

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    cfg = dict(distributed_training=False, episode_length=1000, use_grad_clip=False, grad_clip_norm=0.5,
               device="cuda:0", model_dir="/tmp/xb_models", running_steps=100000, parallels=16, learning_rate=1e-4,
               end_factor_lr_decay=0.5, gamma=0.99, sync_frequency=2, start_training=0, training_frequency=1)
    cfg.update(kw)
    return Namespace(**cfg)


@pytest.mark.parametrize("B,A", [(512, 4), (33, 18), (1000, 6)])
def test_dqn_td_kernel(B, A):
    from xuance_b200 import _lib
    rng = np.random.default_rng(B)
    qe = rng.normal(size=(B, A)).astype(np.float32)
    qn = rng.normal(size=(B, A)).astype(np.float32)
    act = rng.integers(0, A, B).astype(np.float32)
    rew = rng.normal(size=B).astype(np.float32)
    ter = (rng.random(B) < 0.3).astype(np.float32)
    qt = torch.tensor(qe, requires_grad=True)
    pred = qt.gather(-1, torch.tensor(act).long().unsqueeze(-1)).squeeze(-1)
    y = torch.tensor(rew) + 0.99 * (1 - torch.tensor(ter)) * torch.tensor(qn).max(-1).values
    loss = nn.functional.mse_loss(pred, y.detach())
    loss.backward()
    dev = torch.device("cuda:0")
    t_in = [torch.tensor(x, device=dev) for x in (qe, qn, act, rew, ter)]
    dq, td, stats = torch.empty((B, A), device=dev), torch.empty(B, device=dev), torch.zeros(4, device=dev)
    scratch = _lib.scratch(dev)
    _lib.call("xb_dqn_td_fwd_bwd", _lib.ptr(t_in[0]), _lib.ptr(t_in[1]), None, *[_lib.ptr(t) for t in t_in[2:]], B, A, B, 0.99, _lib.ptr(dq), _lib.ptr(td),
              _lib.ptr(stats), _lib.ptr(scratch))
    assert np.array_equal(td.cpu().numpy(), (y - pred).detach().numpy())          # same op order: bit-exact
    np.testing.assert_allclose(dq.cpu().numpy(), qt.grad.numpy(), rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(stats[0].item(), loss.item(), rtol=1e-5)
    np.testing.assert_allclose(stats[1].item(), pred.mean().item(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("per,double_q,graph", [(False, False, False), (True, False, False), (False, True, False),
                                                 (True, False, True)])
def test_dqn_learner_matches_oracle(per, double_q, graph):
    from xuance_b200.common import Discrete, BaseCallback
    from xuance_b200.torch.rl_models import Basic_CNN, DeepQNetwork
    from xuance_b200.torch.learners import DQN_Learner, PerDQN_Learner, DDQN_Learner
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    A, B = 5, 64
    om = DeepQNetworkOracle(A)
    rep = Basic_CNN(input_shape=(84, 84, 4), kernels=[8, 4, 3], strides=[4, 2, 1], filters=[32, 64, 64],
                    activation=nn.ReLU, device="cuda:0")
    model = DeepQNetwork(rep, [512], Discrete(A), None, None, nn.ReLU, "cuda:0").to("cuda:0")
    model.load_state_dict(om.state_dict())
    lrn = (DDQN_Learner if double_q else PerDQN_Learner if per else DQN_Learner)(_cfg(use_cuda_graph=graph), model,
                                                                                  BaseCallback())
    orc = DQNLearnerOracle(om, learning_rate=1e-4, sync_frequency=2, end_factor_lr_decay=0.5,
                           total_iters=lrn.total_iters, per=per, double_q=double_q)
    rng = np.random.default_rng(3)
    for it in range(4):     # crosses two target syncs
        s = {"obs": rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8),
             "actions": rng.integers(0, A, size=B).astype(np.float32),
             "obs_next": rng.integers(0, 256, size=(B, 84, 84, 4), dtype=np.uint8),
             "rewards": rng.normal(size=B).astype(np.float32), "terminals": (rng.random(B) < 0.2).astype(np.float32)}
        ro = orc.update(**s)
        sd = dict(s, obs=torch.from_numpy(s["obs"]).cuda(), obs_next=torch.from_numpy(s["obs_next"]).cuda())
        rp = lrn.update(**sd)
        if per:
            (td_o, info_o), (td_p, info_p) = ro, rp
            np.testing.assert_allclose(td_p.cpu().numpy(), td_o, rtol=2e-3, atol=2e-5)
        else:
            info_o, info_p = ro, rp
        np.testing.assert_allclose(info_p["Qloss"], info_o["Qloss"], rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(info_p["predictQ"], info_o["predictQ"], rtol=2e-4, atol=1e-6)
        assert info_p["learning_rate"] == info_o["learning_rate"]
    so, sp = om.state_dict(), model.state_dict()
    for k in so:
        np.testing.assert_allclose(sp[k].cpu().numpy(), so[k].numpy(), rtol=1e-3, atol=1e-4, err_msg=k)


def test_dueling_dqn_update_matches_autograd():
    """DuelDQN (q_head.py:42-80, dueldqn_learner.py): the dueling network through the K6 TD path vs the same update written
    with torch autograd + torch.optim.Adam on a float32 copy of the network."""
    import copy
    from xuance_b200.common import Discrete, Box, BaseCallback
    from xuance_b200.torch.rl_models import Basic_MLP, DuelingDeepQNetwork
    from xuance_b200.torch.learners import DuelDQN_Learner
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(4)
    A, B, D = 6, 128, 17
    rep = Basic_MLP((D,), [64], None, nn.init.orthogonal_, nn.ReLU, "cuda:0")
    model = DuelingDeepQNetwork(rep, [64], Discrete(A), None, nn.init.orthogonal_, nn.ReLU, "cuda:0").to("cuda:0")
    assert model.eval_Q_head.v_model[0].out_features == 32 and model.eval_Q_head.a_model[-1].out_features == A
    ref = copy.deepcopy(model)
    lrn = DuelDQN_Learner(_cfg(), model, BaseCallback())
    opt = torch.optim.Adam(list(ref.representation.parameters()) + list(ref.eval_Q_head.parameters()), 1e-4, eps=1e-5)
    sch = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.5, total_iters=lrn.total_iters)
    rng = np.random.default_rng(5)
    for it in range(4):
        s = {"obs": rng.normal(size=(B, D)).astype(np.float32), "actions": rng.integers(0, A, size=B).astype(np.float32),
             "obs_next": rng.normal(size=(B, D)).astype(np.float32), "rewards": rng.normal(size=B).astype(np.float32),
             "terminals": (rng.random(B) < 0.2).astype(np.float32)}
        t = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
        info = lrn.update(**t)
        q = ref(t["obs"]).values
        pred = q.gather(-1, t["actions"].long().unsqueeze(-1)).squeeze(-1)
        with torch.no_grad():
            y = t["rewards"] + 0.99 * (1 - t["terminals"]) * ref.target(t["obs_next"]).values.max(-1).values
        loss = nn.functional.mse_loss(pred, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sch.step()
        if (it + 1) % 2 == 0:
            ref.copy_target()
        np.testing.assert_allclose(info["Qloss"], loss.item(), rtol=2e-4, atol=1e-6)
    for (k, p), (_, q) in zip(model.state_dict().items(), ref.state_dict().items()):
        np.testing.assert_allclose(p.cpu().numpy(), q.cpu().numpy(), rtol=1e-3, atol=1e-5, err_msg=k)
