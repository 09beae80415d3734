"""GPU half of the drop-in proof (tests/test_dropin_reference.py is the other half): the restatement of the reference's
``OnPolicyAgent.train_epochs`` (oracle/agents.py, pinned call-for-call to the live reference in the build container) drives
the REAL xuance_b200 rollout buffer and PPO learner through their reference-shaped surface - ``memory.sample(indexes)``
with a NumPy index array, ``learner.update(**samples)`` - and must end in the same parameters as the product agent's own
``train_epochs`` (device-side index upload, fused gather, one synchronisation) from the same seed."""
import numpy as np
import pytest
import torch

from helpers import synth_rollout, fill_buffers, build_product_ppo_model, ppo_config
from oracle.agents import reference_train_epochs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _filled_buffer(N, T, A):
    from xuance_b200.common import DummyOnPolicyBuffer_Atari, Box, Discrete
    rng = np.random.default_rng(3)
    ro = synth_rollout(rng, N, T, (84, 84, 4), n_actions=A)
    buf = DummyOnPolicyBuffer_Atari(Box(0, 255, (84, 84, 4), np.uint8), Discrete(A), {"old_logp": ()}, N, T, device=DEV)
    fill_buffers([buf], ro, [(5, 2, np.float32(0.5))])
    return buf


@pytest.mark.parametrize("fused", [False, True])
def test_reference_loop_over_xb200_classes_equals_product_agent(fused):
    from types import SimpleNamespace
    from xuance_b200.common import BaseCallback
    from xuance_b200.torch.learners import PPO_Learner
    from xuance_b200.torch.agents.on_policy import OnPolicyAgent
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    N, T, A, n_mb, n_epochs = 8, 16, 4, 4, 3
    torch.manual_seed(0)
    m0 = build_product_ppo_model(A, DEV)
    results = []
    for mode in ("reference_loop", "product"):
        model = build_product_ppo_model(A, DEV)
        model.load_state_dict(m0.state_dict())
        learner = PPO_Learner(ppo_config(DEV), model, BaseCallback())
        memory = _filled_buffer(N, T, A)
        np.random.seed(21)
        if mode == "reference_loop":
            agent = SimpleNamespace(buffer_size=N * T, batch_size=N * T // n_mb, memory=memory, learner=learner)
            info = reference_train_epochs(agent, n_epochs)
        else:
            agent = SimpleNamespace(buffer_size=N * T, batch_size=N * T // n_mb, memory=memory, learner=learner, model=model,
                                    world_size=1, device=DEV, config=SimpleNamespace(use_cuda_graph=False, fused_sample=fused))
            agent._obs_format = lambda: OnPolicyAgent._obs_format(agent)
            info = OnPolicyAgent.train_epochs(agent, n_epochs)
        results.append((info, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    (ia, pa), (ib, pb) = results
    for k in ("actor_loss", "critic_loss", "entropy", "predict_value"):
        np.testing.assert_allclose(ia[k], ib[k], rtol=1e-5, atol=1e-6, err_msg=k)
    for k in pa:
        # same minibatches in the same order; the fused gather converts u8 -> float32 with the same correctly rounded value
        np.testing.assert_allclose(pb[k].cpu().numpy(), pa[k].cpu().numpy(), rtol=1e-5, atol=1e-6, err_msg=k)
