"""Host logic of the actor packing (no GPU): which policies the fused rollout kernels take, which go to the general engine,
and the shape validation of `GeneralActor` (the reference builds MLPs of any depth / width / activation, RL/apprfunc/mlp.py:18-33)."""
import pytest
import torch

import msacl_b200  # noqa: F401
from msacl_b200.sampler import ActorWeights, GeneralActor, actor_from_policy


def _policy(sizes, act):
    mods = []
    for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        mods += [torch.nn.Linear(a, b), (act() if i < len(sizes) - 2 else torch.nn.Identity())]
    return torch.nn.Sequential(*mods)


def test_default_policy_takes_the_fused_actor():
    a = actor_from_policy(_policy([4, 256, 256, 4], torch.nn.ReLU), device="cpu")
    assert isinstance(a, ActorWeights) and (a.obs_dim, a.act_dim) == (4, 2)
    assert tuple(a.w2t.shape) == (256, 256)


def test_narrow_relu_policy_is_zero_padded_into_the_fused_actor():
    pol = _policy([7, 64, 96, 4], torch.nn.ReLU)
    a = actor_from_policy(pol, device="cpu")
    assert isinstance(a, ActorWeights) and a.hidden_sizes == (64, 96)
    assert tuple(a.w1.shape) == (256, 7) and tuple(a.w2t.shape) == (256, 256) and tuple(a.w3.shape) == (4, 256)
    x = torch.randn(33, 7)
    with torch.no_grad():
        want = pol(x)
        got = torch.relu(torch.relu(x @ a.w1.t() + a.b1) @ a.w2t + a.b2) @ a.w3.t() + a.b3
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
    assert float(a.w1[64:].abs().max()) == 0.0 and float(a.w2t[64:].abs().max()) == 0.0 and float(a.w2t[:, 96:].abs().max()) == 0.0


@pytest.mark.parametrize("sizes,act", [([4, 300, 64, 4], torch.nn.ReLU), ([12, 256, 256, 8], torch.nn.Tanh),
                                       ([2, 256, 256, 256, 2], torch.nn.ReLU), ([7, 300, 4], torch.nn.GELU)])
def test_other_policies_take_the_general_actor(sizes, act):
    pol = _policy(sizes, act)
    a = actor_from_policy(pol, device="cpu")
    assert isinstance(a, GeneralActor)
    assert (a.obs_dim, a.act_dim, len(a.layers)) == (sizes[0], sizes[-1] // 2, len(sizes) - 1)
    assert [type(m) for m in a.activations] == [act] * (len(sizes) - 2) + [torch.nn.Identity]
    with pytest.raises(ValueError):                      # a fused engine requested by name refuses such a policy
        actor_from_policy(pol, device="cpu", strict=True)


def test_general_actor_validates_the_layer_chain():
    w = lambda o, i: (torch.zeros(o, i), torch.zeros(o))
    with pytest.raises(ValueError, match="chain"):
        GeneralActor([w(8, 4), w(4, 9)], [torch.nn.ReLU(), torch.nn.Identity()], device="cpu")
    with pytest.raises(ValueError, match="one activation"):
        GeneralActor([w(8, 4), w(4, 8)], [torch.nn.ReLU()], device="cpu")
    with pytest.raises(ValueError):
        GeneralActor.from_policy(torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Linear(8, 4)), device="cpu")
