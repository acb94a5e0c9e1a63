"""DiM (reference models/dim.py, the nn.MultiheadAttention variant it builds without mamba_ssm) on the DiT engine.
CPU: the oracle restatement against the golden eps of the reference's own DiM; the LayerNorm-affine fold that maps a DiM
state dict onto the DiT plan's tables (models/dim.py here) is exact: DiT-oracle(canonical(sd)) == DiM-oracle(sd); the
reference's key / shape contract (strict load) and default init laws.  GPU: native eps vs the reference golden."""

import numpy as np
import pytest
import torch

from diffusion_models_collection_b200 import synth
from oracle import model_oracle
from tests.golden_cases import DIM_CASES, case_inputs


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def _cfg(c):
    return dict(synth.CIFAR_DIM, hidden_size=c["hidden"], depth=c["depth"])


@pytest.mark.parametrize("name", list(DIM_CASES))
def test_dim_oracle_matches_reference_golden(golden, name):
    c = DIM_CASES[name]
    sd = synth.make_dim_state_dict(_cfg(c), c["num_classes"], seed=c["wseed"])
    x, t, y = case_inputs(c)
    eps = model_oracle.dim_forward(sd, _cfg(c), x, t, y, num_classes=c["num_classes"])
    assert rel_l2(eps, golden["dim"][name]) < 5e-6


@pytest.mark.parametrize("name", list(DIM_CASES))
def test_dim_maps_onto_the_dit_engine_exactly(golden, name):
    """the host-side fold (LayerNorm affine into the adaLN linears, two 3-chunk tables -> one 6-chunk table) in fp64: the DiT
    restatement fed the canonical state reproduces the DiM golden"""
    from diffusion_models_collection_b200.models import DiM

    c = DIM_CASES[name]
    cfg = _cfg(c)
    net = DiM(**cfg, num_classes=c["num_classes"])
    sd = synth.make_dim_state_dict(cfg, c["num_classes"], seed=c["wseed"])
    net.load_state_dict(sd, strict=True)  # the reference's key / shape contract
    canon = net._canonical_state({k: v.detach().clone() for k, v in net.state_dict().items()})
    x, t, y = case_inputs(c)
    eps = model_oracle.dit_forward(canon, dict(cfg, num_heads=8), x, t, y, num_classes=c["num_classes"])
    assert rel_l2(eps, golden["dim"][name]) < 2e-5


def test_dim_constructor_contract_and_init():
    from diffusion_models_collection_b200.models import DiM, DiT

    torch.manual_seed(0)
    net = DiM(hidden_size=256, depth=2, num_classes=10)
    assert isinstance(net, DiT) and net.num_heads == 8 and net.state_size == 16 and net.out_channels == 3
    sd = net.state_dict()
    assert len(sd) == 7 + 1 + 2 * 16 + 6
    assert sd["blocks.0.mamba_block.mamba.in_proj_weight"].shape == (768, 256)
    assert sd["blocks.1.ff_block.mlp.0.weight"].shape == (1024, 256)
    assert float(sd["final_layer.linear.weight"].abs().max()) == 0.0                      # zero-init output (models/dim.py:304-305)
    assert float(sd["blocks.0.mamba_block.adaLN_modulation.1.weight"].abs().max()) == 0.0
    assert torch.equal(sd["blocks.0.ff_block.norm.weight"], torch.ones(256))
    assert float(sd["y_embedder.embedding_table.weight"][0].abs().max()) == 0.0           # padding row
    with pytest.raises(RuntimeError):  # a checkpoint trained WITH mamba_ssm has other keys: refuse loudly
        net.load_state_dict({**sd, "blocks.0.mamba_block.mamba.A_log": torch.zeros(4)}, strict=True)
    from diffusion_models_collection_b200 import _lib

    with pytest.raises(_lib.DmcError):  # no CPU path, like every model here
        net(torch.zeros(1, 3, 32, 32), torch.zeros(1, dtype=torch.long))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(DIM_CASES))
def test_dim_native_eps_vs_reference_golden(golden, name):
    from diffusion_models_collection_b200.models import DiM

    c = DIM_CASES[name]
    cfg = _cfg(c)
    net = DiM(**cfg, num_classes=c["num_classes"])
    net.load_state_dict(synth.make_dim_state_dict(cfg, c["num_classes"], seed=c["wseed"]), strict=True)
    net = net.cuda().eval()
    x, t, y = case_inputs(c)
    with torch.no_grad():
        eps = net(x.cuda(), t.cuda(), None if y is None else y.cuda())
        net.precision = "bf16x3"
        eps32 = net(x.cuda(), t.cuda(), None if y is None else y.cuda())
    ref = torch.from_numpy(golden["dim"][name])
    a, b = rel_l2(eps, ref), rel_l2(eps32, ref)
    print(f"dim {name}: eps rel-L2 bf16 {a:.3e} split-bf16 {b:.3e}")
    assert a < 2e-2 and b < 1e-3


@pytest.mark.gpu
def test_dim_ddim_cfg_sampling_through_the_graph_loop():
    from diffusion_models_collection_b200.diffusion import DDIM
    from diffusion_models_collection_b200.models import DiM

    c = DIM_CASES["cond_h512"]
    net = DiM(**_cfg(c), num_classes=10)
    net.load_state_dict(synth.make_dim_state_dict(_cfg(c), 10, seed=c["wseed"]))
    net = net.cuda().eval()
    y = torch.tensor([1, 10, 3, 5]).cuda()
    d = DDIM(1000, 5, device=torch.device("cuda"))
    d.progress = False
    xT = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(3)).cuda()
    a = d.sample_with_cfg(net, (4, 3, 32, 32), y, cfg_scale=2.0, noise=xT)
    d.use_cuda_graph = False
    b = d.sample_with_cfg(net, (4, 3, 32, 32), y, cfg_scale=2.0, noise=xT)
    assert torch.isfinite(a).all() and torch.equal(a, b)
