"""GPU: the native UNet training step (BASELINE configs[4]) -- forward + backward on the CUDA kernels under
torch autograd -- against autograd through the fp32 oracle restatement of the reference model, same weights / inputs / noise.

Tolerances: activations and activation gradients are bf16 (as in the reference's documented bf16 training mode), weight gradients
are accumulated in fp32.  Per-parameter gradient relative L2 <= 6e-2 for every tensor that carries at least 1e-3 of the total
gradient norm, and <= 3e-2 over all parameters together."""

import os

import pytest
import torch
import torch.nn.functional as F

from diffusion_models_collection_b200.diffusion import DDPM
from diffusion_models_collection_b200.models.unet import UNet
from oracle import model_oracle
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu

TOL_ALL, TOL_PARAM = 3e-2, 6e-2


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _model(nc, seed=0, dropout=0.1):
    torch.manual_seed(seed)
    net = UNet(model_channels=128, num_classes=nc, dropout=dropout).cuda()
    with torch.no_grad():  # GroupNorm affines start at 1 / 0: perturb so their gradients are exercised off the trivial point
        for n, p in net.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    return net


def _batch(B, nc, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, 3, 32, 32, device="cuda", generator=g)
    t = torch.randint(0, 1000, (B,), device="cuda", generator=g)
    y = torch.randint(0, 11, (B,), device="cuda", generator=g) if nc else None
    noise = torch.randn(B, 3, 32, 32, device="cuda", generator=g)
    return x, t, y, noise


@pytest.mark.parametrize("nc,B", [(10, 4), (None, 3), (10, 16)])
def test_gradients_match_autograd_through_the_oracle(nc, B):
    net = _model(nc).eval()  # eval: no dropout, so the two sides compute the same function
    x, t, y, noise = _batch(B, nc)
    eps = net(x, t, y)
    assert eps.requires_grad
    loss = F.mse_loss(noise, eps)
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ref_eps = model_oracle.unet_forward.__wrapped__(sd, net._cfg(), x, t, y, nc)
    ref_loss = F.mse_loss(noise, ref_eps)
    names = [n for n, _ in net.named_parameters()]
    refs = torch.autograd.grad(ref_loss, [sd[n] for n in names])
    assert rel_l2(eps, ref_eps) < 2e-2 and abs(loss.item() - ref_loss.item()) < 2e-2 * ref_loss.item()
    total = sum(float(r.norm()) ** 2 for r in refs) ** 0.5
    num, worst = 0.0, (0.0, "")
    for n, r in zip(names, refs):
        g = net.get_parameter(n).grad
        assert g is not None and g.shape == r.shape and torch.isfinite(g).all(), n
        num += float((g - r).norm()) ** 2
        if float(r.norm()) >= 1e-3 * total:
            e = rel_l2(g, r)
            worst = max(worst, (e, n))
    print(f"\n[train parity nc={nc} B={B}] all-parameter rel_l2 {num ** 0.5 / total:.3e}, worst tensor {worst[1]} {worst[0]:.3e}")
    assert num ** 0.5 / total < TOL_ALL
    assert worst[0] < TOL_PARAM, worst


def test_gradient_accumulation_and_zero_grad():
    """two backward passes accumulate into .grad like autograd does; parameters that were not used get no gradient"""
    net = _model(10).eval()
    x, t, y, noise = _batch(4, 10)
    F.mse_loss(noise, net(x, t, y)).backward()
    g1 = {n: p.grad.clone() for n, p in net.named_parameters()}
    F.mse_loss(noise, net(x, t, y)).backward()
    for n, p in net.named_parameters():
        assert torch.allclose(p.grad, 2 * g1[n], rtol=1e-5, atol=1e-12), n  # deterministic kernels: exactly twice
    net.zero_grad(set_to_none=True)
    F.mse_loss(noise, net(x, t, None)).backward()  # unconditional call of a conditional model (models/unet.py:256)
    assert net.get_parameter("label_embed.weight").grad is None
    assert net.get_parameter("time_embed.1.weight").grad is not None


def test_dropout_follows_the_torch_seed():
    net = _model(10).train()
    x, t, y, noise = _batch(4, 10)

    def grads(seed):
        net.zero_grad(set_to_none=True)
        torch.manual_seed(seed)
        loss = F.mse_loss(noise, net(x, t, y))
        loss.backward()
        return loss.item(), net.get_parameter("middle_block.0.conv2.3.weight").grad.clone()

    l1, a = grads(5)
    l2, b = grads(5)
    l3, c = grads(6)
    assert l1 == l2 and torch.equal(a, b)
    assert l1 != l3 and not torch.equal(a, c)
    net.eval()
    net.zero_grad(set_to_none=True)
    l4 = F.mse_loss(noise, net(x, t, y)).item()
    with torch.no_grad():
        l5 = F.mse_loss(noise, net(x, t, y)).item()  # the sampling plan (phase-decomposed Upsample): same function
    assert abs(l4 - l5) < 5e-3 * l5 and l4 != l1


def test_stale_graph_is_reported():
    net = _model(None).eval()
    x, t, _, noise = _batch(2, None)
    l1 = F.mse_loss(noise, net(x, t))
    l2 = F.mse_loss(noise, net(x, t))
    l2.backward()
    with pytest.raises(RuntimeError, match="overwritten"):
        l1.backward()


def test_training_loop_learns_and_sampling_sees_the_new_weights():
    """the trainer's inner loop (utils/trainer.py:244-262): p_losses, backward, clip_grad_norm_, AdamW -- the loss on a fixed
    batch falls, and the sampling path picks the updated weights up (in-place re-pack, no plan rebuild)"""
    net = _model(10, dropout=0.0).train()
    ddpm = DDPM(num_timesteps=1000, device="cuda")
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4)
    g = torch.Generator(device="cuda").manual_seed(3)
    x0 = torch.randn(16, 3, 32, 32, device="cuda", generator=g).clamp(-1, 1)
    y = torch.randint(1, 11, (16,), device="cuda", generator=g)
    t = torch.randint(0, 1000, (16,), device="cuda", generator=g)
    noise = torch.randn(16, 3, 32, 32, device="cuda", generator=g)
    with torch.no_grad():
        before = net(x0, t, y).clone()
    nplans = len(net._plans)
    losses = []
    for _ in range(12):
        loss = ddpm.p_losses(net, x0, t, y, noise=noise)
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        assert torch.isfinite(gn)
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert losses[-1] < 0.8 * losses[0], losses
    with torch.no_grad():
        after = net(x0, t, y)
    assert len(net._plans) == nplans  # same plan object, re-packed weights
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    ref = model_oracle.unet_forward(sd, net._cfg(), x0, t, y, 10)
    assert rel_l2(after, ref) < 2e-2 and rel_l2(after, before) > 5e-2


def test_ddp_wrapper_single_process():
    """DistributedDataParallel(model) as utils/trainer.py:58-61 wraps it: bucket hooks fire from the per-entry autograd nodes"""
    import torch.distributed as dist

    if dist.is_initialized():
        pytest.skip("process group already initialised")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        net = _model(10).eval()
        x, t, y, noise = _batch(4, 10)
        F.mse_loss(noise, net(x, t, y)).backward()
        want = {n: p.grad.clone() for n, p in net.named_parameters()}
        net.zero_grad(set_to_none=True)
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        F.mse_loss(noise, ddp(x, t, y)).backward()
        for n, p in net.named_parameters():
            assert torch.equal(p.grad, want[n]), n
    finally:
        dist.destroy_process_group()


def test_cuda_graph_replay_equals_eager_launches():
    """from the third step on the engine replays CUDA graphs (forward + one per UNet entry); with the same torch seed they
    reproduce the eager launch lists bit for bit, dropout mask included"""
    net = _model(10).train()
    x, t, y, noise = _batch(8, 10)

    def grads():
        net.zero_grad(set_to_none=True)
        torch.manual_seed(11)
        loss = F.mse_loss(noise, net(x, t, y))
        loss.backward()
        return loss.item(), [p.grad.clone() for p in net.parameters()]

    l_eager, g_eager = grads()
    grads()
    eng = next(iter(net._train_engines.values()))
    assert eng.graphs is None
    l_graph, g_graph = grads()
    from diffusion_models_collection_b200.models import unet_train
    if unet_train.USE_GRAPHS:
        assert eng.graphs is not None and len(eng.graphs) == len(eng.segs) + 1
    assert l_eager == l_graph
    for a, b in zip(g_eager, g_graph):
        assert torch.equal(a, b)


def test_native_gradient_allreduce_matches_ddp_on_two_gpus():
    """UNet.set_gradient_allreduce() (the DDP-free data-parallel mode) under torchrun on 2 GPUs: same gradients as DDP(model)"""
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(root, "tools", "check_native_allreduce.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "native all-reduce == DDP" in r.stdout
