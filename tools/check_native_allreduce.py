"""torchrun --nproc-per-node 2 tools/check_native_allreduce.py
Gradients of UNet.set_gradient_allreduce() (one NCCL all-reduce per UNet entry from the engine's flat buffers, no DDP wrapper)
against DistributedDataParallel(model) on the same per-rank batches."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_models_collection_b200.models.unet import UNet  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    a = UNet(model_channels=128, num_classes=10, dropout=0.0).to(dev).eval()
    b = UNet(model_channels=128, num_classes=10, dropout=0.0).to(dev).eval()
    b.load_state_dict(a.state_dict())
    c = UNet(model_channels=128, num_classes=10, dropout=0.0).to(dev).eval()
    c.load_state_dict(a.state_dict())
    g = torch.Generator(device=dev).manual_seed(100 + rank)  # a different batch on every rank
    x = torch.randn(4, 3, 32, 32, device=dev, generator=g)
    t = torch.randint(0, 1000, (4,), device=dev, generator=g)
    y = torch.randint(0, 11, (4,), device=dev, generator=g)
    noise = torch.randn(4, 3, 32, 32, device=dev, generator=g)
    os.environ["DMC_DDP_NATIVE"] = "0"  # a: stock DDP over every parameter (per-parameter bucket copies)
    ddp = torch.nn.parallel.DistributedDataParallel(a)
    assert len(ddp._module_parameters) == len(list(a.parameters())) and a._grad_allreduce is None
    b.set_gradient_allreduce()
    os.environ["DMC_DDP_NATIVE"] = "1"  # c: the same DDP(model) line, the UNet hands DDP all parameters but one to ignore
    ddp_c = torch.nn.parallel.DistributedDataParallel(c)
    assert len(ddp_c._module_parameters) == 1 and c._grad_allreduce is ddp_c.process_group
    worst = 0.0
    for it in range(4):  # eager launches, then CUDA-graph replays; the last pass accumulates on top of the third
        if it < 3:
            a.zero_grad(set_to_none=True)
            b.zero_grad(set_to_none=True)
            c.zero_grad(set_to_none=True)
        F.mse_loss(noise, ddp(x, t, y)).backward()
        F.mse_loss(noise, b(x, t, y)).backward()
        F.mse_loss(noise, ddp_c(x, t, y)).backward()
        for (n, p), q, r in zip(a.named_parameters(), b.parameters(), c.parameters()):
            assert q.grad is not None and p.grad is not None and r.grad is not None, n
            for other in (q, r):
                err = float((p.grad - other.grad).norm() / p.grad.norm().clamp_min(1e-30))
                worst = max(worst, err)
                assert err < 1e-5, (it, n, err)
        # every rank holds the same averaged gradient
        chk = torch.stack([q.grad.double().sum() for q in b.parameters()]).sum()
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert all(torch.equal(both[0], v) for v in both)
    if rank == 0:
        print(f"native all-reduce == DDP == DDP(model) with the ignore list: worst relative difference {worst:.2e} over 4 backward passes", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
