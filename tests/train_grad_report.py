"""Per-parameter gradient error of the native UNet training step against autograd through the fp32 oracle (GPU box tool).
   python tests/train_grad_report.py [B] [num_classes|none]"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from diffusion_models_collection_b200.models.unet import UNet  # noqa: E402
from oracle import model_oracle  # noqa: E402  (checker only)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    nc = None if (len(sys.argv) > 2 and sys.argv[2] == "none") else 10
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = UNet(model_channels=128, num_classes=nc, dropout=0.1).cuda().eval()
    with torch.no_grad():  # make GroupNorm affine / zero-initialised tensors non-trivial
        for n, p in net.named_parameters():
            if n.endswith(".0.weight") or n.endswith("norm.weight"):
                p.add_(0.1 * torch.randn_like(p))
            elif p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(B, 3, 32, 32, device="cuda")
    t = torch.randint(0, 1000, (B,), device="cuda")
    y = torch.randint(0, 11, (B,), device="cuda") if nc else None
    noise = torch.randn_like(x)
    eps = net(x, t, y)
    loss = F.mse_loss(noise, eps)
    loss.backward()
    torch.cuda.synchronize()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ref_eps = model_oracle.unet_forward.__wrapped__(sd, net._cfg(), x, t, y, nc)
    ref_loss = F.mse_loss(noise, ref_eps)
    names = [n for n, _ in net.named_parameters()]
    refs = torch.autograd.grad(ref_loss, [sd[n] for n in names], allow_unused=True)
    print(f"loss {loss.item():.6f} ref {ref_loss.item():.6f}  eps rel_l2 {((eps - ref_eps).norm() / ref_eps.norm()).item():.3e}")
    num = den = 0.0
    for n, r in zip(names, refs):
        g = net.get_parameter(n).grad
        if r is None or g is None:
            print(f"{n:48s} ours={'None' if g is None else 'set'} ref={'None' if r is None else 'set'}")
            continue
        e = (g - r).norm().item()
        num += e * e
        den += r.norm().item() ** 2
        print(f"{n:48s} {tuple(g.shape)!s:20s} rel {e / max(r.norm().item(), 1e-30):.3e}  |ref| {r.norm().item():.3e}")
    print(f"ALL rel_l2 {(num / den) ** 0.5:.3e}")


if __name__ == "__main__":
    main()
