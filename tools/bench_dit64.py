import sys, json, torch
sys.path.insert(0, '.')
from diffusion_models_collection_b200 import synth
from diffusion_models_collection_b200.models import DiT
cfg = dict(synth.CIFAR_DIT, img_size=(64, 64))
net = DiT(**cfg, num_classes=None); net.load_state_dict(synth.make_dit_state_dict(cfg, None, seed=42)); net = net.cuda().eval()
B = 256
x = torch.randn(B, 3, 64, 64, device='cuda'); t = torch.full((B,), 500, device='cuda')
with torch.no_grad(), net.uniform_timesteps():
    net(x, t)
    plan = net.plan_info(B)
    ops = plan.time_ops(iters=3)
fam = {}
for o in ops: fam[o['kind']] = fam.get(o['kind'], 0) + o['ms']
tot = sum(fam.values())
print('DiT-64 B=%d forward %.2f ms (%.1f us/img)' % (B, tot, 1e3 * tot / B), {k: round(v, 3) for k, v in fam.items()})
att = [o for o in ops if o['kind'] == 'attention'][0]
print('attention L=1024: %.3f ms %.0f TFLOP/s' % (att['ms'], att['flops'] / att['ms'] / 1e9))
