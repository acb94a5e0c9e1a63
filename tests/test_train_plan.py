"""CPU: host logic of the native UNet training engine (models/unet_train.py) with the C library replaced by a recorder -- no
kernel runs.  Checks what a GPU is not needed for: the backward launch list derived from the recorded forward ops covers every
parameter exactly once, reads no activation gradient before a producer wrote it, the autograd chain hands a gradient to every
parameter, and the DDP-free data-parallel mode averages the per-entry flat buffers over a world of 2 (gloo)."""

import contextlib
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_models_collection_b200 import _lib


class _Recorder:
    """stands in for libdmc_b200.so: every dmc_* call is recorded and reports success"""

    def __init__(self):
        self.calls, self.nops = [], 0

    def __getattr__(self, name):
        if not name.startswith("dmc_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append((name, args))
            if name.startswith("dmc_plan_add"):
                self.nops += 1
                return self.nops - 1
            if name == "dmc_conv_wgrad_splits":
                return 4
            if name == "dmc_gn_backward_scratch":
                return 1024
            if name == "dmc_plan_gemm_flops":
                return 0.0
            return 0

        fn.__name__ = name
        self.__dict__[name] = fn
        return fn


@pytest.fixture
def fake_lib(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(_lib, "load", lambda: rec)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_lib, "check", lambda rc, what="": rc)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    from diffusion_models_collection_b200.models import unet_train

    monkeypatch.setattr(unet_train, "USE_GRAPHS", False)  # CUDA-graph capture needs a device
    return rec


def _net(nc):
    from diffusion_models_collection_b200.models.unet import UNet

    torch.manual_seed(0)
    return UNet(model_channels=128, num_classes=nc, dropout=0.1).train()


@pytest.mark.parametrize("nc", [10, None])
def test_backward_launch_list_covers_every_parameter(fake_lib, nc):
    net = _net(nc)
    B = 2
    x, t = torch.randn(B, 3, 32, 32), torch.randint(0, 1000, (B,))
    y = torch.randint(0, 11, (B,)) if nc else None
    eps = net._run_train(x, t, y)
    assert eps.requires_grad and eps.shape == (B, 3, 32, 32)
    eng = next(iter(net._train_engines.values()))
    names = [n for n, _ in net.named_parameters()]
    # every parameter gets its gradient either from a kernel (a view of a per-entry flat buffer) or from the conditioning replay
    assert sorted(list(eng.gview) + eng.cond_names) == sorted(names)
    assert len(eng.segs) == 26 and eng.segs[0] == "stem" and eng.segs[-1] == "output"
    # one weight-gradient launch per convolution weight (+ one per fused-shortcut source), one GroupNorm backward per GroupNorm
    kinds = [m["kind"] for s in eng.segs for _, _, m in eng.bwd[s]]
    n_gn = sum(1 for n in names if n.endswith(".0.weight") or n.endswith(".norm.weight"))
    assert kinds.count("gn_backward") == n_gn
    n_conv_w = sum(1 for n in names if n.endswith(".weight") and net.get_parameter(n).dim() == 4)
    n_sc_extra = sum(1 for n in names if n.startswith("up_blocks") and n.endswith(".shortcut.weight"))  # two raw sources each
    assert kinds.count("conv_wgrad") == n_conv_w + n_sc_extra
    assert kinds.count("attention_backward") == sum(1 for n in names if n.endswith(".qkv.weight"))
    assert len(eng.drop_ops) == sum(1 for n in names if n.endswith(".conv2.0.weight"))
    # the backward pass hands a gradient to every parameter and runs the entries in reverse order
    fake_lib.calls.clear()
    eps.square().mean().backward()
    assert all(p.grad is not None and p.grad.shape == p.shape for p in net.parameters())
    ran = [a[1] for n, a in fake_lib.calls if n == "dmc_plan_run_op"]
    assert len(ran) == kinds.count("conv_dgrad")
    # a second forward before backward() invalidates the first graph
    e1 = net._run_train(x, t, y)
    net._run_train(x, t, y)
    with pytest.raises(RuntimeError, match="overwritten"):
        e1.sum().backward()


def test_weight_repack_items_describe_every_gemm_operand(fake_lib):
    """dmc_pack_weights tables: forward operands (K-concatenated conv2 + shortcut as two items of one matrix) and the transposed,
    tap-flipped input-gradient operands; geometry only (the kernel itself is checked on the GPU)"""
    net = _net(10)
    pk = net._ensure_packed(torch.device("cpu"))
    items = net._refresh_lists(pk)["items"]
    rows = {}
    for it in items:
        assert it.mode == 0 and it.taps in (1, 9) and it.cin == it.cin_total
        rows.setdefault(it.dst, []).append(it)
    names = [n for n, p in net.named_parameters() if p.dim() == 4 and not n.startswith("input_conv")]
    # + the input convolution twice: [W | W | 0], the matrix of the opt-in gathered-columns stem GEMM (hi / lo parts of the input)
    assert len(items) == len(names) + 2
    x, t, y = torch.randn(2, 3, 32, 32), torch.randint(0, 1000, (2,)), torch.randint(0, 11, (2,))
    net._run_train(x, t, y)
    eng = next(iter(net._train_engines.values()))
    for tns, kind, wkey, extra in eng.dgrad_items:
        w = net.get_parameter(wkey)
        if kind == "3x3":
            assert tns.shape[0] == w.shape[1] and tns.shape[1] % 9 == 0 and tns.shape[1] // 9 >= w.shape[0]
        else:
            assert tns.shape == (extra[1], w.shape[0]) and extra[0] + extra[1] <= w.shape[1]


def test_dit_training_engine_hands_a_gradient_to_every_parameter(fake_lib):
    """host logic of models/dit_train.py with the library replaced by the recorder: every linear has a forward, an input-gradient
    and a weight-gradient launch, the attention backward runs once per block, every parameter receives a gradient of its shape"""
    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.models.dit import DiT

    cfg = dict(synth.CIFAR_DIT, depth=2)
    net = DiT(**cfg, num_classes=10).eval()
    net.load_state_dict(synth.make_dit_state_dict(cfg, 10, seed=1))
    x, t, y = torch.randn(2, 3, 32, 32), torch.randint(0, 1000, (2,)), torch.tensor([0, 7])
    eps = net._run_train(x, t, y)
    assert eps.shape == x.shape
    eng = next(iter(net._train_engines.values()))
    for lay in eng.layers.values():  # what the kernels would have written
        lay.y.zero_(), lay.dw.fill_(1.0), lay.db.fill_(2.0)
    for buf in list(eng._dx.values()) + list(eng._dy.values()):
        buf.zero_()
    eps = net._run_train(x, t, y)
    eps.sum().backward()
    for n, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, n
    calls = [c[0] for c in fake_lib.calls]
    assert calls.count("dmc_conv_wgrad") == 4 * 2 and calls.count("dmc_attention_backward") == 2
    assert calls.count("dmc_channel_sum") == 4 * 2
    assert float(net.get_parameter("blocks.0.mlp.0.weight").grad.mean()) == 1.0
    assert net.get_parameter("blocks.0.mlp.0.weight").grad.data_ptr() != eng.layers["blocks.0.fc1"].dw.data_ptr()
    from diffusion_models_collection_b200.models.dim import DiM

    with pytest.raises(NotImplementedError):
        DiM(hidden_size=512, depth=1)._run_train(x, t, None)


def _native_allreduce_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec = _Recorder()
        _lib.load, _lib.stream_ptr, _lib.check = (lambda: rec), (lambda: 0), (lambda rc, what="": rc)
        torch.cuda.device = lambda d: contextlib.nullcontext()
        from diffusion_models_collection_b200.models import unet_train

        unet_train.USE_GRAPHS = False
        net = _net(None).eval()
        with torch.no_grad():
            for p in net.parameters():
                p.add_(float(rank))  # different parameters per rank: set_gradient_allreduce must broadcast rank 0's
        net.set_gradient_allreduce()
        first = next(net.parameters()).detach().clone()
        x, t = torch.randn(2, 3, 32, 32), torch.randint(0, 1000, (2,))
        eps = net._run_train(x, t, None)
        eng = next(iter(net._train_engines.values()))
        for s in eng.segs:  # what the kernels would have written: rank-dependent constants
            eng.flat[s].fill_(float(rank + 1))
        for _, dc in eng.dcond_parts:
            dc.zero_()
        eps.sum().backward()
        # (head / stem / fused-shortcut weights pass through staging tensors the recorder never fills: left out of the value check)
        kernel_params = [net.get_parameter(n) for n in eng.gview
                         if not (n.startswith("output.2") or n == "input_conv.weight" or n.endswith(".shortcut.weight"))]
        ok = all(torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2)) for p in kernel_params)
        ok = ok and all(p.grad is not None for p in net.parameters()) and net.module is net
        # second backward pass accumulates the averaged gradient on top
        eps = net._run_train(x, t, None)
        for s in eng.segs:
            eng.flat[s].fill_(float(rank + 1))
        for _, dc in eng.dcond_parts:
            dc.zero_()
        eps.sum().backward()
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.0 + world)) for p in kernel_params)
        out[rank] = (bool(ok), first)
    finally:
        dist.destroy_process_group()


def test_native_gradient_allreduce_world_2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_native_allreduce_worker, args=(world, 29655, out), nprocs=world, join=True)
        assert all(out[r][0] for r in range(world))
        assert torch.equal(out[0][1], out[1][1])  # parameters were broadcast from rank 0


def _ddp_wrap_worker(rank, world, port, out):
    """the reference trainer's own line, `DDP(model)` (utils/trainer.py:58-61), around the native UNet: DDP is handed every parameter
    but the sentinel to ignore, the engine averages the rest itself, the sentinel's gradient goes through DDP's reducer"""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec = _Recorder()
        _lib.load, _lib.stream_ptr, _lib.check = (lambda: rec), (lambda: 0), (lambda rc, what="": rc)
        torch.cuda.device = lambda d: contextlib.nullcontext()
        from torch.nn.parallel import DistributedDataParallel as DDP

        from diffusion_models_collection_b200.models import unet_train

        unet_train.USE_GRAPHS = False
        net = _net(None).train()
        net.dropout = 0.0
        assert not hasattr(net, "_ddp_params_and_buffers_to_ignore")  # nobody but a DDP constructor sees the attribute
        with torch.no_grad():
            for p in net.parameters():
                p.add_(float(rank))
        net._run = lambda x, t, y, cfg: net._run_train(x, t, y)  # (the CUDA-only check of forward(); no device here)
        model = DDP(net)
        ok = net._grad_allreduce is model.process_group and net._ddp_sentinel == "input_conv.bias"
        ok = ok and len(model._module_parameters) == 1 and model._module_parameters[0] is net.get_parameter("input_conv.bias")
        ok = ok and len(model.parameters_to_ignore) == len(list(net.parameters())) - 1 and model.module is net
        first = next(net.parameters()).detach().clone()
        x, t = torch.randn(2, 3, 32, 32), torch.randint(0, 1000, (2,))
        kernel_params = None
        for it in range(2):  # two steps: DDP's reducer must be satisfied after the first (else the second forward raises)
            eps = model(x, t, None)
            eng = next(iter(net._train_engines.values()))
            for s in eng.segs:
                eng.flat[s].fill_(float(rank + 1))
            for _, dc in eng.dcond_parts:
                dc.zero_()
            eps.sum().backward()
            kernel_params = [net.get_parameter(n) for n in eng.gview
                             if not (n.startswith("output.2") or n == "input_conv.weight" or n.endswith(".shortcut.weight"))]
            ok = ok and all(torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2)) for p in kernel_params)
            ok = ok and all(p.grad is not None for p in net.parameters())
            ok = ok and any(p is net.get_parameter("input_conv.bias") for p in kernel_params)
            for p in net.parameters():
                p.grad = None
        os.environ["DMC_DDP_NATIVE"] = "0"  # stock DDP over every parameter
        plain = _net(None)
        ok = ok and len(DDP(plain)._module_parameters) == len(list(plain.parameters())) and plain._grad_allreduce is None
        out[rank] = (bool(ok), first)
    finally:
        dist.destroy_process_group()


def test_ddp_wrapper_hands_the_gradients_to_the_native_allreduce_world_2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ddp_wrap_worker, args=(world, 29657, out), nprocs=world, join=True)
        assert all(out[r][0] for r in range(world))
        assert torch.equal(out[0][1], out[1][1])  # rank 0's parameters everywhere, like DDP's own broadcast


def test_models_deepcopy_and_pickle_without_their_native_caches():
    """EMA helpers deep-copy the model (torch.optim.swa_utils.AveragedModel does) and torch.save(model) pickles it: plans, engines
    and packed operands hold native handles / device pointers and must not travel; the copy rebuilds them on its first forward"""
    import copy
    import io

    from diffusion_models_collection_b200 import synth
    from diffusion_models_collection_b200.models.dit import DiT

    net = _net(10)
    net._plans, net._packed = {"k": object()}, {"stale": 1}
    twin = copy.deepcopy(net)
    assert twin._plans == {} and twin._train_engines == {} and twin._packed is None and twin._plist is None
    assert net._plans != {}  # the original keeps its caches
    assert [k for k in twin.state_dict()] == [k for k in net.state_dict()]
    assert all(torch.equal(a, b) and a is not b for a, b in zip(net.parameters(), twin.parameters()))
    net._plans, net._packed = {}, None
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    assert len(torch.load(buf, weights_only=False).state_dict()) == 357
    dit = DiT(**synth.CIFAR_DIT, num_classes=None)
    assert copy.deepcopy(dit)._plans == {}


def test_fused_adamw_host_side_contract():
    """FusedAdamW without a device: torch.optim.AdamW hyper-parameters / state keys, torch AdamW checkpoints load (and reset the
    pointer tables), EMA pairing is checked, and step() refuses to run without the CUDA path"""
    from diffusion_models_collection_b200.optim import FusedAdamW

    ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 3))]
    opt = FusedAdamW(ps, lr=1e-3, weight_decay=0.1, max_grad_norm=1.0)
    assert opt.param_groups[0]["betas"] == (0.9, 0.999) and opt.param_groups[0]["weight_decay"] == 0.1
    ref_ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 3))]
    ref = torch.optim.AdamW(ref_ps, lr=5e-4)
    for p in ref_ps:
        p.grad = torch.ones_like(p)
    ref.step()
    opt._tables = "stale"
    opt.load_state_dict(ref.state_dict())
    assert opt._tables is None and opt.param_groups[0]["lr"] == 5e-4
    assert set(opt.state[ps[0]]) == {"step", "exp_avg", "exp_avg_sq"} and float(opt.state[ps[0]]["step"]) == 1.0
    assert not opt.state[ps[0]]["step"].is_cuda
    with pytest.raises(ValueError):
        FusedAdamW(ps, ema_params=[torch.zeros(4)])
    with pytest.raises(ValueError):
        FusedAdamW(ps, lr=-1.0)
    for p in ps:
        p.grad = torch.ones_like(p)
    with pytest.raises(_lib.DmcError):
        opt.step()
