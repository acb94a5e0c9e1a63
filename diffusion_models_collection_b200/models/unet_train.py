"""Training step of the native UNet (BASELINE configs[4]): ``loss = diffusion.p_losses(model, x0, t, y)``
followed by ``loss.backward()`` -- what utils/trainer.py:249-254 of the reference does every iteration -- with the forward AND the
backward pass running on the hand-written sm_100a kernels behind the C ABI (include/dmc.h):

  forward   the same launch list as sampling, recorded with every activation kept alive, ResidualBlock dropout
            (models/unet.py:53) folded into the conv2 GroupNorm+SiLU pass, Upsample as upsample + 3x3 conv
  backward  derived from the recorded forward ops, walked in reverse:
              conv      input gradient  = the forward tcgen05 implicit-GEMM kernel over dY with transposed, tap-flipped weights
                        weight gradient = dmc_conv_wgrad (tcgen05, K = output pixels, both operands straight from NHWC)
                        bias / per-image conditioning gradient = dmc_channel_sum
              GroupNorm(+SiLU)(+dropout) = dmc_gn_backward (recomputes the normalised value and the dropout mask)
              attention = dmc_attention_backward;  nearest-2x upsample = dmc_block_sum2x2;  stride-2 conv = dmc_conv_dgrad_strided
            activation gradients are bf16 NHWC, parameter gradients fp32 in the reference's parameter layout.

Autograd sees a *chain* of ``torch.autograd.Function`` nodes, one per UNet entry (output head, every up / middle / down entry, stem +
conditioning MLPs), each taking that entry's parameters as inputs.  The gradients of an entry are therefore handed to autograd -- and
to DistributedDataParallel's bucket hooks -- as soon as that entry's backward kernels have been enqueued, so the NCCL all-reduce of
the late layers' buckets overlaps the backward kernels of the earlier layers exactly like it does for the reference's module tree
(utils/trainer.py:58-61).  Optimizer, gradient clipping, EMA and checkpointing keep working on ordinary ``nn.Parameter``s.

The small conditioning path (time_embed, label_embed, the 22 time_mlp / label_proj rows: 0.1 % of the FLOPs) is differentiated by
replaying it in PyTorch from the per-image sums of the conv1 output gradients.  There is no CPU / eager fallback for the rest.
"""

from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn.functional as F

from .. import _lib
from ..synth import unet_block_structure
from .unet import layer_seed

USE_GRAPHS = os.environ.get("DMC_TRAIN_GRAPHS", "1") != "0"


def _seg_of(name):
    parts = name.split(".")
    if parts[0] in ("down_blocks", "up_blocks"):
        return parts[0] + "." + parts[1]
    if parts[0] == "middle_block":
        return "middle_block"
    if parts[0] == "output":
        return "output"
    return "stem"


def _is_cond_param(name):
    return (name.startswith("time_embed.") or name.startswith("label_embed.") or ".time_mlp." in name or ".label_proj." in name
            or name.endswith(".conv1.2.bias"))


class _Segment(torch.autograd.Function):
    """One entry of the UNet in the autograd graph: forward is a no-op (the whole native forward has already been enqueued), backward
    enqueues the entry's backward kernels and returns its parameter gradients."""

    @staticmethod
    def forward(ctx, eng, k, token, *params):
        ctx.eng, ctx.k, ctx.step = eng, k, eng.step_id
        ctx.mask = [isinstance(p, torch.Tensor) and p.requires_grad for p in params]
        if k == len(eng.segs) - 1:
            return eng.eps_out
        return torch.empty(0, device=eng.device, dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        eng = ctx.eng
        if ctx.step != eng.step_id:
            raise RuntimeError("UNet backward: the activations of this forward have been overwritten by a later forward of the same "
                               "batch size (run backward() before the next training forward)")
        grads = eng.backward_segment(ctx.k, g, ctx.mask)
        grads = [gr if m else None for gr, m in zip(grads, ctx.mask)]
        tok = None if ctx.k == 0 else torch.zeros(0, device=eng.device, dtype=torch.float32)
        return (None, None, tok, *grads)


class UNetTrainEngine:
    """forward + backward launch lists of one (batch size, conditional?, dropout?) signature"""

    def __init__(self, net, device, B, has_y, drop_p):
        from .unet import _UNetPlan

        lib = self.lib = _lib.load()
        self.net, self.device, self.B, self.has_y, self.drop_p = net, device, B, has_y, float(drop_p)
        mc = net.model_channels
        if mc % 128 != 0 or any((mc * m) % 128 for m in net.channel_mult):
            raise NotImplementedError("native UNet training: channel counts must be multiples of 128 (weight-gradient tiles)")
        if B > 4096:
            raise NotImplementedError("native UNet training: at most 4096 images per step per GPU")
        pk = self.pk = net._ensure_packed(device, training=True)
        self.seed_dev = torch.zeros(1, dtype=torch.int32, device=device)  # per-step dropout seed (read by the kernels)
        self.fwd = _UNetPlan(net, pk, device, B, B, has_y, False, keep=True, drop_p=self.drop_p, seed_dev=self.seed_dev)
        self.step_id = 0
        self.eps_out = None
        self._keep = []
        down, middle, up, _ = unet_block_structure(net._cfg())
        self.segs = (["stem"] + [f"down_blocks.{i}" for i in range(len(down))] + ["middle_block"] +
                     [f"up_blocks.{i}" for i in range(len(up))] + ["output"])
        # parameters per segment; the conditioning parameters all belong to the stem segment (their gradient needs every block)
        self.seg_params = {s: [] for s in self.segs}
        self.cond_names = []
        for name, _ in net.named_parameters():
            if _is_cond_param(name):
                self.cond_names.append(name)
            else:
                self.seg_params[_seg_of(name)].append(name)
        self.flat, self.gview, self.layout = {}, {}, {}
        for s in self.segs:
            off, lay = 0, []
            for name in self.seg_params[s]:
                p = net.get_parameter(name)
                lay.append((name, off, p.numel(), tuple(p.shape)))
                off += (p.numel() + 63) // 64 * 64
            self.flat[s] = torch.zeros(max(off, 64), dtype=torch.float32, device=device)
            self.layout[s] = lay
            for name, o, n, shp in lay:
                self.gview[name] = self.flat[s][o: o + n].view(shp)
        Hh, Ww = net._hw
        self.xpad = torch.zeros((B, Hh, Ww, 64), dtype=torch.bfloat16, device=device)
        self.deps_pad = torch.zeros((B, Hh, Ww, 128), dtype=torch.bfloat16, device=device)
        # static staging buffers: the launch lists (and the CUDA graphs made of them) never see a caller-owned pointer
        self.x_static = torch.zeros((B, net.in_channels, Hh, Ww), dtype=torch.float32, device=device)
        self.deps_static = torch.zeros((B, net.out_channels, Hh, Ww), dtype=torch.float32, device=device)
        _lib.check(lib.dmc_plan_rebind(self.fwd.handle, self.fwd.stem_idx, 0, self.x_static.data_ptr()), "rebind x")
        self._build_backward()
        self._dgrad_ver, self._dgrad_table = None, None
        self._seg_param_objs = None
        self._cond_param_objs = None
        self.graphs = None
        self._pending = []

    # ------------------------------------------------------------------------------------------------------------------
    def _build_backward(self):
        net, lib, pl, pk, B, device = self.net, self.lib, self.fwd, self.pk, self.B, self.device
        b = pl.builder
        sd = pk["sd"]
        wsp = pl.ws.data_ptr()
        Hh, Ww = net._hw
        bf16, f32 = torch.bfloat16, torch.float32

        def ap(a):
            return wsp + a.blk[0]

        G, written = {}, set()

        def gbuf(a):
            t = G.get(id(a))
            if t is None:
                t = G[id(a)] = torch.empty((B, a.H, a.W, a.C), dtype=bf16, device=device)
            return t

        def contribute(a):
            t = gbuf(a)
            acc = 1 if id(a) in written else 0
            written.add(id(a))
            return t.data_ptr(), acc

        def dy_of(a):
            assert id(a) in written, "backward: gradient read before any producer wrote it"
            return G[id(a)].data_ptr()

        self.bwd = {s: [] for s in self.segs}
        handle = C.c_void_p()
        _lib.check(lib.dmc_plan_create(C.byref(handle)), "dmc_plan_create")
        self.bplan = handle
        self.dgrad_items = []   # (dst bf16 matrix, kind, parameter name, extra)
        self.dcond_parts = []   # (column, fp32 [B, cout]) per-image sums of the conv1 output gradients
        self.drop_ops = []      # (forward plan op index, backward descriptor, layer number)
        wg_descs, gn_descs = [], []
        gn_scratch_need = 64
        cur = None

        def emit(fn, *args, kind="", name="", flops=0.0):
            cur.append((fn, args, dict(kind=kind, name=name, flops=float(flops))))

        def emit_py(fn, name=""):
            cur.append((fn, None, dict(kind="torch_copy", name=name, flops=0.0)))

        def dgrad_matrix(kind, wkey, rows, K, extra=None):
            t = torch.zeros((rows, K), dtype=bf16, device=device)
            self.dgrad_items.append((t, kind, wkey, extra))
            return t

        def conv_dgrad(dyp, dyC, taps, Ho, Wo, wmat, dst, name, real_c=None, flops_scale=1.0):
            ptr, acc = contribute(dst)
            d = _lib.ConvDesc()
            d.nsrc = 1
            d.src[0], d.src_c[0], d.src_taps[0] = dyp, dyC, taps
            d.B, d.Hin, d.Win, d.stride, d.up_phase = B, Ho, Wo, 1, -1
            d.weight, d.Cout, d.Cout_pad, d.Ktot = wmat.data_ptr(), dst.C, dst.C, taps * dyC
            d.out_bf16 = ptr
            if acc:
                d.residual = ptr  # accumulate: out = conv + out (every element is read, then written, by the same thread)
            idx = _lib.check(lib.dmc_plan_add_conv(handle, C.byref(d)), "dgrad conv")
            emit(lib.dmc_plan_run_op, handle, idx, kind="conv_dgrad", name=name,
                 flops=2.0 * B * Ho * Wo * dst.C * taps * (real_c or dyC) * flops_scale)

        def wgrad(xp, xC, dyp, dyC, H, W, stride, taps, dw, name, real=None, window=None):
            d = _lib.WgradDesc()
            d.x, d.dy, d.B, d.Hin, d.Win, d.Cin, d.Cout, d.stride, d.taps = xp, dyp, B, H, W, xC, dyC, stride, taps
            d.splits = _lib.check(lib.dmc_conv_wgrad_splits(C.byref(d)), "dmc_conv_wgrad_splits")
            d.dw, d.accumulate = dw.data_ptr(), 0
            if window is not None:  # (rows, columns, columns of the parameter, first column): dw is a window of a parameter
                d.dw_cout, d.dw_cin, d.dw_cin_total, d.dw_ci0 = window
            wg_descs.append(d)
            self._keep.append(dw)
            emit(lib.dmc_conv_wgrad, C.byref(d), kind="conv_wgrad", name=name,
                 flops=2.0 * B * (H // stride) * (W // stride) * taps * (real or xC * dyC))

        cs_scratch = torch.empty(B * 1024, dtype=f32, device=device)  # [B, C <= 1024] per-image partial sums
        self._keep.append(cs_scratch)

        def channel_sum(dyp, out, HW, Cc, per_image):
            assert Cc <= 1024
            self._keep.append(out)
            emit(lib.dmc_channel_sum, dyp, out.data_ptr(), B, HW, Cc, per_image, 0, cs_scratch.data_ptr(), kind="channel_sum")

        def copy_op(dst, src):
            emit_py(lambda: dst.copy_(src))

        cur_seg = "output"
        cur = self.bwd["output"]
        emit(lib.dmc_nchw_f32_to_nhwc_bf16, self.deps_static.data_ptr(), self.deps_pad.data_ptr(), B, net.out_channels, Hh * Ww, 128,
             kind="pack", name="d(eps)")
        for i in reversed(range(len(b.ops))):
            kind, o = b.ops[i]
            nm = o.get("wname") or o.get("prefix")
            if kind == "stem":
                cur_seg = "stem"
            elif nm:
                cur_seg = _seg_of(nm)
            cur = self.bwd[cur_seg]

            if kind == "conv":
                wname, srcs, taps = o["wname"], o["srcs"], o["taps"]
                H, W, stride = o["H"], o["W"], o["stride"]
                Ho, Wo = H // stride, W // stride
                head = o["out_nchw"]
                if head:
                    dyp, dyC = self.deps_pad.data_ptr(), 128
                else:
                    dyp, dyC = dy_of(o["out"]), o["out"].C
                p = wname.rsplit(".", 1)[0]
                fused_sc = wname.endswith(".conv2+sc")
                if wname.endswith(".conv1"):
                    wkey, bkeys = p + ".conv1.2.weight", []
                elif wname.endswith(".conv2") or fused_sc:
                    wkey, bkeys = p + ".conv2.3.weight", [p + ".conv2.3.bias"] + ([p + ".shortcut.bias"] if fused_sc else [])
                else:  # .qkv / .proj / Downsample / Upsample .conv / output.2
                    wkey, bkeys = wname + ".weight", [wname + ".bias"]
                # --- bias (or per-image conditioning) gradient
                if o["cond_col"] is not None:
                    dc = torch.empty((B, dyC), dtype=f32, device=device)
                    self.dcond_parts.append((o["cond_col"], dc))
                    channel_sum(dyp, dc, Ho * Wo, dyC, 1)
                if head:
                    tmpb = torch.empty(128, dtype=f32, device=device)
                    channel_sum(dyp, tmpb, Ho * Wo, 128, 0)
                    copy_op(self.gview[bkeys[0]], tmpb[: net.out_channels])
                elif bkeys:
                    channel_sum(dyp, self.gview[bkeys[0]], Ho * Wo, dyC, 0)
                    for extra in bkeys[1:]:
                        copy_op(self.gview[extra], self.gview[bkeys[0]])
                # --- weight gradients
                a_in = srcs[0]
                if head:
                    wgrad(ap(a_in), a_in.C, dyp, 128, H, W, 1, 9, self.gview[wkey], wname, real=a_in.C * net.out_channels,
                          window=(net.out_channels, a_in.C, a_in.C, 0))  # dY is zero-padded 3 -> 128 channels
                else:
                    wgrad(ap(a_in), a_in.C, dyp, dyC, H, W, stride, taps[0], self.gview[wkey], wname)
                off = 0
                cin_sc = sum(s.C for s in srcs[1:])
                for s_ in srcs[1:]:  # fused 1x1 shortcut over the raw block inputs: one slice of shortcut.weight per source
                    wgrad(ap(s_), s_.C, dyp, dyC, H, W, 1, 1, self.gview[p + ".shortcut.weight"], p + ".shortcut",
                          window=(dyC, s_.C, cin_sc, off))
                    off += s_.C
                # --- input gradients
                off = 0
                for s_ in srcs[1:]:
                    wm = dgrad_matrix("1x1", p + ".shortcut.weight", s_.C, dyC, (off, s_.C))
                    conv_dgrad(dyp, dyC, 1, Ho, Wo, wm, s_, p + ".shortcut")
                    off += s_.C
                if stride == 2:
                    # Downsample: spread dY onto the input grid (zeros between), then the ordinary stride-1 input-gradient GEMM
                    dil = torch.empty((B, H, W, dyC), dtype=bf16, device=device)
                    self._keep.append(dil)
                    emit(lib.dmc_dilate2x, dyp, dil.data_ptr(), B, Ho, Wo, dyC, kind="dilate2x", name=wname)
                    wm = dgrad_matrix("3x3", wkey, a_in.C, 9 * dyC, None)
                    conv_dgrad(dil.data_ptr(), dyC, 9, H, W, wm, a_in, wname, flops_scale=0.25)
                elif taps[0] == 9:
                    wm = dgrad_matrix("3x3", wkey, a_in.C, 9 * dyC, None)
                    conv_dgrad(dyp, dyC, 9, Ho, Wo, wm, a_in, wname, real_c=net.out_channels if head else None)
                else:
                    wm = dgrad_matrix("1x1", wkey, a_in.C, dyC, (0, a_in.C))
                    conv_dgrad(dyp, dyC, 1, Ho, Wo, wm, a_in, wname)
                # --- identity residual branch (models/unet.py:72 with an Identity shortcut, :99)
                res = o["residual"]
                if res is not None:
                    if id(res) not in written:  # first contribution: share the buffer, later ops accumulate into it
                        G[id(res)] = G[id(o["out"])]
                        written.add(id(res))
                    else:
                        emit(lib.dmc_add_bf16, gbuf(res).data_ptr(), dyp, B * res.H * res.W * res.C, 1, kind="add")

            elif kind == "gn_apply":
                out, srcs = o["out"], o["srcs"]
                d = _lib.GnBwdDesc()
                d.nsrc = len(srcs)
                Ctot = 0
                for k_, s_ in enumerate(srcs):
                    ptr, acc = contribute(s_)
                    d.src[k_], d.src_c[k_], d.stats[k_], d.stats_slots[k_] = ap(s_), s_.C, wsp + s_.stats[0], s_.slots
                    d.dsrc[k_], d.accumulate[k_] = ptr, acc
                    Ctot += s_.C
                HW = srcs[0].H * srcs[0].W
                d.dout, d.B, d.HW, d.groups = dy_of(out), B, HW, 8
                d.gamma, d.beta = sd[o["prefix"] + ".weight"].data_ptr(), sd[o["prefix"] + ".bias"].data_ptr()
                d.eps, d.silu, d.drop_p, d.seed = 1e-5, o["silu"], o["drop_p"], 0
                if o["drop_p"] > 0:  # the same (constant + device seed) pair as the forward pass
                    d.seed, d.seed_dev = layer_seed(pl.op_index[i]), self.seed_dev.data_ptr()
                d.dgamma, d.dbeta = self.gview[o["prefix"] + ".weight"].data_ptr(), self.gview[o["prefix"] + ".bias"].data_ptr()
                gn_scratch_need = max(gn_scratch_need, int(lib.dmc_gn_backward_scratch(C.byref(d))))
                gn_descs.append(d)
                if o["drop_p"] > 0:
                    self.drop_ops.append(pl.op_index[i])
                emit(lib.dmc_gn_backward, C.byref(d), kind="gn_backward", name=o["prefix"])

            elif kind == "attention":
                qkv, ao = o["qkv"], o["out"]
                ptr, acc = contribute(qkv)
                assert acc == 0
                d = _lib.AttnBwdDesc()
                d.qkv, d.out, d.dout, d.dqkv = ap(qkv), ap(ao), dy_of(ao), ptr
                d.B, d.L, d.heads, d.C = B, o["L"], 4, o["C"]
                self._keep.append(d)
                emit(lib.dmc_attention_backward, C.byref(d), kind="attention_backward")

            elif kind == "upsample":
                src, upb = o["src"], o["out"]
                dyp = dy_of(upb)
                ptr, acc = contribute(src)
                emit(lib.dmc_block_sum2x2, dyp, ptr, B, src.H, src.W, src.C, acc, kind="block_sum2x2")

            elif kind == "stem":
                h0 = o["out"]
                dyp = dy_of(h0)
                channel_sum(dyp, self.gview["input_conv.bias"], Hh * Ww, h0.C, 0)
                wgrad(self.xpad.data_ptr(), 64, dyp, h0.C, Hh, Ww, 1, 9, self.gview["input_conv.weight"], "input_conv",
                      real=net.in_channels * h0.C, window=(h0.C, net.in_channels, net.in_channels, 0))  # x is zero-padded 3 -> 64

        need = max(d.splits * d.Cout * d.taps * d.Cin for d in wg_descs)
        self.wg_partial = torch.empty(need, dtype=f32, device=device)
        for d in wg_descs:
            d.partial = self.wg_partial.data_ptr()
        self._keep.append(wg_descs)
        self.gn_scratch = torch.empty(gn_scratch_need, dtype=f32, device=device)
        for d in gn_descs:
            d.scratch = self.gn_scratch.data_ptr()
        self._keep.append(gn_descs)
        self.G = G
        self.dcond_parts.sort(key=lambda e: e[0])
        self.num_backward_ops = sum(len(v) for v in self.bwd.values())

    # ------------------------------------------------------------------------------------------------------------------
    def _refresh_dgrad(self):
        """transposed, tap-flipped bf16 copies of the convolution weights for the input-gradient GEMMs (one pack-kernel launch)"""
        ver = self.net._param_version()
        if self._dgrad_ver == ver:
            return
        if self._dgrad_table is None:
            sd = self.pk["sd"]
            items = []
            for t, kind, wkey, extra in self.dgrad_items:
                w = sd[wkey]
                it = _lib.PackItem()
                it.src, it.dst = w.data_ptr(), t.data_ptr()
                it.cout, it.cin_total, it.mode, it.ld, it.col0 = w.shape[0], w.shape[1], 1, t.shape[1], 0
                if kind == "3x3":  # [ci, (2-r, 2-s), co_pad] <- w[co, ci, r, s]
                    it.ci0, it.cin, it.taps, it.cpad = 0, w.shape[1], 9, t.shape[1] // 9
                else:             # [c, co] <- w[co, off + c]
                    it.ci0, it.cin, it.taps, it.cpad = extra[0], extra[1], 1, t.shape[1]
                items.append(it)
            self._dgrad_table = _lib.pack_table(items, self.device)
        _lib.check(self.lib.dmc_pack_weights(self._dgrad_table[0].data_ptr(), self._dgrad_table[1], _lib.stream_ptr()),
                   "dmc_pack_weights")
        self._dgrad_ver = ver

    def forward(self, x, t, y):
        lib, net, B = self.lib, self.net, self.B
        self._refresh_dgrad()  # (the caller has just re-packed the forward weights: UNet._run_train)
        if self.drop_ops:
            base = int(torch.empty((), dtype=torch.int64).random_(0, 2 ** 31 - 1).item())  # CPU generator: follows torch.manual_seed
            self.seed_dev.fill_(base)
        pl = self.fwd
        self.x_static.copy_(x)
        pl.t_stage.copy_(t)
        if self.has_y:
            pl.y_stage.copy_(y)
        if USE_GRAPHS and self.graphs is None and self.step_id >= 2:
            self._capture()
        if self.graphs is not None:
            self.graphs["forward"].replay()
        else:
            self._forward_ops()
        self.step_id += 1
        self.eps_out = pl.eps.clone()
        return self.eps_out

    def _forward_ops(self):
        lib, net, pl = self.lib, self.net, self.fwd
        Hh, Ww = net._hw
        st = _lib.stream_ptr()
        _lib.check(lib.dmc_nchw_f32_to_nhwc_bf16(self.x_static.data_ptr(), self.xpad.data_ptr(), self.B, net.in_channels, Hh * Ww,
                                                 64, st), "pack x")
        _lib.check(lib.dmc_plan_run(pl.handle, st), "dmc_plan_run")

    def _capture(self):
        """after two eager steps (every lazily initialised kernel attribute is set): the forward launch list and the backward
        launch list of every segment become CUDA graphs -- all their pointers are static, the dropout seed lives in device
        memory -- so a step costs 1 + 26 graph launches instead of ~550 kernel launches from Python"""
        graphs = {}
        torch.cuda.synchronize()
        pool = None
        for name, fn in [("forward", self._forward_ops)] + [(s, (lambda s=s: self._run_ops(self.bwd[s]))) for s in reversed(self.segs)]:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                fn()
            pool = g.pool()
            graphs[name] = g
        self.graphs = graphs

    def attach(self):
        """wires the autograd chain for the forward that has just been enqueued; returns eps with a grad_fn"""
        if self._seg_param_objs is None:  # (parameter objects are stable: a storage change rebuilds the engine)
            net = self.net
            self._seg_param_objs = [[net.get_parameter(n) for n in self.seg_params[s] + (self.cond_names if k == 0 else [])]
                                    for k, s in enumerate(self.segs)]
        tok = None
        for k, params in enumerate(self._seg_param_objs):
            tok = _Segment.apply(self, k, tok, *params)
        return tok

    # ------------------------------------------------------------------------------------------------------------------
    def _run_ops(self, ops):
        st = _lib.stream_ptr()
        lib = self.lib
        for fn, args, _ in ops:
            if args is None:
                fn()
            elif fn(*args, st) < 0:
                raise _lib.DmcError(f"UNet backward: {lib.dmc_last_error().decode()}")

    def _cond_torch(self, P, t, y):
        """the conditioning table cond[n, :] of dmc_cond_desc as a differentiable PyTorch expression (models/unet.py:18-25,
        167-172, 256-260 and the time_mlp / label_proj of every ResidualBlock, :40-48, 65-68)"""
        net = self.net
        freqs = self.pk["freqs"]
        arg = t[:, None] * freqs[None, :]
        e = F.linear(torch.cat((arg.sin(), arg.cos()), dim=-1), P["time_embed.1.weight"], P["time_embed.1.bias"])
        temb = F.linear(F.silu(e), P["time_embed.3.weight"], P["time_embed.3.bias"])
        blocks = [n[: -len(".time_mlp.1.weight")] for n in self.cond_names if n.endswith(".time_mlp.1.weight")]
        wt = torch.cat([P[p + ".time_mlp.1.weight"] for p in blocks], dim=0)
        bt = torch.cat([P[p + ".time_mlp.1.bias"] + P[p + ".conv1.2.bias"] for p in blocks], dim=0)
        cond = F.linear(F.silu(temb), wt, bt)
        if self.has_y:
            yemb = F.embedding(torch.clamp(y, 0, net.num_classes), P["label_embed.weight"], padding_idx=0)
            wy = torch.cat([P[p + ".label_proj.1.weight"] for p in blocks], dim=0)
            cond = cond + F.linear(F.silu(yemb), wy)
        return cond

    def backward_segment(self, k, g, mask):
        seg = self.segs[k]
        if seg == "output":
            self.deps_static.copy_(g)
        if self.graphs is not None:
            self.graphs[seg].replay()
        else:
            self._run_ops(self.bwd[seg])
        flat = self.flat[seg].clone()
        grads = [flat[o: o + n].view(shp) for _, o, n, shp in self.layout[seg]]
        ar = self.net._grad_allreduce
        if ar is None:
            if k == 0:
                grads += self._cond_grads()
            return grads
        # native data-parallel mode (UNet.set_gradient_allreduce): ONE asynchronous NCCL all-reduce per UNet entry over its flat
        # gradient buffer, launched as soon as the entry's backward kernels are enqueued (it overlaps the earlier entries'
        # kernels); .grad is assigned by the engine after the last entry -- no per-parameter bucket copies, no hooks
        params = self._seg_param_objs[k]
        self._pending.append((self._allreduce_mean(flat, ar, True), params, grads, mask))
        if k == 0:
            cg = self._cond_grads()
            idx = [i for i, t in enumerate(cg) if t is not None]
            if idx:
                cflat = torch.cat([cg[i].reshape(-1) for i in idx])
                self._allreduce_mean(cflat, ar, False)
                off = 0
                for i in idx:
                    n = cg[i].numel()
                    cg[i] = cflat[off: off + n].view(cg[i].shape)
                    off += n
            grads += cg
            # under a DistributedDataParallel wrapper (UNet._ddp_params_and_buffers_to_ignore) one sentinel parameter stays with
            # DDP: its averaged gradient is returned through autograd below instead of being assigned here
            sentinel = getattr(self.net, "_ddp_sentinel", None)
            sent_p = self.net.get_parameter(sentinel) if sentinel else None
            fresh_p, acc_p, acc_g = [], [], []
            for work, ps, gs, mk in self._pending:
                work.wait()  # the current stream waits for the collective; the host does not
                for p_, g_, m_ in zip(ps, gs, mk):
                    if not m_ or g_ is None or p_ is sent_p:
                        continue
                    if p_.grad is None:
                        p_.grad = g_
                    else:
                        acc_p.append(p_.grad)
                        acc_g.append(g_)
            if acc_p:
                torch._foreach_add_(acc_p, acc_g)
            self._pending = []
            if sent_p is not None:
                out = [None] * len(mask)
                for i, p_ in enumerate(params):
                    if p_ is sent_p and mask[i]:
                        out[i] = grads[i]
                return out
        return [None] * len(mask)

    @staticmethod
    def _allreduce_mean(buf, group, async_op):
        """in-place mean over the group: ReduceOp.AVG where the backend has it (NCCL), else SUM of pre-divided values (gloo)"""
        import torch.distributed as dist

        if dist.get_backend(group) == "nccl":
            return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        buf.div_(dist.get_world_size(group))
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def _cond_grads(self):
        net = self.net
        dcond = torch.cat([dc for _, dc in self.dcond_parts], dim=1)
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.enable_grad():
                if self._cond_param_objs is None:
                    self._cond_param_objs = [net.get_parameter(n) for n in self.cond_names]
                P = {n: p.detach().float().requires_grad_(True) for n, p in zip(self.cond_names, self._cond_param_objs)}
                cond = self._cond_torch(P, self.fwd.t_stage, self.fwd.y_stage if self.has_y else None)
                grads = torch.autograd.grad(cond, [P[n] for n in self.cond_names], dcond, allow_unused=True)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return list(grads)

    def describe(self):
        """launch and FLOP counts of one training step (bench.py's gpu_launches / utilisation claims)"""
        per = {"conv_wgrad": 2, "gn_backward": 2, "torch_copy": 0}
        ops = [m for s in self.segs for _, _, m in self.bwd[s]]
        return dict(forward_launches=self.fwd.num_launches + 1,
                    backward_launches=sum(per.get(m["kind"], 1) for m in ops) + 1,
                    gemm_flops=float(self.fwd.gemm_flops) + sum(m["flops"] for m in ops))

    def time_ops(self, iters=3):
        """device time of every forward and backward op run alone (CUDA events on the launching stream); call after a step"""
        out = [dict(o, phase="forward") for o in self.fwd.time_ops(iters)]
        st = _lib.stream_ptr()
        for s in reversed(self.segs):
            for fn, args, m in self.bwd[s]:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for it in range(iters + 1):
                    if it == 1:
                        e0.record()
                    if args is None:
                        fn()
                    elif fn(*args, st) < 0:
                        raise _lib.DmcError(f"time_ops: {self.lib.dmc_last_error().decode()}")
                e1.record()
                e1.synchronize()
                out.append(dict(name=m["name"], kind=m["kind"], ms=e0.elapsed_time(e1) / iters, flops=m["flops"], bytes=0.0,
                                phase="backward"))
        return out

    def destroy(self):
        if getattr(self, "bplan", None) is not None:
            self.lib.dmc_plan_destroy(self.bplan)
            self.bplan = None
        self.fwd.destroy()

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass
