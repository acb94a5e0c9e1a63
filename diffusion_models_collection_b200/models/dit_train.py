"""Training step of the native DiT (missing item 7; reference: models/dit.py:111-132 under autograd, driven by
utils/trainer.py:249-255): ``loss = diffusion.p_losses(model, x0, t, y); loss.backward()``.

Where the FLOPs are (99.3 % of a block: qkv / out_proj / fc1 / fc2 and the attention core) the forward AND the backward run on
the hand-written sm_100a kernels behind the C ABI:

  linear forward       tcgen05 implicit GEMM (dmc_plan_add_conv, 1x1 over the token grid), bias in the epilogue
  linear input grad    the same kernel over dY with the weight matrix packed transposed
  linear weight grad   dmc_conv_wgrad (tcgen05, K = tokens, both operands straight from the token matrices), fp32, parameter layout
  linear bias grad     dmc_channel_sum
  attention            tcgen05 forward (P in tensor memory), dmc_attention_backward

The memory-bound glue between them is fused into three more native kernels (csrc/dit_train_ops.cu), each an autograd node too:

  gated residual + LayerNorm + adaLN modulate   x_out = x_in + gate * y;  h = LN(x_out) * (1 + scale) + shift -> bf16, written straight
                                                into the next GEMM's operand buffer; its backward also yields the per-image gradients of
                                                gate / shift / scale (one CTA per image, fixed summation order)
  GELU forward / backward                       between fc1 and fc2 (erf form)

What stays in PyTorch (differentiable, < 1 % of the FLOPs and of the bytes): patch embedding (K = 12) + pos_embed, the timestep /
label conditioning MLPs and the adaLN_modulation linears ([B, hidden] matrices), the 12-column output head and unpatchify.  With
`DMC_DIT_TRAIN_GLUE=torch` the LayerNorm / GELU / residual glue runs in PyTorch as well (the first form of this step: 39.9 ms
per step at batch 128 against the fused kernels' figure in DESIGN.md section 10; kept as a cross-check).  Dropout: the MLP dropouts
of the block (models/dit.py:97,100) are applied in ``.train()`` mode inside the GELU and residual kernels (counter-based mask, regenerated
by the backward kernels); the dropout on the attention probabilities inside nn.MultiheadAttention is not (the attention kernel has none).

There is no CPU fallback: every native op raises when the CUDA library or device is missing."""

from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn.functional as F

from .. import _lib


class _Linear(torch.autograd.Function):
    """y = x W^T + b on bf16 token matrices [B, L, cin] -> [B, L, cout]; `w` / `b` are the fp32 parameters (autograd edges only:
    the kernels read the engine's packed bf16 copies)."""

    @staticmethod
    def forward(ctx, eng, key, x, w, b):
        lay = eng.layers[key]
        if x.data_ptr() != lay.x.data_ptr():
            lay.x.copy_(x)
        eng.run_op(lay.fwd)
        ctx.eng, ctx.key, ctx.step = eng, key, eng.step_id
        return lay.y.view_as(lay.y)  # a fresh alias per call: autograd attaches this node to it, not to the static buffer

    @staticmethod
    def backward(ctx, dy):
        eng, lay = ctx.eng, ctx.eng.layers[ctx.key]
        eng.check_step(ctx.step)
        dyb = eng.dy_buf(lay.cout)
        if dy.data_ptr() != dyb.data_ptr():
            dyb.copy_(dy)
        eng.run_op(lay.dgrad)                      # dx = dy W
        eng.wgrad(lay, dyb)                        # dW = dy^T x  (fp32, [cout, cin])
        eng.bias_grad(lay, dyb)                    # db = sum over tokens
        # (clones: autograd may keep a returned tensor as .grad; the static buffers are rewritten by the next backward pass)
        dx = eng.dx_buf(lay.cin)
        return None, None, dx.view_as(dx), lay.dw.clone(), lay.db.clone()


class _Attention(torch.autograd.Function):
    """softmax(Q K^T / 8) V on the packed [B, L, 3C] bf16 projection (head dim 64)."""

    @staticmethod
    def forward(ctx, eng, i, qkv):
        blk = eng.attn[i]
        if qkv.data_ptr() != blk.qkv.data_ptr():
            blk.qkv.copy_(qkv)
        eng.run_op(blk.fwd)
        ctx.eng, ctx.i, ctx.step = eng, i, eng.step_id
        return blk.out.view_as(blk.out)

    @staticmethod
    def backward(ctx, dout):
        eng, blk = ctx.eng, ctx.eng.attn[ctx.i]
        eng.check_step(ctx.step)
        d = _lib.AttnBwdDesc()
        dout = dout.contiguous()
        d.qkv, d.out, d.dout, d.dqkv = blk.qkv.data_ptr(), blk.out.data_ptr(), dout.data_ptr(), eng.dqkv.data_ptr()
        d.B, d.L, d.heads, d.C = eng.B, eng.L, eng.net.num_heads, eng.hs
        if eng.lib.dmc_attention_backward(C.byref(d), _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT attention backward: {eng.lib.dmc_last_error().decode()}")
        eng.keep.append(dout)
        return None, None, eng.dqkv.view_as(eng.dqkv)


def _row_ptr(t):
    """(pointer, row stride in elements) of a [B, C] fp32 chunk of the adaLN table"""
    assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.float32
    return t.data_ptr(), t.stride(0)


class _GateLnMod(torch.autograd.Function):
    """x_out = x_in + gate * y (when y is given);  h = LayerNorm(x_out) * (1 + scale) + shift -> bf16 into `h_buf`.
    Returns (x_out, h) with y, h alone without."""

    @staticmethod
    def forward(ctx, eng, slot, h_buf, x_in, shift, scale, y=None, gate=None, drop=0.0):
        x_in = x_in.contiguous()
        d = _lib.DitGlmDesc()
        d.x_in, d.h = x_in.data_ptr(), h_buf.data_ptr()
        d.shift, ms = _row_ptr(shift)
        d.scale, ms2 = _row_ptr(scale)
        assert ms == ms2
        d.mod_stride, d.B, d.L, d.C, d.eps = ms, eng.B, eng.L, eng.hs, 1e-6
        x_out = x_in
        if y is not None:
            y = y.contiguous()
            x_out = eng.stream_buf(slot)
            d.y, d.x_out = y.data_ptr(), x_out.data_ptr()
            d.gate, d.gate_stride = _row_ptr(gate)
            d.drop_p, d.seed = drop, eng.op_seed(slot)
        if eng.lib.dmc_dit_gate_ln_mod(C.byref(d), _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT training: {eng.lib.dmc_last_error().decode()}")
        ctx.eng, ctx.step, ctx.has_y = eng, eng.step_id, y is not None
        ctx.drop, ctx.seed = (drop, eng.op_seed(slot)) if y is not None else (0.0, 0)
        ctx.keep = (x_out, y, gate, scale, shift)
        ctx.set_materialize_grads(False)
        h = h_buf.view_as(h_buf)
        return (x_out.view_as(x_out), h) if y is not None else h

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.eng
        eng.check_step(ctx.step)
        dx_out, dh = grads if ctx.has_y else (None, grads[0])
        x_out, y, gate, scale, shift = ctx.keep
        B, hs = eng.B, eng.hs
        if dh is None:
            raise RuntimeError("DiT training: the modulated LayerNorm output received no gradient")
        dh = dh.contiguous()
        dx_in = dx_out.contiguous() if dx_out is not None else torch.empty_like(x_out)  # (in place on the later layers' gradient)
        sums = torch.empty((3, B, hs), dtype=torch.float32, device=x_out.device)
        d = _lib.DitGlmBwdDesc()
        d.x, d.dh, d.dx_in = x_out.data_ptr(), dh.data_ptr(), dx_in.data_ptr()
        d.dx_out = dx_out.data_ptr() if dx_out is not None else None
        d.scale, d.mod_stride = _row_ptr(scale)
        d.dshift, d.dscale = sums[1].data_ptr(), sums[2].data_ptr()
        d.B, d.L, d.C, d.eps = B, eng.L, hs, 1e-6
        d.scratch = eng.glm_scratch.data_ptr()
        dy = None
        if ctx.has_y:
            dyb = eng.dy_buf(hs)
            d.y, d.dy, d.dgate = y.data_ptr(), dyb.data_ptr(), sums[0].data_ptr()
            d.gate, d.gate_stride = _row_ptr(gate)
            d.drop_p, d.seed = ctx.drop, ctx.seed
            dy = dyb.view_as(dyb)
        if eng.lib.dmc_dit_gate_ln_mod_backward(C.byref(d), _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT training: {eng.lib.dmc_last_error().decode()}")
        eng.keep.append((dh, dx_in))
        return None, None, None, dx_in, sums[1], sums[2], dy, (sums[0] if ctx.has_y else None), None


class _Gelu(torch.autograd.Function):
    """m = gelu(u) (erf form) from the fc1 output into the fc2 operand buffer"""

    @staticmethod
    def forward(ctx, eng, i, u, drop=0.0):
        m = eng.layers[f"blocks.{i}.fc2"].x
        u = u.contiguous()
        seed = eng.op_seed(1000 + i)
        if eng.lib.dmc_gelu_forward(u.data_ptr(), m.data_ptr(), u.numel(), drop, seed, _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT training: {eng.lib.dmc_last_error().decode()}")
        ctx.eng, ctx.step, ctx.u, ctx.drop, ctx.seed = eng, eng.step_id, u, drop, seed
        return m.view_as(m)

    @staticmethod
    def backward(ctx, dm):
        eng = ctx.eng
        eng.check_step(ctx.step)
        dm = dm.contiguous()
        du = eng.dy_buf(ctx.u.shape[-1])
        if eng.lib.dmc_gelu_backward(ctx.u.data_ptr(), dm.data_ptr(), du.data_ptr(), dm.numel(), ctx.drop, ctx.seed,
                                     _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT training: {eng.lib.dmc_last_error().decode()}")
        eng.keep.append(dm)
        return None, None, du.view_as(du), None


class _Layer:
    pass


class DiTTrainEngine:
    """static buffers + launch plans of one (device, batch size) signature"""

    def __init__(self, net, device, B):
        lib = self.lib = _lib.load()
        self.net, self.device, self.B = net, device, B
        hs = self.hs = net.hidden_size
        self.L = L = net.h_tokens * net.w_tokens
        hid = int(hs * net.mlp_ratio)
        if hs % 128 != 0 or hid % 128 != 0:
            raise NotImplementedError("native DiT training: hidden sizes must be multiples of 128 (weight-gradient tiles)")
        if hs // net.num_heads != 64 or L > 256:
            raise NotImplementedError("native DiT training: head dim 64 and at most 256 tokens (the attention backward kernel)")
        self.step_id = 0
        self.keep = []
        bf16, f32 = torch.bfloat16, torch.float32
        handle = C.c_void_p()
        _lib.check(lib.dmc_plan_create(C.byref(handle)), "dmc_plan_create")
        self.handle = handle
        self._dy, self._dx = {}, {}
        self.dqkv = torch.empty((B, L, 3 * hs), dtype=bf16, device=device)
        self._dy[3 * hs] = self.dqkv  # the attention backward writes the qkv GEMM's output gradient in place
        self.layers, self.attn = {}, []
        self.wbuf = {}
        shapes = dict(qkv=(hs, 3 * hs), out=(hs, hs), fc1=(hs, hid), fc2=(hid, hs))
        self.part = None
        part_need = 0
        for i in range(net.depth):
            for nm, (cin, cout) in shapes.items():
                key = f"blocks.{i}.{nm}"
                lay = _Layer()
                lay.key, lay.cin, lay.cout = key, cin, cout
                lay.x = torch.empty((B, L, cin), dtype=bf16, device=device)
                lay.y = torch.empty((B, L, cout), dtype=bf16, device=device)
                lay.w = torch.empty((cout, cin), dtype=bf16, device=device)      # forward operand  [N = cout, K = cin]
                lay.wt = torch.empty((cin, cout), dtype=bf16, device=device)     # input-gradient operand [N = cin, K = cout]
                lay.bias = torch.empty((cout,), dtype=f32, device=device)
                lay.dw = torch.empty((cout, cin), dtype=f32, device=device)
                lay.db = torch.empty((cout,), dtype=f32, device=device)
                lay.fwd = self._add_gemm(lay.x, cin, lay.w, cout, lay.bias, lay.y, key)
                lay.dgrad = self._add_gemm(self.dy_buf(cout), cout, lay.wt, cin, None, self.dx_buf(cin), key + ".dgrad")
                wd = _lib.WgradDesc()
                wd.x, wd.B, wd.Hin, wd.Win, wd.Cin, wd.Cout = lay.x.data_ptr(), B, net.h_tokens, net.w_tokens, cin, cout
                wd.dy = self.dy_buf(cout).data_ptr()
                wd.stride, wd.taps, wd.accumulate, wd.dw = 1, 1, 0, lay.dw.data_ptr()
                wd.splits = lib.dmc_conv_wgrad_splits(C.byref(wd))
                if wd.splits <= 0:
                    raise _lib.DmcError(f"DiT training: weight gradient of {key}: {lib.dmc_last_error().decode()}")
                part_need = max(part_need, wd.splits * cout * cin)
                lay.wd = wd
                self.layers[key] = lay
            blk = _Layer()
            blk.qkv = self.layers[f"blocks.{i}.qkv"].y
            blk.out = self.layers[f"blocks.{i}.out"].x   # the attention output IS the out_proj operand: no staging copy
            a = _lib.AttnDesc()
            a.qkv, a.out, a.B, a.L, a.heads, a.C, a.impl = blk.qkv.data_ptr(), blk.out.data_ptr(), B, L, net.num_heads, hs, 0
            blk.fwd = _lib.check(lib.dmc_plan_add_attention(handle, C.byref(a)), "attention")
            self.attn.append(blk)
        self.part = torch.empty((part_need,), dtype=f32, device=device)
        for lay in self.layers.values():
            lay.wd.partial = self.part.data_ptr()
        self.sum_scratch = torch.empty((B * max(3 * hs, hid),), dtype=f32, device=device)
        self._packed_ver = None
        self._stream = {}   # fp32 token stream after every gated residual add (kept for the LayerNorm backward)
        self.h_final = torch.empty((B, L, hs), dtype=bf16, device=device)
        self.glm_scratch = torch.empty((B * _lib.DIT_GLM_BWD_SLICES * 3 * hs,), dtype=f32, device=device)
        self.native_glue = os.environ.get("DMC_DIT_TRAIN_GLUE", "native") != "torch"

    def op_seed(self, slot):
        """dropout seed of one op of the current forward (base drawn once per forward from torch's CPU generator)"""
        return (self.seed_base + 0x9E3779B1 * (slot + 1)) & 0xFFFFFFFF

    def stream_buf(self, slot):
        if slot not in self._stream:
            self._stream[slot] = torch.empty((self.B, self.L, self.hs), dtype=torch.float32, device=self.device)
        return self._stream[slot]

    # ------------------------------------------------------------------------------------------------------------------
    def dy_buf(self, c):
        if c not in self._dy:
            self._dy[c] = torch.empty((self.B, self.L, c), dtype=torch.bfloat16, device=self.device)
        return self._dy[c]

    def dx_buf(self, c):
        if c not in self._dx:
            self._dx[c] = torch.empty((self.B, self.L, c), dtype=torch.bfloat16, device=self.device)
        return self._dx[c]

    def _add_gemm(self, src, cin, wmat, cout, bias, out, name):
        d = _lib.ConvDesc()
        d.B, d.Hin, d.Win, d.stride, d.up_phase = self.B, self.net.h_tokens, self.net.w_tokens, 1, -1
        d.nsrc = 1
        d.src[0], d.src_c[0], d.src_taps[0] = src.data_ptr(), cin, 1
        d.weight, d.Cout, d.Cout_pad, d.Ktot = wmat.data_ptr(), cout, cout, cin
        d.bias = bias.data_ptr() if bias is not None else None
        d.out_bf16 = out.data_ptr()
        return _lib.check(self.lib.dmc_plan_add_conv(self.handle, C.byref(d)), name)

    def run_op(self, idx):
        if self.lib.dmc_plan_run_op(self.handle, idx, _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT training: {self.lib.dmc_last_error().decode()}")

    def check_step(self, step):
        if step != self.step_id:
            raise RuntimeError("DiT backward: the activations of this forward have been overwritten by a later forward of the same "
                               "batch size (run backward() before the next training forward)")

    def wgrad(self, lay, dyb):
        assert dyb.data_ptr() == lay.wd.dy
        if self.lib.dmc_conv_wgrad(C.byref(lay.wd), _lib.stream_ptr()) < 0:
            raise _lib.DmcError(f"DiT training: weight gradient of {lay.key}: {self.lib.dmc_last_error().decode()}")

    def bias_grad(self, lay, dyb):
        rc = self.lib.dmc_channel_sum(dyb.data_ptr(), lay.db.data_ptr(), self.B, self.L, lay.cout, 0, 0,
                                      self.sum_scratch.data_ptr(), _lib.stream_ptr())
        if rc < 0:
            raise _lib.DmcError(f"DiT training: bias gradient of {lay.key}: {self.lib.dmc_last_error().decode()}")

    def pack(self):
        """bf16 GEMM operands (forward and transposed) from the fp32 parameters; redone when a parameter changed"""
        net = self.net
        ver = net._param_version()
        if ver == self._packed_ver:
            return
        names = dict(qkv=("attn.in_proj_weight", "attn.in_proj_bias"), out=("attn.out_proj.weight", "attn.out_proj.bias"),
                     fc1=("mlp.0.weight", "mlp.0.bias"), fc2=("mlp.3.weight", "mlp.3.bias"))
        with torch.no_grad():
            for i in range(net.depth):
                for nm, (wn, bn) in names.items():
                    lay = self.layers[f"blocks.{i}.{nm}"]
                    w = net.get_parameter(f"blocks.{i}.{wn}")
                    lay.w.copy_(w)
                    lay.wt.copy_(w.t())
                    lay.bias.copy_(net.get_parameter(f"blocks.{i}.{bn}"))
        self._packed_ver = ver

    # ------------------------------------------------------------------------------------------------------------------
    def forward(self, x, t, y):
        """models/dit.py:263-295 with the four linears and the attention core of every block on the native kernels"""
        net, hs, B, L = self.net, self.hs, self.B, self.L
        self.pack()
        self.step_id += 1
        self.keep = []
        self.seed_base = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if (net.training and net.dropout > 0) else 0
        P = dict(net.named_parameters())
        p = net.patch_size
        # PatchEmbed + pos_embed (models/dit.py:23-27, 271)
        tok = F.conv2d(x, P["x_embedder.proj.weight"], P["x_embedder.proj.bias"], stride=p).flatten(2).transpose(1, 2) + P["pos_embed"]
        # TimestepEmbedder (models/dit.py:42-55): the reference's own expression for the frequencies (CPU arange, then moved)
        half = 128
        freqs = net._ensure_packed(self.device)["freqs"]
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        c = F.linear(F.silu(F.linear(emb, P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"])),
                     P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"])
        if net.num_classes is not None and y is not None:
            c = c + F.embedding(torch.clamp(y, 0, net.num_classes), P["y_embedder.embedding_table.weight"], padding_idx=0)
        sc = F.silu(c)
        drop = float(net.dropout) if net.training else 0.0
        if not self.native_glue:
            return self._forward_torch_glue(tok, sc, P, drop)
        x, y_prev, g_prev = tok.contiguous(), None, None
        for i in range(net.depth):
            b = f"blocks.{i}"
            mod = F.linear(sc, P[b + ".adaLN_modulation.1.weight"], P[b + ".adaLN_modulation.1.bias"])
            sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
            hbuf = self.layers[b + ".qkv"].x
            if y_prev is None:
                h = _GateLnMod.apply(self, 2 * i, hbuf, x, sh1, sc1)
            else:
                x, h = _GateLnMod.apply(self, 2 * i, hbuf, x, sh1, sc1, y_prev, g_prev, drop)
            qkv = _Linear.apply(self, b + ".qkv", h, P[b + ".attn.in_proj_weight"], P[b + ".attn.in_proj_bias"])
            ao = _Attention.apply(self, i, qkv)
            y1 = _Linear.apply(self, b + ".out", ao, P[b + ".attn.out_proj.weight"], P[b + ".attn.out_proj.bias"])
            x, h = _GateLnMod.apply(self, 2 * i + 1, self.layers[b + ".fc1"].x, x, sh2, sc2, y1, g1)
            u = _Linear.apply(self, b + ".fc1", h, P[b + ".mlp.0.weight"], P[b + ".mlp.0.bias"])
            m = _Gelu.apply(self, i, u, drop)                 # GELU + Dropout (models/dit.py:96-97)
            y2 = _Linear.apply(self, b + ".fc2", m, P[b + ".mlp.3.weight"], P[b + ".mlp.3.bias"])
            y_prev, g_prev = y2, g2                            # (the Dropout after fc2, :100, is applied by the next fused kernel)
        mod = F.linear(sc, P["final_layer.adaLN_modulation.1.weight"], P["final_layer.adaLN_modulation.1.bias"])
        shift, scale = mod.chunk(2, dim=-1)
        x, h = _GateLnMod.apply(self, 2 * net.depth, self.h_final, x, shift, scale, y_prev, g_prev, drop)
        return self._head(h.float(), P)

    def _head(self, h, P):
        """final linear (12 columns) + unpatchify (models/dit.py:150, 249-261)"""
        net, B = self.net, self.B
        p = net.patch_size
        o = F.linear(h, P["final_layer.linear.weight"], P["final_layer.linear.bias"])
        co = net.out_channels
        o = o.reshape(B, net.h_tokens, net.w_tokens, p, p, co)
        return torch.einsum("nhwpqc->nchpwq", o).reshape(B, co, net.h_tokens * p, net.w_tokens * p)

    def _forward_torch_glue(self, tok, sc, P, drop):
        """the first form of this step: LayerNorm / modulate / GELU / gated residual as differentiable PyTorch"""
        net, hs = self.net, self.hs
        for i in range(net.depth):
            b = f"blocks.{i}"
            mod = F.linear(sc, P[b + ".adaLN_modulation.1.weight"], P[b + ".adaLN_modulation.1.bias"])
            sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
            h = (F.layer_norm(tok, (hs,), eps=1e-6) * (1 + sc1[:, None]) + sh1[:, None]).to(torch.bfloat16)
            qkv = _Linear.apply(self, b + ".qkv", h, P[b + ".attn.in_proj_weight"], P[b + ".attn.in_proj_bias"])
            ao = _Attention.apply(self, i, qkv)
            y1 = _Linear.apply(self, b + ".out", ao, P[b + ".attn.out_proj.weight"], P[b + ".attn.out_proj.bias"])
            tok = tok + g1[:, None] * y1.float()
            h = (F.layer_norm(tok, (hs,), eps=1e-6) * (1 + sc2[:, None]) + sh2[:, None]).to(torch.bfloat16)
            u = _Linear.apply(self, b + ".fc1", h, P[b + ".mlp.0.weight"], P[b + ".mlp.0.bias"])
            m = F.gelu(u.float())
            if drop > 0:
                m = F.dropout(m, drop, True)
            y2 = _Linear.apply(self, b + ".fc2", m.to(torch.bfloat16), P[b + ".mlp.3.weight"], P[b + ".mlp.3.bias"]).float()
            if drop > 0:
                y2 = F.dropout(y2, drop, True)
            tok = tok + g2[:, None] * y2
        mod = F.linear(sc, P["final_layer.adaLN_modulation.1.weight"], P["final_layer.adaLN_modulation.1.bias"])
        shift, scale = mod.chunk(2, dim=-1)
        h = F.layer_norm(tok, (hs,), eps=1e-6) * (1 + scale[:, None]) + shift[:, None]
        return self._head(h, P)

    def destroy(self):
        if getattr(self, "handle", None) is not None:
            self.lib.dmc_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass
