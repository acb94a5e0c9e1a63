"""UNet denoiser drop-in: same constructor, attributes, parameter names/shapes (state_dict keys) and
``forward(x, t, y=None)`` contract as /root/reference/models/unet.py:126-292 -- but the forward is a *plan* of
hand-written sm_100a kernels replayed through the C ABI (include/dmc.h):

  conditioning table (1 fused group of launches for time_embed + all 22 time_mlp/label_proj projections)
  stem conv (fp32 NCHW -> bf16 NHWC), GroupNorm statistics / apply(+SiLU, + skip concat as one tensor),
  every 3x3 / 1x1 convolution as a tcgen05 implicit GEMM fed by TMA with bias + conditioning + residual
  + the 1x1 shortcut (extra K columns) fused, flash-style attention, head conv writing fp32 NCHW eps.

There is no PyTorch / CPU fallback: the forward raises when the CUDA library or device is missing.
Parameters stay ordinary ``nn.Parameter``s (optimizer, EMA, DDP and checkpoints see the reference layout); the
bf16 K-major packed copies the kernels read are rebuilt whenever a parameter version changes.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import math
import os
from typing import Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..synth import unet_block_structure

_ALIGN = 1024


def _round_up(v, a=_ALIGN):
    return (v + a - 1) // a * a


def phase_weights(w, phase):
    """Weights of one phase of "nearest-2x upsample then conv3x3" (models/unet.py:118-120) as a 2x2 convolution on the
    low-resolution tensor: output pixel (2i+ph, 2j+pw) only ever sees low-res rows {i-1+ph, i+ph} and columns
    {j-1+pw, j+pw}; the 3x3 taps that land on the same low-res pixel are summed (fp32) -- 2.25x fewer MACs.
    w: [Cout, Cin, 3, 3] fp32 -> [Cout, 4*Cin] fp32, tap-major (r, s), channel-minor; phase = 2*ph + pw."""
    ph, pw = phase >> 1, phase & 1
    groups = ([[0], [1, 2]], [[0, 1], [2]])
    taps = []
    for r in range(2):
        for s_ in range(2):
            acc = 0
            for kh in groups[ph][r]:
                for kw in groups[pw][s_]:
                    acc = acc + w[:, :, kh, kw]
            taps.append(acc)
    return torch.stack(taps, dim=1).reshape(w.shape[0], -1)


class _Node(nn.Module):
    """Anonymous container: gives parameters the reference's dotted names."""


def _register(root: nn.Module, name: str, tensor: torch.Tensor):
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], nn.Parameter(tensor))


class _Arena:
    """First-fit offset allocator with coalescing; activations are reused as soon as their last reader has been
    enqueued (stream order makes that safe)."""

    def __init__(self):
        self.free = [(0, 1 << 62)]
        self.peak = 0

    def alloc(self, nbytes):
        nbytes = _round_up(nbytes)
        for i, (off, size) in enumerate(self.free):
            if size >= nbytes:
                if size == nbytes:
                    self.free.pop(i)
                else:
                    self.free[i] = (off + nbytes, size - nbytes)
                self.peak = max(self.peak, off + nbytes)
                return off, nbytes
        raise MemoryError

    def release(self, blk):
        off, size = blk
        self.free.append((off, size))
        self.free.sort()
        merged = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


class _Act:
    """A bf16 NHWC activation [B, H, W, C] living in the plan workspace.  Offsets are assigned AFTER the op list is final
    (_PlanBuilder._assign_memory): a tensor lives from the first op that touches it to the last one."""

    __slots__ = ("blk", "lo", "C", "H", "W", "stats", "slots", "nbytes", "split", "raw")

    def __init__(self, nbytes, Cc, H, W, split=False):
        self.blk, self.lo, self.C, self.H, self.W = None, None, Cc, H, W  # lo: low parts in split-bf16 mode
        self.stats, self.slots = None, 0  # GroupNorm partial sums [B, slots, C/8, 2] fp32 (workspace block)
        self.nbytes, self.split = nbytes, split
        self.raw = True  # False: nothing reads the tensor itself (only its fused normalised versions): not stored


class _PlanBuilder:
    """Walks the UNet topology once (dry run -> workspace size, then for real) and records the op list."""

    def __init__(self, net: "UNet", nimg, x_batch, has_y, uniform_t, conv_impl, split=False, keep=False, drop_p=0.0):
        self.net, self.B, self.x_batch = net, nimg, x_batch
        self.has_y, self.uniform_t, self.conv_impl = has_y, uniform_t, conv_impl
        self.split = split  # split-bf16 ("bf16x3") accuracy mode: every activation is a (hi, lo) pair of bf16 tensors
        # training forward: every activation stays alive for the backward pass (no workspace reuse), Upsample materialises the
        # upsampled tensor (its weight gradient reads it), ResidualBlock dropout (models/unet.py:53) runs in the conv2 GroupNorm pass
        self.keep, self.drop_p = keep, drop_p
        self.arena = _Arena()
        self.ops = []  # (kind, dict)

    # -- workspace ---------------------------------------------------------------------------------
    def act(self, Cc, H, W):
        return _Act(self.B * H * W * Cc * 2, Cc, H, W, self.split)

    def free(self, a: _Act):
        """(lifetimes are derived from the final op list in _assign_memory; kept as a no-op marker of the topology)"""

    def _alloc_stats(self, a: _Act, slots):
        a.slots = slots

    def stats_of(self, a: _Act):
        """GroupNorm partial sums of `a`: written by the producing conv's epilogue when it has one, else by a
        stand-alone pass (stem output; every tensor when the CUDA-core debug conv is selected)."""
        if a.slots == 0:
            self._alloc_stats(a, (a.H * a.W + 127) // 128)
            self.ops.append(("gn_stats", dict(src=a)))
        return a

    # -- op list post-passes -----------------------------------------------------------------------------
    @staticmethod
    def _touched(kind, o):
        """activations an op reads or writes"""
        if kind in ("stem", "stem_cols", "gn_stats", "head", "head_taps"):
            return [o.get("out") or o.get("src")]
        if kind == "gn_apply":
            return list(o["srcs"]) + [o["out"]]
        if kind == "conv":
            t = list(o["srcs"])
            if o["residual"] is not None:
                t.append(o["residual"])
            if o["out"] is not None:
                t.append(o["out"])
            t += [v["dst"] for v in o.get("gn", ())]
            if o.get("aff") is not None:
                t.append(o["aff"])
            return t
        if kind == "gn_coeff":
            return [o["src"], o["out"]]
        if kind == "attention":
            return [o["qkv"], o["out"]]
        if kind == "upsample":
            return [o["src"], o["out"]]
        return []

    def _fuse_groupnorm(self):
        """Moves GroupNorm(+SiLU) passes into the epilogue of the convolution that produces their input (north_star:
        "GroupNorm+SiLU ... fused into the conv prologue/epilogue"; models/unet.py:35-36,51-52,80,238-239).  A gn_apply op
        disappears when EVERY source is the output of one tcgen05 convolution whose tile geometry can complete the statistics
        in its epilogue (dmc_conv_gn_supported), no statistics group straddles two sources of a concat (models/unet.py:284) and
        the producer has fewer than two fused versions already; that convolution then writes the normalised tensor (its channel
        slice of the concat) itself.  What stays a stand-alone pass: inputs produced by the stem, by the four Upsample phase
        convolutions (four launches write one tensor) and the 384-channel concats whose 48-channel groups straddle."""
        lib = _lib.load()
        producer = {}
        writers = {}
        for i, (kind, o) in enumerate(self.ops):
            if kind == "conv" and o["out"] is not None:
                writers[id(o["out"])] = writers.get(id(o["out"]), 0) + 1
                producer[id(o["out"])] = i
            elif kind in ("stem", "stem_cols", "upsample", "attention", "gn_apply", "gn_coeff"):
                writers[id(o["out"])] = writers.get(id(o["out"]), 0) + 99
        readers = {}  # how many ops read each tensor
        for kind, o in self.ops:
            rd = []
            if kind == "conv":
                rd = list(o["srcs"]) + ([o["residual"]] if o["residual"] is not None else [])
            elif kind == "gn_apply":
                rd = list(o["srcs"])
            elif kind in ("head", "head_taps", "upsample", "gn_coeff"):
                rd = [o["src"]]
            elif kind == "attention":
                rd = [o["qkv"]]
            for a_ in rd:
                readers[id(a_)] = readers.get(id(a_), 0) + 1
        drop = set()
        for gi, (kind, g) in enumerate(self.ops):
            if kind != "gn_apply" or g["drop_p"] > 0:
                continue
            ct = sum(s_.C for s_ in g["srcs"])
            gsz = ct // 8
            if gsz not in (16, 32, 64):
                continue
            plan, off, ok = [], 0, True
            for s_ in g["srcs"]:
                pi = producer.get(id(s_))
                if pi is None or writers.get(id(s_)) != 1 or off % gsz or s_.C % gsz:
                    ok = False
                    break
                po = self.ops[pi][1]
                pend = sum(1 for q_ in plan if q_[0] is po)
                kblocks = sum(t_ * a_.C for t_, a_ in zip(po["taps"], po["srcs"])) // 64
                if kblocks < self.net.fuse_gn_min_kblocks:  # short K loop: the two-pass epilogue would outlast the tile's MMAs
                    ok = False
                    break
                if self.net.fuse_gn_scope == "conv1" and (readers.get(id(s_), 0) != 1 or po["residual"] is not None):
                    # measured (profiles/r02_conv_gn_micro_*): an epilogue that ALSO writes the raw tensor (residual / skip /
                    # shortcut readers) costs more than the stand-alone pass it replaces; conv1 -> conv2.0 pairs do not
                    ok = False
                    break
                if (po["out_nchw"] or po["up_phase"] >= 0 or not po["want_stats"] or len(po.get("gn", ())) + pend >= 2 or
                        not lib.dmc_conv_gn_supported(self.B, s_.H, s_.W, s_.C, max([gsz] + [v["gsize"] for v in po.get("gn", ())]))):
                    ok = False
                    break
                plan.append((po, off, s_))
                off += s_.C
            if not ok:
                continue
            for po, off, s_ in plan:
                po.setdefault("gn", []).append(dict(dst=g["out"], coff=off, prefix=g["prefix"], gsize=gsz, silu=g["silu"]))
            drop.add(gi)
        if not drop:
            return
        self.ops = [op for i, op in enumerate(self.ops) if i not in drop]
        # a raw output nothing reads any more (e.g. conv1 of a ResidualBlock: only its normalised version is consumed) is not stored
        read = set()
        for kind, o in self.ops:
            if kind == "conv":
                read.update(id(s_) for s_ in o["srcs"])
                if o["residual"] is not None:
                    read.add(id(o["residual"]))
            elif kind == "gn_apply":
                read.update(id(s_) for s_ in o["srcs"])
            elif kind in ("gn_stats", "gn_coeff", "head", "head_taps", "upsample"):
                read.add(id(o["src"]))
            elif kind == "attention":
                read.add(id(o["qkv"]))
        for kind, o in self.ops:
            if kind == "conv" and o.get("gn") and o["out"] is not None and id(o["out"]) not in read:
                o["out"].raw = False

    def _assign_memory(self):
        """workspace offsets from lifetimes: every activation (with its statistics block) is allocated right before the first op
        that touches it and released right after the last one (stream order makes reuse safe); training plans keep everything"""
        first, last, acts = {}, {}, {}
        for i, (kind, o) in enumerate(self.ops):
            for a in self._touched(kind, o):
                first.setdefault(id(a), i)
                last[id(a)] = i
                acts[id(a)] = a
        begin, end = {}, {}
        for k, a in acts.items():
            begin.setdefault(first[k], []).append(a)
            end.setdefault(last[k], []).append(a)
        for i in range(len(self.ops)):
            for a in begin.get(i, ()):
                if a.raw:
                    a.blk = self.arena.alloc(a.nbytes)
                    if a.split:
                        a.lo = self.arena.alloc(a.nbytes)
                if a.slots:
                    a.stats = self.arena.alloc(self.B * a.slots * (a.C // 8) * 2 * 4)
            if self.keep:
                continue
            for a in end.get(i, ()):
                for blk in (a.blk, a.lo, a.stats):
                    if blk is not None:
                        self.arena.release(blk)

    # -- layers --------------------------------------------------------------------------------------
    def gn_apply(self, srcs, prefix, silu):
        H, W = srcs[0].H, srcs[0].W
        out = self.act(sum(s.C for s in srcs), H, W)
        for s in srcs:
            self.stats_of(s)
        drop = self.drop_p if prefix.endswith(".conv2.0") else 0.0
        self.ops.append(("gn_apply", dict(srcs=list(srcs), prefix=prefix, silu=silu, out=out, drop_p=drop)))
        return out

    def conv(self, srcs, taps, wname, Cout, H, W, stride=1, bias=None, cond_col=None, residual=None, out_nchw=False,
             up_phase=-1, out=None, want_stats=True, sc_slice=None, out_f32=False, aff=None):
        Ho, Wo = H // stride, W // stride
        if up_phase >= 0:
            Ho, Wo = 2 * H, 2 * W
        if out is None and not out_nchw:
            out = self.act(Cout, Ho, Wo)
        if out is not None and want_stats and self.conv_impl == 0 and out.slots == 0:
            ppi = (H // stride) * (W // stride)  # iteration pixels per image: one slot per 32-pixel epilogue warp
            self._alloc_stats(out, max(1, ppi // 32) * (4 if up_phase >= 0 else 1))
        parts = ["hi"] * len(srcs)
        if self.split:  # hi*W_hi + lo*W_hi + hi*W_lo as three K segments (weights packed [W_hi | W_hi | W_lo])
            assert len(srcs) == 1, "split-bf16 mode: one logical source per launch"
            srcs, taps, parts = [srcs[0]] * 3, [taps[0]] * 3, ["hi", "lo", "hi"]
        self.ops.append(("conv", dict(srcs=list(srcs), taps=list(taps), parts=parts, wname=wname, Cout=Cout, H=H, W=W,
                                      stride=stride, bias=bias, cond_col=cond_col, residual=residual, out=out,
                                      out_nchw=out_nchw, up_phase=up_phase, want_stats=want_stats, sc_slice=sc_slice,
                                      out_f32=out_f32, aff=aff)))
        return out

    def fused_head(self, h):
        """the fused output-head kernel covers this geometry (and we are not in the split-bf16 / debug modes)"""
        if self.split or self.keep or self.conv_impl != 0 or not self.net.fuse_head:
            return False
        return h.C in (64, 128) and h.W in (16, 32, 64) and self.net.out_channels <= 8 and (h.C // 8) % 8 == 0

    def resblock(self, srcs, prefix, cout, cond_col):
        cin = sum(s.C for s in srcs)
        H, W = srcs[0].H, srcs[0].W
        a1 = self.gn_apply(srcs, prefix + ".conv1.0", 1)
        h1 = self.conv([a1], [9], prefix + ".conv1", cout, H, W, cond_col=cond_col)  # bias folded into cond table
        self.free(a1)
        a2 = self.gn_apply([h1], prefix + ".conv2.0", 1)
        self.free(h1)
        if cin != cout and self.split:
            # accuracy mode: conv2, then the 1x1 shortcut of every raw source as its own launch chained through the residual
            out = self.conv([a2], [9], prefix + ".conv2", cout, H, W, bias=prefix + ".conv2+sc", want_stats=False)
            off = 0
            for i, s_ in enumerate(srcs):
                nxt = self.conv([s_], [1], f"{prefix}.sc.{i}", cout, H, W, residual=out, want_stats=(i == len(srcs) - 1),
                                sc_slice=(off, s_.C))
                off += s_.C
                self.free(out)
                out = nxt
        elif cin != cout:  # 1x1 shortcut fused as extra K columns over the raw (un-normalised) inputs
            out = self.conv([a2] + list(srcs), [9] + [1] * len(srcs), prefix + ".conv2+sc", cout, H, W,
                            bias=prefix + ".conv2+sc")
        else:
            assert len(srcs) == 1
            out = self.conv([a2], [9], prefix + ".conv2", cout, H, W, bias=prefix + ".conv2", residual=srcs[0])
        self.free(a2)
        return out

    def attnblock(self, x, prefix):
        if (self.net.fuse_norm_qkv and not (self.split or self.keep or self.conv_impl != 0) and (x.C // 8) in (16, 32, 64) and
                _lib.load().dmc_conv_affine_supported(self.B, x.H, x.W, x.C, 3 * x.C)):
            # GroupNorm (no activation) of the block input applied to the A operand of the qkv GEMM inside the GEMM kernel
            # (models/unet.py:80-81,86-87): a tiny per-image coefficient kernel replaces the whole normalisation pass
            co = _Act(self.B * x.C * 8, x.C, 1, 1)  # fp32 [B, C, 2] (scale, shift)
            self.ops.append(("gn_coeff", dict(src=self.stats_of(x), prefix=prefix + ".norm", out=co)))
            qkv = self.conv([x], [1], prefix + ".qkv", 3 * x.C, x.H, x.W, bias=prefix + ".qkv", want_stats=False, aff=co)
        else:
            an = self.gn_apply([x], prefix + ".norm", 0)
            qkv = self.conv([an], [1], prefix + ".qkv", 3 * x.C, x.H, x.W, bias=prefix + ".qkv", want_stats=False)
            self.free(an)
        ao = self.act(x.C, x.H, x.W)
        self.ops.append(("attention", dict(qkv=qkv, out=ao, L=x.H * x.W, C=x.C)))
        self.free(qkv)
        out = self.conv([ao], [1], prefix + ".proj", x.C, x.H, x.W, bias=prefix + ".proj", residual=x)
        self.free(ao)
        return out

    def build(self):
        net = self.net
        H, W = net._hw
        mc = net.model_channels
        down, middle, up, out_ch = unet_block_structure(net._cfg())
        self.ops.append(("cond", {}))
        if self.net.stem_gemm and not (self.split or self.keep or self.conv_impl != 0) and 18 * net.in_channels <= 64 and mc % 32 == 0:
            # input conv on the tensor cores: gather (tap, channel) columns of the fp32 input as a bf16 (hi, lo) pair, then ONE
            # 1x1 GEMM with K = 64 (bias, GroupNorm partial sums from its epilogue: no stand-alone statistics pass)
            xc = self.act(64, H, W)
            self.ops.append(("stem_cols", dict(out=xc)))
            h = self.conv([xc], [1], "input_conv.cols", mc, H, W, bias="input_conv")
        else:
            h = self.act(mc, H, W)
            self.ops.append(("stem", dict(out=h)))
        hs = [h]
        col = 0

        def run_entry(prefix, layers, h, srcs_first):
            nonlocal col
            cur = h
            for j, l in enumerate(layers):
                p = f"{prefix}.{j}"
                if l[0] == "res":
                    srcs = srcs_first if j == 0 else [cur]
                    new = self.resblock(srcs, p, l[2], col)
                    col += l[2]
                elif l[0] == "attn":
                    new = self.attnblock(cur, p)
                elif l[0] == "down":
                    new = self.conv([cur], [9], p + ".conv", l[1], cur.H, cur.W, stride=2, bias=p + ".conv")
                elif l[0] == "up" and self.net.upsample_phases and not self.keep:
                    # four 2x2 phase convolutions on the low-res tensor, scattered into one output (no upsampled copy)
                    new = self.act(l[1], 2 * cur.H, 2 * cur.W)
                    for ph in range(4):
                        self.conv([cur], [4], f"{p}.conv.ph{ph}", l[1], cur.H, cur.W, bias=p + ".conv", up_phase=ph,
                                  out=new)
                elif l[0] == "up":
                    upb = self.act(cur.C, 2 * cur.H, 2 * cur.W)
                    self.ops.append(("upsample", dict(src=cur, out=upb)))
                    new = self.conv([upb], [9], p + ".conv", l[1], 2 * cur.H, 2 * cur.W, bias=p + ".conv")
                    self.free(upb)
                else:
                    continue
                if j > 0:  # an intermediate of this entry (not a skip): dead once consumed
                    self.free(cur)
                cur = new
            return cur

        for i, layers in enumerate(down):
            h = run_entry(f"down_blocks.{i}", layers, h, [h])
            hs.append(h)
        h = run_entry("middle_block", [l for l in middle], h, [h])  # its input is the last skip: stays alive
        for i, layers in enumerate(up):
            skip = hs.pop()
            new = run_entry(f"up_blocks.{i}", layers, h, [h, skip])
            self.free(h)     # previous up / middle output: only this block's concat read it
            self.free(skip)  # popped skip: dead after the concat
            h = new
        assert not hs
        if self.fused_head(h):
            # GroupNorm + SiLU + conv 3x3 -> fp32 NCHW in ONE kernel (one read of h)
            self.ops.append(("head", dict(src=self.stats_of(h))))
            self.free(h)
        elif self.net.head_taps and not (self.split or self.keep or self.conv_impl != 0) and 9 * net.out_channels <= 64:
            # conv3x3 C -> 3 as ONE 1x1 GEMM with 27 (tap, cout) columns (every input pixel read once, not 3 - 9 times) and a
            # 9-tap fp32 gather that writes the NCHW eps
            a = self.gn_apply([h], "output.0", 1)
            ypitch = _round_up(9 * net.out_channels, 32)
            yact = _Act(self.B * H * W * ypitch * 4, ypitch, H, W)  # fp32 [B, H, W, ypitch]
            self.conv([a], [1], "output.2.taps", ypitch, H, W, out=yact, want_stats=False, out_f32=True)
            self.ops.append(("head_taps", dict(src=yact, ypitch=ypitch)))
        else:
            a = self.gn_apply([h], "output.0", 1)
            self.free(h)
            self.conv([a], [9], "output.2", net.out_channels, H, W, bias="output.2", out_nchw=True, want_stats=False)
            self.free(a)
        self.ncols = col
        if self.net.fuse_groupnorm and not (self.split or self.keep or self.conv_impl != 0):
            self._fuse_groupnorm()
        self._assign_memory()
        return self


class UNet(nn.Module):
    """UNet model for diffusion (reference: models/unet.py:126-292); see the module docstring."""

    graph_capturable = True  # a forward is a fixed, allocation-free, sync-free launch list (samplers capture it)
    # activations of larger batches are processed in chunks of this many images (workspace ~1.7 MB per image: 7 GB); 4096 instead of
    # round 1's 2048 is bit-identical (batch invariance) and 1.3 % faster per image (run 39: the small layers fill the chip better)
    max_images_per_launch = int(os.environ.get("DMC_MAX_IMAGES_PER_LAUNCH", "4096"))
    upsample_phases = os.environ.get("DMC_UPSAMPLE_PHASES", "1") != "0"  # Upsample as four 2x2 phase convolutions
    # output GroupNorm + SiLU + conv as one mma.sync kernel: opt-in -- measured 0.725 ms against 0.20 + 0.53 ms for the
    # two-kernel path at 2048 images (legacy mma.sync issues ~1 per 80 clk per SM sub-partition on sm_100a), no gain
    fuse_head = os.environ.get("DMC_FUSED_HEAD", "0") != "0"
    # output conv3x3 (C -> 3) as a 1x1 GEMM over 27 (tap, cout) columns + a 9-tap gather (dmc_head_taps_desc); 0: the padded-N
    # 3x3 implicit GEMM (which re-reads its input 4.5x through the slab path: 0.55 ms against ~0.2 ms at 2048 images)
    head_taps = os.environ.get("DMC_HEAD_TAPS", "1") != "0"
    # input conv3x3 (3 -> C) as a gather of 27 (tap, channel) columns (bf16 hi + lo of the fp32 input) + a 1x1 tcgen05 GEMM whose
    # epilogue also writes the GroupNorm partial sums.  OPT-IN: measured (run 12, 2048 images) 0.45 ms gather + 0.25 ms GEMM
    # against 0.49 + 0.10 ms for the fp32 CUDA-core stem kernel + the stand-alone statistics pass, and the bf16 weights of the
    # first layer cost parity (whole-model eps error at batch 16: 6.8e-3 against 6.2e-3) -- the exact fp32 stem stays the default
    stem_gemm = os.environ.get("DMC_STEM_GEMM", "0") != "0"
    # AttentionBlock: GroupNorm (no activation) of the block input applied to the A operand of the qkv 1x1 GEMM by two otherwise
    # idle warps of that kernel (bit-identical to the stand-alone pass, which disappears).  OPT-IN: measured (run 13, 2048 images)
    # the two transform warps are ~4x too slow for the operand stream -- qkv 0.26 -> 1.10 ms per 16x16 layer against the 0.09 ms
    # pass it removes; a transform needs a warpgroup or more, which the kernel's register budget (384 threads x 168) does not have
    fuse_norm_qkv = os.environ.get("DMC_FUSE_NORM_QKV", "0") != "0"
    # GroupNorm(+SiLU) applied by the epilogue of the convolution that produces the tensor (see _PlanBuilder._fuse_groupnorm);
    # DMC_FUSE_GN=0 keeps the stand-alone gn_apply passes everywhere (A/B measurements, tests)
    fuse_groupnorm = os.environ.get("DMC_FUSE_GN", "1") != "0"
    # ... only into convolutions whose K loop has at least this many 64-element blocks: the epilogue of tile k (two passes +
    # the statistics hand-shake) overlaps the MMAs of tile k+1, which a 1x1 convolution (4 - 8 blocks) finishes long before
    fuse_gn_min_kblocks = int(os.environ.get("DMC_FUSE_GN_MIN_KB", "30"))
    # "conv1": only where the producer's raw output then disappears (conv1 -> conv2.0 of a ResidualBlock); "all": wherever
    # the geometry allows (every mode is parity-tested; "conv1" is what the measurements favour)
    fuse_gn_scope = os.environ.get("DMC_FUSE_GN_SCOPE", "conv1")

    def __init__(self, image_size: Tuple[int, int] = (32, 32), in_channels=3, model_channels=128, out_channels=3,
                 num_res_blocks=2, attention_resolutions=(16, 8), dropout=0.1, channel_mult=(1, 2, 2, 2),
                 num_classes=None, use_attention=True):
        super().__init__()
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.num_classes = num_classes
        self.use_attention = use_attention
        self._hw = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
        self._uniform_t = False
        self._plans = {}
        self._train_engines = {}
        self._plist = None
        self._grad_allreduce = None  # process group of the native data-parallel mode (set_gradient_allreduce)
        self._ddp_sentinel = None    # name of the parameter left to a DistributedDataParallel wrapper (see the property below)
        self._packed = None
        self._packed_version = self._packed_ids = None
        self._init_parameters()

    # ------------------------------------------------------------------------------------------------
    def _cfg(self):
        return dict(image_size=self._hw, in_channels=self.in_channels, model_channels=self.model_channels,
                    out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                    attention_resolutions=tuple(self.attention_resolutions), channel_mult=tuple(self.channel_mult),
                    use_attention=self.use_attention)

    def _init_parameters(self):
        """Registers every tensor of the reference's state_dict with PyTorch's default init laws
        (Kaiming-uniform a=sqrt(5) conv/linear, GroupNorm 1/0, Embedding N(0,1) with the padding row zeroed)."""
        mc, temb = self.model_channels, self.model_channels * 4

        def conv(name, cout, cin, k):
            bound = 1.0 / math.sqrt(cin * k * k)
            _register(self, name + ".weight", torch.empty(cout, cin, k, k).uniform_(-bound, bound))
            _register(self, name + ".bias", torch.empty(cout).uniform_(-bound, bound))

        def linear(name, cout, cin, bias=True):
            bound = 1.0 / math.sqrt(cin)
            _register(self, name + ".weight", torch.empty(cout, cin).uniform_(-bound, bound))
            if bias:
                _register(self, name + ".bias", torch.empty(cout).uniform_(-bound, bound))

        def gn(name, c):
            _register(self, name + ".weight", torch.ones(c))
            _register(self, name + ".bias", torch.zeros(c))

        linear("time_embed.1", temb, mc)
        linear("time_embed.3", temb, temb)
        if self.num_classes is not None:
            w = torch.randn(self.num_classes + 1, temb)
            w[0].zero_()
            _register(self, "label_embed.weight", w)
        conv("input_conv", mc, self.in_channels, 3)

        def entry(prefix, layers):
            for j, l in enumerate(layers):
                p = f"{prefix}.{j}"
                if l[0] == "res":
                    gn(p + ".conv1.0", l[1])
                    conv(p + ".conv1.2", l[2], l[1], 3)
                    linear(p + ".time_mlp.1", l[2], temb)
                    if self.num_classes is not None:
                        linear(p + ".label_proj.1", l[2], temb, bias=False)
                    gn(p + ".conv2.0", l[2])
                    conv(p + ".conv2.3", l[2], l[2], 3)
                    if l[1] != l[2]:
                        conv(p + ".shortcut", l[2], l[1], 1)
                elif l[0] == "attn":
                    gn(p + ".norm", l[1])
                    conv(p + ".qkv", 3 * l[1], l[1], 1)
                    conv(p + ".proj", l[1], l[1], 1)
                elif l[0] in ("down", "up"):
                    conv(p + ".conv", l[1], l[1], 3)

        down, middle, up, out_ch = unet_block_structure(self._cfg())
        for i, layers in enumerate(down):
            entry(f"down_blocks.{i}", layers)
        entry("middle_block", middle)
        for i, layers in enumerate(up):
            entry(f"up_blocks.{i}", layers)
        gn("output.0", out_ch)
        conv("output.2", self.out_channels, out_ch, 3)

    # ------------------------------------------------------------------------------------------------
    # weight packing (one-time / on parameter change; plain torch ops -- not on the hot path)
    # ------------------------------------------------------------------------------------------------
    def __getstate__(self):
        """copy.deepcopy(model) / torch.save(model): the launch plans, training engines and packed operands are caches bound to
        native handles and device pointers -- a copy starts without them and rebuilds them on its first forward"""
        st = self.__dict__.copy()
        st.update(_plans={}, _train_engines={}, _plist=None, _packed=None, _packed_version=None, _packed_ids=None,
                  _grad_allreduce=None, _ddp_sentinel=None, _dmc_graph_token=None)
        return st

    def parameters(self, recurse: bool = True):
        """nn.Module.parameters() walks the module tree (one small container per dotted-name component: ~1.5 ms for the 357
        tensors); the trainer calls it several times per step (optimizer, clip_grad_norm_), so the flat list is cached."""
        if not recurse:
            return super().parameters(recurse)
        if self._plist is None:
            self._plist = list(super().parameters())
        return iter(self._plist)

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self._plist = None
        return r

    def _param_ids(self):
        return tuple([p.data_ptr() for p in self.parameters()])

    def _param_version(self):
        return tuple([p._version for p in self.parameters()])

    def _ensure_packed(self, device, training=False):
        """packed copies of the parameters.  A change of parameter *storage* (load onto another device, .to(), a new
        tensor assigned) rebuilds everything and drops the plans; a change of parameter *values* only (optimizer step, EMA
        copy_, load_state_dict) re-packs in place, so plans, TMA descriptors and CUDA graphs stay valid.  `training` skips
        what only the sampling plans read (Upsample phase weights)."""
        ids, ver = (str(device), self._param_ids()), self._param_version()
        pk = self._packed
        if pk is not None and self._packed_ids == ids and (self._packed_version == ver or not pk["w3"]):
            if pk["ver_main"] != ver:
                self._refresh_packed(pk, phases=False)
                pk["ver_main"] = ver
            if not training and pk["ver_phases"] != ver:
                self._refresh_packed(pk, phases=True)
                pk["ver_phases"] = ver
            self._packed_version = ver
            return pk
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in self.state_dict().items()}
        down, middle, up, out_ch = unet_block_structure(self._cfg())
        temb = self.model_channels * 4
        W, Bv = {}, {}  # packed bf16 [Cout_pad, K] matrices and fp32 bias vectors by logical name

        def pack3(w):  # [Cout, Cin, 3, 3] -> [Cout, 9*Cin], tap-major (r, s), channel-minor
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

        wt, bt, wy = [], [], []

        def entry(prefix, layers):
            for j, l in enumerate(layers):
                p = f"{prefix}.{j}"
                if l[0] == "res":
                    W[p + ".conv1"] = pack3(sd[p + ".conv1.2.weight"])
                    wt.append(sd[p + ".time_mlp.1.weight"])
                    bt.append(sd[p + ".time_mlp.1.bias"] + sd[p + ".conv1.2.bias"])
                    if self.num_classes is not None:
                        wy.append(sd[p + ".label_proj.1.weight"])
                    if l[1] != l[2]:
                        W[p + ".conv2+sc"] = torch.cat([pack3(sd[p + ".conv2.3.weight"]),
                                                        sd[p + ".shortcut.weight"].reshape(l[2], l[1])], dim=1)
                        Bv[p + ".conv2+sc"] = sd[p + ".conv2.3.bias"] + sd[p + ".shortcut.bias"]
                    else:
                        W[p + ".conv2"] = pack3(sd[p + ".conv2.3.weight"])
                        Bv[p + ".conv2"] = sd[p + ".conv2.3.bias"]
                elif l[0] == "attn":
                    W[p + ".qkv"] = sd[p + ".qkv.weight"].reshape(3 * l[1], l[1])
                    Bv[p + ".qkv"] = sd[p + ".qkv.bias"]
                    W[p + ".proj"] = sd[p + ".proj.weight"].reshape(l[1], l[1])
                    Bv[p + ".proj"] = sd[p + ".proj.bias"]
                elif l[0] in ("down", "up"):
                    W[p + ".conv"] = pack3(sd[p + ".conv.weight"])
                    Bv[p + ".conv"] = sd[p + ".conv.bias"]
                    if l[0] == "up":
                        for ph in range(4):
                            W[f"{p}.conv.ph{ph}"] = phase_weights(sd[p + ".conv.weight"], ph)

        for i, layers in enumerate(down):
            entry(f"down_blocks.{i}", layers)
        entry("middle_block", middle)
        for i, layers in enumerate(up):
            entry(f"up_blocks.{i}", layers)
        if 18 * self.in_channels <= 64:
            # the input convolution as a 1x1 GEMM over gathered (tap, channel) columns (dmc_stem_cols_desc): [W | W | 0], the
            # second copy multiplies the bf16 rounding remainders of the input
            ws = pack3(sd["input_conv.weight"])
            W["input_conv.cols"] = torch.cat([ws, ws, ws.new_zeros(ws.shape[0], 64 - 2 * ws.shape[1])], dim=1)
            Bv["input_conv"] = sd["input_conv.bias"]
        head = pack3(sd["output.2.weight"])
        W["output.2"] = torch.cat([head, head.new_zeros(32 - head.shape[0], head.shape[1])], dim=0)
        Bv["output.2"] = sd["output.2.bias"]
        if 9 * self.out_channels <= 64:
            # the head as a 1x1 GEMM over (tap, cout) columns + a 9-tap gather (dmc_head_taps_desc): row tap * Cout + co
            hw = sd["output.2.weight"]
            ht = hw.permute(2, 3, 0, 1).reshape(9 * self.out_channels, hw.shape[1])
            W["output.2.taps"] = torch.cat([ht, ht.new_zeros(_round_up(ht.shape[0], 32) - ht.shape[0], ht.shape[1])], dim=0)

        # one bf16 blob for all GEMM weights (1 KiB aligned slices: TMA needs 16 B), one fp32 blob for the rest
        offs, total = {}, 0
        for k, v in W.items():
            offs[k] = total
            total += _round_up(v.numel() * 2)
        wblob = torch.zeros(total // 2, dtype=torch.bfloat16, device=device)
        for k, v in W.items():
            wblob[offs[k] // 2: offs[k] // 2 + v.numel()] = v.reshape(-1).to(torch.bfloat16)
        wt_all = torch.cat(wt, dim=0).contiguous()
        bt_all = torch.cat(bt, dim=0).contiguous()
        ytab = None
        if self.num_classes is not None:
            wy_all = torch.cat(wy, dim=0)
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                ytab = (torch.nn.functional.silu(sd["label_embed.weight"]) @ wy_all.t()).contiguous()
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        half = self.model_channels // 2
        # models/unet.py:20-22 -- exponent divisor (half - 1); computed with the reference's own expression
        freqs = torch.exp(torch.arange(half, device=device) * -(math.log(10000) / (half - 1))).float().contiguous()
        # split-bf16 mode packs [W_hi | W_hi | W_lo] per logical convolution on demand (see _split_weight)
        wlog = dict(W)
        for i_, layers_ in [(f"down_blocks.{i}", l) for i, l in enumerate(down)] + [("middle_block", middle)] + \
                [(f"up_blocks.{i}", l) for i, l in enumerate(up)]:
            for j, l in enumerate(layers_):
                if l[0] == "res" and l[1] != l[2]:
                    wlog[f"{i_}.{j}.conv2"] = pack3(sd[f"{i_}.{j}.conv2.3.weight"])
                    wlog[f"{i_}.{j}.sc"] = sd[f"{i_}.{j}.shortcut.weight"].reshape(l[2], l[1])
        self._packed = dict(
            wlog=wlog, w3={},
            sd=sd, wblob=wblob, woffs=offs, wshape={k: tuple(v.shape) for k, v in W.items()},
            bias={k: v.contiguous() for k, v in Bv.items()}, wt_all=wt_all, bt_all=bt_all, ytab=ytab, freqs=freqs,
            ncols=wt_all.shape[0], temb=temb, ver_main=ver, ver_phases=ver, refresh=None,
        )
        self._packed_version, self._packed_ids = ver, ids
        for pl in list(self._plans.values()) + list(self._train_engines.values()):
            pl.destroy()
        self._plans, self._train_engines = {}, {}
        return self._packed

    def _refresh_packed(self, pk, phases):
        """re-pack the bf16 GEMM weights / fused biases / conditioning tables in place from the current parameter values"""
        sd, device = pk["sd"], pk["wblob"].device
        with torch.no_grad():
            for k, v in self.state_dict().items():  # no-op aliases for fp32 parameters already on the device
                if sd[k].data_ptr() != v.data_ptr():
                    sd[k].copy_(v)
            if pk["refresh"] is None:
                pk["refresh"] = self._refresh_lists(pk)
            r = pk["refresh"]
            if phases:
                for dst, wkey, ph in r["phases"]:
                    dst.copy_(phase_weights(sd[wkey], ph))
                if "output.2.taps" in pk["woffs"]:  # (sampling plans only, like the phase weights)
                    hw = sd["output.2.weight"]
                    rows, K = pk["wshape"]["output.2.taps"]
                    o_ = pk["woffs"]["output.2.taps"] // 2
                    pk["wblob"][o_: o_ + rows * K].view(rows, K)[: 9 * self.out_channels].copy_(
                        hw.permute(2, 3, 0, 1).reshape(9 * self.out_channels, hw.shape[1]))
                return
            if device.type == "cuda":  # all GEMM weights in one launch of the native pack kernel
                if r["table"] is None:
                    r["table"] = _lib.pack_table(r["items"], device)
                _lib.check(_lib.load().dmc_pack_weights(r["table"][0].data_ptr(), r["table"][1], _lib.stream_ptr()), "dmc_pack_weights")
            else:
                torch._foreach_copy_(r["w_dst"], r["w_src"])
            torch._foreach_copy_(r["b_dst"], r["b_src"])
            if r["b2_dst"]:
                torch._foreach_add_(r["b2_dst"], r["b2_src"])
            torch.cat(r["wt"], dim=0, out=pk["wt_all"])
            torch.cat(r["bt"], dim=0, out=pk["bt_all"])
            torch._foreach_add_(r["bt_dst"], r["bt_add"])
            if pk["ytab"] is not None:
                prev = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                try:
                    torch.mm(torch.nn.functional.silu(sd["label_embed.weight"]), torch.cat(r["wy"], dim=0).t(), out=pk["ytab"])
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = prev

    def _refresh_lists(self, pk):
        """(destination view, source view) pairs of the in-place re-pack: one strided converting copy per weight"""
        sd = pk["sd"]
        down, middle, up, out_ch = unet_block_structure(self._cfg())
        r = dict(w_dst=[], w_src=[], b_dst=[], b_src=[], b2_dst=[], b2_src=[], wt=[], bt=[], bt_dst=[], bt_add=[], wy=[],
                 phases=[], items=[], table=None)

        def item(dst, w, taps):  # forward layout: dst[co, tap * cin + ci], dst possibly a column range of a wider matrix
            it = _lib.PackItem()
            it.src, it.dst = w.data_ptr(), dst.data_ptr()
            it.cout, it.cin_total, it.ci0, it.cin, it.taps, it.mode = w.shape[0], w.shape[1], 0, w.shape[1], taps, 0
            it.ld, it.col0, it.cpad = dst.stride(0), 0, 0
            r["items"].append(it)

        def wview(name):
            rows, K = pk["wshape"][name]
            o = pk["woffs"][name] // 2
            return pk["wblob"][o: o + rows * K].view(rows, K)

        def w3x3(dst, w):  # dst [Cout, 9*Cin] (possibly a column range of a wider matrix) <- [Cout, Cin, 3, 3]
            co, ci = w.shape[0], w.shape[1]
            r["w_dst"].append(dst.view(co, 3, 3, ci))
            r["w_src"].append(w.permute(0, 2, 3, 1))
            item(dst, w, 9)

        def w1x1(dst, w):
            r["w_dst"].append(dst)
            r["w_src"].append(w.view(w.shape[0], w.shape[1]))
            item(dst, w, 1)

        col = 0

        def entry(prefix, layers):
            nonlocal col
            for j, l in enumerate(layers):
                p = f"{prefix}.{j}"
                if l[0] == "res":
                    w3x3(wview(p + ".conv1"), sd[p + ".conv1.2.weight"])
                    r["wt"].append(sd[p + ".time_mlp.1.weight"])
                    r["bt"].append(sd[p + ".time_mlp.1.bias"])
                    r["bt_dst"].append(pk["bt_all"][col: col + l[2]])
                    r["bt_add"].append(sd[p + ".conv1.2.bias"])
                    col += l[2]
                    if self.num_classes is not None:
                        r["wy"].append(sd[p + ".label_proj.1.weight"])
                    if l[1] != l[2]:
                        v = wview(p + ".conv2+sc")
                        w3x3(v[:, : 9 * l[2]], sd[p + ".conv2.3.weight"])
                        w1x1(v[:, 9 * l[2]:], sd[p + ".shortcut.weight"])
                        r["b_dst"].append(pk["bias"][p + ".conv2+sc"])
                        r["b_src"].append(sd[p + ".conv2.3.bias"])
                        r["b2_dst"].append(pk["bias"][p + ".conv2+sc"])
                        r["b2_src"].append(sd[p + ".shortcut.bias"])
                    else:
                        w3x3(wview(p + ".conv2"), sd[p + ".conv2.3.weight"])
                        r["b_dst"].append(pk["bias"][p + ".conv2"])
                        r["b_src"].append(sd[p + ".conv2.3.bias"])
                elif l[0] == "attn":
                    for nm in (".qkv", ".proj"):
                        w1x1(wview(p + nm), sd[p + nm + ".weight"])
                        r["b_dst"].append(pk["bias"][p + nm])
                        r["b_src"].append(sd[p + nm + ".bias"])
                elif l[0] in ("down", "up"):
                    w3x3(wview(p + ".conv"), sd[p + ".conv.weight"])
                    r["b_dst"].append(pk["bias"][p + ".conv"])
                    r["b_src"].append(sd[p + ".conv.bias"])
                    if l[0] == "up":
                        for ph in range(4):
                            r["phases"].append((wview(f"{p}.conv.ph{ph}"), p + ".conv.weight", ph))

        for i, layers in enumerate(down):
            entry(f"down_blocks.{i}", layers)
        entry("middle_block", middle)
        for i, layers in enumerate(up):
            entry(f"up_blocks.{i}", layers)
        if "input_conv.cols" in pk["woffs"]:
            k27 = 9 * self.in_channels
            v = wview("input_conv.cols")
            w3x3(v[:, :k27], sd["input_conv.weight"])
            w3x3(v[:, k27: 2 * k27], sd["input_conv.weight"])
            r["b_dst"].append(pk["bias"]["input_conv"])
            r["b_src"].append(sd["input_conv.bias"])
        w3x3(wview("output.2")[: self.out_channels], sd["output.2.weight"])
        r["b_dst"].append(pk["bias"]["output.2"])
        r["b_src"].append(sd["output.2.bias"])
        return r

    @staticmethod
    def _split_weight(pk, wname, sc_slice=None):
        """bf16 [rows, 3K] = [W_hi | W_hi | W_lo] of one logical convolution (split-bf16 mode), cached per packing"""
        key = (wname, sc_slice)
        w3 = pk["w3"].get(key)
        if w3 is None:
            if sc_slice is not None:  # the columns of one raw source of a 1x1 shortcut
                w = pk["wlog"][wname.rsplit(".sc.", 1)[0] + ".sc"][:, sc_slice[0]: sc_slice[0] + sc_slice[1]]
            else:
                w = pk["wlog"][wname]
            w = w.float()
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            w3 = pk["w3"][key] = torch.cat([hi, hi, lo], dim=1).contiguous()
        return w3

    # ------------------------------------------------------------------------------------------------
    # plans
    # ------------------------------------------------------------------------------------------------
    precision = os.environ.get("DMC_PRECISION", "bf16")  # "bf16" (default) or "bf16x3": split-bf16, fp32-level accuracy

    def _get_plan(self, device, nimg, x_batch, has_y, uniform_t):
        if self.precision not in ("bf16", "bf16x3"):
            raise ValueError(f"UNet.precision must be 'bf16' or 'bf16x3', got {self.precision!r}")
        key = (str(device), nimg, x_batch, has_y, uniform_t, self.precision)
        pl = self._plans.get(key)
        if pl is None:
            pk = self._ensure_packed(device)
            pl = _UNetPlan(self, pk, device, nimg, x_batch, has_y, uniform_t)
            self._plans[key] = pl
        return pl

    def _run(self, x, t, y, cfg):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise _lib.DmcError("UNet.forward: CUDA tensors only -- the B200 hot path has no CPU / PyTorch fallback")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("the native UNet does not differentiate with respect to its image input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if cfg:
                raise NotImplementedError("forward_cfg is a sampling call: wrap it in torch.no_grad()")
            return self._run_train(x, t, y)
        _lib.load()
        device = x.device
        self._ensure_packed(device)
        Hh, Ww = self._hw
        if x.dim() != 4 or x.shape[1] != self.in_channels or tuple(x.shape[2:]) != (Hh, Ww):
            raise ValueError(f"UNet.forward: expected x of shape [B, {self.in_channels}, {Hh}, {Ww}], got {tuple(x.shape)}")
        B = x.shape[0]
        if t.shape[0] != B or (y is not None and y.shape[0] != B):
            raise ValueError("UNet.forward: t / y batch size mismatch")
        x = x.contiguous().float()
        t = t.to(device=device, dtype=torch.long).contiguous()
        has_y = self.num_classes is not None and y is not None
        if cfg and not has_y:
            raise ValueError("forward_cfg needs a conditional model and labels")
        if has_y:
            y = y.to(device=device, dtype=torch.long).contiguous()
        mult = 2 if cfg else 1
        out = torch.empty((mult * B, self.out_channels, Hh, Ww), device=device, dtype=torch.float32)
        cb = max(1, min(B, self.max_images_per_launch // mult))
        with torch.cuda.device(device):
            for s in range(0, B, cb):
                n = min(cb, B - s)
                pl = self._get_plan(device, mult * n, n, has_y, bool(self._uniform_t))
                pl.run(x[s:s + n], t[s:s + n], y[s:s + n] if has_y else None, cfg,
                       out[s:s + n], out[B + s:B + s + n] if cfg else None)
        return out

    def _run_train(self, x, t, y):
        """training forward (autograd enabled): native forward with every activation kept, wired into autograd as a chain of
        per-entry nodes whose backward runs the native backward kernels (models/unet_train.py)"""
        from .unet_train import UNetTrainEngine

        _lib.load()
        device = x.device
        Hh, Ww = self._hw
        if x.dim() != 4 or x.shape[1] != self.in_channels or tuple(x.shape[2:]) != (Hh, Ww):
            raise ValueError(f"UNet.forward: expected x of shape [B, {self.in_channels}, {Hh}, {Ww}], got {tuple(x.shape)}")
        B = x.shape[0]
        if t.shape[0] != B or (y is not None and y.shape[0] != B):
            raise ValueError("UNet.forward: t / y batch size mismatch")
        if self.precision != "bf16":
            raise NotImplementedError("native UNet training runs in the bf16 mode only")
        x = x.detach().contiguous().float()
        t = t.to(device=device, dtype=torch.long).contiguous()
        has_y = self.num_classes is not None and y is not None
        if has_y:
            y = y.to(device=device, dtype=torch.long).contiguous()
        drop_p = float(self.dropout) if self.training else 0.0
        with torch.cuda.device(device):
            self._ensure_packed(device, training=True)  # before the engine lookup: a storage change drops the engines too
            key = (str(device), B, has_y, drop_p)
            eng = self._train_engines.get(key)
            if eng is None:
                eng = self._train_engines[key] = UNetTrainEngine(self, device, B, has_y, drop_p)
            eng.forward(x, t, y)
            return eng.attach()

    @property
    def module(self):
        """DDP-compatibility alias, only in the native data-parallel mode: the reference's trainer reaches the wrapped model as
        `self.model.module` when distributed (utils/trainer.py:127,159-162,193,335)"""
        if self._grad_allreduce is None:
            raise AttributeError("module")
        return self

    # The one parameter DistributedDataParallel keeps managing when it wraps this model (see below): DDP refuses a module
    # without a single managed parameter, and its reducer wants that parameter's gradient from autograd every backward pass.
    _DDP_SENTINEL = "input_conv.bias"

    @property
    def _ddp_params_and_buffers_to_ignore(self):
        """`DDP(model)` of the reference's trainer (utils/trainer.py:58-61), unmodified, at the speed of the native all-reduce.
        DistributedDataParallel.__init__ asks the module it wraps for this attribute (torch/nn/parallel/distributed.py:
        `if hasattr(module, "_ddp_params_and_buffers_to_ignore")`) -- its documented way to leave parameters to someone else.  When the
        reader is a DDP constructor over a group of two or more ranks, the UNet switches on `set_gradient_allreduce(ddp.process_group)`
        (rank 0's parameters are broadcast like DDP would) and hands DDP every parameter name but one small sentinel: gradients are
        then averaged by one NCCL all-reduce per UNet entry straight from the engine's flat buffers instead of DDP's per-parameter
        bucket copies (714 small kernels per step: 6.6x against 7.7x at 8 GPUs, profiles/r02_train_n8_*_run9.json); the sentinel's
        already averaged gradient goes through autograd and DDP's reducer (a 512-byte all-reduce of identical values), which keeps
        DDP's own bookkeeping consistent.  `DMC_DDP_NATIVE=0` restores stock DDP over all parameters; any other reader sees no such
        attribute.  Like set_gradient_allreduce there is no `no_sync()`: every backward pass is averaged."""
        import sys

        if os.environ.get("DMC_DDP_NATIVE", "1") == "0":
            raise AttributeError("_ddp_params_and_buffers_to_ignore")
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel

        frame = sys._getframe(1)
        owner = frame.f_locals.get("self")
        if not (isinstance(owner, DistributedDataParallel) and frame.f_code.co_name == "__init__"
                and dist.is_available() and dist.is_initialized()):
            raise AttributeError("_ddp_params_and_buffers_to_ignore")
        group = getattr(owner, "process_group", None)
        if group is None:
            group = dist.group.WORLD
        names = [n for n, _ in self.named_parameters()]
        if dist.get_world_size(group) < 2 or self._DDP_SENTINEL not in names:
            raise AttributeError("_ddp_params_and_buffers_to_ignore")
        if self._grad_allreduce is not group:
            self.set_gradient_allreduce(group)
        self._ddp_sentinel = self._DDP_SENTINEL
        return [n for n in names if n != self._DDP_SENTINEL]

    def set_gradient_allreduce(self, process_group=None, enabled=True, broadcast_parameters=True):
        """Native data-parallel training WITHOUT the DistributedDataParallel wrapper: every backward pass averages the gradients
        over `process_group` (default: the world) with one asynchronous NCCL all-reduce per UNet entry, issued straight from
        the engine's flat per-entry gradient buffers while the earlier entries' backward kernels run, and assigns `.grad`
        itself.  It replaces `DDP(model)` (utils/trainer.py:58-61) -- use one or the other, not both: DDP's per-parameter
        bucket copies (714 small kernels and their hooks per step) cost 1.5 - 3 ms of a 15 ms step.  Gradient accumulation
        over several backward passes averages every pass (there is no no_sync())."""
        import torch.distributed as dist

        self._ddp_sentinel = None
        if not enabled:
            self._grad_allreduce = None
            return self
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("set_gradient_allreduce: torch.distributed is not initialised")
        group = process_group if process_group is not None else dist.group.WORLD
        if broadcast_parameters:  # what DDP does when it wraps the model: rank 0's parameters everywhere
            with torch.no_grad():
                for p in self.parameters():
                    # p.detach() shares the version counter with p (p.data does not): ranks that packed their bf16
                    # operands before the broadcast re-pack them on the next forward
                    dist.broadcast(p.detach(), src=dist.get_global_rank(group, 0), group=group)
        self._grad_allreduce = group
        return self

    @contextlib.contextmanager
    def uniform_timesteps(self):
        """Promise that every t[n] of the calls inside the block is the same value (what the samplers do): the time
        part of the conditioning table is then computed once per forward instead of once per image."""
        prev = self._uniform_t
        self._uniform_t = True
        try:
            yield self
        finally:
            self._uniform_t = prev

    def forward(self, x, t, y=None):
        """eps = UNet(x, t, y): x fp32 [B, C, H, W], t int64 [B], y int64 [B] in [0, num_classes] (0 = null) or None."""
        return self._run(x, t, y, cfg=False)

    def forward_cfg(self, x, t, y):
        """(eps(x, t, y), eps(x, t, 0)) computed as ONE batch of 2B images (the reference runs two full forwards,
        diffusion/ddim.py:300-301; samples never interact inside the model, so this is the same arithmetic)."""
        B = x.shape[0]
        out = self._run(x, t, y, cfg=True)
        return out[:B], out[B:]

    def launches_per_forward(self, batch, cfg=False):
        """kernel launches (+ memsets) of OUR library for one model call on `batch` images (all chunks), plus the
        fused scheduler kernel that follows it -- bench.py's gpu_launches claim"""
        mult = 2 if cfg else 1
        cb = max(1, min(batch, self.max_images_per_launch // mult))
        n = 0
        for s0 in range(0, batch, cb):
            nn_ = min(cb, batch - s0)
            n += self.plan_info(nn_, cfg=cfg).num_launches
        return n + 1

    def plan_info(self, batch, cfg=False, device=None):
        """(plan, ops) for introspection / profiling: builds (or reuses) the plan for a batch of `batch` images."""
        device = device or next(self.parameters()).device
        _lib.load()
        mult = 2 if cfg else 1
        self._ensure_packed(torch.device(device))
        return self._get_plan(torch.device(device), mult * batch, batch, self.num_classes is not None,
                              bool(self._uniform_t))


def layer_seed(op_index):
    """dropout seed constant of the GroupNorm pass that is op `op_index` of the forward plan"""
    return ((op_index + 1) * 0x9E3779B1) & 0xFFFFFFFF


class _UNetPlan:
    """Owns one dmc_plan (C side), its workspace and the small staging tensors of one (batch, mode) signature."""

    def __init__(self, net: UNet, pk, device, nimg, x_batch, has_y, uniform_t, keep=False, drop_p=0.0, seed_dev=None):
        lib = _lib.load()
        self.lib = lib
        self.nimg, self.x_batch, self.has_y = nimg, x_batch, has_y
        conv_impl = int(os.environ.get("DMC_DEBUG_CONV_IMPL", "0"))
        split = net.precision == "bf16x3" and not keep
        b = _PlanBuilder(net, nimg, x_batch, has_y, uniform_t, conv_impl, split=split, keep=keep, drop_p=drop_p).build()
        self.builder = b if keep else None  # the training engine derives the backward pass from the recorded forward ops
        self.op_index = []                  # plan op index of every builder op
        self.workspace_bytes = b.arena.peak
        Hh, Ww = net._hw
        R = 1 if uniform_t else nimg
        temb, ncols = pk["temb"], pk["ncols"]
        assert ncols == b.ncols
        self.ws = torch.empty(max(b.arena.peak, 16), dtype=torch.uint8, device=device)
        self.cond = torch.empty((nimg, ncols), dtype=torch.float32, device=device)
        self.cond_scratch = torch.empty((2 * R * temb + R * ncols,), dtype=torch.float32, device=device)
        self.t_stage = torch.zeros((nimg,), dtype=torch.long, device=device)
        self.y_stage = torch.zeros((nimg,), dtype=torch.long, device=device) if has_y else None
        self.eps = torch.empty((nimg, net.out_channels, Hh, Ww), dtype=torch.float32, device=device)
        self.x_keepalive = None
        handle = C.c_void_p()
        _lib.check(lib.dmc_plan_create(C.byref(handle)), "dmc_plan_create")
        self.handle = handle
        self.op_names = []
        wsp = self.ws.data_ptr()
        sd = pk["sd"]

        def ap(a):
            return wsp + a.blk[0]

        def ap_lo(a):
            return (wsp + a.lo[0]) if a.lo is not None else None

        def add(fn, desc, name):
            idx = _lib.check(fn(handle, C.byref(desc)), name)
            self.op_names.append(name)
            return idx

        self.stem_idx = self.cond_idx = self.head_idx = -1
        self.gn_counters = []
        self.fused_gn = sum(len(o.get("gn", ())) for kind, o in b.ops if kind == "conv")  # normalised versions written by convs
        for kind, o in b.ops:
            self.op_index.append(len(self.op_names))
            if kind == "cond":
                d = _lib.CondDesc()
                d.t, d.y = self.t_stage.data_ptr(), (self.y_stage.data_ptr() if has_y else None)
                d.B, d.uniform_t = nimg, 1 if uniform_t else 0
                d.num_classes = net.num_classes if net.num_classes is not None else 0
                d.half, d.temb, d.ncols = net.model_channels // 2, temb, ncols
                d.freqs = pk["freqs"].data_ptr()
                d.w1, d.b1 = sd["time_embed.1.weight"].data_ptr(), sd["time_embed.1.bias"].data_ptr()
                d.w2, d.b2 = sd["time_embed.3.weight"].data_ptr(), sd["time_embed.3.bias"].data_ptr()
                d.wt_all, d.bt_all = pk["wt_all"].data_ptr(), pk["bt_all"].data_ptr()
                d.ytab = pk["ytab"].data_ptr() if (has_y and pk["ytab"] is not None) else None
                d.scratch, d.cond = self.cond_scratch.data_ptr(), self.cond.data_ptr()
                self.cond_idx = add(lib.dmc_plan_add_cond, d, "cond")
            elif kind == "stem":
                d = _lib.StemDesc()
                d.x, d.x_batch, d.B = self.eps.data_ptr(), x_batch, nimg  # x is re-bound on every run
                d.Cin, d.H, d.W, d.Cout = net.in_channels, Hh, Ww, net.model_channels
                d.weight, d.bias = sd["input_conv.weight"].data_ptr(), sd["input_conv.bias"].data_ptr()
                d.out, d.out_lo = ap(o["out"]), ap_lo(o["out"])
                self.stem_idx = add(lib.dmc_plan_add_stem, d, "input_conv")
            elif kind == "gn_coeff":
                a = o["src"]
                d = _lib.GnCoeffDesc()
                d.stats, d.stats_slots, d.B, d.HW, d.C, d.groups = wsp + a.stats[0], a.slots, nimg, a.H * a.W, a.C, 8
                d.gamma, d.beta = sd[o["prefix"] + ".weight"].data_ptr(), sd[o["prefix"] + ".bias"].data_ptr()
                d.eps, d.out = 1e-5, ap(o["out"])
                add(lib.dmc_plan_add_gn_coeff, d, o["prefix"] + ".coeff")
            elif kind == "stem_cols":
                d = _lib.StemColsDesc()
                d.x, d.x_batch, d.B = self.eps.data_ptr(), x_batch, nimg  # x is re-bound on every run
                d.Cin, d.H, d.W, d.out = net.in_channels, Hh, Ww, ap(o["out"])
                self.stem_idx = add(lib.dmc_plan_add_stem_cols, d, "input_conv.gather")
            elif kind == "gn_stats":
                a = o["src"]
                d = _lib.GnStatsDesc()
                d.src, d.B, d.HW, d.C, d.stats = ap(a), nimg, a.H * a.W, a.C, wsp + a.stats[0]
                d.src_lo = ap_lo(a)
                add(lib.dmc_plan_add_gn_stats, d, "gn_stats")
            elif kind == "gn_apply":
                d = _lib.GnApplyDesc()
                d.nsrc = len(o["srcs"])
                for i, s in enumerate(o["srcs"]):
                    d.src[i], d.src_c[i], d.stats[i], d.stats_slots[i] = ap(s), s.C, wsp + s.stats[0], s.slots
                    d.src_lo[i] = ap_lo(s)
                d.B, d.HW, d.groups = nimg, o["srcs"][0].H * o["srcs"][0].W, 8
                d.gamma, d.beta = sd[o["prefix"] + ".weight"].data_ptr(), sd[o["prefix"] + ".bias"].data_ptr()
                d.eps, d.silu, d.out, d.out_lo = 1e-5, o["silu"], ap(o["out"]), ap_lo(o["out"])
                d.drop_p = o["drop_p"]
                if o["drop_p"] > 0:  # per-layer constant + the per-step seed the training engine keeps in device memory
                    d.seed = layer_seed(len(self.op_names))
                    d.seed_dev = seed_dev.data_ptr() if seed_dev is not None else None
                add(lib.dmc_plan_add_gn_apply, d, o["prefix"])
            elif kind == "conv":
                d = _lib.ConvDesc()
                d.nsrc = len(o["srcs"])
                for i, s in enumerate(o["srcs"]):
                    d.src[i], d.src_c[i], d.src_taps[i] = (ap_lo(s) if o["parts"][i] == "lo" else ap(s)), s.C, o["taps"][i]
                d.B, d.Hin, d.Win, d.stride, d.up_phase = nimg, o["H"], o["W"], o["stride"], o["up_phase"]
                if split:
                    w3 = net._split_weight(pk, o["wname"], o["sc_slice"])
                    rows, K = w3.shape
                    d.weight = w3.data_ptr()
                else:
                    rows, K = pk["wshape"][o["wname"]]
                    d.weight = pk["wblob"].data_ptr() + pk["woffs"][o["wname"]]
                d.Cout, d.Cout_pad, d.Ktot = o["Cout"], rows, K
                d.bias = pk["bias"][o["bias"]].data_ptr() if o["bias"] is not None else None
                if o["cond_col"] is not None:
                    d.cond, d.cond_stride = self.cond.data_ptr() + 4 * o["cond_col"], ncols
                d.residual = ap(o["residual"]) if o["residual"] is not None else None
                d.residual_lo = ap_lo(o["residual"]) if o["residual"] is not None else None
                if o["out_nchw"]:
                    d.out_f32_nchw = self.eps.data_ptr()
                elif o.get("out_f32"):
                    d.out_f32_nhwc = ap(o["out"])
                elif o["out"].raw:
                    d.out_bf16, d.out_lo = ap(o["out"]), ap_lo(o["out"])
                if o.get("aff") is not None:
                    d.a_affine = ap(o["aff"])
                d.impl = conv_impl
                if conv_impl == 0 and o["out"] is not None and o["out"].stats is not None and o["want_stats"]:
                    d.stats, d.stats_slots = wsp + o["out"].stats[0], o["out"].slots
                gn = o.get("gn", ())
                if gn:  # GroupNorm(+SiLU) of this output, applied by this convolution's epilogue for its consumer(s)
                    d.gn_nver, d.gn_eps = len(gn), 1e-5
                    for vi, v in enumerate(gn):
                        d.gn_out[vi], d.gn_pitch[vi], d.gn_coff[vi] = ap(v["dst"]), v["dst"].C, v["coff"]
                        d.gn_gamma[vi] = sd[v["prefix"] + ".weight"].data_ptr() + 4 * v["coff"]
                        d.gn_beta[vi] = sd[v["prefix"] + ".bias"].data_ptr() + 4 * v["coff"]
                        d.gn_gsize[vi], d.gn_silu[vi] = v["gsize"], v["silu"]
                    cnt = torch.zeros(2 * nimg * max(1, o["Cout"] // 64), dtype=torch.int32, device=device)
                    self.gn_counters.append(cnt)  # self-resetting image counters (images that span several CTAs)
                    d.gn_counters = cnt.data_ptr()
                idx = add(lib.dmc_plan_add_conv, d, o["wname"])
                if o["out_nchw"]:
                    self.head_idx = idx
            elif kind == "attention":
                d = _lib.AttnDesc()
                d.qkv, d.out, d.B, d.L, d.heads, d.C = ap(o["qkv"]), ap(o["out"]), nimg, o["L"], 4, o["C"]
                d.qkv_lo, d.out_lo = ap_lo(o["qkv"]), ap_lo(o["out"])
                d.impl = int(os.environ.get("DMC_DEBUG_ATTN_IMPL", "0"))
                add(lib.dmc_plan_add_attention, d, "attention")
            elif kind == "head":
                a = o["src"]
                d = _lib.HeadDesc()
                d.src, d.stats, d.stats_slots = ap(a), wsp + a.stats[0], a.slots
                d.B, d.H, d.W, d.C, d.Cout, d.groups = nimg, a.H, a.W, a.C, net.out_channels, 8
                d.gamma, d.beta, d.eps = sd["output.0.weight"].data_ptr(), sd["output.0.bias"].data_ptr(), 1e-5
                d.weight, d.bias = sd["output.2.weight"].data_ptr(), sd["output.2.bias"].data_ptr()
                d.out = self.eps.data_ptr()
                self.head_wfrag = torch.empty(9 * (a.C // 16) * 256, dtype=torch.uint8, device=device)
                d.wfrag = self.head_wfrag.data_ptr()
                self.head_idx = add(lib.dmc_plan_add_head, d, "output.head")
            elif kind == "head_taps":
                d = _lib.HeadTapsDesc()
                d.y, d.B, d.H, d.W, d.Cout, d.ypitch = ap(o["src"]), nimg, Hh, Ww, net.out_channels, o["ypitch"]
                d.bias, d.out = sd["output.2.bias"].data_ptr(), self.eps.data_ptr()
                self.head_idx = add(lib.dmc_plan_add_head_taps, d, "output.2.gather")
            elif kind == "upsample":
                d = _lib.UpsampleDesc()
                d.src, d.out, d.B, d.H, d.W, d.C = ap(o["src"]), ap(o["out"]), nimg, o["src"].H, o["src"].W, o["src"].C
                add(lib.dmc_plan_add_upsample, d, "upsample")
        self.num_launches = lib.dmc_plan_num_launches(handle)
        self.gemm_flops = lib.dmc_plan_gemm_flops(handle)

    def run(self, x, t, y, cfg, out_a, out_b):
        """x: [x_batch, C, H, W] fp32 (a contiguous slice), t/y: [x_batch]; out_a/out_b: destination slices."""
        lib, n = self.lib, self.x_batch
        if cfg:
            self.t_stage[:n].copy_(t)
            self.t_stage[n:].copy_(t)
            self.y_stage[:n].copy_(y)  # second half stays 0 = null label (ddim.py:283 y_uncond)
        else:
            self.t_stage.copy_(t)
            if self.has_y:
                self.y_stage.copy_(y)
        _lib.check(lib.dmc_plan_rebind(self.handle, self.stem_idx, 0, x.data_ptr()), "rebind x")
        _lib.check(lib.dmc_plan_run(self.handle, _lib.stream_ptr()), "dmc_plan_run")
        if cfg:
            out_a.copy_(self.eps[:n])
            out_b.copy_(self.eps[n:])
        else:
            out_a.copy_(self.eps)

    def time_ops(self, iters=5):
        n = self.lib.dmc_plan_num_ops(self.handle)
        buf = (C.c_float * n)()
        _lib.check(self.lib.dmc_plan_time_ops(self.handle, _lib.stream_ptr(), iters, buf, n), "dmc_plan_time_ops")
        kinds = [_lib.OP_KINDS[self.lib.dmc_plan_op_kind(self.handle, i)] for i in range(n)]
        flops = [self.lib.dmc_plan_op_flops(self.handle, i) for i in range(n)]
        nbytes = [self.lib.dmc_plan_op_bytes(self.handle, i) for i in range(n)]
        return [dict(name=self.op_names[i], kind=kinds[i], ms=buf[i], flops=flops[i], bytes=nbytes[i]) for i in range(n)]

    def destroy(self):
        if getattr(self, "handle", None) is not None:
            self.lib.dmc_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass
