"""Fused Smart_Tree inference engine: compiles a reference state_dict into a layer plan
(BN folded into a per-channel affine epilogue, weights transposed to [tap, cin, cout], heads
packed) and runs the whole network on libst_b200 kernels.

Architecture follows /root/reference/smart_tree/model/model.py:77-87 and
model_blocks.py:107-243 as found in the shipped checkpoints:
stem 1x1 -> recursive UBlock(Head ResBlock, strided Encode, U, inverse Decode, concat, Tail
ResBlock) -> three heads.  Index structures are built ONCE per level and shared by every conv
of that level (the reference rebuilds the sub-manifold rulebook 14 times per forward)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops

F32 = torch.float32


def _fold_bn(sd, prefix, eps, extra_bias=None):
    """BatchNorm1d(eval) as y = x*scale + shift (model_blocks.py:33; eps=1e-4 in the checkpoints)."""
    w, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    m, v = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    scale = w / torch.sqrt(v + eps)
    shift = b - m * scale
    if extra_bias is not None:
        shift = shift + extra_bias.double() * scale
    return scale.float(), shift.float()


def _conv_w(w):
    """spconv layout [cout, kz, ky, kx, cin] -> [taps, cin, cout]."""
    cout, cin = w.shape[0], w.shape[-1]
    return w.reshape(cout, -1, cin).permute(1, 2, 0).contiguous().float()


@dataclass
class ConvLayer:
    w: torch.Tensor                       # [taps, cin, cout]
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    w_tc: Optional[torch.Tensor] = None   # tensor-core layout (ops.conv_tc_prepare), built lazily on the device
    w_tc_fused: object = None             # same with a fused identity 1x1 conv (tc_weights_fused)

    def to(self, dev):
        return ConvLayer(self.w.to(dev), None if self.scale is None else self.scale.to(dev),
                         None if self.shift is None else self.shift.to(dev))

    def tc_weights(self):
        if self.w_tc is None:
            self.w_tc = ops.conv_tc_prepare(self.w)
        return self.w_tc

    def tc_weights_fused(self, w2):
        """Tensor-core weights with the ResBlock identity 1x1 conv appended as extra K stages (ops.conv_tc_prepare_fused);
        False if unsupported (shape, or a BN scale too close to zero to divide by)."""
        if self.w_tc_fused is None:
            ok = self.scale is None or bool((self.scale.abs() > 1e-6).all())
            self.w_tc_fused = (ops.conv_tc_prepare_fused(self.w, w2, self.scale) if ok else None)
            if self.w_tc_fused is None:
                self.w_tc_fused = False
        return self.w_tc_fused


@dataclass
class ResBlockPlan:
    c1: ConvLayer
    c2: ConvLayer
    ident_w: Optional[torch.Tensor]       # [cin, cout] or None (Identity)


@dataclass
class LevelPlan:
    head: ResBlockPlan
    encode: Optional[ConvLayer] = None
    decode: Optional[ConvLayer] = None
    tail: Optional[ResBlockPlan] = None


class LevelIndex:
    """Coordinate table and gather maps of one resolution level.  `spatial_shape` (int32 device tensor (z,y,x), optional):
    the shape spconv would bound-check neighbour locations against -- strict_spconv_bounds, see build_levels."""

    def __init__(self, coords, spatial_shape=None, brick=False):
        """brick: the rows are in (batch, Z-order) and the level's 3x3x3 convs run on 4x4x4 bricks (ops.brick_plan); the
        [27, n] gather map is then only built if somebody asks for it (`nbr`)."""
        self.coords = coords
        self.n = coords.shape[0]
        self.spatial_shape = spatial_shape
        self.table = ops.CoordTable(coords)
        self.brick = ops.brick_plan(coords, self.table) if (brick and spatial_shape is None and self.n > 0) else None
        self._nbr = None if self.brick is not None else ops.subm_map(coords, self.table, spatial_shape)
        self.down = None
        self.up = None
        self.child = None
        self._plans = {}

    @property
    def nbr(self):
        if self._nbr is None:
            self._nbr = ops.subm_map(self.coords, self.table, self.spatial_shape)
        return self._nbr

    def up_map(self):
        """The unsorted `up` map (built on demand when the level was indexed with the parity-sorted plan only)."""
        if self.up is None and self.child is not None:
            _, self.up = ops.strided_maps(self.coords, self.child.coords, self.child.table)
        return self.up

    def inverse_plan(self):
        """Parity-sorted launch order of this level's inverse conv (ops.inverse_plan), built on first use."""
        p = self._plans.get("inv")
        if p is None:
            p = self._plans["inv"] = ops.inverse_plan(self.coords, self.up_map())
        return p

    def plan(self, name, n_in):
        """Tile plan (distinct source rows + local map per 128-row tile) of gather map `name`, built on first use
        and shared by every conv of the level that gathers through that map."""
        p = self._plans.get(name)
        if p is None:
            p = self._plans[name] = ops.conv_plan_build(self.up_map() if name == "up" else getattr(self, name), n_in)
        return p


def declared_spatial_shape(coords: torch.Tensor) -> torch.Tensor:
    """The spatial_shape the reference hands to spconv: max(coords) per axis over the whole batch -- NOT max + 1
    (/root/reference/smart_tree/model/sparse.py:15-19).  int32 device tensor (z,y,x)."""
    if coords.shape[0] == 0:
        return torch.zeros(3, dtype=torch.int32, device=coords.device)
    return coords[:, 1:].max(0).values.int().contiguous()


def build_levels(coords: torch.Tensor, depth: int, morton: bool = False, inverse_plan: bool = False,
                 spatial_shape: torch.Tensor = None, brick0: bool = False) -> List[LevelIndex]:
    """Index structures of every level.  With `spatial_shape` (strict_spconv_bounds) the maps reproduce spconv's bound
    checks against the declared shape: neighbour locations >= shape are invisible to the sub-manifold convs and strided
    outputs >= out_shape = (shape - 1) // 2 + 1 are never created; each level passes its out_shape on as the next
    level's shape, as spconv does.  Default (None): unbounded grid."""
    levels = [LevelIndex(coords, spatial_shape, brick=brick0 and morton)]
    for _ in range(depth - 1):
        cur = levels[-1]
        out_shape = None
        if cur.spatial_shape is not None:
            out_shape = (torch.div(cur.spatial_shape - 1, 2, rounding_mode="floor") + 1).int().contiguous()
        oc = ops.strided_coords(cur.coords, morton=morton, out_shape=out_shape)
        nxt = LevelIndex(oc, out_shape)
        cur.child = nxt
        if inverse_plan:
            cur.down, cur.up, cur._plans["inv"] = ops.strided_maps(cur.coords, oc, nxt.table, inverse_plan=True)
        else:
            cur.down, cur.up = ops.strided_maps(cur.coords, oc, nxt.table)
        levels.append(nxt)
    levels[0].table.check()          # coordinate range (deeper levels derive from level 0)
    return levels


class SmartTreeEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", eps: float = 1e-4, conv_impl: str = None,
                 strict_spconv_bounds: bool = None):
        import os
        # strict_spconv_bounds: reproduce spconv's bound checks against the reference's declared spatial_shape = max(coords)
        # (see build_levels).  Off by default: the library treats the grid as unbounded.
        if strict_spconv_bounds is None:
            strict_spconv_bounds = bool(int(os.environ.get("ST_STRICT_SPCONV_BOUNDS", "0")))
        self.strict_spconv_bounds = bool(strict_spconv_bounds)
        # 3x3x3 layers with cin * cout <= fma_max run on the FFMA2 kernel under "auto", wider ones on tcgen05
        self.fma_max = int(os.environ.get("ST_CONV_FMA_MAX", "128"))
        # "auto": tensor cores (tcgen05, per-thread gathers) where they win on B200 -- every 3x3x3 layer with >= 16
        # input or output channels; the 8 -> 8 layers stay on the FMA kernel (tools/conv_micro.py: 85 us fma vs 87 tc).
        # "tp": the tile-plan kernel (source rows of a 128-row tile staged in shared memory by bulk copies) for the
        # sub-manifold and inverse convs with <= 32 input channels.  Measured equal to "tc" per launch (both are
        # bound by the producer <-> MMA handshake, not by the gather: DESIGN.md section 4) and it pays for the plan
        # builds, so it is opt-in (ST_CONV_IMPL=tp), kept parity-tested
        conv_impl = conv_impl or os.environ.get("ST_CONV_IMPL", "auto")
        sd = {k: v.detach().cpu() for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
        self.device = torch.device(device)
        self.eps = eps
        self.conv_impl = conv_impl
        self.morton = bool(int(os.environ.get("ST_MORTON", "1")))
        # level-0 sub-manifold convs (8 / 16 -> 8 channels) on 4x4x4 bricks staged in shared memory instead of a gather map
        # (conv_brick.cu).  Opt-in (ST_CONV_BRICK=1 or conv_impl="brick"): first measurement 128 us per 8 -> 8 launch against
        # 72 us of the gather-map kernel -- one warp per brick at 24 warps per SM is latency bound (profiles/README.md)
        self.brick_convs = bool(int(os.environ.get("ST_CONV_BRICK", "0"))) or conv_impl == "brick"
        self.inverse_sorted = bool(int(os.environ.get("ST_INVERSE_SORTED", "1")))
        dev = self.device
        self.stem = ConvLayer(_conv_w(sd["input_conv.sequence.0.weight"]),
                              *_fold_bn(sd, "input_conv.sequence.1", eps)).to(dev)
        self.levels: List[LevelPlan] = []
        pre = "UNet."
        while pre + "Head.sequence.0.weight" in sd:
            lp = LevelPlan(head=self._resblock(sd, pre + "Head."))
            if pre + "Encode.sequence.0.weight" in sd:
                lp.encode = ConvLayer(_conv_w(sd[pre + "Encode.sequence.0.weight"]), *_fold_bn(sd, pre + "Encode.sequence.1", eps)).to(dev)
                lp.decode = ConvLayer(_conv_w(sd[pre + "Decode.sequence.0.weight"]), *_fold_bn(sd, pre + "Decode.sequence.1", eps)).to(dev)
                lp.tail = self._resblock(sd, pre + "Tail.")
            self.levels.append(lp)
            pre += "U."
        self.depth = len(self.levels)
        self.planes = [lp.head.c1.w.shape[2] for lp in self.levels]
        self.in_channels = self.stem.w.shape[1]
        self.heads_packed = self._pack_heads(sd)
        self.head_layers = None if self.heads_packed is not None else {h: self._head_layers(sd, h + "_head.") for h in ("radius", "direction", "class")}
        if self.heads_packed is not None:
            self.heads_packed = self.heads_packed.to(dev)

    # ---- plan construction
    def _resblock(self, sd, pre):
        dev = self.device
        c1 = ConvLayer(_conv_w(sd[pre + "sequence.0.weight"]), *_fold_bn(sd, pre + "sequence.1", self.eps)).to(dev)
        c2 = ConvLayer(_conv_w(sd[pre + "sequence.3.weight"]), *_fold_bn(sd, pre + "sequence.4", self.eps)).to(dev)
        idk = pre + "identity.0.weight"
        ident = _conv_w(sd[idk])[0].contiguous().to(dev) if idk in sd else None
        return ResBlockPlan(c1, c2, ident)

    def _head_layers(self, sd, pre):
        layers, i = [], 0
        while f"{pre}sequence.{i}.weight" in sd:
            w = sd[f"{pre}sequence.{i}.weight"]
            lin_bias = sd.get(f"{pre}sequence.{i}.bias") if w.dim() == 2 else None
            wt = (w.t().contiguous().float()[None] if w.dim() == 2 else _conv_w(w))
            if f"{pre}sequence.{i + 1}.weight" in sd:
                scale, shift = _fold_bn(sd, f"{pre}sequence.{i + 1}", self.eps, extra_bias=lin_bias)
                layers.append((ConvLayer(wt, scale, shift).to(self.device), True))
                i += 3
            else:
                shift = lin_bias.float() if lin_bias is not None else None
                layers.append((ConvLayer(wt, None, shift).to(self.device), False))
                break
        return layers

    def _pack_heads(self, sd):
        """Packed block for st_heads_fused: per head W1[8][8], s1[8], b1[8], W2[8][4], s2[4], b2[4], W3[4][k], b3[k]."""
        chunks = []
        for name, k in (("radius", 1), ("direction", 3), ("class", 2)):
            ls = self._head_layers(sd, name + "_head.")
            if len(ls) != 3:
                return None
            (l1, _), (l2, _), (l3, _) = ls
            if tuple(l1.w.shape) != (1, 8, 8) or tuple(l2.w.shape) != (1, 8, 4) or tuple(l3.w.shape) != (1, 4, k):
                return None
            b3 = l3.shift if l3.shift is not None else torch.zeros(k, device=self.device)
            chunks += [l1.w.reshape(-1), l1.scale, l1.shift, l2.w.reshape(-1), l2.scale, l2.shift, l3.w.reshape(-1), b3]
        return torch.cat([c.float().reshape(-1).to(self.device) for c in chunks]).contiguous()

    # ---- execution
    def _conv(self, x, layer: ConvLayer, nbr, n_out, relu, out=None, residual=None, in2=None, w2=None, lv=None, which=None):
        """`lv`/`which` name the LevelIndex map behind `nbr` ("nbr" | "up"): with them the conv may take the
        tile-plan path (source rows staged in shared memory)."""
        taps, cin, cout = layer.w.shape
        impl, plan = "fma", None
        if (which == "nbr" and lv is not None and getattr(lv, "brick", None) is not None and self.conv_impl in ("auto", "brick")
                and ops.conv_brick_supported(taps, cin, cout)):
            return ops.conv_brick(x, lv.brick, layer.w, n_out, layer.scale, layer.shift, residual=residual, in2=in2, w2=w2, out=out, relu=relu)
        if taps > 1 and nbr is None and which == "nbr" and lv is not None:
            nbr = lv.nbr
        if taps > 1:
            if self.conv_impl == "tp" and lv is not None and ops.conv_tp_supported(taps, cin, cout):
                if which == "up" and nbr is None:
                    nbr = lv.up_map()
                impl, plan = "tp", lv.plan(which, x.shape[0])
            elif (self.conv_impl in ("tc", "tp") or (self.conv_impl == "auto" and cin * cout > self.fma_max)) and ops.conv_tc_supported(taps, cin, cout):
                impl = "tc"
        wtc = layer.tc_weights() if impl != "fma" else None
        sorted_inv = which == "up" and self.inverse_sorted and residual is None and in2 is None and cin <= 64 and lv._plans.get("inv") is not None
        if impl == "fma" and sorted_inv:
            # narrow decoder (16 -> 8 at level 0): FFMA2 kernel on parity-sorted rows, only the occurring taps are visited
            return ops.conv_gather_inv(x, lv.inverse_plan(), layer.w, n_out, layer.scale, layer.shift, out=out, relu=relu)
        if which == "up" and nbr is None:
            if not (impl == "tc" and self.inverse_sorted and residual is None and in2 is None and cin <= 64):
                nbr = lv.up_map()
        if impl == "tc" and which == "up" and self.inverse_sorted and residual is None and in2 is None and cin <= 64:
            # decoder: only the <= 8 taps a fine voxel's parity class allows have work (parity-sorted rows, stage skipping)
            return ops.conv_gather_tc_inv(x, lv.inverse_plan(), wtc, taps, cin, cout, n_out, layer.scale, layer.shift, out=out, relu=relu)
        if impl == "tc" and in2 is not None:
            fused = layer.tc_weights_fused(w2)          # identity 1x1 conv as extra K stages on the tensor cores
            if fused is not False:
                wtc, w2 = fused, None
        return ops.conv_gather(x, nbr, layer.w, n_out, layer.scale, layer.shift, residual=residual, in2=in2, w2=w2,
                               out=out, relu=relu, impl=impl, weight_tc=wtc, plan=plan)

    def _resblock_run(self, x, rb: ResBlockPlan, lv, out):
        n = x.shape[0]
        nbr = None if getattr(lv, "brick", None) is not None and self.conv_impl in ("auto", "brick") else lv.nbr      # (lazy on a brick level)
        t = self._conv(x, rb.c1, nbr, n, relu=True, lv=lv, which="nbr")
        if rb.ident_w is None:
            return self._conv(t, rb.c2, nbr, n, relu=True, out=out, residual=x, lv=lv, which="nbr")
        return self._conv(t, rb.c2, nbr, n, relu=True, out=out, in2=x, w2=rb.ident_w, lv=lv, which="nbr")

    def _ublock(self, x, li, levels, trace, pre="UNet."):
        lp, lv = self.levels[li], levels[li]
        n, c = lv.n, self.planes[li]
        if lp.encode is None:
            y = self._resblock_run(x, lp.head, lv, None)
            if trace is not None:
                trace[pre + "Head"] = y
            return y
        cat = torch.empty((n, 2 * c), dtype=F32, device=x.device)
        skip = cat[:, :c]
        self._resblock_run(x, lp.head, lv, skip)
        nxt = levels[li + 1]
        y = self._conv(skip, lp.encode, lv.down, nxt.n, relu=True)
        if trace is not None:
            trace[pre + "Head"] = skip.clone(); trace[pre + "Encode"] = y
        y = self._ublock(y, li + 1, levels, trace, pre + "U.")
        self._conv(y, lp.decode, lv.up, n, relu=True, out=cat[:, c:], lv=lv, which="up")
        out = self._resblock_run(cat, lp.tail, lv, None)
        if trace is not None:
            trace[pre + "Decode"] = cat[:, c:].clone(); trace[pre + "Tail"] = out
        return out

    def build_levels(self, coords):
        """Index structures of all levels.  With `self.morton` the rows of every level are kept in
        (batch, Z-order): level 0 through a permutation of the caller's rows (undone by the heads
        kernel), deeper levels by construction."""
        coords = coords.contiguous().int()
        inv = self.inverse_sorted and self.conv_impl in ("auto", "tc", "fma", "brick")      # (the tile-plan path wants the unsorted up map)
        shape = declared_spatial_shape(coords) if self.strict_spconv_bounds else None
        if not self.morton:
            return build_levels(coords, self.depth, inverse_plan=inv, spatial_shape=shape)
        perm = ops.morton_perm(coords)
        brick0 = self.conv_impl in ("auto", "brick") and self.planes[0] == 8 and self.brick_convs
        levels = build_levels(ops.gather_rows(coords, perm), self.depth, morton=True, inverse_plan=inv, spatial_shape=shape,      # int32 index: no widening pass
                              brick0=brick0)
        levels[0].perm = perm
        return levels

    @torch.no_grad()
    def forward(self, features: torch.Tensor, coords: torch.Tensor, levels=None, trace=None, fused_outputs=False):
        """features [N, Cin] f32, coords [N,4] i32 (b,z,y,x), both CUDA.  Returns the reference's
        prediction dict (model.py:77-87); with fused_outputs also medial_vector / class index."""
        if not features.is_cuda:
            raise RuntimeError("SmartTreeEngine.forward needs CUDA tensors: there is no CPU path")
        features = features.float()
        coords = coords.contiguous().int()
        n = features.shape[0]
        if levels is None:
            levels = self.build_levels(coords)
        perm = getattr(levels[0], "perm", None)
        scin, scout = self.stem.w.shape[1], self.stem.w.shape[2]
        if scin <= 8 and scout in (8, 16) and features.dim() == 2 and features.stride(1) == 1:
            # stem fused with the Z-order row permutation; reads the caller's rows in place (a column slice is fine)
            x = ops.stem_conv(features, self.stem.w[0], self.stem.scale, self.stem.shift, row_index=perm, relu=True)
        else:
            features = features.contiguous()
            if perm is not None:
                features = ops.gather_rows(features, perm)
            x = self._conv(features, self.stem, None, n, relu=True)
        if trace is not None:
            trace["input_conv"] = x
        x = self._ublock(x, 0, levels, trace)
        if trace is not None and perm is not None:      # level-0 activations back in the caller's row order
            inv = torch.empty_like(perm).long()
            inv[perm.long()] = torch.arange(n, device=perm.device)
            for k, v in list(trace.items()):
                if v.shape[0] == n and k.count("U.") == 0:
                    trace[k] = v[inv]
        if self.heads_packed is not None and x.shape[1] == 8:
            radius, direction, logits, medial, cls = ops.heads_fused(x, self.heads_packed, out_index=perm)
        else:
            if perm is not None:
                inv = torch.empty_like(perm).long()
                inv[perm.long()] = torch.arange(n, device=perm.device)
                x = x[inv]
            outs = {}
            for name, ls in self.head_layers.items():
                h = x
                for layer, relu in ls:
                    h = self._conv(h, layer, None, n, relu=relu)
                outs[name] = h
            radius, logits = outs["radius"], outs["class"]
            direction = torch.nn.functional.normalize(outs["direction"])
            medial = torch.exp(radius) * direction
            cls = torch.argmax(logits, dim=1).int()
        preds = {"radius": radius, "direction": direction, "class_l": logits}
        if fused_outputs:
            preds["medial_vector"] = medial
            preds["class_idx"] = cls
        return preds
