"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A plain PyTorch-fp32 (CPU) functional restatement of the VideoSeal per-frame inference
hot path: embedder U-Net -> JND attenuation / blend / clamp -> ConvNeXt-V2 extractor, plus
the video-mode plumbing around it.  Every function cites the reference file:line it
follows (paths relative to facebookresearch/videoseal @ 871eda0).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (videoseal_b200/) never does and has no CPU
fallback.

Pinning: the reference ships no tests or golden vectors, so this restatement
is pinned against the UNMODIFIED reference modules imported in the build container
(oracle/ref_import.py) by oracle/make_golden.py, which also writes the fixtures under
tests/golden/ that travel to the GPU box.  tests/test_oracle_golden.py re-checks the
restatement against those fixtures everywhere.

The arithmetic leaves are PyTorch ATen CPU kernels (torch>=2.3 per the reference's
pyproject.toml:29); the torch version used is recorded in each golden file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# card -> spec (mirrors videoseal/utils/cfg.py:88-144, models/embedder.py:243-282,
# models/extractor.py:170-213)
# --------------------------------------------------------------------------------------


def spec_from_card(card: dict) -> dict:
    a = card["args"]
    nbits = int(a["nbits"])
    hsm = a.get("hidden_size_multiplier", 2)                      # cfg.py:107-110
    hidden = int(nbits * hsm)                                     # embedder.py:244
    img_size = a["img_size_proc"] if "img_size_proc" in a else a["img_size_extractor"]  # cfg.py:101-104
    emb_name = card["embedder"]["model"]
    u = dict(card["embedder"]["params"]["unet"])
    mp = dict(card["embedder"]["params"]["msg_processor"])
    ext_name = card["extractor"]["model"]
    ep = card["extractor"]["params"]
    spec = {
        "nbits": nbits,
        "hidden": hidden,
        "img_size": int(img_size),
        "scaling_w": float(a["scaling_w"]),
        "scaling_i": float(a["scaling_i"]),
        "attenuation": str(a["attenuation"]),
        "chunk_size": int(a.get("videoseal_chunk_size", a.get("videowam_chunk_size", 8))),  # cfg.py:115-118
        "step_size": int(a.get("videoseal_step_size", a.get("videowam_step_size", 4))),
        "yuv": "yuv" in emb_name,                                   # embedder.py:281
        "msg_type": mp.get("msg_processor_type", "binary+concat"),
        "unet": {
            "in_channels": int(u["in_channels"]), "out_channels": int(u["out_channels"]),
            "z_channels": int(u["z_channels"]), "num_blocks": int(u["num_blocks"]),
            "activation": u["activation"], "normalization": u["normalization"],
            "mults": [int(m) for m in u["z_channels_mults"]], "last_tanh": bool(u.get("last_tanh", True)),
        },
        "ext_kind": "convnext" if ext_name.startswith("convnext") else ("sam" if ext_name.startswith("sam") else ext_name),
    }
    if spec["ext_kind"] == "convnext":
        enc = ep["encoder"]
        dims = [int(d) for d in enc["dims"]]
        if ep.get("proportional_dim", False):                     # extractor.py:193-198
            mult = math.sqrt(nbits / 128)
            dims = [int(d * mult) for d in dims]
        spec["convnext"] = {"depths": [int(d) for d in enc["depths"]], "dims": dims,
                            "stem_stride": int(enc.get("stem_stride", 4))}
    elif spec["ext_kind"] == "sam":
        spec["vit"] = dict(ep["encoder"])
        spec["vit"]["img_size"] = spec["img_size"]                # extractor.py:172
    return spec


# --------------------------------------------------------------------------------------
# synthetic checkpoints in the reference's state_dict layout
# --------------------------------------------------------------------------------------


def _kaiming_uniform(shape, fan_in, g):
    bound = 1.0 / math.sqrt(fan_in)                               # nn.Conv2d/Linear default init == U(-1/sqrt(fan_in), ..)
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def synth_state_dict(spec: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded synthetic `checkpoint['model']` with the reference key names/shapes.
    BN running stats / affines, LN affines and GRN gamma/beta are randomised so that
    folding bugs are visible."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(key, cout, cin, k, bias, groups=1):
        fan_in = (cin // groups) * k * k
        sd[key + ".weight"] = _kaiming_uniform((cout, cin // groups, k, k), fan_in, g)
        if bias:
            sd[key + ".bias"] = _kaiming_uniform((cout,), fan_in, g)

    def linear(key, cout, cin):
        sd[key + ".weight"] = _kaiming_uniform((cout, cin), cin, g)
        sd[key + ".bias"] = _kaiming_uniform((cout,), cin, g)

    def affine(key, c):
        sd[key + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[key + ".bias"] = 0.1 * torch.randn(c, generator=g)

    u = spec["unet"]
    norm = u["normalization"]

    def norm_layer(key, c):
        if norm.startswith("batch"):
            affine(key, c)
            sd[key + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[key + ".running_var"] = 0.5 + torch.rand(c, generator=g)
            sd[key + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)
        elif norm.startswith("rms"):
            sd[key + ".gamma"] = (1 + 0.1 * torch.randn(c, generator=g)).view(c, 1, 1)
        else:
            raise NotImplementedError(norm)

    def resblock(key, cin, cout):
        conv(key + ".double_conv.0", cout, cin, 3, False)
        norm_layer(key + ".double_conv.1", cout)
        conv(key + ".double_conv.3", cout, cout, 3, False)
        norm_layer(key + ".double_conv.4", cout)
        conv(key + ".res_conv", cout, cin, 1, True)

    z = [u["z_channels"] * m for m in u["mults"]]
    P = "embedder.unet."
    emb_table = torch.randn(2 * spec["nbits"], spec["hidden"], generator=g)   # nn.Embedding default N(0,1)
    sd[P + "msg_processor.msg_embeddings.weight"] = emb_table
    resblock(P + "inc", u["in_channels"], z[0])
    for i in range(len(z) - 1):
        conv(P + f"downs.{i}.down", z[i + 1], z[i], 3, True)
        resblock(P + f"downs.{i}.conv", z[i + 1], z[i + 1])
    zb = z[-1] + spec["hidden"]
    for i in range(u["num_blocks"]):
        resblock(P + f"bottleneck.model.{i}", zb, zb)
    zz = list(z)
    zz[-1] = zb
    for j, ii in enumerate(reversed(range(len(zz) - 1))):
        conv(P + f"ups.{j}.up.upsample_block.2", zz[ii], 2 * zz[ii + 1], 3, False)
        affine(P + f"ups.{j}.up.upsample_block.3", zz[ii])
        resblock(P + f"ups.{j}.conv", zz[ii], zz[ii])
    conv(P + "outc", u["out_channels"], zz[0], 1, True)
    sd["embedder.msg_processor.msg_embeddings.weight"] = emb_table     # alias of the same storage in the reference

    if spec["ext_kind"] == "convnext":
        cn = spec["convnext"]
        dims, depths = cn["dims"], cn["depths"]
        Q = "detector.convnext."
        conv(Q + "downsample_layers.0.0", dims[0], 3, 4, True)
        affine(Q + "downsample_layers.0.1", dims[0])
        for i in range(3):
            affine(Q + f"downsample_layers.{i+1}.0", dims[i])
            conv(Q + f"downsample_layers.{i+1}.1", dims[i + 1], dims[i], 2, True)
        for s in range(4):
            for j in range(depths[s]):
                B = Q + f"stages.{s}.{j}."
                conv(B + "dwconv", dims[s], dims[s], 7, True, groups=dims[s])
                affine(B + "norm", dims[s])
                linear(B + "pwconv1", 4 * dims[s], dims[s])
                sd[B + "grn.gamma"] = 0.1 * torch.randn(1, 1, 1, 4 * dims[s], generator=g)
                sd[B + "grn.beta"] = 0.1 * torch.randn(1, 1, 1, 4 * dims[s], generator=g)
                linear(B + "pwconv2", dims[s], 4 * dims[s])
        D = "detector.pixel_decoder."
        conv(D + "output_upscaling.0.upsample_block.2", dims[-1], dims[-1], 3, False)
        affine(D + "output_upscaling.0.upsample_block.3", dims[-1])
        linear(D + "linear", 1 + spec["nbits"], dims[-1])
    elif spec["ext_kind"] == "sam":
        # modules/vit.py:14-143 (ImageEncoderViT), :146-210 (Block), :302-353 (Attention), :494-525 (PatchEmbed)
        v = spec["vit"]
        E, depth, heads = int(v["embed_dim"]), int(v["depth"]), int(v["num_heads"])
        ps, oc = int(v["patch_size"]), int(v["out_chans"])
        grid = spec["img_size"] // ps
        hd = E // heads
        Q = "detector.image_encoder."
        conv(Q + "patch_embed.proj", E, 3, ps, True)
        sd[Q + "pos_embed"] = 0.02 * torch.randn(1, grid, grid, E, generator=g)
        for i in range(depth):
            B = Q + f"blocks.{i}."
            ws = 0 if i in list(v["global_attn_indexes"]) else int(v["window_size"])
            n = grid if ws == 0 else ws
            affine(B + "norm1", E)
            linear(B + "attn.qkv", 3 * E, E)
            linear(B + "attn.proj", E, E)
            if v.get("use_rel_pos", False):      # zero-initialised in the reference; randomised so the term is exercised
                sd[B + "attn.rel_pos_h"] = 0.02 * torch.randn(2 * n - 1, hd, generator=g)
                sd[B + "attn.rel_pos_w"] = 0.02 * torch.randn(2 * n - 1, hd, generator=g)
            affine(B + "norm2", E)
            linear(B + "mlp.lin1", int(E * float(v["mlp_ratio"])), E)
            linear(B + "mlp.lin2", E, int(E * float(v["mlp_ratio"])))
        conv(Q + "neck.0", oc, E, 1, False)
        affine(Q + "neck.1", oc)
        conv(Q + "neck.2", oc, oc, 3, False)
        affine(Q + "neck.3", oc)
        D = "detector.pixel_decoder."
        conv(D + "output_upscaling.0.upsample_block.2", oc, oc, 3, False)
        affine(D + "output_upscaling.0.upsample_block.3", oc)
        linear(D + "linear", 1 + spec["nbits"], oc)
    else:
        raise NotImplementedError("synthetic checkpoints: only convnext / sam extractors")
    sd["rgb2yuv.M"] = torch.tensor([[0.299, 0.587, 0.114], [-0.14713, -0.28886, 0.436],
                                    [0.615, -0.51499, -0.10001]], dtype=torch.float32)
    return sd


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------


def _act(name: str, x: Tensor) -> Tensor:               # modules/common.py:196-208
    if name == "relu":
        return F.relu(x)
    if name == "silu":
        return F.silu(x)
    if name == "gelu":
        return F.gelu(x)
    if name == "leakyrelu":
        return F.leaky_relu(x, 0.2)
    raise NotImplementedError(name)


def _norm(sd, key: str, kind: str, x: Tensor) -> Tensor:   # modules/common.py:182-194
    if kind.startswith("batch"):                          # nn.BatchNorm2d in eval mode, eps 1e-5
        return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                            sd[key + ".weight"], sd[key + ".bias"], False, 0.0, 1e-5)
    if kind.startswith("rms"):                            # ChanRMSNorm, common.py:172-179
        c = x.shape[1]
        return F.normalize(x, dim=1) * (c ** 0.5) * sd[key + ".gamma"]
    raise NotImplementedError(kind)


def layernorm_cf(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """channels-first LayerNorm, modules/common.py:150-155 (biased variance)"""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def resnet_block(sd, key: str, x: Tensor, act: str, norm: str) -> Tensor:
    """modules/unet.py:17-39: act(norm(conv3(act(norm(conv3 x))))) + conv1(x)"""
    h = F.conv2d(x, sd[key + ".double_conv.0.weight"], None, padding=1)
    h = _act(act, _norm(sd, key + ".double_conv.1", norm, h))
    h = F.conv2d(h, sd[key + ".double_conv.3.weight"], None, padding=1)
    h = _act(act, _norm(sd, key + ".double_conv.4", norm, h))
    return h + F.conv2d(x, sd[key + ".res_conv.weight"], sd[key + ".res_conv.bias"])


def upsample_block(sd, key: str, x: Tensor, act: str, up_factor: int = 2) -> Tensor:
    """modules/common.py:45-52: bilinear x up_factor (align_corners=False) -> ReflectionPad2d(1)
    -> conv3x3 valid (no bias) -> channels-first LN -> act"""
    x = F.interpolate(x, scale_factor=up_factor, mode="bilinear", align_corners=False)
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    x = F.conv2d(x, sd[key + ".upsample_block.2.weight"], None)
    x = layernorm_cf(x, sd[key + ".upsample_block.3.weight"], sd[key + ".upsample_block.3.bias"])
    return _act(act, x)


def msg_embed(sd, spec, msgs: Tensor) -> Tensor:
    """modules/msg_processor.py:88-95 (binary): idx = 2*arange(K)+bit; gather; sum over K -> [B, hidden]"""
    table = sd["embedder.unet.msg_processor.msg_embeddings.weight"]
    K = msgs.shape[-1]
    idx = (2 * torch.arange(K).repeat(msgs.shape[0], 1) + msgs.to(torch.long)).long()
    return table[idx].sum(dim=-2)


def unet_forward(sd, spec, x: Tensor, msgs: Tensor, taps: Optional[dict] = None) -> Tensor:
    """models/embedder.py:151-165 (x*2-1) + modules/unet.py:170-197"""
    u = spec["unet"]
    act, norm = u["activation"], u["normalization"]
    P = "embedder.unet."
    x = x * 2 - 1                                                     # embedder.py:23,163
    x1 = resnet_block(sd, P + "inc", x, act, norm)
    hiddens = [x1]
    nd = len(u["mults"]) - 1
    for i in range(nd):                                               # DBlock unet.py:71-84
        d = F.conv2d(hiddens[-1], sd[P + f"downs.{i}.down.weight"], sd[P + f"downs.{i}.down.bias"], stride=2, padding=1)
        hiddens.append(resnet_block(sd, P + f"downs.{i}.conv", d, act, norm))
        if taps is not None:
            taps[f"down{i}"] = hiddens[-1]
    lat = hiddens.pop()
    m = msg_embed(sd, spec, msgs)                                      # b d
    m = m[:, :, None, None].repeat(1, 1, lat.shape[-2], lat.shape[-1])
    lat = torch.cat([lat, m], dim=1)                                   # msg_processor.py:111-115 (msg_mult = 1)
    hiddens.append(lat)
    x = lat
    for i in range(u["num_blocks"]):
        x = resnet_block(sd, P + f"bottleneck.model.{i}", x, act, norm)
        if taps is not None:
            taps[f"bott{i}"] = x
    for j in range(nd):                                               # unet.py:187-191
        x = torch.cat((x, hiddens.pop() * (2 ** -0.5)), dim=1)
        x = upsample_block(sd, P + f"ups.{j}.up", x, act)
        if taps is not None:
            taps[f"up{j}_conv"] = x
        x = resnet_block(sd, P + f"ups.{j}.conv", x, act, norm)
        if taps is not None:
            taps[f"up{j}"] = x
    x = F.conv2d(x, sd[P + "outc.weight"], sd[P + "outc.bias"])
    if u["last_tanh"]:
        x = torch.tanh(x)
    return x


def convnext_block(sd, key: str, x: Tensor) -> Tensor:
    """modules/convnext.py:41-57 + GRN modules/common.py:166-169"""
    inp = x
    C = x.shape[1]
    x = F.conv2d(x, sd[key + "dwconv.weight"], sd[key + "dwconv.bias"], padding=3, groups=C)
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (C,), sd[key + "norm.weight"], sd[key + "norm.bias"], 1e-6)
    x = F.linear(x, sd[key + "pwconv1.weight"], sd[key + "pwconv1.bias"])
    x = F.gelu(x)
    Gx = torch.norm(x, p=2, dim=(1, 2), keepdim=True)
    Nx = Gx / (Gx.mean(dim=-1, keepdim=True) + 1e-6)
    x = sd[key + "grn.gamma"] * (x * Nx) + sd[key + "grn.beta"] + x
    x = F.linear(x, sd[key + "pwconv2.weight"], sd[key + "pwconv2.bias"])
    return inp + x.permute(0, 3, 1, 2)


def convnext_extractor_forward(sd, spec, x: Tensor, taps: Optional[dict] = None) -> Tensor:
    """models/extractor.py:154-167; modules/convnext.py:146-150; modules/pixel_decoder.py:61-83"""
    cn = spec["convnext"]
    Q = "detector.convnext."
    x = x * 2 - 1                                                     # extractor.py:25,164
    for s in range(4):
        if s == 0:
            x = F.conv2d(x, sd[Q + "downsample_layers.0.0.weight"], sd[Q + "downsample_layers.0.0.bias"], stride=cn["stem_stride"])
            x = layernorm_cf(x, sd[Q + "downsample_layers.0.1.weight"], sd[Q + "downsample_layers.0.1.bias"])
        else:
            x = layernorm_cf(x, sd[Q + f"downsample_layers.{s}.0.weight"], sd[Q + f"downsample_layers.{s}.0.bias"])
            x = F.conv2d(x, sd[Q + f"downsample_layers.{s}.1.weight"], sd[Q + f"downsample_layers.{s}.1.bias"], stride=2)
        if taps is not None:
            taps[f"ds{s}"] = x
        for j in range(cn["depths"][s]):
            x = convnext_block(sd, Q + f"stages.{s}.{j}.", x)
        if taps is not None:
            taps[f"stage{s}"] = x
    D = "detector.pixel_decoder."
    # Upsample('bilinear', up_factor=1) is the identity (common.py:47 with scale_factor=1)
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    x = F.conv2d(x, sd[D + "output_upscaling.0.upsample_block.2.weight"], None)
    x = layernorm_cf(x, sd[D + "output_upscaling.0.upsample_block.3.weight"], sd[D + "output_upscaling.0.upsample_block.3.bias"])
    x = F.gelu(x)
    x = x.mean(dim=[-2, -1])                                           # pixel_decoder.py:77
    return F.linear(x, sd[D + "linear.weight"], sd[D + "linear.bias"])  # no sigmoid (all cards)


def _vit_rel_pos(q_size: int, k_size: int, rel_pos: Tensor) -> Tensor:
    """modules/vit.py:398-428 get_rel_pos"""
    max_rel = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel:
        rp = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel, mode="linear")
        rp = rp.reshape(-1, max_rel).permute(1, 0)
    else:
        rp = rel_pos
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rel = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rp[rel.long()]


def _vit_attention(sd, key: str, x: Tensor, heads: int, use_rel_pos: bool) -> Tensor:
    """modules/vit.py:302-353 Attention.forward + :431-470 add_decomposed_rel_pos; x: [B, H, W, C]"""
    B, H, W, C = x.shape
    qkv = F.linear(x, sd[key + "qkv.weight"], sd[key + "qkv.bias"]).reshape(B, H * W, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * heads, H * W, -1).unbind(0)
    scale = (C // heads) ** -0.5
    attn = (q * scale) @ k.transpose(-2, -1)
    if use_rel_pos:
        Rh = _vit_rel_pos(H, H, sd[key + "rel_pos_h"])
        Rw = _vit_rel_pos(W, W, sd[key + "rel_pos_w"])
        r_q = q.reshape(B * heads, H, W, -1)
        rel_h = torch.einsum("bhwc,hkc->bhwk", r_q, Rh)
        rel_w = torch.einsum("bhwc,wkc->bhwk", r_q, Rw)
        attn = (attn.view(B * heads, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(B * heads, H * W, H * W)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).view(B, heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)
    return F.linear(x, sd[key + "proj.weight"], sd[key + "proj.bias"])


def _vit_block(sd, key: str, x: Tensor, heads: int, window: int, use_rel_pos: bool) -> Tensor:
    """modules/vit.py:146-210 Block.forward with window_partition / window_unpartition (:356-395); nn.LayerNorm eps 1e-5"""
    C = x.shape[-1]
    shortcut = x
    x = F.layer_norm(x, (C,), sd[key + "norm1.weight"], sd[key + "norm1.bias"], 1e-5)
    if window > 0:
        B, H, W, _ = x.shape
        ph, pw = (window - H % window) % window, (window - W % window) % window
        if ph or pw:
            x = F.pad(x, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        x = x.view(B, Hp // window, window, Wp // window, window, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window, window, C)
    x = _vit_attention(sd, key + "attn.", x, heads, use_rel_pos)
    if window > 0:
        x = x.view(B, Hp // window, Wp // window, window, window, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
        if Hp > H or Wp > W:
            x = x[:, :H, :W, :].contiguous()
    x = shortcut + x
    h = F.layer_norm(x, (C,), sd[key + "norm2.weight"], sd[key + "norm2.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, sd[key + "mlp.lin1.weight"], sd[key + "mlp.lin1.bias"])),
                 sd[key + "mlp.lin2.weight"], sd[key + "mlp.lin2.bias"])              # common.py:112-125 MLPBlock
    return x + h


def sam_extractor_forward(sd, spec, x: Tensor, taps: Optional[dict] = None) -> Tensor:
    """models/extractor.py:41-69 SegmentationExtractor; modules/vit.py:111-143 ImageEncoderViT.forward;
    modules/pixel_decoder.py:61-83"""
    v = spec["vit"]
    Q = "detector.image_encoder."
    ps = int(v["patch_size"])
    x = x * 2 - 1                                                      # extractor.py:25,64
    x = F.conv2d(x, sd[Q + "patch_embed.proj.weight"], sd[Q + "patch_embed.proj.bias"], stride=ps).permute(0, 2, 3, 1)
    x = x + sd[Q + "pos_embed"]
    glob = [int(i) for i in v["global_attn_indexes"]]
    for i in range(int(v["depth"])):
        x = _vit_block(sd, Q + f"blocks.{i}.", x, int(v["num_heads"]), 0 if i in glob else int(v["window_size"]),
                       bool(v.get("use_rel_pos", False)))
        if taps is not None:
            taps[f"blk{i}"] = x
    x = x.permute(0, 3, 1, 2).contiguous()
    x = F.conv2d(x, sd[Q + "neck.0.weight"], None)                      # neck: 1x1, LN, 3x3 pad 1, LN (vit.py:90-109)
    x = layernorm_cf(x, sd[Q + "neck.1.weight"], sd[Q + "neck.1.bias"])
    x = F.conv2d(x, sd[Q + "neck.2.weight"], None, padding=1)
    x = layernorm_cf(x, sd[Q + "neck.3.weight"], sd[Q + "neck.3.bias"])
    D = "detector.pixel_decoder."
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")                          # Upsample(up_factor=1): identity + reflect pad
    x = F.conv2d(x, sd[D + "output_upscaling.0.upsample_block.2.weight"], None)
    x = layernorm_cf(x, sd[D + "output_upscaling.0.upsample_block.3.weight"], sd[D + "output_upscaling.0.upsample_block.3.bias"])
    x = F.gelu(x).mean(dim=[-2, -1])
    return F.linear(x, sd[D + "linear.weight"], sd[D + "linear.bias"])


def rgb_to_y(x: Tensor) -> Tensor:
    """data/transforms.py:15-27 RGB2YUV, row 0 only (wam.py:168-172)"""
    M = torch.tensor([[0.299, 0.587, 0.114], [-0.14713, -0.28886, 0.436], [0.615, -0.51499, -0.10001]], dtype=torch.float32)
    yuv = torch.matmul(x.permute(0, 2, 3, 1).contiguous(), M.T).permute(0, 3, 1, 2).contiguous()
    return yuv[:, 0:1]


_K_X = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3)
_K_Y = torch.tensor([[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]]).view(1, 1, 3, 3)
_K_L = torch.tensor([[1., 1., 1., 1., 1.], [1., 2., 2., 2., 1.], [1., 2., 0., 2., 1.],
                     [1., 2., 2., 2., 1.], [1., 1., 1., 1., 1.]]).view(1, 1, 5, 5)


def jnd_heatmaps(imgs: Tensor, in_channels: int = 1, out_channels: int = 1, clc: float = 0.3) -> Tensor:
    """modules/jnd.py:63-108 for the jnd_1_1 / jnd_1_3 / jnd_3_3 / jnd_3_1 configs (blue=False)"""
    x = 255 * imgs
    if in_channels == 1:
        x = 0.299 * x[..., 0:1, :, :] + 0.587 * x[..., 1:2, :, :] + 0.114 * x[..., 2:3, :, :]
    g = in_channels
    la = F.conv2d(x, _K_L.repeat(g, 1, 1, 1), padding=2, groups=g) / 32
    mask = la <= 127
    la = torch.where(mask, 17 * (1 - torch.sqrt(la / 127 + 1e-5)), 3 / 128 * (la - 127) + 3)
    gx = F.conv2d(x, _K_X.repeat(g, 1, 1, 1), padding=1, groups=g)
    gy = F.conv2d(x, _K_Y.repeat(g, 1, 1, 1), padding=1, groups=g)
    cm = torch.sqrt(gx ** 2 + gy ** 2)
    cm = 0.117 * (16 * cm ** 2.4 / (cm ** 2 + 26 ** 2))
    h = torch.clamp_min(la + cm - clc * torch.minimum(la, cm), 0)
    if out_channels == 3 and in_channels == 1:
        h = h.repeat(1, 3, 1, 1)
    elif out_channels == 1 and in_channels == 3:
        h = torch.sum(h / 3, dim=1, keepdim=True)
    return h / 255


_JND_CFG = {"jnd_1_1": (1, 1), "jnd_3_3": (3, 3), "jnd_1_3": (1, 3), "jnd_3_1": (3, 1)}   # configs/attenuation.yaml

_DEF_INTERP = {"mode": "bilinear", "align_corners": False, "antialias": True}


def apply_video_mode(preds_w: Tensor, total_frames: int, step_size: int, video_mode: str) -> Tensor:
    """models/videoseal.py:80-118"""
    if video_mode == "repeat":
        preds_w = torch.repeat_interleave(preds_w, step_size, dim=0)
    elif video_mode == "alternate":
        full = torch.zeros((total_frames,) + preds_w.shape[1:])
        full[::step_size] = preds_w
        preds_w = full
    elif video_mode == "interpolate":
        full = torch.zeros((total_frames,) + preds_w.shape[1:])
        alpha = 1 - torch.linspace(0, 1, steps=step_size)
        alpha = alpha.repeat((total_frames - 1) // step_size).view(-1, 1, 1, 1)
        start = torch.repeat_interleave(preds_w[:-1], step_size, dim=0)
        end = torch.repeat_interleave(preds_w[1:], step_size, dim=0)
        interp = alpha * start + (1 - alpha) * end
        last = len(interp)
        full[:last] = interp
        full[last:] = preds_w[-1]
        preds_w = full
    else:
        raise ValueError(video_mode)
    return preds_w[:total_frames]


class OracleModel:
    """Functional stand-in for the reference `Videoseal` object (models/videoseal.py, models/wam.py)."""

    def __init__(self, spec: dict, sd: Dict[str, Tensor]):
        self.spec, self.sd = spec, sd
        self.img_size = spec["img_size"]
        self.scaling_w, self.scaling_i = spec["scaling_w"], spec["scaling_i"]
        self.chunk_size, self.step_size = spec["chunk_size"], spec["step_size"]
        self.video_mode, self.clamp = "repeat", True
        self.attenuation = spec["attenuation"] if spec["attenuation"].lower().startswith("jnd") else None

    # -- leaves ------------------------------------------------------------------------
    def embedder(self, imgs_res: Tensor, msgs: Tensor, taps=None) -> Tensor:
        x = rgb_to_y(imgs_res) if self.spec["yuv"] else imgs_res
        return unet_forward(self.sd, self.spec, x, msgs, taps)

    def detector(self, imgs_res: Tensor, taps=None) -> Tensor:
        if self.spec["ext_kind"] == "sam":
            return sam_extractor_forward(self.sd, self.spec, imgs_res, taps)
        return convnext_extractor_forward(self.sd, self.spec, imgs_res, taps)

    def heatmaps(self, imgs: Tensor) -> Tensor:
        return jnd_heatmaps(imgs, *_JND_CFG[self.attenuation])

    def _resize(self, x, interpolation):
        if x.shape[-2:] != (self.img_size, self.img_size):
            return F.interpolate(x, size=(self.img_size, self.img_size), **interpolation)
        return x.clone()

    # -- models/wam.py:134-204 ------------------------------------------------------------
    def embed_images(self, imgs, msgs, interpolation=_DEF_INTERP, lowres_attenuation=False):
        imgs_res = self._resize(imgs, interpolation)
        preds_w = self.embedder(imgs_res, msgs)
        if self.attenuation is not None and lowres_attenuation:
            preds_w = self.heatmaps(imgs_res) * preds_w
        if imgs.shape[-2:] != (self.img_size, self.img_size):
            preds_w = F.interpolate(preds_w, size=imgs.shape[-2:], **interpolation)
        if self.attenuation is not None and not lowres_attenuation:
            preds_w = self.heatmaps(imgs) * preds_w
        imgs_w = self.scaling_i * imgs + self.scaling_w * preds_w      # blender.py:68
        if self.clamp:
            imgs_w = torch.clamp(imgs_w, 0, 1)
        return {"msgs": msgs, "preds_w": preds_w, "imgs_w": imgs_w}

    # -- models/videoseal.py:258-350 ------------------------------------------------------
    def embed(self, imgs, msgs, is_video=True, interpolation=_DEF_INTERP, lowres_attenuation=False):
        if not is_video:
            return self.embed_images(imgs, msgs, interpolation, lowres_attenuation)
        assert msgs.shape[0] == 1
        msgs = msgs.repeat(self.chunk_size, 1)
        cs, ss = self.chunk_size, self.step_size
        imgs_w = torch.zeros_like(imgs)
        nkey = len(imgs[::ss])
        for ii in range(0, nkey, cs):
            n = min(cs, nkey - ii)
            start, end = ii * ss, ii * ss + n * ss
            ck = imgs[start:end]
            if n < cs:
                msgs = msgs[:n]
            ck_res = self._resize(ck, interpolation)
            key = ck_res[::ss]
            preds_w = self.embedder(key, msgs)
            preds_w = apply_video_mode(preds_w, len(ck), ss, self.video_mode)
            if self.attenuation is not None and lowres_attenuation:
                preds_w = self.heatmaps(ck_res) * preds_w
            if ck.shape[-2:] != (self.img_size, self.img_size):
                preds_w = F.interpolate(preds_w, size=ck.shape[-2:], **interpolation)
            if self.attenuation is not None and not lowres_attenuation:
                preds_w = self.heatmaps(ck) * preds_w
            imgs_w[start:end] = self.scaling_i * ck + self.scaling_w * preds_w
        if self.clamp:
            imgs_w = torch.clamp(imgs_w, 0, 1)
        return {"imgs_w": imgs_w, "msgs": msgs[0:1].repeat(len(imgs), 1)}

    # -- models/wam.py:206-234, models/videoseal.py:352-388 --------------------------------
    def detect(self, imgs, is_video=True, interpolation=_DEF_INTERP):
        if not is_video:
            return {"preds": self.detector(self._resize(imgs, _DEF_INTERP))}   # videoseal.py:374 drops `interpolation`
        outs = []
        for ii in range(0, len(imgs), self.chunk_size):
            outs.append(self.detector(self._resize(imgs[ii:ii + self.chunk_size], interpolation)))
        return {"preds": torch.cat(outs, dim=0)}

    # -- models/videoseal.py:390-428 ------------------------------------------------------
    def extract_message(self, imgs, aggregation="avg",
                        interpolation={"mode": "bilinear", "align_corners": False, "antialias": False}):
        preds = self.detect(imgs, True, interpolation)["preds"]
        bit_preds = preds[:, 1:]
        if aggregation is None:
            dec = bit_preds
        elif aggregation == "avg":
            dec = bit_preds.mean(dim=0)
        elif aggregation == "squared_avg":
            dec = (bit_preds * bit_preds.abs()).mean(dim=0)
        elif aggregation == "l1norm_avg":
            dec = (bit_preds * torch.norm(bit_preds, p=1, dim=1).unsqueeze(1)).mean(dim=0)
        elif aggregation == "l2norm_avg":
            dec = (bit_preds * torch.norm(bit_preds, p=2, dim=1).unsqueeze(1)).mean(dim=0)
        else:
            raise ValueError(aggregation)
        return (dec > 0).squeeze().unsqueeze(0)


# -- evals/metrics.py:22-36,150-178 ---------------------------------------------------------


def psnr(x: Tensor, y: Tensor) -> Tensor:
    delta = 255 * (x - y)
    delta = delta.reshape(-1, x.shape[-3], x.shape[-2], x.shape[-1])
    return 20 * math.log10(255) - 10 * torch.log10(torch.mean(delta ** 2, dim=(1, 2, 3)))


def bit_accuracy(preds: Tensor, msgs: Tensor) -> Tensor:
    """preds: [B, 1+K] logits (column 0 = detection bit), msgs: [B, K] in {0,1}"""
    p = preds[:, 1:] > 0
    return (p == (msgs > 0.5)).float().mean(dim=1)
