"""fp32 CPU restatement of the RSPrompter inference algorithm (TEST INFRASTRUCTURE, see
oracle/__init__.py).  Plain functions over plain state dicts -- no modules, no registry --
so every step the CUDA path fuses or re-orders is visible here in reference order.

Citations: HF: = transformers/models/sam/modeling_sam.py (5.5.0 copy in this image),
VS: = mmpretrain/models/backbones/vit_sam.py, M: = mmdet/rsprompter/models.py, other paths
relative to the reference tree.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# =============================================================================================
# SAM ViT image encoder
# =============================================================================================


def rel_pos_gather(table: torch.Tensor, q_size: int, k_size: int) -> torch.Tensor:
    """R[q, k] = table[q - k + (k_size - 1)] (HF:729-758 get_rel_pos, VS:78-114).

    When the table length differs from 2*max(q,k)-1 it is linearly interpolated first."""
    L = 2 * max(q_size, k_size) - 1
    if table.shape[0] != L:
        table = F.interpolate(table.t()[None], size=L, mode="linear")[0].t()
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    idx = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return table[idx.long()]  # [q, k, hd]


def vit_attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rel_h: torch.Tensor,
                       rel_w: torch.Tensor, S: int) -> torch.Tensor:
    """q,k,v [N, T, hd] with T = S*S.  softmax((q*scale) k^T + rel_h + rel_w) v
    (HF:814-829; the bias uses the UNscaled q, HF:760-801)."""
    N, T, hd = q.shape
    scale = hd ** -0.5
    attn = (q * scale) @ k.transpose(-2, -1)
    bias = decomposed_rel_pos_bias(q, rel_h, rel_w, S)
    attn = torch.softmax((attn + bias).float(), dim=-1).to(q.dtype)
    return attn @ v


def decomposed_rel_pos_bias(q: torch.Tensor, rel_h: torch.Tensor, rel_w: torch.Tensor, S: int) -> torch.Tensor:
    """bias[n, (qh,qw), (kh,kw)] = q . Rh[qh-kh+S-1] + q . Rw[qw-kw+S-1] with the UNscaled q
    (HF:760-801 get_decomposed_rel_pos, VS:117-157 add_decomposed_rel_pos).  q [N, S*S, hd]."""
    N, T, hd = q.shape
    Rh = rel_pos_gather(rel_h, S, S)  # [qh, kh, hd]
    Rw = rel_pos_gather(rel_w, S, S)
    q4 = q.reshape(N, S, S, hd)
    bh = torch.einsum("bhwc,hkc->bhwk", q4, Rh)  # [N, qh, qw, kh]
    bw = torch.einsum("bhwc,wkc->bhwk", q4, Rw)  # [N, qh, qw, kw]
    return (bh[:, :, :, :, None] + bw[:, :, :, None, :]).reshape(N, T, T)


def window_partition(x: torch.Tensor, ws: int):
    """[B,H,W,C] -> [B*nW, ws, ws, C], zero padded to a multiple of ws (HF:900-922)."""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.reshape(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(w: torch.Tensor, ws: int, padded, orig) -> torch.Tensor:
    """Inverse of window_partition, cropping the padding (HF:925-952)."""
    Hp, Wp = padded
    H, W = orig
    B = w.shape[0] // ((Hp // ws) * (Wp // ws))
    x = w.reshape(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, Hp, Wp, -1)[:, :H, :W, :]


def vit_layer(sd: dict, p: str, x: torch.Tensor, window: int, heads: int, eps: float) -> torch.Tensor:
    """One SamVisionLayer (HF:954-972) / TransformerEncoderLayer (VS:298-313)."""
    B, H, W, D = x.shape
    hd = D // heads
    res = x
    y = F.layer_norm(x, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
    if window > 0:
        y, padded = window_partition(y, window)
    n, h, w, _ = y.shape
    qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    if p + "attn.qkv.lora_A" in sd:     # peft LoRA on qkv, UNMERGED (M:785-797): y = Wx + b + (alpha/r) B(A x)
        A, Bm, scale = sd[p + "attn.qkv.lora_A"], sd[p + "attn.qkv.lora_B"], sd[p + "attn.qkv.lora_scale"]
        qkv = qkv + float(scale) * F.linear(F.linear(y, A), Bm)
    qkv = qkv.reshape(n, h * w, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, n * heads, h * w, hd)
    o = vit_attention_core(qkv[0], qkv[1], qkv[2], sd[p + "attn.rel_pos_h"], sd[p + "attn.rel_pos_w"], h)
    o = o.reshape(n, heads, h, w, hd).permute(0, 2, 3, 1, 4).reshape(n, h, w, D)
    o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    if window > 0:
        o = window_unpartition(o, window, padded, (H, W))
    x = res + o
    y = F.layer_norm(x, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
    y = F.linear(y, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
    return x + y


def layer_norm_channels_first(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    """LN over C of NCHW (SamLayerNorm channels_first HF:147-170; LayerNorm2d norm.py:64-89;
    LN2d M:45-50 -- the explicit mean / biased-variance formula)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[None, :, None, None] * x + b[None, :, None, None]


def vit_encoder(sd: dict, arch, pixel_values: torch.Tensor):
    """HF SamVisionEncoder.forward (HF:1058-1072) with hidden-state capture.
    -> (embeddings [B,C,g,g], [L+1 hidden states [B,g,g,D]])."""
    x = F.conv2d(pixel_values, sd["patch_embed.projection.weight"], sd["patch_embed.projection.bias"],
                 stride=arch.patch_size).permute(0, 2, 3, 1)
    x = x + sd["pos_embed"]
    hidden = [x]
    for i in range(arch.num_layers):
        win = 0 if i in arch.global_attn_indexes else arch.window_size
        x = vit_layer(sd, f"layers.{i}.", x, win, arch.num_heads, arch.layer_norm_eps)
        hidden.append(x)
    y = x.permute(0, 3, 1, 2)
    y = F.conv2d(y, sd["neck.conv1.weight"])
    y = layer_norm_channels_first(y, sd["neck.layer_norm1.weight"], sd["neck.layer_norm1.bias"], 1e-6)
    y = F.conv2d(y, sd["neck.conv2.weight"], padding=1)
    y = layer_norm_channels_first(y, sd["neck.layer_norm2.weight"], sd["neck.layer_norm2.bias"], 1e-6)
    return y, hidden


# =============================================================================================
# positional embedding, prompt encoder pieces, mask decoder
# =============================================================================================


def image_wide_positional_embedding(gauss: torch.Tensor, size: int) -> torch.Tensor:
    """get_image_wide_positional_embeddings (M:85-95) + SamPositionalEmbedding.forward (HF:552-566).
    gauss [2, F] -> [1, 2F, size, size]."""
    grid = torch.ones(size, size, dtype=gauss.dtype)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    c = torch.stack([x, y], dim=-1)
    c = (2 * c - 1) @ gauss
    c = 2 * math.pi * c
    pe = torch.cat([torch.sin(c), torch.cos(c)], dim=-1)
    return pe.permute(2, 0, 1)[None]


def embed_boxes(gauss: torch.Tensor, point_embed_2: torch.Tensor, point_embed_3: torch.Tensor, boxes: torch.Tensor,
                image_size: int) -> torch.Tensor:
    """SamPromptEncoder._embed_boxes (HF:636-645) + SamPositionalEmbedding.forward (HF:552-566): boxes [B, nb, 4]
    (image-space xyxy) -> sparse prompt embeddings [B, nb, 2, C]."""
    coords = (boxes + 0.5).reshape(boxes.shape[0], boxes.shape[1], 2, 2).clone()
    coords[..., 0] = coords[..., 0] / image_size
    coords[..., 1] = coords[..., 1] / image_size
    c = (2 * coords - 1) @ gauss
    c = 2 * math.pi * c
    emb = torch.cat([torch.sin(c), torch.cos(c)], dim=-1)
    emb[:, :, 0, :] += point_embed_2.reshape(-1)
    emb[:, :, 1, :] += point_embed_3.reshape(-1)
    return emb


def sam_mask_embedding(sd: dict, masks: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """SamMaskEmbedding.forward (HF:583-593): conv2x2s2 -> LN -> GELU -> conv2x2s2 -> LN -> GELU -> conv1x1."""
    p = "mask_embed."
    h = F.conv2d(masks, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=2)
    h = F.gelu(layer_norm_channels_first(h, sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps))
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2)
    h = F.gelu(layer_norm_channels_first(h, sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps))
    return F.conv2d(h, sd[p + "conv3.weight"], sd[p + "conv3.bias"])


def _sam_attention(sd: dict, p: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """SamAttention.forward (HF:231-270): projections, per-head softmax(q k^T / sqrt(c)) v, out_proj.
    q [N, Tq, C], k/v [N, Tk, C]."""
    q = F.linear(q, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(k, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(v, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    N, Tq, Ci = q.shape
    c = Ci // heads
    sp = lambda t: t.reshape(N, t.shape[1], heads, c).transpose(1, 2)  # noqa: E731
    q, k, v = sp(q), sp(k), sp(v)
    a = torch.softmax((q @ k.transpose(2, 3)) * (c ** -0.5), dim=-1, dtype=torch.float32).to(q.dtype)
    o = (a @ v).transpose(1, 2).reshape(N, Tq, Ci)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ln(sd: dict, p: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def two_way_transformer(sd: dict, arch, tokens: torch.Tensor, src: torch.Tensor, pos: torch.Tensor):
    """SamTwoWayTransformer.forward (HF:365-405) over SamTwoWayAttentionBlock (HF:306-348).
    tokens [N, Tt, C] (point embeddings), src / pos [N, HW, C] -> (queries, keys)."""
    H, eps = arch.num_heads, arch.layer_norm_eps
    queries, keys = tokens, src
    for i in range(arch.num_layers):
        p = f"transformer.layers.{i}."
        if i == 0:  # skip_first_layer_pe: self-attention output replaces the queries
            queries = _sam_attention(sd, p + "self_attn.", queries, queries, queries, H)
        else:
            q = queries + tokens
            queries = queries + _sam_attention(sd, p + "self_attn.", q, q, queries, H)
        queries = _ln(sd, p + "layer_norm1", queries, eps)
        q = queries + tokens
        k = keys + pos
        queries = queries + _sam_attention(sd, p + "cross_attn_token_to_image.", q, k, keys, H)
        queries = _ln(sd, p + "layer_norm2", queries, eps)
        m = F.linear(queries, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])
        m = F.linear(F.relu(m), sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
        queries = _ln(sd, p + "layer_norm3", queries + m, eps)
        q = queries + tokens
        k = keys + pos
        keys = keys + _sam_attention(sd, p + "cross_attn_image_to_token.", k, q, queries, H)
        keys = _ln(sd, p + "layer_norm4", keys, eps)
    q = queries + tokens
    k = keys + pos
    queries = queries + _sam_attention(sd, "transformer.final_attn_token_to_image.", q, k, keys, H)
    queries = _ln(sd, "transformer.layer_norm_final_attn", queries, 1e-5)  # nn.LayerNorm default (HF:363)
    return queries, keys


def _feed_forward(sd: dict, p: str, x: torch.Tensor, n_hidden: int) -> torch.Tensor:
    """SamFeedForward (HF:408-429), ReLU between layers, no final activation."""
    x = F.relu(F.linear(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"]))
    for k in range(n_hidden):
        x = F.relu(F.linear(x, sd[p + f"layers.{k}.weight"], sd[p + f"layers.{k}.bias"]))
    return F.linear(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def mask_decoder(sd: dict, arch, image_embeddings: torch.Tensor, image_pe: torch.Tensor,
                 sparse: torch.Tensor, dense: torch.Tensor, multimask_output: bool = False):
    """SamMaskDecoder.forward (HF:461-543) with point_batch_size = 1 as RSPrompter calls it
    (M:369-378, M:1685-1694).  image_embeddings / image_pe / dense [N, C, h, w]; sparse [N, 1, P, C].
    -> masks [N, 1, n_out, 4h, 4w], iou [N, 1, n_out]."""
    N, C, h, w = image_embeddings.shape
    nm = arch.num_multimask_outputs + 1
    out_tokens = torch.cat([sd["iou_token.weight"], sd["mask_tokens.weight"]], dim=0)
    tokens = torch.cat([out_tokens[None].expand(N, -1, -1), sparse[:, 0]], dim=1)  # [N, 1+nm+P, C]
    src = (image_embeddings + dense).flatten(2).permute(0, 2, 1)
    pos = image_pe.flatten(2).permute(0, 2, 1)
    queries, keys = two_way_transformer(sd, arch, tokens, src, pos)
    iou_tok = queries[:, 0]
    mask_tok = queries[:, 1:1 + nm]
    up = keys.transpose(1, 2).reshape(N, C, h, w)
    up = F.conv_transpose2d(up, sd["upscale_conv1.weight"], sd["upscale_conv1.bias"], stride=2)
    up = F.gelu(layer_norm_channels_first(up, sd["upscale_layer_norm.weight"], sd["upscale_layer_norm.bias"], 1e-6))
    up = F.gelu(F.conv_transpose2d(up, sd["upscale_conv2.weight"], sd["upscale_conv2.bias"], stride=2))
    hyper = torch.stack([_feed_forward(sd, f"output_hypernetworks_mlps.{i}.", mask_tok[:, i], 1)
                         for i in range(nm)], dim=1)  # [N, nm, C/8]
    c8, H4, W4 = up.shape[1:]
    masks = (hyper @ up.reshape(N, c8, H4 * W4)).reshape(N, 1, nm, H4, W4)
    iou = _feed_forward(sd, "iou_prediction_head.", iou_tok, arch.iou_head_depth - 2)[:, None]
    sl = slice(1, None) if multimask_output else slice(0, 1)
    return masks[:, :, sl], iou[:, :, sl]
