"""CPU oracle: a functional restatement of the reference's VITS inference path.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import it; the product path
(wetts_b200/*) never does and has no CPU fallback.

It restates, with plain `torch` CPU ops on a flat state dict (no nn.Module), what the
reference computes in wetts/vits/model/models.py:228-280 (`SynthesizerTrn.infer`) and
the sub-modules it calls.  Every function cites the reference lines it follows.  The
reference has NO test of this path, so the oracle is pinned differently:
oracle/gen_golden.py imports the *real* reference from /root/reference (authoring
container only), runs it on the seeded synthetic checkpoints of wetts_b200/synth.py and
commits its outputs under tests/golden/; tests/test_oracle_golden.py checks this file
against those vectors on any box.  Parity status: pinned by reference-generated fixtures.
"""
import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules.py:7


# --------------------------------------------------------------------------- weights
def fold_weight_norm(sd):
    """w = g * v / ||v||, norm over all dims but 0 (torch weight_norm dim=0).
    Conv1d: v [C_out,C_in,k], g [C_out,1,1]; ConvTranspose1d: v [C_in,C_out,k],
    g [C_in,1,1] -- i.e. per INPUT channel (decoders.py:41-48; SURVEY §0 finding 10)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            continue
        if k.endswith(".weight_v"):
            g = sd[k[:-1] + "g"]
            n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            out[k[: -len("_v")]] = v * (g / n)
        else:
            out[k] = v
    return out


def _plain(cfg):
    """the model section as a plain dict (accepts the HParams attribute bag, which has no .get)"""
    return cfg.to_dict() if hasattr(cfg, "to_dict") else cfg


def seq_mask(lengths, max_len):
    """commons.py:113-117"""
    return (torch.arange(max_len)[None, :] < lengths[:, None]).to(torch.float32)


def channel_layer_norm(x, gamma, beta, eps=1e-5):
    """normalization.py:16-19 -- LN over the channel dim of [B,C,T]."""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma[None, :, None] + beta[None, :, None]


# --------------------------------------------------------------------------- text encoder
def rel_attention(q, k, v, emb_k, emb_v, mask2d, n_heads, window=4):
    """attentions.py:232-282 with the pad/reshape skew trick (:302-358) replaced by
    its meaning: banded key bias and banded value term, |j-i|<=window.
    q,k,v [B,C,T]; emb_* [1,2w+1,dk]; mask2d [B,T] (1 valid)."""
    B, C, T = q.shape
    dk = C // n_heads
    qh = q.view(B, n_heads, dk, T).transpose(2, 3) / math.sqrt(dk)  # [B,h,T,dk]
    kh = k.view(B, n_heads, dk, T).transpose(2, 3)
    vh = v.view(B, n_heads, dk, T).transpose(2, 3)
    scores = qh @ kh.transpose(-2, -1)  # [B,h,T,T]
    idx = torch.arange(T)
    rel = idx[None, :] - idx[:, None]  # j - i
    band = rel.abs() <= window
    rel_c = (rel + window).clamp(0, 2 * window)
    logits_rel = qh @ emb_k[0].t()  # [B,h,T,2w+1]
    bias = torch.gather(logits_rel, 3, rel_c[None, None].expand(B, n_heads, T, T))
    scores = scores + bias * band
    pair = mask2d[:, None, :, None] * mask2d[:, None, None, :]
    scores = scores.masked_fill(pair == 0, -1e4)  # attentions.py:262
    p = torch.softmax(scores, dim=-1)
    out = p @ vh
    # relative values: sum_{|j-i|<=w} p[i,j] * E_v[j-i+w]
    pw = torch.zeros(B, n_heads, T, 2 * window + 1)
    pw.scatter_add_(3, rel_c[None, None].expand(B, n_heads, T, T), p * band)
    out = out + pw @ emb_v[0]
    return out.transpose(2, 3).reshape(B, C, T)


def text_encoder(w, cfg, x_ids, x_lengths):
    """encoders.py:47-57 + attentions.py:70-87,225-231,403-411.  Returns h, m, logs, x_mask."""
    H = cfg["hidden_channels"]
    Tx = x_ids.shape[1]
    mask = seq_mask(x_lengths, Tx)  # [B,T]
    m3 = mask[:, None, :]
    h = (F.embedding(x_ids, w["enc_p.emb.weight"]) * math.sqrt(H)).transpose(1, 2) * m3
    ks = cfg["kernel_size"]
    pl, pr = (ks - 1) // 2, ks // 2
    for i in range(cfg["n_layers"]):
        a = f"enc_p.encoder.attn_layers.{i}"
        q = F.conv1d(h, w[a + ".conv_q.weight"], w[a + ".conv_q.bias"])
        k = F.conv1d(h, w[a + ".conv_k.weight"], w[a + ".conv_k.bias"])
        v = F.conv1d(h, w[a + ".conv_v.weight"], w[a + ".conv_v.bias"])
        y = rel_attention(q, k, v, w[a + ".emb_rel_k"], w[a + ".emb_rel_v"], mask, cfg["n_heads"])
        y = F.conv1d(y, w[a + ".conv_o.weight"], w[a + ".conv_o.bias"])
        n1 = f"enc_p.encoder.norm_layers_1.{i}"
        h = channel_layer_norm(h + y, w[n1 + ".gamma"], w[n1 + ".beta"])
        f = f"enc_p.encoder.ffn_layers.{i}"
        y = F.conv1d(F.pad(h * m3, (pl, pr)), w[f + ".conv_1.weight"], w[f + ".conv_1.bias"])
        y = torch.relu(y)
        y = F.conv1d(F.pad(y * m3, (pl, pr)), w[f + ".conv_2.weight"], w[f + ".conv_2.bias"]) * m3
        n2 = f"enc_p.encoder.norm_layers_2.{i}"
        h = channel_layer_norm(h + y, w[n2 + ".gamma"], w[n2 + ".beta"])
    h = h * m3
    stats = F.conv1d(h, w["enc_p.proj.weight"], w["enc_p.proj.bias"]) * m3
    C = cfg["inter_channels"]
    return h, stats[:, :C], stats[:, C:], m3


# --------------------------------------------------------------------------- durations
def duration_predictor(w, h, m3, g):
    """duration_predictors.py:297-311 (deterministic DP, v3 configs)."""
    x = h
    if g is not None:
        x = x + F.conv1d(g, w["dp.cond.weight"], w["dp.cond.bias"])
    x = torch.relu(F.conv1d(x * m3, w["dp.conv_1.weight"], w["dp.conv_1.bias"], padding=1))
    x = channel_layer_norm(x, w["dp.norm_1.gamma"], w["dp.norm_1.beta"])
    x = torch.relu(F.conv1d(x * m3, w["dp.conv_2.weight"], w["dp.conv_2.bias"], padding=1))
    x = channel_layer_norm(x, w["dp.norm_2.gamma"], w["dp.norm_2.beta"])
    return F.conv1d(x * m3, w["dp.proj.weight"], w["dp.proj.bias"]) * m3


def dds_conv(w, prefix, x, m3, g=None):
    """duration_predictors.py:45-57"""
    if g is not None:
        x = x + g
    C = x.shape[1]
    for i in range(3):
        d = 3 ** i
        y = F.conv1d(x * m3, w[f"{prefix}.convs_sep.{i}.weight"], w[f"{prefix}.convs_sep.{i}.bias"],
                     padding=d, dilation=d, groups=C)
        y = F.gelu(channel_layer_norm(y, w[f"{prefix}.norms_1.{i}.gamma"], w[f"{prefix}.norms_1.{i}.beta"]))
        y = F.conv1d(y, w[f"{prefix}.convs_1x1.{i}.weight"], w[f"{prefix}.convs_1x1.{i}.bias"])
        y = F.gelu(channel_layer_norm(y, w[f"{prefix}.norms_2.{i}.gamma"], w[f"{prefix}.norms_2.{i}.beta"]))
        x = x + y
    return x * m3


def rqs_inverse(y, uw, uh, ud, bound=5.0, min_w=1e-3, min_h=1e-3, min_d=1e-3):
    """transforms.py:47-97 (linear tails) + :100-187 inverse branch, written densely:
    every element is evaluated, then `|y|>bound` elements pass through.  y [...],
    uw/uh [...,K], ud [...,K-1]."""
    K = uw.shape[-1]
    inside = (y >= -bound) & (y <= bound)
    const = math.log(math.exp(1 - min_d) - 1)
    ud = F.pad(ud, (1, 1), value=const)
    widths = min_w + (1 - min_w * K) * torch.softmax(uw, dim=-1)
    cw = F.pad(torch.cumsum(widths, dim=-1), (1, 0)) * (2 * bound) - bound
    cw[..., 0] = -bound
    cw[..., -1] = bound
    widths = cw[..., 1:] - cw[..., :-1]
    derivs = min_d + F.softplus(ud)
    heights = min_h + (1 - min_h * K) * torch.softmax(uh, dim=-1)
    ch = F.pad(torch.cumsum(heights, dim=-1), (1, 0)) * (2 * bound) - bound
    ch[..., 0] = -bound
    ch[..., -1] = bound
    heights = ch[..., 1:] - ch[..., :-1]
    yc = y.clamp(-bound, bound)
    edges = ch.clone()
    edges[..., -1] += 1e-6  # transforms.py:42-44
    bin_idx = ((yc[..., None] >= edges).sum(dim=-1) - 1).clamp(0, K - 1)[..., None]

    def pick(t):
        return t.gather(-1, bin_idx)[..., 0]

    in_cw, in_w, in_ch, in_h = pick(cw), pick(widths), pick(ch), pick(heights)
    delta = pick(heights / widths)
    d0, d1 = pick(derivs), pick(derivs[..., 1:])
    t = (yc - in_ch) * (d0 + d1 - 2 * delta)
    a = t + in_h * (delta - d0)
    b = in_h * d0 - t
    c = -delta * (yc - in_ch)
    disc = b * b - 4 * a * c
    root = (2 * c) / (-b - torch.sqrt(disc))
    x = root * in_w + in_cw
    return torch.where(inside, x, y)


def conv_flow_reverse(w, prefix, z, m3, cond, H):
    """duration_predictors.py:90-122 with reverse=True"""
    x0, x1 = z[:, :1], z[:, 1:]
    h = F.conv1d(x0, w[prefix + ".pre.weight"], w[prefix + ".pre.bias"])
    h = dds_conv(w, prefix + ".convs", h, m3, g=cond)
    h = F.conv1d(h, w[prefix + ".proj.weight"], w[prefix + ".proj.bias"]) * m3  # [B,29,T]
    hp = h.transpose(1, 2)  # [B,T,29]
    s = math.sqrt(H)
    x1n = rqs_inverse(x1[:, 0], hp[..., :10] / s, hp[..., 10:20] / s, hp[..., 20:])
    return torch.cat([x0, x1n[:, None]], dim=1) * m3


def sdp_reverse(w, cfg, h, m3, g, noise_w, noise_scale_w):
    """duration_predictors.py:213-219,254-263.  noise_w [B,2,Tx] ~ N(0,1)."""
    H = cfg["hidden_channels"]
    x = F.conv1d(h, w["dp.pre.weight"], w["dp.pre.bias"])
    if g is not None:
        x = x + F.conv1d(g, w["dp.cond.weight"], w["dp.cond.bias"])
    x = dds_conv(w, "dp.convs", x, m3)
    x = F.conv1d(x, w["dp.proj.weight"], w["dp.proj.bias"]) * m3
    z = noise_w * noise_scale_w
    # reversed(flows)[:-2] + [flows[0]]: Flip, CF7, Flip, CF5, Flip, CF3, Flip, EA0
    for j in (7, 5, 3):
        z = torch.flip(z, [1])
        z = conv_flow_reverse(w, f"dp.flows.{j}", z, m3, x, H)
    z = torch.flip(z, [1])
    z = (z - w["dp.flows.0.m"]) * torch.exp(-w["dp.flows.0.logs"]) * m3
    return z[:, :1]


# --------------------------------------------------------------------------- regulation
def length_regulate(logw, m3, length_scale, durations=None):
    """models.py:254-265 + commons.py:120-136.  The one-hot `attn @ m_p^T` equals a
    repeat-by-duration gather.  Returns w_ceil [B,1,Tx], y_lengths [B],
    frame->phoneme index [B,Ty] (-1 on padding)."""
    if durations is None:
        w_ceil = torch.ceil(torch.exp(logw) * m3 * length_scale)
    else:
        w_ceil = durations.to(torch.float32).reshape(logw.shape) * m3
    y_lengths = torch.clamp_min(w_ceil.sum(dim=[1, 2]), 1).long()
    Ty = int(y_lengths.max())
    cum = torch.cumsum(w_ceil[:, 0], dim=-1)  # [B,Tx]
    frames = torch.arange(Ty, dtype=torch.float32)
    idx = (frames[None, :, None] >= cum[:, None, :]).sum(dim=-1)  # [B,Ty]
    covered = frames[None, :] < cum[:, -1:]
    y_mask = seq_mask(y_lengths, Ty)
    idx = torch.where(covered & (y_mask > 0), idx, torch.full_like(idx, -1))
    return w_ceil, y_lengths, idx, y_mask[:, None, :]


def expand_by_index(x, idx):
    """x [B,C,Tx], idx [B,Ty] -> [B,C,Ty], zeros where idx<0."""
    safe = idx.clamp_min(0)
    out = torch.gather(x, 2, safe[:, None, :].expand(-1, x.shape[1], -1))
    return out * (idx >= 0)[:, None, :].to(x.dtype)


def dense_path(idx, Tx):
    """The reference's attn tensor [B,1,Ty,Tx] (one-hot rows)."""
    return (idx[:, :, None] == torch.arange(Tx)[None, None, :]).to(torch.float32)[:, None]


# --------------------------------------------------------------------------- VITS2 transformer flow pieces
def plain_encoder(w, prefix, x, mask, n_layers, n_heads, ks):
    """attentions.Encoder.forward (attentions.py:70-87) with window_size=None: plain multi-head attention
    (attentions.py:232-246,262-272, no relative-position terms), FFN k=ks (:403-411), channel LayerNorms.
    x [B,C,T], mask [B,T]."""
    B, C, T = x.shape
    dk = C // n_heads
    m3 = mask[:, None, :]
    pl, pr = (ks - 1) // 2, ks // 2
    x = x * m3
    pair = mask[:, None, :, None] * mask[:, None, None, :]
    for i in range(n_layers):
        a = f"{prefix}.attn_layers.{i}"
        q = F.conv1d(x, w[a + ".conv_q.weight"], w[a + ".conv_q.bias"])
        k = F.conv1d(x, w[a + ".conv_k.weight"], w[a + ".conv_k.bias"])
        v = F.conv1d(x, w[a + ".conv_v.weight"], w[a + ".conv_v.bias"])
        qh = q.view(B, n_heads, dk, T).transpose(2, 3) / math.sqrt(dk)
        kh = k.view(B, n_heads, dk, T).transpose(2, 3)
        vh = v.view(B, n_heads, dk, T).transpose(2, 3)
        scores = (qh @ kh.transpose(-2, -1)).masked_fill(pair == 0, -1e4)
        y = (torch.softmax(scores, dim=-1) @ vh).transpose(2, 3).reshape(B, C, T)
        y = F.conv1d(y, w[a + ".conv_o.weight"], w[a + ".conv_o.bias"])
        n1 = f"{prefix}.norm_layers_1.{i}"
        x = channel_layer_norm(x + y, w[n1 + ".gamma"], w[n1 + ".beta"])
        f = f"{prefix}.ffn_layers.{i}"
        y = torch.relu(F.conv1d(F.pad(x * m3, (pl, pr)), w[f + ".conv_1.weight"], w[f + ".conv_1.bias"]))
        y = F.conv1d(F.pad(y * m3, (pl, pr)), w[f + ".conv_2.weight"], w[f + ".conv_2.bias"]) * m3
        n2 = f"{prefix}.norm_layers_2.{i}"
        x = channel_layer_norm(x + y, w[n2 + ".gamma"], w[n2 + ".beta"])
    return x * m3


# --------------------------------------------------------------------------- flow
def wn(w, prefix, x, m3, g, H, n_layers=4):
    """modules.py:60-87"""
    out = torch.zeros_like(x)
    if g is not None:
        gc = F.conv1d(g, w[prefix + ".cond_layer.weight"], w[prefix + ".cond_layer.bias"])
    for i in range(n_layers):
        a = F.conv1d(x, w[f"{prefix}.in_layers.{i}.weight"], w[f"{prefix}.in_layers.{i}.bias"], padding=2)
        if g is not None:
            a = a + gc[:, 2 * H * i: 2 * H * (i + 1)]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
        rs = F.conv1d(acts, w[f"{prefix}.res_skip_layers.{i}.weight"], w[f"{prefix}.res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = (x + rs[:, :H]) * m3
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out * m3


def flow_reverse(w, cfg, z, y_m3, g):
    """flows.py:442-449 reversed list + flows.py:494-513 (mean_only, reverse)."""
    cfg = _plain(cfg)
    H = cfg["hidden_channels"]
    half = cfg["inter_channels"] // 2
    tflow = cfg.get("use_transformer_flows", False)      # 'pre_conv' type: flows.py:89-176
    for f in (6, 4, 2, 0):
        z = torch.flip(z, [1])
        p = f"flow.flows.{f}"
        x0, x1 = z[:, :half], z[:, half:]
        xin = x0
        if tflow:   # flows.py:148-151: x0_ = pre_transformer(x0 * mask, mask) + x0 ; pre() sees x0_
            xin = plain_encoder(w, p + ".pre_transformer", x0 * y_m3, y_m3[:, 0], 2, 2, 3) + x0
        h = F.conv1d(xin, w[p + ".pre.weight"], w[p + ".pre.bias"]) * y_m3
        h = wn(w, p + ".enc", h, y_m3, g, H)
        m = F.conv1d(h, w[p + ".post.weight"], w[p + ".post.bias"]) * y_m3
        x1 = (x1 - m) * y_m3
        z = torch.cat([x0, x1], dim=1)
    return z


# --------------------------------------------------------------------------- Vocos generator
def istft_hann(re, im, n_fft, hop):
    """torch.istft(center=True, periodic hann window, normalized=False, onesided) restated: per-frame inverse real DFT,
    windowing, overlap-add, division by the overlap-added squared window, n_fft/2 trimmed at both ends
    (what torchaudio's InverseSpectrogram, decoders.py:277-282,303-306, evaluates).  re, im [B, n_fft/2+1, F]."""
    B, _, Fr = re.shape
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * k[:, None] * n[None, :] / n_fft                       # [K, N]
    ck = torch.full((n_fft // 2 + 1,), 2.0, dtype=torch.float64)
    ck[0] = ck[-1] = 1.0
    cosb, sinb = (ck[:, None] * torch.cos(ang)) / n_fft, (ck[:, None] * torch.sin(ang)) / n_fft
    sinb[0] = sinb[-1] = 0.0                                                   # imaginary parts of DC / Nyquist are ignored
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
    frames = (re.double().transpose(1, 2) @ cosb - im.double().transpose(1, 2) @ sinb) * win    # [B, F, N]
    L = hop * (Fr - 1) + n_fft
    out = torch.zeros(B, L, dtype=torch.float64)
    env = torch.zeros(L, dtype=torch.float64)
    for f in range(Fr):
        out[:, f * hop: f * hop + n_fft] += frames[:, f]
        env[f * hop: f * hop + n_fft] += win * win
    half = n_fft // 2
    return (out[:, half: L - half] / env[half: L - half]).float()


def vocos_generator(w, cfg, z, g):
    """VocosGenerator.forward (decoders.py:287-307) with ConvNeXtLayer.forward (:239-247)."""
    cfg = _plain(cfg)
    x = F.pad(z, (1, 0), mode="reflect")
    x = F.conv1d(x, w["dec.in_conv.weight"], w["dec.in_conv.bias"])
    if g is not None:
        x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    x = channel_layer_norm(x, w["dec.norm_pre.gamma"], w["dec.norm_pre.beta"])
    C = x.shape[1]
    for i in range(cfg.get("vocos_num_layers", 8)):
        p = f"dec.layers.{i}"
        r = x
        x = F.conv1d(x, w[p + ".dw_conv.weight"], w[p + ".dw_conv.bias"], padding=1, groups=C)
        x = channel_layer_norm(x, w[p + ".norm.gamma"], w[p + ".norm.beta"])
        x = F.gelu(F.conv1d(x, w[p + ".pw_conv1.weight"], w[p + ".pw_conv1.bias"]))
        x = F.conv1d(x, w[p + ".pw_conv2.weight"], w[p + ".pw_conv2.bias"])
        x = r + w[p + ".scale"] * x
    x = channel_layer_norm(x, w["dec.norm_post.gamma"], w["dec.norm_post.beta"])
    x = F.conv1d(x, w["dec.out_conv.weight"], w["dec.out_conv.bias"])
    mag, phase = x.chunk(2, dim=1)
    mag = mag.exp().clamp_max(1e2)
    ic = cfg.get("vocos_istft_config", {"n_fft": 1024, "hop_length": 256})
    return istft_hann(mag * phase.cos(), mag * phase.sin(), ic["n_fft"], ic["hop_length"])[:, None, :]


# --------------------------------------------------------------------------- generator
def generator(w, cfg, z, g):
    """decoders.py:63-82 + ResBlock1 :157-170 / ResBlock2 :205-214."""
    cfg = _plain(cfg)
    if cfg.get("vocoder_type", "hifigan") == "vocos":
        return vocos_generator(w, cfg, z, g)
    x = F.conv1d(z, w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
    if g is not None:
        x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    ks, ds = cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]
    nk = len(ks)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, w[f"dec.ups.{i}.weight"], w[f"dec.ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            p = f"dec.resblocks.{i * nk + j}"
            r = x
            if str(cfg["resblock"]) == "1":
                for n, d in enumerate(ds[j]):
                    t = F.conv1d(F.leaky_relu(r, LRELU_SLOPE), w[f"{p}.convs1.{n}.weight"],
                                 w[f"{p}.convs1.{n}.bias"], dilation=d, padding=(ks[j] - 1) * d // 2)
                    t = F.conv1d(F.leaky_relu(t, LRELU_SLOPE), w[f"{p}.convs2.{n}.weight"],
                                 w[f"{p}.convs2.{n}.bias"], padding=(ks[j] - 1) // 2)
                    r = t + r
            else:
                for n, d in enumerate(ds[j]):
                    t = F.conv1d(F.leaky_relu(r, LRELU_SLOPE), w[f"{p}.convs.{n}.weight"],
                                 w[f"{p}.convs.{n}.bias"], dilation=d, padding=(ks[j] - 1) * d // 2)
                    r = t + r
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (decoders.py:78)
    x = F.conv1d(x, w["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------- end to end
def infer(sd, cfg, x_ids, x_lengths, sid=None, noise_scale=1.0, length_scale=1.0, noise_scale_w=1.0,
          max_len=None, noise_w=None, noise_z=None, durations=None, folded=False):
    """models.py:228-280.  `sd` is a reference-format state dict (weight_g/weight_v pairs
    unless folded=True).  noise_w [B,2,Tx], noise_z [B,192,>=Ty] are explicit standard
    normal draws (SURVEY §0 finding 7); `durations` teacher-forces w_ceil (finding 8).
    Returns dict with o, attn_idx, y_mask, y_lengths, z, z_p, m_p, logs_p, logw, w_ceil, h."""
    cfg = cfg.to_dict() if hasattr(cfg, "to_dict") else dict(cfg)
    w = sd if folded else fold_weight_norm(sd)
    g = None
    if "emb_g.weight" in w and sid is not None:
        g = F.embedding(sid, w["emb_g.weight"])[:, :, None]
    h, m_p, logs_p, m3 = text_encoder(w, cfg, x_ids, x_lengths)
    if cfg.get("use_sdp", True):
        if noise_w is None:
            noise_w = torch.randn(x_ids.shape[0], 2, x_ids.shape[1])
        logw = sdp_reverse(w, cfg, h, m3, g, noise_w, noise_scale_w)
    else:
        logw = duration_predictor(w, h, m3, g)
    w_ceil, y_lengths, idx, y_m3 = length_regulate(logw, m3, length_scale, durations)
    m_e, logs_e = expand_by_index(m_p, idx), expand_by_index(logs_p, idx)
    Ty = idx.shape[1]
    if noise_z is None:
        noise_z = torch.randn(x_ids.shape[0], m_e.shape[1], Ty)
    z_p = m_e + noise_z[:, :, :Ty] * torch.exp(logs_e) * noise_scale
    z = flow_reverse(w, cfg, z_p, y_m3, g)
    o = generator(w, cfg, (z * y_m3)[:, :, :max_len], g)
    return dict(o=o, attn_idx=idx, y_mask=y_m3, y_lengths=y_lengths, z=z, z_p=z_p, m_p=m_e,
                logs_p=logs_e, logw=logw, w_ceil=w_ceil, h=h, m_p_tx=m_p, logs_p_tx=logs_p, x_mask=m3)
