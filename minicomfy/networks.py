"""Random-initialised stand-in networks for bench / tests (ComfyUI's `comfy.ldm.*` in the real thing).

There are no checkpoints in this sandbox, and the diffusion network is not part of the path this repository
replaces: the engine calls whatever PyTorch module ComfyUI hands it.  `DiTStandIn` is a Flux-shaped transformer
(2x2 patchify of a 16-channel latent, adaLN-modulated blocks, bf16 `scaled_dot_product_attention`, flow-matching
velocity output) small enough to run 72 times per job in a bench, large enough that one forward is hundreds of
kernels and evicts L2 -- the regime a real inpainting run is in.  Test / bench infrastructure only.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _sincos_2d(h, w, dim, device):
    """fixed 2-D sin/cos position table [1, h*w, dim]"""
    def axis(n, d):
        pos = torch.arange(n, device=device, dtype=torch.float32)[:, None]
        freq = torch.exp(torch.arange(0, d, 2, device=device, dtype=torch.float32) * (-math.log(10000.0) / d))
        return torch.cat([torch.sin(pos * freq), torch.cos(pos * freq)], dim=1)
    ey, ex = axis(h, dim // 2), axis(w, dim // 2)
    return torch.cat([ey[:, None, :].expand(h, w, -1), ex[None, :, :].expand(h, w, -1)], dim=-1).reshape(1, h * w, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio):
        super().__init__()
        self.heads = heads
        self.mod = nn.Linear(dim, 6 * dim)
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)
        self.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.fc2 = nn.Linear(int(dim * mlp_ratio), dim)

    def forward(self, h, vec):
        B, T, D = h.shape
        s1, b1, g1, s2, b2, g2 = self.mod(F.silu(vec))[:, None, :].chunk(6, dim=-1)
        a = F.layer_norm(h, (D,)) * (1 + s1) + b1
        q, k, v = self.qkv(a).view(B, T, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, D)
        h = h + g1 * self.proj(a)
        m = F.layer_norm(h, (D,)) * (1 + s2) + b2
        return h + g2 * self.fc2(F.gelu(self.fc1(m), approximate="tanh"))


class DiTStandIn(nn.Module):
    """`denoiser(x, sigma, cond) -> x0` (minicomfy.BaseModel protocol) for a flow model: x0 = x - sigma * v(x, sigma, cond),
    ComfyUI's `ModelSamplingCONST.calculate_denoised`, with the network's output widened to fp32 like ComfyUI does."""

    def __init__(self, in_ch=16, patch=2, hidden=1024, depth=8, heads=16, mlp_ratio=4.0, dtype=torch.bfloat16):
        super().__init__()
        self.in_ch, self.patch, self.hidden, self.dtype = in_ch, patch, hidden, dtype
        self.proj_in = nn.Linear(in_ch * patch * patch, hidden)
        self.t_embed = nn.Sequential(nn.Linear(256, hidden), nn.SiLU(), nn.Linear(hidden, hidden))
        self.c_embed = nn.Sequential(nn.Linear(1, hidden), nn.SiLU(), nn.Linear(hidden, hidden))
        self.blocks = nn.ModuleList([_Block(hidden, heads, mlp_ratio) for _ in range(depth)])
        self.final_mod = nn.Linear(hidden, 2 * hidden)
        self.proj_out = nn.Linear(hidden, in_ch * patch * patch)
        nn.init.normal_(self.proj_out.weight, std=0.02)
        nn.init.zeros_(self.proj_out.bias)
        self._pos = {}
        self.calls = 0
        self.to(dtype)

    def n_params(self):
        return sum(p.numel() for p in self.parameters())

    def _timestep_embedding(self, t):
        half = 128
        freq = torch.exp(torch.arange(half, device=t.device, dtype=torch.float32) * (-math.log(10000.0) / half))
        a = (t.float() * 1000.0)[:, None] * freq[None]
        return torch.cat([torch.cos(a), torch.sin(a)], dim=-1).to(self.dtype)

    def forward(self, x, sigma, cond):
        self.calls += 1
        B, C, H, W = x.shape
        p = self.patch
        hh, ww = H // p, W // p
        tok = x.reshape(B, C, hh, p, ww, p).permute(0, 2, 4, 1, 3, 5).reshape(B, hh * ww, C * p * p).to(self.dtype)
        key = (hh, ww, x.device)
        pos = self._pos.get(key)
        if pos is None:
            pos = self._pos[key] = _sincos_2d(hh, ww, self.hidden, x.device).to(self.dtype)
        sigma = sigma.reshape(-1).expand(B) if sigma.numel() == 1 else sigma.reshape(B)
        vec = self.t_embed(self._timestep_embedding(sigma))
        vec = vec + self.c_embed(torch.full((B, 1), float(cond), device=x.device, dtype=self.dtype))
        h = self.proj_in(tok) + pos
        for blk in self.blocks:
            h = blk(h, vec)
        shift, scale = self.final_mod(F.silu(vec))[:, None, :].chunk(2, dim=-1)
        h = self.proj_out(F.layer_norm(h, (self.hidden,)) * (1 + scale) + shift)
        v = h.reshape(B, hh, ww, C, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, C, H, W).float()
        return x - sigma.view(B, 1, 1, 1).float() * v
