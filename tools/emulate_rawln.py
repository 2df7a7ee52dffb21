#!/usr/bin/env python
"""CPU emulation: does feeding the RAW (un-normalised) rows to the bf16x3 tensor-core GEMM and applying the LayerNorm
in the epilogue (y = rstd * (x (W*w)^T - mu * s) + c,  s_n = sum_k w_k W_nk,  c_n = sum_k b_k W_nk) keep the 1e-4 gate?
Compares, on the fixture dialogue through the real checkpoint: plain fp32 oracle, the shipped scheme (LayerNorm in fp32,
then hi/lo split) and the raw-A scheme.  Build container only (test infrastructure)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.vap_oracle import D, EPS, OracleState, VapOracle  # noqa: E402
from vap_realtime_b200 import weights  # noqa: E402


def split(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def mm3(x, W):
    xh, xl = split(x)
    wh, wl = split(W)
    return (xl @ wh.T + xh @ wl.T) + xh @ wh.T


class Emu(VapOracle):
    def __init__(self, *a, raw=False, **k):
        super().__init__(*a, **k)
        self.raw = raw

    def lin(self, x, W):
        return mm3(x, W)

    def ln_lin(self, x, name, Ws):
        """LN(x) @ W^T for each W in Ws."""
        w, b = self.w[name + ".weight"], self.w[name + ".bias"]
        if not self.raw:
            z = F.layer_norm(x, (D,), w, b, EPS)
            return [mm3(z, W) for W in Ws]
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        rstd = 1.0 / torch.sqrt(var + EPS)
        outs = []
        for W in Ws:
            Wp = W * w[None, :]
            s = Wp.sum(1)
            c = W @ b
            acc = mm3(x, Wp)
            outs.append(rstd * (acc - mu * s[None, :]) + c[None, :])
        return outs

    def attn(self, prefix, q, k, v):
        N, t, _ = q.shape
        H, HD = 4, 64
        q = q.view(N, t, H, HD).transpose(1, 2)
        k = k.view(N, t, H, HD).transpose(1, 2)
        v = v.view(N, t, H, HD).transpose(1, 2)
        att = torch.einsum("bhid,bhjd->bhij", q, k) * (1.0 / 16.0)
        j = torch.arange(t, dtype=torch.float32)
        att = att + self.w[prefix + "m"].view(1, H, 1, 1) * j.view(1, 1, 1, t) + torch.full((t, t), float("-inf")).triu(1)
        att = F.softmax(att, dim=-1)
        y = (att @ v).transpose(1, 2).reshape(N, t, D)
        return y

    def layer(self, prefix, x, src):
        N, t, _ = x.shape
        x2 = x.reshape(N * t, D)
        q, k, v = self.ln_lin(x2, prefix + "ln_self_attn", [self.w[prefix + "mha.query.weight"], self.w[prefix + "mha.key.weight"], self.w[prefix + "mha.value.weight"]])
        o = self.attn(prefix + "mha.", q.view(N, t, D), k.view(N, t, D), v.view(N, t, D))
        x2 = x2 + self.lin(o.reshape(N * t, D), self.w[prefix + "mha.proj.weight"])
        if src is not None:
            (q,) = self.ln_lin(x2, prefix + "ln_src_attn", [self.w[prefix + "mha_cross.query.weight"]])
            s2 = src.reshape(N * t, D)
            k = self.lin(s2, self.w[prefix + "mha_cross.key.weight"])
            v = self.lin(s2, self.w[prefix + "mha_cross.value.weight"])
            o = self.attn(prefix + "mha_cross.", q.view(N, t, D), k.view(N, t, D), v.view(N, t, D))
            x2 = x2 + self.lin(o.reshape(N * t, D), self.w[prefix + "mha_cross.proj.weight"])
        (h,) = self.ln_lin(x2, prefix + "ln_ffnetwork", [self.w[prefix + "ffnetwork.0.weight"]])
        x2 = x2 + self.lin(F.gelu(h), self.w[prefix + "ffnetwork.3.weight"])
        return x2.view(N, t, D)


def main():
    torch.set_num_threads(8)
    w = weights.load(os.path.join(ROOT, "assets/_built/vap_jp_20hz_2500msec.vapw"))
    d = np.load(os.path.join(ROOT, "tests/golden/ref_vap_ctx2500.npz"))
    audio = d["audio"].astype(np.float32) / 32768.0
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    variants = {"fp32": VapOracle(w, 20, 50, "vap"), "ln_then_split": Emu(w, 20, 50, "vap", raw=False), "raw_A": Emu(w, 20, 50, "vap", raw=True)}
    cases = {"as recorded": audio, "right silent": audio * np.array([[1.0], [0.0]], dtype=np.float32), "-60 dB": audio * np.float32(1e-3)}
    for cname, a in cases.items():
        sts = {k: OracleState(1) for k in variants}
        worst = {k: 0.0 for k in variants}
        stats = []
        for i in range(n):
            c = a[None, :, 800 * i: 800 * i + 1120]
            outs = {k: v.step(c, sts[k]).numpy()[0] for k, v in variants.items()}
            for k in variants:
                worst[k] = max(worst[k], float(np.abs(outs[k] - outs["fp32"]).max()))
        print(cname, {k: f"{v:.2e}" for k, v in worst.items()})


if __name__ == "__main__":
    main()
