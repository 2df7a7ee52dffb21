"""CPU ORACLE for the VAP streaming step -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product path
(``vap_realtime_b200``) never does: it fails loudly when the CUDA library is
missing.

What it is
----------
A from-scratch, batched, functional restatement of one ``process_vap`` step of
inokoj/VAP-Realtime (``rvap/vap_main/vap_main.py:249-335``) in fp32 on the CPU.
The reference's arithmetic lives in a third-party dependency, PyTorch ATen
(``torch>=2.2.0``, reference ``requirements.txt:2``; this image has 2.11.0):
conv1d / addmm / softmax / layer_norm / gelu.  The oracle calls the same ATen
primitives through ``torch.nn.functional`` (no ``nn.Module`` of the reference is
used, nothing is imported from ``/root/reference``), so its rounding behaviour
is that of the reference's CPU path, and it doubles as the "port" CPU baseline
in ``bench.py``.

Parity pinning
--------------
Pinned (see ``tools/make_golden.py`` and ``tests/test_oracle_golden.py``):
  * against the reference's only golden vector
    ``rvap/vap_main/output_offline.txt`` (first rows committed under
    ``tests/golden/``; all 5 312 rows when ``assets/_built`` is present);
  * against outputs of the unmodified reference imported in the build container
    (vap head + vad, bc head, ctx 2.5 s and 5.0 s), committed as fixtures.

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional

import numpy as np
import torch
import torch.nn.functional as F

D = 256          # model width            (vap_main.py:50)
FF = 768         # FFN width, dff_k = 3   (vap_main.py:110, modules.py:322)
H = 4            # heads                  (vap_main.py:53)
HD = D // H
PAD = 320        # frame_contxt_padding   (vap_main.py:224)
SR = 16000       # sampling rate          (vap_main.py:223)
EPS = 1e-5

# (Cin, Cout, kernel, stride, padding)   encoder_components.py:83-92
CONV_SPECS = [(1, D, 10, 5, 3), (D, D, 8, 4, 2), (D, D, 4, 2, 1), (D, D, 4, 2, 1), (D, D, 4, 2, 1)]

G = "encoder.encoder.gEncoder."
AR = "encoder.encoder.gAR.baseNet."


def chunk_samples(frame_hz: int) -> int:
    """audio_frame_size (vap_main.py:230)."""
    return SR // frame_hz + PAD


def conv_lengths(n: int) -> List[int]:
    out = []
    for (_, _, k, s, p) in CONV_SPECS:
        n = (n + 2 * p - k) // s + 1
        out.append(n)
    return out


def codebook_abp(from_bin: int, to_bin: int) -> torch.Tensor:
    """abp[c, s] = sum_{bin in [from,to]} bit(c, 4*s + bin).

    Codebook.create_code_vectors / single_idx_to_onehot (objective.py:93-110):
    row c of the embedding is the binary expansion of c, LSB first; decode()
    reshapes the 8 digits to (speaker=2, bin=4) (objective.py:141-143);
    probs_next_speaker_aggregate sums bins from..to (objective.py:196-201).
    """
    c = torch.arange(256)
    abp = torch.zeros(256, 2)
    for s in range(2):
        for b in range(from_bin, to_bin + 1):
            abp[:, s] += ((c >> (4 * s + b)) & 1).float()
    return abp


class OracleState:
    """Per-batch stream state: LSTM (h, c) kept forever (encoder.py:27,
    encoder_components.py:148-153) and the list of past embeddings
    (vap_main.py:243-244, 274-280)."""

    def __init__(self, batch: int):
        self.h = torch.zeros(batch, 2, D)
        self.c = torch.zeros(batch, 2, D)
        self.ring: List[torch.Tensor] = []   # each [B, 2, D], oldest first
        self.count = 0


class VapOracle:
    def __init__(self, tensors: Mapping[str, np.ndarray], frame_hz: int = 20,
                 ctx_frames: int = 50, head: str = "vap"):
        self.w: Dict[str, torch.Tensor] = {k: torch.from_numpy(np.array(v, dtype=np.float32, copy=True))
                                           for k, v in tensors.items()}
        self.frame_hz = frame_hz
        self.T = int(ctx_frames)
        self.head = head
        self.S = chunk_samples(frame_hz)
        self.lens = conv_lengths(self.S)
        self.n_lstm = self.lens[-1] - 2
        kd = self.w["encoder.downsample.1.weight"].shape[2]
        if kd != self.n_lstm:
            raise ValueError(f"downsample kernel {kd} does not match {self.n_lstm} LSTM frames at {frame_hz} Hz")
        self.abp_now = codebook_abp(0, 1)      # BINS_P_NOW    vap_main.py:187
        self.abp_fut = codebook_abp(2, 3)      # BINS_PFUTURE  vap_main.py:188

    # ------------------------------------------------------------------ encoder
    @staticmethod
    def channel_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """ChannelNorm.forward (encoder_components.py:62-70): statistics over the
        channel dim, UNBIASED variance (torch.var default), eps inside rsqrt."""
        mean = x.mean(dim=1, keepdim=True)
        var = x.var(dim=1, keepdim=True)
        return (x - mean) * torch.rsqrt(var + EPS) * w + b

    def conv_stack(self, audio: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """CPCEncoder.forward (encoder_components.py:98-104). audio [N,1,S] -> [N,256,L5]."""
        x = audio
        for i, (_, _, k, s, p) in enumerate(CONV_SPECS):
            x = F.conv1d(x, self.w[f"{G}conv{i}.weight"], self.w[f"{G}conv{i}.bias"], stride=s, padding=p)
            x = F.relu(self.channel_norm(x, self.w[f"{G}batchNorm{i}.weight"], self.w[f"{G}batchNorm{i}.bias"]))
            if taps is not None:
                taps[f"conv{i}"] = x.transpose(1, 2).contiguous()      # channels-last [N, L, 256]
        return x

    def lstm(self, z: torch.Tensor, h: torch.Tensor, c: torch.Tensor):
        """CPCAR.forward over nn.LSTM(256,256,1) (encoder_components.py:120-123,
        140-153). z [N,n,256]; gate order i,f,g,o (torch LSTM convention)."""
        w_ih, w_hh = self.w[AR + "weight_ih_l0"], self.w[AR + "weight_hh_l0"]
        b_ih, b_hh = self.w[AR + "bias_ih_l0"], self.w[AR + "bias_hh_l0"]
        ys = []
        for t in range(z.shape[1]):
            g = F.linear(z[:, t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            ys.append(h)
        return torch.stack(ys, dim=1), h, c

    def downsample(self, y: torch.Tensor) -> torch.Tensor:
        """get_cnn_layer Sequential (encoder_components.py:496-511) with the
        weights patched in at vap_main.py:203-212: Conv1d over exactly k frames
        -> LayerNorm(256) -> GELU(erf).  y [N,k,256] -> [N,256]."""
        x = F.conv1d(y.transpose(1, 2), self.w["encoder.downsample.1.weight"], self.w["encoder.downsample.1.bias"])
        x = x[:, :, 0]
        x = F.layer_norm(x, (D,), self.w["encoder.downsample.2.ln.weight"], self.w["encoder.downsample.2.ln.bias"], EPS)
        return F.gelu(x)

    def encode(self, audio: torch.Tensor, st: OracleState, taps: Optional[dict] = None) -> torch.Tensor:
        """VapGPT.encode_audio -> EncoderCPC.forward (vap_main.py:175-180,
        encoder.py:58-80). audio [B,2,S] -> e [B,2,256]; updates (h,c)."""
        B = audio.shape[0]
        x = self.conv_stack(audio.reshape(B * 2, 1, self.S), taps)
        z = x.transpose(1, 2)[:, 1:-1, :]                  # encoder.py:75-76
        y, h, c = self.lstm(z, st.h.reshape(B * 2, D), st.c.reshape(B * 2, D))
        st.h, st.c = h.reshape(B, 2, D), c.reshape(B, 2, D)
        e = self.downsample(y).reshape(B, 2, D)
        if taps is not None:
            taps["lstm_out"] = y.reshape(B, 2, -1, D)
            taps["e"] = e
        return e

    # -------------------------------------------------------------- transformer
    def mha(self, prefix: str, q_in, k_in, v_in):
        """MultiHeadAttention.forward + MultiHeadAttentionAlibi.mask_scores
        (modules.py:82-110, 170-212): scale 1/sqrt(dim)=1/16 (modules.py:52),
        bias m_h * j plus causal -inf, softmax over keys, bias-free projections."""
        N, t, _ = q_in.shape
        q = F.linear(q_in, self.w[prefix + "query.weight"]).view(N, t, H, HD).transpose(1, 2)
        k = F.linear(k_in, self.w[prefix + "key.weight"]).view(N, t, H, HD).transpose(1, 2)
        v = F.linear(v_in, self.w[prefix + "value.weight"]).view(N, t, H, HD).transpose(1, 2)
        att = torch.einsum("bhid,bhjd->bhij", q, k) * (1.0 / math.sqrt(D))
        j = torch.arange(t, dtype=torch.float32)
        alibi = self.w[prefix + "m"].view(1, H, 1, 1) * j.view(1, 1, 1, t)
        causal = torch.full((t, t), float("-inf")).triu(1)
        att = F.softmax(att + alibi + causal, dim=-1)
        y = (att @ v).transpose(1, 2).reshape(N, t, D)
        return F.linear(y, self.w[prefix + "proj.weight"])

    def ln(self, x, name):
        return F.layer_norm(x, (D,), self.w[name + ".weight"], self.w[name + ".bias"], EPS)

    def layer(self, prefix: str, x: torch.Tensor, src: Optional[torch.Tensor]):
        """TransformerLayer.forward (modules.py:257-286); cross-attention K/V are
        the UN-normalised sibling input (modules.py:276-283)."""
        z = self.ln(x, prefix + "ln_self_attn")
        x = x + self.mha(prefix + "mha.", z, z, z)
        if src is not None:
            z = self.ln(x, prefix + "ln_src_attn")
            x = x + self.mha(prefix + "mha_cross.", z, src, src)
        z = self.ln(x, prefix + "ln_ffnetwork")
        hdn = F.gelu(F.linear(z, self.w[prefix + "ffnetwork.0.weight"]))
        return x + F.linear(hdn, self.w[prefix + "ffnetwork.3.weight"])

    def transformer(self, X: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """ar_channel (GPT, modules.py:356-372) on each channel with shared
        weights (vap_main.py:285-286), then GPTStereo (modules.py:395-423) with
        the Combinator (modules.py:461-464), heads and aggregation
        (vap_main.py:290-317; objective.py:186-206).  X [B,2,t,256] -> [B,6]."""
        B, _, t, _ = X.shape
        a = self.layer("ar_channel.layers.0.", X.reshape(B * 2, t, D), None).reshape(B, 2, t, D)
        if taps is not None:
            taps["chan_out"] = a
        x1, x2 = a[:, 0], a[:, 1]
        for li in range(3):
            p = f"ar.layers.{li}."
            z1 = self.layer(p, x1, x2)          # TransformerStereoLayer.forward (modules.py:289-300)
            z2 = self.layer(p, x2, x1)
            x1, x2 = z1, z2
            if taps is not None:
                taps[f"cross{li}_out"] = torch.stack([x1, x2], dim=1)
        ha = F.gelu(self.ln(F.linear(x1[:, -1], self.w["ar.combinator.h0_a.weight"]), "ar.combinator.ln"))
        hb = F.gelu(self.ln(F.linear(x2[:, -1], self.w["ar.combinator.h0_b.weight"]), "ar.combinator.ln"))
        hc = ha + hb
        if taps is not None:
            taps["comb"] = hc
        out = torch.zeros(B, 6)
        if self.head == "vap":
            logits = F.linear(hc, self.w["vap_head.weight"], self.w["vap_head.bias"])        # vap_main.py:290
            probs = logits.softmax(dim=-1)                                                    # vap_main.py:295
            for col, abp in ((0, self.abp_now), (2, self.abp_fut)):
                p = probs @ abp                                                               # objective.py:203
                p = p / (p.sum(-1, keepdim=True) + EPS)                                       # objective.py:205
                out[:, col:col + 2] = p
            # vad from the ar_channel outputs (vap_main.py:292-293, 313-314)
            va = F.linear(a[:, :, -1, :], self.w["va_classifier.weight"], self.w["va_classifier.bias"])
            out[:, 4:6] = torch.sigmoid(va[..., 0])
            if taps is not None:
                taps["logits"] = logits
        else:
            bc = F.linear(hc, self.w["bc_head.weight"], self.w["bc_head.bias"])              # vap_bc_main.py:272
            pb = bc.softmax(dim=-1)
            out[:, 0] = pb[:, 1]                                                              # vap_bc_main.py:276
            out[:, 1] = pb[:, 2]                                                              # vap_bc_main.py:277
            if taps is not None:
                taps["logits"] = bc
        return out

    # --------------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, audio, st: OracleState, taps: Optional[dict] = None) -> torch.Tensor:
        """One process_vap call for a batch of streams that share the same
        frame count (vap_main.py:249-335). audio [B,2,S] float32."""
        audio = torch.as_tensor(audio, dtype=torch.float32)
        e = self.encode(audio, st, taps)
        st.ring.append(e)
        if len(st.ring) > self.T:                         # vap_main.py:277-280
            st.ring = st.ring[-self.T:]
        st.count += 1
        X = torch.stack(st.ring, dim=2)                   # [B,2,t,256]  (vap_main.py:282-283)
        return self.transformer(X, taps)


def synthetic_audio(stream: int, n_frames: int, frame_hz: int = 20) -> np.ndarray:
    """SURVEY 8(d) synthetic input: seed 1234+stream, clamp(0.05*randn, -1, 1),
    [2, shift*n + 320] fp32; chunk n = samples [shift*n, shift*n + S)."""
    g = torch.Generator().manual_seed(1234 + stream)
    shift = SR // frame_hz
    a = torch.randn(2, shift * n_frames + PAD, generator=g) * 0.05
    return a.clamp_(-1, 1).numpy()


def chunks_from_audio(audio: np.ndarray, n: int, frame_hz: int = 20) -> np.ndarray:
    """audio [..., 2, L] -> chunk n [..., 2, S] (vap_offline.py:51-61)."""
    shift = SR // frame_hz
    return audio[..., shift * n: shift * n + shift + PAD]
