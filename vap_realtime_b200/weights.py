"""Weight blob ("VAPW") reader/writer for the B200 VAP streaming path.

The reference loads two pickled state-dicts at start-up
(rvap/vap_main/vap_main.py:199-212): the VAP file (transformer, heads and the
``encoder.downsample.*`` tensors) and the CPC file (``['weights']``: conv stack
+ LSTM, rvap/vap_main/encoder_components.py:179-405).  The C-ABI library takes
ONE flat little-endian blob instead, so that no pickle parsing happens on the
native side.  Tensors keep the reference's own names and shapes (fp32, row
major); every re-layout the kernels want (transposes, bf16 hi/lo planes) is
done on the device by ``vapb_create``.

Blob layout (all little endian):

    0   8s   magic  b"VAPW0001"
    8   u32  n_tensors
    12  u32  table_bytes (n_tensors * 96)
    16  table entries, 96 bytes each:
          64s name (NUL padded) | u32 ndim | 4 x u32 dims | u64 offset | u32 nbytes
        (offset is from the start of the blob, 256-byte aligned)
    ... fp32 payloads
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, Mapping

import numpy as np

MAGIC = b"VAPW0001"
_ENTRY = struct.Struct("<64sI4IQI")  # 96 bytes
assert _ENTRY.size == 96
_ALIGN = 256

CPC_PREFIX = "encoder.encoder."

# Keys of the VAP state-dict that the streaming path never reads
# (rvap/vap_main/vap_main.py:201 loads with strict=False; the codebook is
# recomputed, objective.py:93-110; zero_shot is training-only).
_SKIP_PREFIXES = ("zero_shot.", "objective.")


def collect_tensors(vap_sd: Mapping, cpc_weights: Mapping | None) -> "OrderedDict[str, np.ndarray]":
    """Merge the two state-dicts the way VAPRealTime.__init__ does.

    CPC conv/LSTM tensors come from the CPC file (rvap/vap_main/encoder.py:21-24),
    everything else from the VAP file; ``encoder.downsample.*`` is the manual
    patch at vap_main.py:203-212.
    """
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def to_np(t):
        if hasattr(t, "detach"):
            t = t.detach().cpu().float().numpy()
        return np.ascontiguousarray(np.asarray(t, dtype=np.float32))

    if cpc_weights is not None:
        for k, v in cpc_weights.items():
            if k.startswith("gEncoder.") or k.startswith("gAR."):
                out[CPC_PREFIX + k] = to_np(v)
    for k, v in vap_sd.items():
        if k.startswith(_SKIP_PREFIXES):
            continue
        if k.startswith(CPC_PREFIX):
            if cpc_weights is None:
                out[k] = to_np(v)
            continue
        out[k] = to_np(v)
    return out


def load_reference_checkpoints(vap_model: str, cpc_model: str) -> "OrderedDict[str, np.ndarray]":
    """Read the reference's ``.pt`` files (needs torch; host side only)."""
    import torch

    sd = torch.load(vap_model, map_location="cpu")
    cpc = torch.load(cpc_model, map_location="cpu")
    if isinstance(cpc, dict) and "weights" in cpc:
        cpc = cpc["weights"]
    return collect_tensors(sd, cpc)


def pack(tensors: Mapping[str, np.ndarray]) -> bytes:
    names = list(tensors.keys())
    n = len(names)
    table_bytes = n * _ENTRY.size
    cursor = 16 + table_bytes
    entries = []
    payloads = []
    for name in names:
        a = np.ascontiguousarray(tensors[name], dtype="<f4")
        if a.ndim > 4:
            raise ValueError(f"{name}: ndim {a.ndim} > 4")
        enc = name.encode("ascii")
        if len(enc) > 63:
            raise ValueError(f"tensor name too long: {name}")
        cursor = (cursor + _ALIGN - 1) // _ALIGN * _ALIGN
        dims = list(a.shape) + [1] * (4 - a.ndim)
        entries.append(_ENTRY.pack(enc, a.ndim, *dims, cursor, a.nbytes))
        payloads.append((cursor, a.tobytes()))
        cursor += a.nbytes
    buf = bytearray(cursor)
    buf[0:8] = MAGIC
    struct.pack_into("<II", buf, 8, n, table_bytes)
    off = 16
    for e in entries:
        buf[off:off + _ENTRY.size] = e
        off += _ENTRY.size
    for o, p in payloads:
        buf[o:o + len(p)] = p
    return bytes(buf)


def unpack(blob: bytes) -> "OrderedDict[str, np.ndarray]":
    if blob[:8] != MAGIC:
        raise ValueError("not a VAPW blob")
    n, table_bytes = struct.unpack_from("<II", blob, 8)
    if table_bytes != n * _ENTRY.size:
        raise ValueError("corrupt VAPW table")
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for i in range(n):
        name, ndim, d0, d1, d2, d3, off, nbytes = _ENTRY.unpack_from(blob, 16 + i * _ENTRY.size)
        shape = (d0, d1, d2, d3)[:ndim]
        a = np.frombuffer(blob, dtype="<f4", count=nbytes // 4, offset=off).reshape(shape)
        out[name.rstrip(b"\0").decode("ascii")] = a
    return out


def save(path: str, tensors: Mapping[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(pack(tensors))


def load(path: str) -> "OrderedDict[str, np.ndarray]":
    with open(path, "rb") as f:
        return unpack(f.read())


def infer_frame_hz(tensors: Mapping[str, np.ndarray]) -> int:
    """20/10/5 Hz checkpoints carry a downsample kernel of 5/10/20 taps
    (vap_main.py:203-212)."""
    k = tensors["encoder.downsample.1.weight"].shape[2]
    return {5: 20, 10: 10, 20: 5}[k]


def head_kind(tensors: Mapping[str, np.ndarray]) -> int:
    """0 = vap head (vap_main.py:142), 1 = backchannel head (vap_bc_main.py:137)."""
    return 1 if "bc_head.weight" in tensors else 0


def random_tensors(seed: int = 0, frame_hz: int = 20, bc: bool = False) -> "OrderedDict[str, np.ndarray]":
    """Random-init weights with the reference architecture's shapes (for tests
    and for benches when the real checkpoints are not on the box).  Scales are
    chosen so activations stay O(1) like the trained model."""
    rng = np.random.default_rng(seed)
    D, F = 256, 768
    t: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def w(*shape, fan_in):
        return (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)

    def small(*shape, s=0.1):
        return (rng.standard_normal(shape) * s).astype(np.float32)

    g = CPC_PREFIX + "gEncoder."
    ks = [10, 8, 4, 4, 4]
    for i, k in enumerate(ks):
        cin = 1 if i == 0 else D
        t[f"{g}conv{i}.weight"] = w(D, cin, k, fan_in=cin * k)
        t[f"{g}conv{i}.bias"] = small(D)
        t[f"{g}batchNorm{i}.weight"] = (1.0 + small(1, D, 1)).astype(np.float32)
        t[f"{g}batchNorm{i}.bias"] = small(1, D, 1)
    a = CPC_PREFIX + "gAR.baseNet."
    t[a + "weight_ih_l0"] = w(4 * D, D, fan_in=D)
    t[a + "weight_hh_l0"] = w(4 * D, D, fan_in=D)
    t[a + "bias_ih_l0"] = small(4 * D)
    t[a + "bias_hh_l0"] = small(4 * D)
    kd = {20: 5, 10: 10, 5: 20}[frame_hz]
    t["encoder.downsample.1.weight"] = w(D, D, kd, fan_in=D * kd)
    t["encoder.downsample.1.bias"] = small(D)
    t["encoder.downsample.2.ln.weight"] = (1.0 + small(D)).astype(np.float32)
    t["encoder.downsample.2.ln.bias"] = small(D)

    slopes = np.array([2.0 ** -2, 2.0 ** -4, 2.0 ** -6, 2.0 ** -8], dtype=np.float32)

    def layer(prefix, cross):
        for ln in ["ln_self_attn", "ln_ffnetwork"] + (["ln_src_attn"] if cross else []):
            t[f"{prefix}{ln}.weight"] = (1.0 + small(D)).astype(np.float32)
            t[f"{prefix}{ln}.bias"] = small(D)
        for mha in ["mha"] + (["mha_cross"] if cross else []):
            t[f"{prefix}{mha}.m"] = slopes.copy()
            for p in ["key", "query", "value", "proj"]:
                t[f"{prefix}{mha}.{p}.weight"] = w(D, D, fan_in=D)
        t[f"{prefix}ffnetwork.0.weight"] = w(F, D, fan_in=D)
        t[f"{prefix}ffnetwork.3.weight"] = w(D, F, fan_in=F)

    layer("ar_channel.layers.0.", False)
    for i in range(3):
        layer(f"ar.layers.{i}.", True)
    t["ar.combinator.h0_a.weight"] = w(D, D, fan_in=D)
    t["ar.combinator.h0_b.weight"] = w(D, D, fan_in=D)
    t["ar.combinator.ln.weight"] = (1.0 + small(D)).astype(np.float32)
    t["ar.combinator.ln.bias"] = small(D)
    t["va_classifier.weight"] = w(1, D, fan_in=D)
    t["va_classifier.bias"] = small(1)
    t["vap_head.weight"] = w(D, D, fan_in=D)
    t["vap_head.bias"] = small(D)
    if bc:
        t["bc_head.weight"] = w(3, D, fan_in=D)
        t["bc_head.bias"] = small(3)
    return t
