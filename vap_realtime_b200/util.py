"""Wire codec of the VAP TCP surface, byte-compatible with the reference's
``rvap/common/util.py`` (twin ``vap_realtime/util.py:78-385``) but vectorised
with numpy instead of one ``struct`` call per sample.

Input  (port 50007): packets of 160 x (f64 LE ch1, f64 LE ch2) = 2 560 bytes
                     (reference util.py:43-62 encode, :93-106 decode).
Output (port 50008): u32 LE length prefix (added by the server, vap_main.py:446-448), then
                     f64 t | u32 n, n x f64 x1 | u32 n, n x f64 x2 | u32 2, p_now | u32 2, p_future | u32 2, vad
                     (reference util.py:122-143 encode, :148-188 decode);
                     bc:  ... | u32 1, p_bc_react | u32 1, p_bc_emo (util.py:193-211).
Function names mirror the reference so callers can switch imports.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

BYTE_ORDER = "little"


def _f64(a) -> np.ndarray:
    if hasattr(a, "detach"):                      # torch tensors (result_vad holds tensors, vap_main.py:320)
        a = a.detach().cpu().numpy()
    if isinstance(a, (list, tuple)):
        a = [float(x.detach().cpu().reshape(-1)[0]) if hasattr(x, "detach") else float(np.asarray(x).reshape(-1)[0])
             if np.ndim(x) else float(x) for x in a]
    return np.asarray(a, dtype="<f8").reshape(-1)


# ---- input packets ------------------------------------------------------------------------
def conv_2floatarray_2_bytearray(arr1: Sequence[float], arr2: Sequence[float]) -> bytes:
    a1, a2 = _f64(arr1), _f64(arr2)
    if len(a1) != len(a2):
        raise ValueError("Two arrays must have the same length")
    return np.stack([a1, a2], axis=1).astype("<f8").tobytes()


def conv_bytearray_2_2floatarray(barr: bytes) -> Tuple[np.ndarray, np.ndarray]:
    a = np.frombuffer(barr, dtype="<f8", count=(len(barr) // 16) * 2).reshape(-1, 2)
    return a[:, 0].copy(), a[:, 1].copy()


def conv_floatarray_2_byte(arr: Sequence[float]) -> bytes:
    return _f64(arr).tobytes()


def conv_bytearray_2_floatarray(barr: bytes) -> List[float]:
    return np.frombuffer(barr, dtype="<f8", count=len(barr) // 8).tolist()


# ---- result packets -----------------------------------------------------------------------
def _pack_fields(t: float, fields: Sequence[Sequence[float]]) -> bytes:
    parts = [struct.pack("<d", float(t))]
    for f in fields:
        a = _f64(f)
        parts.append(struct.pack("<I", a.size))
        parts.append(a.tobytes())
    return b"".join(parts)


def _unpack_fields(barr: bytes, names: Sequence[str]) -> Dict[str, object]:
    out: Dict[str, object] = {"t": struct.unpack_from("<d", barr, 0)[0]}
    idx = 8
    for name in names:
        n = struct.unpack_from("<I", barr, idx)[0]
        idx += 4
        out[name] = np.frombuffer(barr, dtype="<f8", count=n, offset=idx).tolist()
        idx += 8 * n
    return out


_VAP_FIELDS = ("x1", "x2", "p_now", "p_future", "vad")
_BC_FIELDS = ("x1", "x2", "p_bc_react", "p_bc_emo")
_NOD_FIELDS = ("x1", "x2", "p_bc", "p_nod_short", "p_nod_long", "p_nod_long_p")


def conv_vapresult_2_bytearray(vap_result: Dict) -> bytes:
    return _pack_fields(vap_result["t"], [vap_result[k] for k in _VAP_FIELDS])


def conv_bytearray_2_vapresult(barr: bytes) -> Dict:
    return _unpack_fields(barr, _VAP_FIELDS)


def conv_vapresult_2_bytearray_bc(vap_result: Dict) -> bytes:
    return _pack_fields(vap_result["t"], [vap_result[k] for k in _BC_FIELDS])


def conv_bytearray_2_vapresult_bc(barr: bytes) -> Dict:
    return _unpack_fields(barr, _BC_FIELDS)


def conv_vapresult_2_bytearray_nod(vap_result: Dict) -> bytes:
    return _pack_fields(vap_result["t"], [vap_result[k] for k in _NOD_FIELDS])


def conv_bytearray_2_vapresult_nod(barr: bytes) -> Dict:
    return _unpack_fields(barr, _NOD_FIELDS)


def frame_result(payload: bytes) -> bytes:
    """Length prefix the server puts in front of every result (vap_main.py:446-448)."""
    return len(payload).to_bytes(4, BYTE_ORDER) + payload


# ---- scalar helpers kept for import compatibility (reference util.py:13-50, 64-91) -----------
def conv_2int16_2_byte(val1: int, val2: int) -> bytes:
    return int(val1).to_bytes(2, BYTE_ORDER) + int(val2).to_bytes(2, BYTE_ORDER)


def conv_2int16array_2_bytearray(arr1, arr2) -> bytes:
    if len(arr1) != len(arr2):
        raise ValueError("Two arrays must have the same length")
    return b"".join(conv_2int16_2_byte(int(a), int(b)) for a, b in zip(arr1, arr2))


def conv_2float_2_byte(val1: float, val2: float) -> bytes:
    return struct.pack("<dd", val1, val2)


conv_float32_2_byte = conv_2float_2_byte


def conv_byte_2_2float(b1: bytes, b2: bytes) -> Tuple[float, float]:
    return struct.unpack("<d", b1)[0], struct.unpack("<d", b2)[0]
