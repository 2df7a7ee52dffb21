#!/usr/bin/env python
"""Build-time asset preparation (run by ``__graft_entry__.build()`` in the
build container, where the reference checkout is mounted).

Reads the reference's DATA files (checkpoints, sample wavs, the golden output
table) and writes them in this repo's own formats under ``assets/_built/``.
That directory is git-ignored (kept out of history like the built ``.so``) but
NOT gpurun-ignored, so it travels to the GPU box, where ``/root/reference`` does
not exist.  No reference source code is read or copied.

  vap_jp_20hz_2500msec.vapw        jp_20hz_2500msec VAP + CPC weights (VAPW blob)
  vap_bc_erica_20hz_5000msec.vapw  backchannel model (config 5)
  vap_state_dict_jp_{10hz_5000msec,5hz_3000msec,10hz_5000msec_MC}.vapw, vap-bc_state_dict_erica_20hz_3000msec.vapw
                                   10 Hz / 5 Hz / multi-condition / 3 s backchannel checkpoints
  jpn_pair_16k.npz                 int16 L/R of jpn_inoue / jpn_sumida (golden input)
  golden_offline.npy               rvap/vap_main/output_offline.txt as float64 [5312, 5]
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REF = os.environ.get("VAP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "assets", "_built")

CHECKPOINTS = {
    "vap_jp_20hz_2500msec.vapw": "asset/vap/vap_state_dict_jp_20hz_2500msec.pt",
    "vap_bc_erica_20hz_5000msec.vapw": "asset/vap_bc/vap-bc_state_dict_erica_20hz_5000msec.pt",
    # the other rates / modes pinned by tests/golden/ref_rates.npz (tools/make_golden_rates.py); named like the reference's
    # files so that vap_realtime_b200.model._find_weights resolves them the way load_vap_model names them
    "vap_state_dict_jp_10hz_5000msec.vapw": "asset/vap/vap_state_dict_jp_10hz_5000msec.pt",
    "vap_state_dict_jp_5hz_3000msec.vapw": "asset/vap/vap_state_dict_jp_5hz_3000msec.pt",
    "vap_state_dict_jp_10hz_5000msec_MC.vapw": "asset/vap/vap_state_dict_jp_10hz_5000msec_MC.pt",
    "vap-bc_state_dict_erica_20hz_3000msec.vapw": "asset/vap_bc/vap-bc_state_dict_erica_20hz_3000msec.pt",
}
CPC = "asset/cpc/60k_epoch4-d0f474de.pt"


def main(force: bool = False) -> int:
    if not os.path.isdir(REF):
        print(f"[prepare_assets] {REF} not present; keeping whatever is in {OUT}")
        return 0
    from scipy.io import wavfile
    from vap_realtime_b200 import weights

    os.makedirs(OUT, exist_ok=True)
    for name, rel in CHECKPOINTS.items():
        dst = os.path.join(OUT, name)
        if os.path.exists(dst) and not force:
            continue
        tensors = weights.load_reference_checkpoints(os.path.join(REF, rel), os.path.join(REF, CPC))
        weights.save(dst, tensors)
        print(f"[prepare_assets] wrote {dst} ({os.path.getsize(dst) / 1e6:.1f} MB, {len(tensors)} tensors)")

    dst = os.path.join(OUT, "jpn_pair_16k.npz")
    if not os.path.exists(dst) or force:
        srl, left = wavfile.read(os.path.join(REF, "input/wav_sample/jpn_inoue_16k.wav"))
        srr, right = wavfile.read(os.path.join(REF, "input/wav_sample/jpn_sumida_16k.wav"))
        assert srl == srr == 16000 and left.dtype == np.int16 and right.dtype == np.int16
        n = min(len(left), len(right))
        np.savez_compressed(dst, left=left[:n], right=right[:n])
        print(f"[prepare_assets] wrote {dst}")

    dst = os.path.join(OUT, "golden_offline.npy")
    if not os.path.exists(dst) or force:
        g = np.loadtxt(os.path.join(REF, "rvap/vap_main/output_offline.txt"), delimiter=",", skiprows=1)
        np.save(dst, g)
        print(f"[prepare_assets] wrote {dst} {g.shape}")
    return 0


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
