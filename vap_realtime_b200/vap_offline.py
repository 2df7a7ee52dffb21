"""Drop-in for rvap/vap_main/vap_offline.py: replays two 16 kHz wav files through
``VAPRealTime.process_vap`` (frame 1 120, shift 800, :51-73) and writes the same CSV
(``time_sec,p_now(0=left),...``, :76-86) that output/offline_prediction_visualizer consumes."""
from __future__ import annotations

import argparse

import numpy as np


def read_wav_float32(path):
    from scipy.io import wavfile

    sr, x = wavfile.read(path)
    if sr != 16000:
        raise ValueError(f"{path}: expected 16 kHz")
    return x.astype(np.float32) / 32768.0 if x.dtype == np.int16 else x.astype(np.float32)


def run(vap, data_left, data_right):
    frame_size = vap.audio_frame_size
    shift = vap.audio_frame_size - vap.frame_contxt_padding
    result = []
    for i in range(0, len(data_left), shift):
        if i + frame_size > len(data_left):
            break
        vap.process_vap(data_left[i:i + frame_size], data_right[i:i + frame_size])
        result.append({"t": float(i + frame_size) / vap.sampling_rate, "p_now": vap.result_p_now, "p_future": vap.result_p_future})
    return result


def run_bulk(vap, data_left, data_right, max_batch: int = 256):
    """Same result list as ``run`` for a whole recording, computed in bulk (``vapb_score_offline``): the conv stack
    over many chunks at once, the LSTM sequentially, one transformer window per frame with ``max_batch`` windows per
    launch -- instead of one batch-1 step per frame.  ``vap`` supplies the weights and the geometry; its own stream
    state is not touched (the bulk pass starts from a fresh state like a new ``VAPRealTime``)."""
    from .engine import VapEngine

    eng = getattr(vap, "_bulk_engine", None)
    if eng is None:
        src = vap.engine
        eng = VapEngine(src._weights_blob, frame_hz=vap.frame_rate, ctx_frames=vap.audio_context_len, max_streams=max_batch,
                        head=src.head, device=src.device)
        vap._bulk_engine = eng
    n = min(len(data_left), len(data_right))
    audio = np.stack([np.asarray(data_left[:n], dtype=np.float32), np.asarray(data_right[:n], dtype=np.float32)])
    out = eng.score_offline(audio)
    frame_size = vap.audio_frame_size
    shift = frame_size - vap.frame_contxt_padding
    return [{"t": float(shift * i + frame_size) / vap.sampling_rate, "p_now": [float(o[0]), float(o[1])], "p_future": [float(o[2]), float(o[3])]}
            for i, o in enumerate(out)]


def write_csv(path, result):
    with open(path, "w") as f:
        f.write("time_sec,p_now(0=left),p_now(1=right),p_future(0=left),p_future(1=right)\n")
        for r in result:
            f.write(f"{r['t']},{r['p_now'][0]},{r['p_now'][1]},{r['p_future'][0]},{r['p_future'][1]}\n")


def main(argv=None):
    import torch

    from .vap_main import VAPRealTime

    p = argparse.ArgumentParser()
    p.add_argument("--vap_model", type=str, default="../../asset/vap/vap_state_dict_jp_20hz_2500msec.pt")
    p.add_argument("--cpc_model", type=str, default="../../asset/cpc/60k_epoch4-d0f474de.pt")
    p.add_argument("--filename_output", type=str, default="output_offline.txt")
    p.add_argument("--input_wav_left", type=str, default="../../input/wav_sample/jpn_inoue_16k.wav")
    p.add_argument("--input_wav_right", type=str, default="../../input/wav_sample/jpn_sumida_16k.wav")
    p.add_argument("--vap_process_rate", type=int, default=20)
    p.add_argument("--context_len_sec", type=float, default=2.5)
    p.add_argument("--gpu", action="store_true")
    p.add_argument("--frame_by_frame", action="store_true", help="replay with one process_vap call per frame (the reference's loop) instead of the bulk scorer")
    a = p.parse_args(argv)
    vap = VAPRealTime(a.vap_model, a.cpc_model, torch.device("cuda"), a.vap_process_rate, a.context_len_sec)
    left, right = read_wav_float32(a.input_wav_left), read_wav_float32(a.input_wav_right)
    res = run(vap, left, right) if a.frame_by_frame else run_bulk(vap, left, right)
    write_csv(a.filename_output, res)
    print("Generated output file: ", a.filename_output)


if __name__ == "__main__":
    main()
