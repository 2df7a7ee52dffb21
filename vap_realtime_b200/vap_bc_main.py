"""Drop-in for ``rvap/vap_bc/vap_bc_main.py``: the backchannel twin of VAPRealTime
(``process_vap`` :241-300; results ``result_p_bc_react`` / ``result_p_bc_emo`` as
one-element lists of tensors, :283-284; wire ``conv_vapresult_2_bytearray_bc``)."""
from __future__ import annotations

import argparse
import copy
import threading
import time

from . import util
from .vap_main import _StreamFrontEnd, proc_serv_in, proc_serv_out, proc_serv_out_dist


class VAPRealTime(_StreamFrontEnd):
    HEAD = "bc"

    def __init__(self, vap_model, cpc_model, device, frame_rate, context_len_sec, **kw):
        super().__init__(vap_model, cpc_model, device, frame_rate, context_len_sec, **kw)
        self.result_p_bc_react = 0.
        self.result_p_bc_emo = 0.

    def process_vap(self, x1, x2):
        time_start = time.time()
        self.current_x1_audio = x1[self.frame_contxt_padding:]
        self.current_x2_audio = x2[self.frame_contxt_padding:]
        o = self._run(x1, x2)
        torch = self._torch
        self.result_p_bc_react = [torch.tensor([float(o[0])])]
        self.result_p_bc_emo = [torch.tensor([float(o[1])])]
        self.result_last_time = time.time()
        self._tick(time_start)


def _result_dict_bc(vap):
    return {
        "t": copy.copy(vap.result_last_time),
        "x1": copy.copy(vap.current_x1_audio), "x2": copy.copy(vap.current_x2_audio),
        "p_bc_react": copy.copy(vap.result_p_bc_react), "p_bc_emo": copy.copy(vap.result_p_bc_emo),
    }


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--vap_model", type=str, default='../../asset/vap_bc/vap-bc_state_dict_erica_20hz_5000msec.pt')
    parser.add_argument("--cpc_model", type=str, default='../../asset/cpc/60k_epoch4-d0f474de.pt')
    parser.add_argument("--port_num_in", type=int, default=50007)
    parser.add_argument("--port_num_out", type=int, default=50008)
    parser.add_argument("--vap_process_rate", type=int, default=20)
    parser.add_argument("--context_len_sec", type=float, default=5)
    parser.add_argument("--gpu", action='store_true')
    args = parser.parse_args(argv)

    import torch
    device = torch.device('cuda')
    print('Device: ', device)
    vap = VAPRealTime(args.vap_model, args.cpc_model, device, args.vap_process_rate, args.context_len_sec)
    list_socket_out = []
    threading.Thread(target=proc_serv_out, args=(list_socket_out, args.port_num_out), daemon=True).start()
    threading.Thread(target=proc_serv_out_dist, args=(list_socket_out, vap, _result_dict_bc, util.conv_vapresult_2_bytearray_bc),
                     daemon=True).start()
    proc_serv_in(args.port_num_in, vap)


if __name__ == "__main__":
    main()
