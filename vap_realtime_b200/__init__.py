"""vap_realtime_b200: B200-native (sm_100a) implementation of VAP-Realtime's streaming step.

Public surface (mirrors the reference's):
    VapEngine                      batched C-ABI engine (one call = one step of B dialogues)
    vap_main.VAPRealTime           drop-in for rvap/vap_main/vap_main.py (class + TCP server)
    vap_bc_main.VAPRealTime        drop-in for rvap/vap_bc/vap_bc_main.py
    Vap / VapModel, VapInput       drop-in for the ``vap_realtime`` (maai) library API
    server.BatchedVapServer        many dialogues on one GPU, reference wire format
    dist.ShardedVap                one process per GPU, NCCL scatter / gather
"""
from . import input as VapInput  # noqa: N812  (the reference exports the module under this name)


def __getattr__(name):
    # lazy: importing the package must not require torch / CUDA (weights packing, codec tests)
    if name == "VapEngine":
        from .engine import VapEngine
        return VapEngine
    if name in ("Vap", "VapModel"):
        from .model import Vap
        return Vap
    raise AttributeError(name)
