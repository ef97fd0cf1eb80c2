"""vsdeoldify_b200 - B200-native (sm_100a) implementation of HAVC's per-frame colorization hot path.

Public surface (same names and argument meaning as dan64/vs-deoldify for this path):
    HAVC_main, HAVC_colorizer, HAVC_deoldify, HAVC_ddeoldify, HAVC_merge, HAVC_stabilizer, ModelImageRender
"""
__version__ = "0.1.0"


def __getattr__(name):   # lazy: importing the package must not import torch / load the CUDA library
    if name in ("HAVC_main", "HAVC_colorizer", "HAVC_deoldify", "HAVC_ddeoldify", "HAVC_merge", "HAVC_stabilizer", "ModelImageRender", "register_state_dict",
                "ddeoldify_main", "ddeoldify", "ddeoldify_stabilizer", "vs_chroma_stabilizer_ex"):
        from . import havc
        return getattr(havc, name)
    raise AttributeError(name)
