"""Drop-in module for the reference's `elastic_diffusion_w_controlnet.py`: same class name and signatures, B200-native
hot path (see elasticdiffusion-official_b200/controlnet.py)."""
import importlib as _il
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)
_cn = _il.import_module("elasticdiffusion-official_b200.controlnet")

ElasticDiffusion = _cn.ElasticDiffusion
CosineScheduler, LinearScheduler, ConstScheduler = _cn.CosineScheduler, _cn.LinearScheduler, _cn.ConstScheduler
TimeIt, timelog = _cn.TimeIt, _cn.timelog

if __name__ == '__main__':   # same flags / outputs as the twin's command line (elastic_diffusion_w_controlnet.py:1342-1436)
    _il.import_module("elasticdiffusion-official_b200.cli").main(ElasticDiffusion, timelog, twin=True)
