"""Drop-in module: `from elastic_diffusion import ElasticDiffusion` works as with the reference's file of the same
name (/root/reference/elastic_diffusion.py), but the class is the B200-native implementation in
`elasticdiffusion-official_b200/`."""
import importlib as _il
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)
_pkg = _il.import_module("elasticdiffusion-official_b200")

ElasticDiffusion = _pkg.ElasticDiffusion
CosineScheduler = _pkg.CosineScheduler
LinearScheduler = _pkg.LinearScheduler
ConstScheduler = _pkg.ConstScheduler
TimeIt = _pkg.TimeIt
timelog = _pkg.timelog
package = _pkg

if __name__ == '__main__':   # same flags / outputs as the reference's command line (elastic_diffusion.py:1134-1210)
    _il.import_module("elasticdiffusion-official_b200.cli").main(ElasticDiffusion, timelog)
