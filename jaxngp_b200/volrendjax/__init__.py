"""Host-side mirror of the reference's ``volrendjax`` package
(deps/volume-rendering-jax/src/volrendjax/__init__.py:1-15): same function names, argument names,
argument order, shapes and dtypes; arrays are torch CUDA tensors instead of jax Arrays, and uint32
arrays are carried as torch.int32 with identical bits (torch has no uint32 arithmetic).

Each function validates its inputs exactly like the reference's abstract-eval rules
(marching/abstract.py, integrating/abstract.py, packbits/abstract.py, morton3d/abstract.py), builds
the same opaque descriptor and issues one custom call into libngp_b200.so.
"""
from .marching import march_rays, march_rays_inference
from .integrating import integrate_rays, integrate_rays_inference
from .packbits import packbits
from .morton3d import morton3d, morton3d_invert

__all__ = [
    "integrate_rays",
    "integrate_rays_inference",
    "march_rays",
    "march_rays_inference",
    "morton3d",
    "morton3d_invert",
    "packbits",
]
