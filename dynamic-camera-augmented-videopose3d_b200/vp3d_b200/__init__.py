"""B200-native (sm_100a) hot path of Dynamic-Camera-Augmented-VideoPose3D.

Importable as `vp3d_b200` with this directory's parent (dynamic-camera-augmented-videopose3d_b200/) on sys.path;
the drop-in replacements of the reference modules live beside it under `common/`.
"""
from . import native  # noqa: F401

__all__ = ['native']
