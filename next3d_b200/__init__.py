"""next3d_b200 -- Blackwell-native (sm_100a) implementation of the Next3D generator forward hot path
(`TriPlaneGenerator.synthesis`), behind the reference's torch_utils.ops / TriPlaneGenerator API.

Nothing in this package imports `oracle/`; the CUDA extension is mandatory (see _lib.py)."""
__version__ = '0.1.0'
