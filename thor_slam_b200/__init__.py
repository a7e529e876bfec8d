"""thor-slam ingest path, B200-native.

Only the per-frame-set ingest stage of WT-MM/thor-slam lives here (see
DESIGN.md): the ``CameraSource`` / ``CameraRig`` / ``SlamEngine`` API mirror on
the host and, behind it, hand-written sm_100a CUDA kernels reached through the
C-ABI library ``libthoringest.so`` (``include/thoringest.h``).
"""

__version__ = "0.1.0"
