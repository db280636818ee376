"""spectrogram-rs_b200: B200-native (sm_100a) frame transform behind spectrogram-rs's
AudioTransform interface.  csrc/ holds the CUDA kernels and the C ABI (include/sbx.h); the Python
modules here are the thin host-side mirror of the reference interface used by tests and bench.py.
Import it as `spectrogram_rs_b200` (see the alias package at the repository root)."""
