"""aesr_b200: B200-native (sm_100a) implementation of the ae_combined slice-synthesis hot path of
qurAI-amsterdam/SuperResolution_aniso_MRI.  Host side = Python/PyTorch plumbing; arithmetic = libaesr_b200.so."""
__all__ = ["ops", "synthesis", "build"]
