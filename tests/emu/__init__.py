"""TEST INFRASTRUCTURE: host emulation of the CUDA engine (see cuda_emu.h, build_emu.py)."""
