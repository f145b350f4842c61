/* empty on purpose: tests/emu/cuda_emu.h (force-included) stands in for the CUDA headers */
