// tests/emu/emu_lib.cpp -- TEST INFRASTRUCTURE.  Builds libvor_kernel_emu.so: the kernel bodies and host round
// loop of voronoids_b200/csrc compiled for the CPU (sequential "launches") so their logic can be unit-tested in
// the GPU-less container.  Exposes the same C ABI as the product, but it is only ever loaded explicitly by
// tests/test_emu_*.py; voronoids_b200/_lib.py loads libvoronoids_b200.so and nothing else.
#define VOR_EMU 1
#include "backend_emu.h"
#include "engine.cuh"
#include "capi.inl"
