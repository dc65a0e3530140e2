// bt_kernel.cuh -- Belytschko-Tsay shell CFORC3 (placeholder until the kernel lands).
#pragma once
#include "shell_common.cuh"
static void launch_bt_forces(const ShellParams&, int, cudaStream_t) {}
