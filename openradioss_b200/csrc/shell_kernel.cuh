// shell_kernel.cuh -- 4-node shell super-groups (placeholder: the QEPH/BT kernels land next).
#pragma once
#include <vector>
#include "common.cuh"
struct HostShellGroup { int nel, nft, law; orgpu_prop_shell prop; orgpu_law2 m2; orgpu_law36 m36; };
struct ShellSGHost { int first_elem = 0; std::vector<void*> owned; };
static int shell_add_group(std::vector<HostShellGroup>&, int, int, int, const void*, const orgpu_prop_shell*)
{ orgpu_set_error("shell groups are not built yet"); return -5; }
static int shell_build_supergroups(std::vector<HostShellGroup>&, std::vector<ShellSGHost>&, const std::vector<int>&,
                                   const std::vector<int>&, const std::vector<int>&, const std::vector<double>&,
                                   const orgpu_control&, int&, int&, FinalizeArgs&) { return 0; }
static void launch_shell_forces(ShellSGHost&, const DevNodes&, double*, int, CycleState*, const DtBlocks&, const FinalizeArgs&, cudaStream_t) {}
static int shell_download_state(std::vector<ShellSGHost>&, int, int, double*) { orgpu_set_error("no shell groups"); return -5; }
