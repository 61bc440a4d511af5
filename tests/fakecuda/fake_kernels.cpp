// fake_kernels.cpp -- the FUNCTIONAL kernel launchers of the fake CUDA build (tests/test_engine_cpu_twin.py): whole images go through the
// real engine host code (csrc/engine.cu) on the CPU.
//   * conv_first / upsample2x / tta_gather / alpha_bicubic4x: the real kernel SOURCE (csrc/misc_kernels.cu) under the host emulation
//     of tests/emu/cuda_host_emu.h;
//   * conv3x3: fake_conv.cpp (a CPU model of the kernel's contract).
#include "../emu/cuda_host_emu.h"
#define RSR_HOST_EMULATION 1
#include "../../realsr-ncnn-vulkan_b200/csrc/misc_kernels.cu"

#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace rsr {

cudaError_t launch_conv_first(const FirstArgs& a, cudaStream_t)
{
    dim3 grid, block;
    conv_first_launch_dims(a, grid, block);
    cuda_emu::launch(grid, block, [=] { conv_first_kernel(a); });
    return cudaSuccess;
}
cudaError_t launch_upsample2x(const __half* in, int in_cs, __half* out, int B, int H, int W, cudaStream_t)
{
    cuda_emu::launch(dim3(2), dim3(256), [=] { upsample2x_kernel(reinterpret_cast<const uint4*>(in), in_cs / 8, reinterpret_cast<uint4*>(out), B, H, W); });
    return cudaSuccess;
}
cudaError_t launch_tta_gather(const TtaGatherArgs& a, int max_crop_w, int max_crop_h, cudaStream_t)
{
    cuda_emu::launch(dim3((max_crop_w + 31) / 32, (max_crop_h + 7) / 8, a.ngroups), dim3(32, 8), [=] { tta_gather_kernel(a); });
    return cudaSuccess;
}
cudaError_t launch_alpha_bicubic4x(const AlphaArgs& a, cudaStream_t)
{
    cuda_emu::launch(dim3((a.img_w * 4 + 31) / 32, (a.img_h * 4 + 7) / 8), dim3(32, 8), [=] { alpha_bicubic4x_kernel(a); });
    return cudaSuccess;
}

}  // namespace rsr
