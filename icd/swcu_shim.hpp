// swcu_shim.hpp — the piece of host C++ that is compiled INTO the reference ICD (oracle/build_cuda_icd.sh) to put the CUDA draw
// path behind sw::Renderer::draw.  Everything above stays the reference's own code: vkCmdDraw recording, CmdDrawBase::draw
// (src/Vulkan/VkCommandBuffer.cpp:958-1013), the state gathering of Renderer::draw (src/Device/Renderer.cpp:183-487).  The patch
// (icd/swiftshader_cuda.patch) replaces three things:
//   * DrawCall::run at the end of Renderer::draw (Renderer.cpp:489)            -> swcu_shim::draw      -> swcu_draw
//   * Renderer::synchronize (Renderer.cpp:664-671)                             -> swcu_shim::synchronize -> downloads + swcu_sync
//   * DeviceMemory::allocateBuffer / freeBuffer (src/Vulkan/VkDeviceMemory.cpp:340-356) -> swcu_mem_register / _unregister
// libswcuda.so is loaded with dlopen at the first use (SWCU_LIB, or next to the ICD): the ICD has no link-time CUDA dependency.
// SWCU_ICD=0 leaves the reference's own routines in charge (one binary, A/B by environment, never a silent fallback: a draw the
// CUDA path rejects aborts with its error text).
#pragma once

#include <cstddef>
#include <cstdint>

#include "Vulkan/VulkanPlatform.hpp"  // the reference's own way to pull in vulkan_core.h (its handle types are redefined there)

namespace vk {
class Device;
class GraphicsPipeline;
class GraphicsState;
struct Inputs;
}  // namespace vk

namespace sw {
struct DrawData;
}

namespace swcu_shim {

bool enabled();

// vk::DeviceMemory (host malloc): a device shadow per allocation
void onAllocate(void *base, size_t bytes);
void onFree(void *base);

struct DrawArgs
{
	const vk::Device *device;
	const vk::GraphicsPipeline *pipeline;
	const vk::GraphicsState *state;
	const vk::Inputs *inputs;
	const sw::DrawData *data;  // what Renderer::draw has gathered: viewport constants are re-derived by the library, the pointers are used
	unsigned int count;
	int baseVertex;
	const void *indexBuffer;
	VkIndexType indexType;
	VkRect2D renderArea;
	int layer;
};
void draw(const DrawArgs &args);

// Renderer::synchronize: everything drawn so far is in host memory again when this returns
void synchronize();

}  // namespace swcu_shim
