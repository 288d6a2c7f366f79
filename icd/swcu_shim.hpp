// swcu_shim.hpp — the piece of host C++ that is compiled INTO the reference ICD (oracle/build_cuda_icd.sh) to put the CUDA draw
// path behind sw::Renderer::draw.  Everything above stays the reference's own code: vkCmdDraw recording, CmdDrawBase::draw
// (src/Vulkan/VkCommandBuffer.cpp:958-1013), the state gathering of Renderer::draw (src/Device/Renderer.cpp:183-487).  The patch
// (icd/swiftshader_cuda.patch) replaces three things:
//   * DrawCall::run at the end of Renderer::draw (Renderer.cpp:489)            -> swcu_shim::draw      -> swcu_draw
//   * Renderer::synchronize (Renderer.cpp:664-671)                             -> swcu_shim::synchronize -> downloads + swcu_sync
//   * DeviceMemory::allocateBuffer / freeBuffer (src/Vulkan/VkDeviceMemory.cpp:340-356) -> swcu_mem_register / _unregister
//   * DeviceMemory::getOffsetPointer / map (VkDeviceMemory.cpp:293-309): whoever asks for a host pointer into an allocation gets
//     host memory that is current — attachments the device has drawn to come down first (and only then)
//   * ImageView::resolve (src/Vulkan/VkImageView.cpp:284-312), the end-of-pass multisample resolve    -> swcu_resolve on the shadows
// libswcuda.so is loaded with dlopen at the first use (SWCU_LIB, or next to the ICD): the ICD has no link-time CUDA dependency.
// SWCU_ICD=0 leaves the reference's own routines in charge (one binary, A/B by environment, never a silent fallback: a draw the
// CUDA path rejects aborts with its error text).
#pragma once

#include <cstddef>
#include <cstdint>

#include "Vulkan/VulkanPlatform.hpp"  // the reference's own way to pull in vulkan_core.h (its handle types are redefined there)

namespace vk {
class Device;
class ImageView;
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

// Renderer::synchronize: every draw issued so far has finished; attachments in host-MAPPED allocations are current in host memory
// (an application reads those through its own pointer).  Other attachments stay resident on the device until somebody asks for a
// pointer into their allocation (hostAccess).
void synchronize();

// DeviceMemory::getOffsetPointer: the caller is about to read or write [base, base + bytes) on the host
void hostAccess(const void *base, size_t bytes);
void onMap(const void *base);
// while one is alive on this thread, hostAccess does nothing: the pointers Renderer::draw gathers are handed to the device, not read
struct DrawScope
{
	DrawScope();
	~DrawScope();
};

// ImageView::resolve: true if the 4x -> 1x resolve has been done on the device shadows (Blitter::fastResolve's arithmetic)
bool resolve(vk::ImageView *src, vk::ImageView *dst);

}  // namespace swcu_shim
