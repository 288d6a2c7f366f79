// swcu_shim.cpp — see swcu_shim.hpp.  Compiled with the reference's own include paths (it reads the reference's state objects
// through their public accessors) plus this repository's include/.
#include "swcu_shim.hpp"

#include "swcu.h"

#include "Device/Context.hpp"
#include "Device/Renderer.hpp"
#include "Pipeline/SpirvShader.hpp"
#include "Vulkan/VkDescriptorSetLayout.hpp"
#include "Vulkan/VkDevice.hpp"
#include "Vulkan/VkImageView.hpp"
#include "Vulkan/VkPipeline.hpp"
#include "Vulkan/VkPipelineLayout.hpp"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct Api
{
	int (*create)(swcu_ctx **, int);
	const char *(*last_error)(swcu_ctx *);
	int (*mem_register)(swcu_ctx *, const void *, size_t);
	int (*mem_unregister)(swcu_ctx *, const void *);
	int (*mem_upload)(swcu_ctx *, const void *, size_t);
	int (*mem_download)(swcu_ctx *, void *, size_t);
	int (*draw)(swcu_ctx *, const swcu_draw_desc *);
	int (*sync)(swcu_ctx *);
	int (*resolve)(swcu_ctx *, const swcu_attachment *, uint32_t, const swcu_attachment *);
	int (*shader_translate)(const uint32_t *, uint32_t, swcu_shader_info *, char *, size_t);
};

struct Range
{
	uintptr_t lo, hi;
};

struct State
{
	std::mutex mutex;
	bool tried = false, on = false;
	Api api{};
	swcu_ctx *ctx = nullptr;
	std::map<uintptr_t, size_t> allocations;  // registered host ranges (base -> bytes)
	std::map<uintptr_t, bool> mapped;         // allocations the application holds a pointer into (vkMapMemory)
	std::vector<Range> deviceNewer;           // attachment ranges the device has drawn to since their last download
};

State &S()
{
	static State *s = new State();  // never destroyed: the queue thread may outlive static destruction
	return *s;
}

[[noreturn]] void die(const char *what, const char *detail)
{
	fprintf(stderr, "swiftshader-cuda: %s: %s\n", what, detail ? detail : "");
	abort();  // the CUDA draw path has no CPU fallback (north_star); SWCU_ICD=0 selects the reference's own routines explicitly
}

std::string libraryPath()
{
	if(const char *e = getenv("SWCU_LIB")) return e;
	Dl_info info;
	if(dladdr((void *)&libraryPath, &info) && info.dli_fname)
	{
		std::string p = info.dli_fname;  // .../oracle/_cuda/libvk_swiftshader_cuda.so
		const size_t cut = p.rfind('/');
		if(cut != std::string::npos) return p.substr(0, cut) + "/../../swiftshader_b200/csrc/libswcuda.so";
	}
	return "libswcuda.so";
}

// with the lock held
bool start()
{
	State &s = S();
	if(s.tried) return s.on;
	s.tried = true;
	const char *e = getenv("SWCU_ICD");
	if(e && !strcmp(e, "0")) return false;
	const std::string path = libraryPath();
	void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
	if(!h) die("cannot load the CUDA draw path", dlerror());
#define SYM(field, name) \
	s.api.field = (decltype(s.api.field))dlsym(h, name); \
	if(!s.api.field) die("missing symbol", name)
	SYM(create, "swcu_create");
	SYM(last_error, "swcu_last_error");
	SYM(mem_register, "swcu_mem_register");
	SYM(mem_unregister, "swcu_mem_unregister");
	SYM(mem_upload, "swcu_mem_upload");
	SYM(mem_download, "swcu_mem_download");
	SYM(draw, "swcu_draw");
	SYM(sync, "swcu_sync");
	SYM(resolve, "swcu_resolve");
	SYM(shader_translate, "swcu_shader_translate");
#undef SYM
	const char *ord = getenv("SWCU_DEVICE");
	if(s.api.create(&s.ctx, ord ? atoi(ord) : 0) != SWCU_OK) die("swcu_create", s.api.last_error(nullptr));
	s.on = true;
	return true;
}

void check(int rc, const char *what)
{
	if(rc != SWCU_OK) die(what, S().api.last_error(S().ctx));
}

// the registered allocation that holds p (with the lock held)
bool findAllocation(uintptr_t p, uintptr_t &base, size_t &bytes)
{
	State &s = S();
	auto it = s.allocations.upper_bound(p);
	if(it == s.allocations.begin()) return false;
	--it;
	if(p >= it->first + it->second) return false;
	base = it->first;
	bytes = it->second;
	return true;
}

bool deviceIsNewer(uintptr_t lo, uintptr_t hi);

// host -> device for [p, p + bytes), clipped to the allocation that holds p
void uploadRange(const void *p, size_t bytes)
{
	if(!p || !bytes) return;
	if(deviceIsNewer((uintptr_t)p, (uintptr_t)p + bytes)) return;  // (a render target read as an input: the device copy IS the current one)
	uintptr_t base;
	size_t size;
	if(!findAllocation((uintptr_t)p, base, size)) die("a draw reads host memory that is not a vk::DeviceMemory allocation", "");
	const size_t room = base + size - (uintptr_t)p;
	check(S().api.mem_upload(S().ctx, p, bytes < room ? bytes : room), "swcu_mem_upload");
}

bool deviceIsNewer(uintptr_t lo, uintptr_t hi)
{
	for(const Range &r : S().deviceNewer)
		if(r.lo <= lo && hi <= r.hi) return true;
	return false;
}

// An attachment about to be drawn to: if the host copy is the current one (a load-op clear by Blitter, an upload, a previous
// synchronize), it goes up first; from here on the device copy is the current one until the next synchronize.
void acquireAttachment(const void *p, size_t bytes)
{
	if(!p || !bytes) return;
	const uintptr_t lo = (uintptr_t)p, hi = lo + bytes;
	if(deviceIsNewer(lo, hi)) return;
	uploadRange(p, bytes);
	S().deviceNewer.push_back({ lo, hi });
}

thread_local int tl_drawScope = 0;

// device -> host for what the device has drawn inside [lo, hi); with the lock held
void makeHostCurrent(uintptr_t lo, uintptr_t hi)
{
	State &s = S();
	bool any = false;
	for(size_t i = 0; i < s.deviceNewer.size();)
	{
		const Range r = s.deviceNewer[i];
		if(r.lo < hi && r.hi > lo)
		{
			check(s.api.mem_download(s.ctx, (void *)r.lo, r.hi - r.lo), "swcu_mem_download");
			s.deviceNewer.erase(s.deviceNewer.begin() + i);
			any = true;
		}
		else i++;
	}
	if(any) check(s.api.sync(s.ctx), "swcu_sync");
}

swcu_attachment attachment(vk::ImageView *view, VkImageAspectFlagBits aspect, int layer, int samples)
{
	swcu_attachment a = {};
	if(!view) return a;
	a.buffer = view->getOffsetPointer({ 0, 0, 0 }, aspect, 0, layer);
	a.format = (uint32_t)(VkFormat)view->getFormat(aspect);
	a.pitchB = view->rowPitchBytes(aspect, 0);
	a.sliceB = view->slicePitchBytes(aspect, 0);
	const VkExtent2D e = view->getMipLevelExtent(0);
	a.width = e.width;
	a.height = e.height;
	acquireAttachment(a.buffer, (size_t)a.sliceB * (size_t)samples);
	return a;
}

}  // namespace

namespace swcu_shim {

bool enabled()
{
	std::lock_guard<std::mutex> lock(S().mutex);
	return start();
}

void onAllocate(void *base, size_t bytes)
{
	std::lock_guard<std::mutex> lock(S().mutex);
	if(!start() || !base || !bytes) return;
	check(S().api.mem_register(S().ctx, base, bytes), "swcu_mem_register");
	S().allocations[(uintptr_t)base] = bytes;
}

void onFree(void *base)
{
	std::lock_guard<std::mutex> lock(S().mutex);
	State &s = S();
	if(!s.on || !base) return;
	auto it = s.allocations.find((uintptr_t)base);
	if(it == s.allocations.end()) return;
	const uintptr_t lo = it->first, hi = lo + it->second;
	for(size_t i = 0; i < s.deviceNewer.size();)
	{
		if(s.deviceNewer[i].lo >= lo && s.deviceNewer[i].hi <= hi) s.deviceNewer.erase(s.deviceNewer.begin() + i);
		else i++;
	}
	check(s.api.mem_unregister(s.ctx, base), "swcu_mem_unregister");
	s.allocations.erase(it);
	s.mapped.erase((uintptr_t)base);
}

DrawScope::DrawScope() { tl_drawScope++; }
DrawScope::~DrawScope() { tl_drawScope--; }

void hostAccess(const void *base, size_t bytes)
{
	if(tl_drawScope > 0 || !base) return;
	std::lock_guard<std::mutex> lock(S().mutex);
	if(!S().on || S().deviceNewer.empty()) return;
	makeHostCurrent((uintptr_t)base, (uintptr_t)base + bytes);
}

void onMap(const void *base)
{
	std::lock_guard<std::mutex> lock(S().mutex);
	if(S().on && base) S().mapped[(uintptr_t)base] = true;
}

bool resolve(vk::ImageView *src, vk::ImageView *dst)
{
	DrawScope scope;
	std::lock_guard<std::mutex> lock(S().mutex);
	State &s = S();
	if(!s.on || !src || !dst) return false;
	const VkFormat fs = (VkFormat)src->getFormat(VK_IMAGE_ASPECT_COLOR_BIT), fd = (VkFormat)dst->getFormat(VK_IMAGE_ASPECT_COLOR_BIT);
	// Blitter::fastResolve's formats, and the ones whose resolve is the generic blit (swcu_resolve restates both)
	if(fs != fd || (fs != VK_FORMAT_R8G8B8A8_UNORM && fs != VK_FORMAT_B8G8R8A8_UNORM && fs != VK_FORMAT_R8G8B8A8_SRGB && fs != VK_FORMAT_B8G8R8A8_SRGB &&
	                fs != VK_FORMAT_R16G16B16A16_SFLOAT && fs != VK_FORMAT_R32G32B32A32_SFLOAT)) return false;
	if(src->getSampleCount() != 4 || dst->getSampleCount() != 1) return false;
	if(src->getSubresourceRange().layerCount != 1 || dst->getSubresourceRange().layerCount != 1) return false;
	const VkExtent2D es = src->getMipLevelExtent(0), ed = dst->getMipLevelExtent(0);
	if(es.width != ed.width || es.height != ed.height) return false;
	swcu_attachment a = {}, b = {};
	a.buffer = src->getOffsetPointer({ 0, 0, 0 }, VK_IMAGE_ASPECT_COLOR_BIT, 0, 0);
	a.format = (uint32_t)fs; a.pitchB = src->rowPitchBytes(VK_IMAGE_ASPECT_COLOR_BIT, 0); a.sliceB = src->slicePitchBytes(VK_IMAGE_ASPECT_COLOR_BIT, 0);
	a.width = es.width; a.height = es.height;
	b.buffer = dst->getOffsetPointer({ 0, 0, 0 }, VK_IMAGE_ASPECT_COLOR_BIT, 0, 0);
	b.format = (uint32_t)fd; b.pitchB = dst->rowPitchBytes(VK_IMAGE_ASPECT_COLOR_BIT, 0); b.sliceB = dst->slicePitchBytes(VK_IMAGE_ASPECT_COLOR_BIT, 0);
	b.width = ed.width; b.height = ed.height;
	const size_t srcBytes = (size_t)a.sliceB * 4, dstBytes = (size_t)b.pitchB * b.height;
	if(!deviceIsNewer((uintptr_t)a.buffer, (uintptr_t)a.buffer + srcBytes)) uploadRange(a.buffer, srcBytes);  // (the source was last written on the host)
	check(s.api.resolve(s.ctx, &a, 4, &b), "swcu_resolve");
	// every pixel of the target is written: the device copy is the current one from here on
	const uintptr_t lo = (uintptr_t)b.buffer, hi = lo + dstBytes;
	if(!deviceIsNewer(lo, hi)) s.deviceNewer.push_back({ lo, hi });
	return true;
}

void draw(const DrawArgs &args)
{
	DrawScope scope;
	std::lock_guard<std::mutex> lock(S().mutex);
	State &s = S();
	const vk::GraphicsState &state = *args.state;
	const vk::VertexInputInterfaceState &vii = state.getVertexInputInterfaceState();
	const vk::PreRasterizationState &pre = state.getPreRasterizationState();
	if(pre.hasRasterizerDiscard()) return;
	const vk::FragmentState &frag = state.getFragmentState();
	const vk::FragmentOutputInterfaceState &out = state.getFragmentOutputInterfaceState();
	const vk::Attachments attachments = args.pipeline->getAttachments();
	const sw::SpirvShader *vs = args.pipeline->getShader(VK_SHADER_STAGE_VERTEX_BIT).get();
	const sw::SpirvShader *fs = args.pipeline->getShader(VK_SHADER_STAGE_FRAGMENT_BIT).get();
	if(!vs || !fs) die("swcu_draw", "a draw without a vertex or a fragment shader is outside the subset");

	swcu_draw_desc d = {};
	d.structSize = sizeof(d);
	d.topology = (uint32_t)vii.getTopology();
	d.provokingVertexMode = pre.getProvokingVertexMode() == VK_PROVOKING_VERTEX_MODE_LAST_VERTEX_EXT ? 1u : 0u;
	d.indexType = args.indexBuffer ? (args.indexType == VK_INDEX_TYPE_UINT16 ? 2u : (args.indexType == VK_INDEX_TYPE_UINT32 ? 4u : 1u)) : 0u;
	d.indexBuffer = args.indexBuffer;
	d.primitiveCount = args.count;
	d.baseVertex = args.baseVertex;
	if(args.indexBuffer)
	{
		const size_t c = args.count;
		size_t n;
		switch(vii.getTopology())  // indices the draw fetches (setBatchIndices, Renderer.cpp:50-145)
		{
		case VK_PRIMITIVE_TOPOLOGY_POINT_LIST: n = c; break;
		case VK_PRIMITIVE_TOPOLOGY_LINE_LIST: n = c * 2; break;
		case VK_PRIMITIVE_TOPOLOGY_LINE_STRIP: n = c + 1; break;
		case VK_PRIMITIVE_TOPOLOGY_TRIANGLE_LIST: n = c * 3; break;
		default: n = c + 2; break;
		}
		uploadRange(args.indexBuffer, n * d.indexType);
	}
	for(int i = 0; i < SWCU_MAX_INPUTS; i++)  // DrawData::input / stride / robustnessSize (Renderer.cpp:282-288)
	{
		const sw::Stream &st = args.inputs->getStream(i);
		if(!st.buffer || st.format == VK_FORMAT_UNDEFINED) continue;
		d.input[i].buffer = st.buffer;
		d.input[i].robustnessSize = st.robustnessSize;
		d.input[i].vertexStride = (uint32_t)args.inputs->getVertexStride(i);
		d.input[i].format = (uint32_t)st.format;
		uploadRange(st.buffer, st.robustnessSize);
	}
	// push constants as Renderer::draw copied them for this draw (Renderer.cpp:484-486), the line width of the pipeline (:276)
	d.pushConstants = &args.data->pushConstants;
	d.pushConstantBytes = (uint32_t)sizeof(args.data->pushConstants);
	d.lineWidth = args.data->lineWidth;
	d.vertexShader = vs->insns.data();    // SpirvShader.hpp:164 — what the pipeline holds after spirv-opt (VkPipeline.cpp:42-107)
	d.vertexShaderWords = (uint32_t)vs->insns.size();
	d.fragmentShader = fs->insns.data();
	d.fragmentShaderWords = (uint32_t)fs->insns.size();
	const VkViewport &vp = pre.getViewport();
	d.viewportX = vp.x; d.viewportY = vp.y; d.viewportWidth = vp.width; d.viewportHeight = vp.height;
	d.viewportMinDepth = vp.minDepth; d.viewportMaxDepth = vp.maxDepth;
	const VkRect2D &sc = pre.getScissor();
	d.scissor = { sc.offset.x, sc.offset.y, sc.extent.width, sc.extent.height };
	d.renderArea = { args.renderArea.offset.x, args.renderArea.offset.y, args.renderArea.extent.width, args.renderArea.extent.height };
	d.cullMode = (uint32_t)pre.getCullMode();
	d.frontFace = (uint32_t)pre.getFrontFace();
	d.depthClipEnable = pre.getDepthClipEnable() ? 1u : 0u;
	d.depthClampEnable = pre.getDepthClampEnable() ? 1u : 0u;
	d.depthBiasConstant = pre.getConstantDepthBias();
	d.depthBiasSlope = pre.getSlopeDepthBias();
	d.depthBiasClamp = pre.getDepthBiasClamp();
	const int ms = out.getSampleCount();
	d.sampleCount = (uint32_t)ms;
	d.sampleMask = out.getMultiSampleMask();
	d.alphaToCoverageEnable = out.hasAlphaToCoverage() ? 1u : 0u;
	d.depthTestEnable = frag.depthTestActive(attachments) ? 1u : 0u;
	d.depthWriteEnable = frag.depthWriteActive(attachments) ? 1u : 0u;
	d.depthCompareOp = (uint32_t)frag.getDepthCompareMode();
	d.stencilTestEnable = frag.stencilActive(attachments) ? 1u : 0u;
	auto face = [](const VkStencilOpState &f) {
		return swcu_stencil_face{ (uint32_t)f.failOp, (uint32_t)f.passOp, (uint32_t)f.depthFailOp, (uint32_t)f.compareOp, f.compareMask, f.writeMask, f.reference };
	};
	d.front = face(frag.getFrontStencil());
	d.back = face(frag.getBackStencil());
	d.depthBoundsTestEnable = frag.depthBoundsTestActive(attachments) ? 1u : 0u;
	d.minDepthBounds = frag.getMinDepthBounds();
	d.maxDepthBounds = frag.getMaxDepthBounds();
	const vk::BlendState b = out.getBlendState(0, attachments, true);  // already folded (Context.cpp:1090-1147); the library folds idempotently
	d.blendEnable = b.alphaBlendEnable ? 1u : 0u;
	d.srcColorBlendFactor = (uint32_t)b.sourceBlendFactor; d.dstColorBlendFactor = (uint32_t)b.destBlendFactor; d.colorBlendOp = (uint32_t)b.blendOperation;
	d.srcAlphaBlendFactor = (uint32_t)b.sourceBlendFactorAlpha; d.dstAlphaBlendFactor = (uint32_t)b.destBlendFactorAlpha; d.alphaBlendOp = (uint32_t)b.blendOperationAlpha;
	d.colorWriteMask = (uint32_t)out.colorWriteActive(0, attachments);
	memcpy(d.blendConstants, &out.getBlendConstants(), 16);
	d.color = attachment(attachments.colorBuffer[0], VK_IMAGE_ASPECT_COLOR_BIT, args.layer, ms);  // Renderer.cpp:442-474
	d.depth = attachment(attachments.depthBuffer, VK_IMAGE_ASPECT_DEPTH_BIT, args.layer, ms);
	d.stencil = attachment(attachments.stencilBuffer, VK_IMAGE_ASPECT_STENCIL_BIT, args.layer, ms);
	for(int i = 1; i < sw::MAX_COLOR_BUFFERS; i++)
		if(attachments.colorBuffer[i]) die("swcu_draw", "more than one colour attachment is outside the subset");

	// the combined image sampler the fragment shader reads: the sw::Texture + vk::SamplerState its descriptor holds
	// (Sampler.hpp:24-50, VkDescriptorSetLayout.cpp:299-336, SpirvShaderSampling.cpp:49-128)
	swcu_shader_info info;
	char err[256] = "";
	if(s.api.shader_translate(fs->insns.data(), (uint32_t)fs->insns.size(), &info, err, sizeof(err)) != SWCU_OK) die("fragment shader outside the subset", err);
	if(info.usesTexture)
	{
		const vk::PipelineLayout *layout = frag.getPipelineLayout();
		if(info.textureSet >= layout->getDescriptorSetCount() || info.textureBinding >= layout->getBindingCount(info.textureSet) ||
		   layout->getDescriptorType(info.textureSet, info.textureBinding) != VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER)
			die("swcu_draw", "the fragment shader's sampler is not a combined image sampler of the pipeline layout");
		const uint8_t *set = args.data->descriptorSets[info.textureSet];
		const vk::SampledImageDescriptor *sd = (const vk::SampledImageDescriptor *)(set + layout->getBindingOffset(info.textureSet, info.textureBinding));
		const vk::SamplerState *sampler = args.device->findSampler(sd->samplerId);
		if(!sampler || !sd->memoryOwner) die("swcu_draw", "unbound combined image sampler");
		swcu_sampled_image &si = d.sampledImage[d.sampledImageCount++];
		si.set = info.textureSet;
		si.binding = info.textureBinding;
		si.format = (uint32_t)(VkFormat)sd->memoryOwner->getFormat(VK_IMAGE_ASPECT_COLOR_BIT);
		si.levelCount = (uint32_t)sd->mipLevels;
		for(int l = 0; l < SWCU_MIPMAP_LEVELS; l++)
		{
			const sw::Mipmap &m = sd->texture.mipmap[l];
			si.level[l].buffer = m.buffer;
			si.level[l].width = m.width[0];
			si.level[l].height = m.height[0];
			si.level[l].pitchP = m.pitchP[0];
			if(l < sd->mipLevels) uploadRange(m.buffer, (size_t)m.pitchP[0] * (m.height[0] - 1) * 4 + (size_t)m.width[0] * 4);
		}
		si.magFilter = (uint32_t)sampler->magFilter; si.minFilter = (uint32_t)sampler->minFilter; si.mipmapMode = (uint32_t)sampler->mipmapMode;
		si.addressModeU = (uint32_t)sampler->addressModeU; si.addressModeV = (uint32_t)sampler->addressModeV;
		si.mipLodBias = sampler->mipLodBias; si.minLod = sampler->minLod; si.maxLod = sampler->maxLod;
		si.anisotropyEnable = sampler->anisotropyEnable; si.compareEnable = sampler->compareEnable; si.unnormalizedCoordinates = sampler->unnormalizedCoordinates;
	}
	// the uniform buffers the vertex shader reads: the vk::BufferDescriptor { ptr, sizeInBytes } at the binding's offset in the bound
	// set (VkDescriptorSetLayout.hpp:75-82; a dynamic one moved on by its dynamic offset, SpirvShaderMemory.cpp / EmitDescriptor*).
	// Host memory the application wrote: the library reads the words the vertex program uses during swcu_draw.
	swcu_shader_info vinfo;
	if(s.api.shader_translate(vs->insns.data(), (uint32_t)vs->insns.size(), &vinfo, err, sizeof(err)) != SWCU_OK) die("vertex shader outside the subset", err);
	for(uint32_t u = 0; u < vinfo.uniformCount && u < SWCU_MAX_UNIFORM_BUFFERS; u++)
	{
		const vk::PipelineLayout *layout = pre.getPipelineLayout();
		const uint32_t set = vinfo.uniformSet[u], binding = vinfo.uniformBinding[u];
		if(set >= layout->getDescriptorSetCount() || binding >= layout->getBindingCount(set)) die("swcu_draw", "the vertex shader's uniform block is not in the pipeline layout");
		const VkDescriptorType type = layout->getDescriptorType(set, binding);
		if(type != VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER && type != VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC) die("swcu_draw", "the vertex shader's uniform block is not bound as a uniform buffer");
		const uint8_t *setData = args.data->descriptorSets[set];
		if(!setData) die("swcu_draw", "unbound descriptor set");
		const vk::BufferDescriptor *bd = (const vk::BufferDescriptor *)(setData + layout->getBindingOffset(set, binding));
		uint32_t dynamic = 0;
		if(type == VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC) dynamic = args.data->descriptorDynamicOffsets[layout->getDynamicOffsetIndex(set, binding)];
		swcu_uniform_buffer &ub = d.uniformBuffer[d.uniformBufferCount++];
		ub.set = set;
		ub.binding = binding;
		ub.data = bd->ptr ? (const uint8_t *)bd->ptr + dynamic : nullptr;
		ub.bytes = (uint32_t)(bd->sizeInBytes > 0 ? bd->sizeInBytes : 0);
	}
	check(s.api.draw(s.ctx, &d), "swcu_draw");
}

void synchronize()
{
	std::lock_guard<std::mutex> lock(S().mutex);
	State &s = S();
	if(!s.on) return;
	// an application reads a mapped allocation through its own pointer once the fence has signalled: those come down here; every
	// other attachment stays on the device until somebody asks for a host pointer into its allocation (hostAccess)
	for(size_t i = 0; i < s.deviceNewer.size();)
	{
		const Range r = s.deviceNewer[i];
		uintptr_t base;
		size_t bytes;
		if(findAllocation(r.lo, base, bytes) && s.mapped.count(base))
		{
			check(s.api.mem_download(s.ctx, (void *)r.lo, r.hi - r.lo), "swcu_mem_download");
			s.deviceNewer.erase(s.deviceNewer.begin() + i);
		}
		else i++;
	}
	check(s.api.sync(s.ctx), "swcu_sync");
}

}  // namespace swcu_shim
