// refrender.cpp — drives the REFERENCE implementation (the unmodified SwiftShader ICD built from
// /root/reference into oracle/_ref/libvk_swiftshader.so) through plain Vulkan calls, for two purposes:
//   (1) golden generation / oracle pinning: render a scene file, dump colour/depth/stencil;
//   (2) the CPU baseline of bench.py (--time N): time vkQueueSubmit -> vkQueueWaitIdle of the draw-only
//       render pass (loadOp LOAD), after one warm-up frame that JIT-compiles the routines
//       (tests/VulkanBenchmarks/TriangleBenchmarks.cpp:29-30).
// TEST INFRASTRUCTURE ONLY — nothing in the product path links or calls this.
//
// The ICD is dlopen'ed and every entry point is resolved through the exported vkGetInstanceProcAddr
// (like tests/VulkanWrapper/VulkanTester.cpp:201-231), so no Vulkan loader is needed.
//
// usage: refrender <libvk_swiftshader.so> <scene.bin> <out.bin> [--time N] [--warmup W]
#define VK_NO_PROTOTYPES
#include <vulkan/vulkan.h>

#include "scene_format.h"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x)                                                              \
	do {                                                                      \
		VkResult _r = (x);                                                    \
		if(_r != VK_SUCCESS)                                                  \
		{                                                                     \
			fprintf(stderr, "%s -> %d (line %d)\n", #x, (int)_r, __LINE__);   \
			exit(2);                                                          \
		}                                                                     \
	} while(0)

#define FNS(X)                                                                                                                      \
	X(vkEnumeratePhysicalDevices) X(vkCreateDevice) X(vkGetDeviceQueue) X(vkCreateImage) X(vkGetImageMemoryRequirements)            \
	X(vkAllocateMemory) X(vkBindImageMemory) X(vkCreateImageView) X(vkCreateRenderPass) X(vkCreateFramebuffer) X(vkCreateBuffer)    \
	X(vkGetBufferMemoryRequirements) X(vkBindBufferMemory) X(vkMapMemory) X(vkCreateShaderModule) X(vkCreatePipelineLayout)         \
	X(vkCreateGraphicsPipelines) X(vkCreateCommandPool) X(vkAllocateCommandBuffers) X(vkBeginCommandBuffer) X(vkEndCommandBuffer)   \
	X(vkCmdBeginRenderPass) X(vkCmdEndRenderPass) X(vkCmdBindPipeline) X(vkCmdBindVertexBuffers) X(vkCmdBindIndexBuffer)            \
	X(vkCmdDraw) X(vkCmdDrawIndexed) X(vkCmdPushConstants) X(vkCmdCopyImageToBuffer) X(vkQueueSubmit) X(vkQueueWaitIdle) X(vkCmdPipelineBarrier)          \
	X(vkCreateSampler) X(vkCreateDescriptorSetLayout) X(vkCreateDescriptorPool) X(vkAllocateDescriptorSets)                         \
	X(vkUpdateDescriptorSets) X(vkCmdBindDescriptorSets) X(vkCmdCopyBufferToImage) X(vkGetPhysicalDeviceProperties)
#define DECL(n) static PFN_##n n;
FNS(DECL)
static PFN_vkCreateInstance vkCreateInstance;

static VkDevice dev;

static void mkBuffer(VkDeviceSize size, VkBufferUsageFlags usage, VkBuffer &buf, void *&ptr)
{
	VkBufferCreateInfo bi{ VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO };
	bi.size = size ? size : 4;
	bi.usage = usage;
	CHECK(vkCreateBuffer(dev, &bi, nullptr, &buf));
	VkMemoryRequirements mr;
	vkGetBufferMemoryRequirements(dev, buf, &mr);
	VkMemoryAllocateInfo ai{ VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO };
	ai.allocationSize = mr.size;
	ai.memoryTypeIndex = 0;
	VkDeviceMemory m;
	CHECK(vkAllocateMemory(dev, &ai, nullptr, &m));
	CHECK(vkBindBufferMemory(dev, buf, m, 0));
	CHECK(vkMapMemory(dev, m, 0, VK_WHOLE_SIZE, 0, &ptr));
}

static void mkImage(uint32_t w, uint32_t h, uint32_t levels, VkSampleCountFlagBits samples, VkFormat fmt, VkImageUsageFlags usage,
                    VkImageAspectFlags aspect, VkImage &img, VkImageView &view)
{
	VkImageCreateInfo ii{ VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO };
	ii.imageType = VK_IMAGE_TYPE_2D;
	ii.format = fmt;
	ii.extent = { w, h, 1 };
	ii.mipLevels = levels;
	ii.arrayLayers = 1;
	ii.samples = samples;
	ii.tiling = VK_IMAGE_TILING_OPTIMAL;
	ii.usage = usage;
	CHECK(vkCreateImage(dev, &ii, nullptr, &img));
	VkMemoryRequirements mr;
	vkGetImageMemoryRequirements(dev, img, &mr);
	VkMemoryAllocateInfo ai{ VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO };
	ai.allocationSize = mr.size;
	ai.memoryTypeIndex = 0;
	VkDeviceMemory m;
	CHECK(vkAllocateMemory(dev, &ai, nullptr, &m));
	CHECK(vkBindImageMemory(dev, img, m, 0));
	VkImageViewCreateInfo vi{ VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO };
	vi.image = img;
	vi.viewType = VK_IMAGE_VIEW_TYPE_2D;
	vi.format = fmt;
	vi.subresourceRange = { aspect, 0, levels, 0, 1 };
	CHECK(vkCreateImageView(dev, &vi, nullptr, &view));
}

struct DrawObjects
{
	VkPipeline pipeline[2]; // [0] for the clearing pass, [1] for the LOAD pass (render-pass compatible, but keep it explicit)
	VkPipelineLayout layout;
	VkDescriptorSet dset = VK_NULL_HANDLE;
	VkBuffer vb, ib = VK_NULL_HANDLE, instb = VK_NULL_HANDLE;
};

int main(int argc, char **argv)
{
	if(argc < 4)
	{
		fprintf(stderr, "usage: refrender <icd.so> <scene.bin> <out.bin> [--time N]\n");
		return 1;
	}
	int timing = 0, warmup = 1;
	for(int i = 4; i + 1 < argc; i++)
		if(!strcmp(argv[i], "--time")) timing = atoi(argv[i + 1]);
		else if(!strcmp(argv[i], "--warmup")) warmup = atoi(argv[i + 1]);

	// ---- scene ----
	FILE *fi = fopen(argv[2], "rb");
	if(!fi) { perror("scene"); return 1; }
	fseek(fi, 0, SEEK_END);
	size_t fsize = (size_t)ftell(fi);
	fseek(fi, 0, SEEK_SET);
	std::vector<uint8_t> file(fsize);
	if(fread(file.data(), 1, fsize, fi) != fsize) return 1;
	fclose(fi);
	const SceneHeader *hdr = (const SceneHeader *)file.data();
	if(hdr->magic != SCENE_MAGIC || hdr->version != SCENE_VERSION) { fprintf(stderr, "bad scene file\n"); return 1; }
	const SceneDraw *draws = (const SceneDraw *)(hdr + 1);
	const SceneBlob *blobs = (const SceneBlob *)(draws + hdr->numDraws);
	auto blobPtr = [&](uint32_t id) { return file.data() + blobs[id].offset; };
	auto blobSize = [&](uint32_t id) { return (size_t)blobs[id].size; };
	const uint32_t W = hdr->width, H = hdr->height;
	const VkSampleCountFlagBits S = (VkSampleCountFlagBits)hdr->samples;
	const bool ms = hdr->samples > 1;

	// ---- ICD ----
	void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if(!lib) { fprintf(stderr, "%s\n", dlerror()); return 1; }
	auto gipa = (PFN_vkGetInstanceProcAddr)dlsym(lib, "vkGetInstanceProcAddr");
	vkCreateInstance = (PFN_vkCreateInstance)gipa(nullptr, "vkCreateInstance");
	VkApplicationInfo app{ VK_STRUCTURE_TYPE_APPLICATION_INFO };
	app.apiVersion = VK_API_VERSION_1_1;
	VkInstanceCreateInfo ici{ VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO };
	ici.pApplicationInfo = &app;
	VkInstance inst;
	CHECK(vkCreateInstance(&ici, nullptr, &inst));
#define LOAD(n)                                                  \
	n = (PFN_##n)gipa(inst, #n);                                 \
	if(!n) { fprintf(stderr, "missing %s\n", #n); return 1; }
	FNS(LOAD)
	uint32_t npd = 1;
	VkPhysicalDevice pd;
	vkEnumeratePhysicalDevices(inst, &npd, &pd);
	float prio = 1;
	VkDeviceQueueCreateInfo qci{ VK_STRUCTURE_TYPE_DEVICE_QUEUE_CREATE_INFO };
	qci.queueCount = 1;
	qci.pQueuePriorities = &prio;
	VkDeviceCreateInfo dci{ VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO };
	dci.queueCreateInfoCount = 1;
	dci.pQueueCreateInfos = &qci;
	VkPhysicalDeviceFeatures feat{};
	feat.depthBounds = VK_TRUE;
	feat.depthClamp = VK_TRUE;
	dci.pEnabledFeatures = &feat;
	CHECK(vkCreateDevice(pd, &dci, nullptr, &dev));
	VkQueue queue;
	vkGetDeviceQueue(dev, 0, 0, &queue);

	// ---- attachments ----
	const VkFormat cfmt = (VkFormat)hdr->colorFormat;
	const bool hasDS = hdr->hasDepth || hdr->hasStencil;
	// hasDepth: 0 = none, 1 = D32_SFLOAT, 2 = D16_UNORM (without stencil)
	const bool d16 = hdr->hasDepth == 2 && !hdr->hasStencil;
	const VkFormat dsfmt = hdr->hasStencil ? VK_FORMAT_D32_SFLOAT_S8_UINT : (d16 ? VK_FORMAT_D16_UNORM : VK_FORMAT_D32_SFLOAT);
	const size_t depthBpp = d16 ? 2 : 4;
	const VkImageAspectFlags dsAspect = VK_IMAGE_ASPECT_DEPTH_BIT | (hdr->hasStencil ? VK_IMAGE_ASPECT_STENCIL_BIT : 0);
	VkImage cimg, rimg = VK_NULL_HANDLE, dimg = VK_NULL_HANDLE;
	VkImageView cview, rview = VK_NULL_HANDLE, dview = VK_NULL_HANDLE;
	mkImage(W, H, 1, S, cfmt, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, VK_IMAGE_ASPECT_COLOR_BIT, cimg, cview);
	if(ms) mkImage(W, H, 1, VK_SAMPLE_COUNT_1_BIT, cfmt, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, VK_IMAGE_ASPECT_COLOR_BIT, rimg, rview);
	if(hasDS) mkImage(W, H, 1, S, dsfmt, VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, dsAspect, dimg, dview);

	// ---- render passes: [0] clears, [1] loads (timed) ----
	VkRenderPass rp[2];
	VkFramebuffer fb[2];
	for(int pass = 0; pass < 2; pass++)
	{
		VkAttachmentDescription att[3]{};
		uint32_t n = 0;
		VkAttachmentLoadOp lop = pass == 0 ? VK_ATTACHMENT_LOAD_OP_CLEAR : VK_ATTACHMENT_LOAD_OP_LOAD;
		VkImageLayout cInit = pass == 0 ? VK_IMAGE_LAYOUT_UNDEFINED : VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL;
		att[n].format = cfmt; att[n].samples = S; att[n].loadOp = lop; att[n].storeOp = VK_ATTACHMENT_STORE_OP_STORE;
		att[n].stencilLoadOp = VK_ATTACHMENT_LOAD_OP_DONT_CARE; att[n].stencilStoreOp = VK_ATTACHMENT_STORE_OP_DONT_CARE;
		att[n].initialLayout = cInit; att[n].finalLayout = VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL;
		VkAttachmentReference cref{ n, VK_IMAGE_LAYOUT_COLOR_ATTACHMENT_OPTIMAL };
		n++;
		VkAttachmentReference dref{}, rref{};
		if(hasDS)
		{
			att[n].format = dsfmt; att[n].samples = S; att[n].loadOp = lop; att[n].storeOp = VK_ATTACHMENT_STORE_OP_STORE;
			att[n].stencilLoadOp = lop; att[n].stencilStoreOp = VK_ATTACHMENT_STORE_OP_STORE;
			att[n].initialLayout = cInit; att[n].finalLayout = VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL;
			dref = { n, VK_IMAGE_LAYOUT_DEPTH_STENCIL_ATTACHMENT_OPTIMAL };
			n++;
		}
		if(ms)
		{
			att[n].format = cfmt; att[n].samples = VK_SAMPLE_COUNT_1_BIT; att[n].loadOp = VK_ATTACHMENT_LOAD_OP_DONT_CARE; att[n].storeOp = VK_ATTACHMENT_STORE_OP_STORE;
			att[n].stencilLoadOp = VK_ATTACHMENT_LOAD_OP_DONT_CARE; att[n].stencilStoreOp = VK_ATTACHMENT_STORE_OP_DONT_CARE;
			att[n].initialLayout = VK_IMAGE_LAYOUT_UNDEFINED; att[n].finalLayout = VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL;
			rref = { n, VK_IMAGE_LAYOUT_COLOR_ATTACHMENT_OPTIMAL };
			n++;
		}
		VkSubpassDescription sp{};
		sp.pipelineBindPoint = VK_PIPELINE_BIND_POINT_GRAPHICS;
		sp.colorAttachmentCount = 1;
		sp.pColorAttachments = &cref;
		if(hasDS) sp.pDepthStencilAttachment = &dref;
		if(ms) sp.pResolveAttachments = &rref;
		VkRenderPassCreateInfo rpi{ VK_STRUCTURE_TYPE_RENDER_PASS_CREATE_INFO };
		rpi.attachmentCount = n;
		rpi.pAttachments = att;
		rpi.subpassCount = 1;
		rpi.pSubpasses = &sp;
		CHECK(vkCreateRenderPass(dev, &rpi, nullptr, &rp[pass]));
		VkImageView views[3];
		uint32_t nv = 0;
		views[nv++] = cview;
		if(hasDS) views[nv++] = dview;
		if(ms) views[nv++] = rview;
		VkFramebufferCreateInfo fbi{ VK_STRUCTURE_TYPE_FRAMEBUFFER_CREATE_INFO };
		fbi.renderPass = rp[pass];
		fbi.attachmentCount = nv;
		fbi.pAttachments = views;
		fbi.width = W;
		fbi.height = H;
		fbi.layers = 1;
		CHECK(vkCreateFramebuffer(dev, &fbi, nullptr, &fb[pass]));
	}

	VkCommandPoolCreateInfo cpi{ VK_STRUCTURE_TYPE_COMMAND_POOL_CREATE_INFO };
	VkCommandPool pool;
	CHECK(vkCreateCommandPool(dev, &cpi, nullptr, &pool));
	VkCommandBufferAllocateInfo cai{ VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO };
	cai.commandPool = pool;
	cai.level = VK_COMMAND_BUFFER_LEVEL_PRIMARY;
	cai.commandBufferCount = 3;
	VkCommandBuffer cmd[3]; // 0: clear+draw+readback, 1: draw only (timed), 2: uploads
	CHECK(vkAllocateCommandBuffers(dev, &cai, cmd));
	VkCommandBufferBeginInfo cbi{ VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO };

	// ---- per-draw objects ----
	std::vector<DrawObjects> objs(hdr->numDraws);
	CHECK(vkBeginCommandBuffer(cmd[2], &cbi));
	VkDescriptorPoolSize psz[2] = { { VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, hdr->numDraws + 1 }, { VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, hdr->numDraws + 1 } };
	VkDescriptorPoolCreateInfo dpi{ VK_STRUCTURE_TYPE_DESCRIPTOR_POOL_CREATE_INFO };
	dpi.maxSets = hdr->numDraws + 1;
	dpi.poolSizeCount = 2;
	dpi.pPoolSizes = psz;
	VkDescriptorPool dpool;
	CHECK(vkCreateDescriptorPool(dev, &dpi, nullptr, &dpool));

	for(uint32_t di = 0; di < hdr->numDraws; di++)
	{
		const SceneDraw &d = draws[di];
		DrawObjects &o = objs[di];
		void *p;
		mkBuffer(blobSize(d.vertexBlob), VK_BUFFER_USAGE_VERTEX_BUFFER_BIT, o.vb, p);
		memcpy(p, blobPtr(d.vertexBlob), blobSize(d.vertexBlob));
		if(d.indexType)
		{
			mkBuffer(blobSize(d.indexBlob), VK_BUFFER_USAGE_INDEX_BUFFER_BIT, o.ib, p);
			memcpy(p, blobPtr(d.indexBlob), blobSize(d.indexBlob));
		}
		if(d.numInstanceAttribs)
		{
			mkBuffer(blobSize(d.instanceBlob), VK_BUFFER_USAGE_VERTEX_BUFFER_BIT, o.instb, p);
			memcpy(p, blobPtr(d.instanceBlob), blobSize(d.instanceBlob));
		}
		VkDescriptorSetLayout dsl = VK_NULL_HANDLE;
		// the bindings of descriptor set 0: the fragment shader's combined image sampler and / or the vertex shader's uniform buffer
		std::vector<VkDescriptorSetLayoutBinding> lbs;
		VkDescriptorImageInfo dii{};
		VkDescriptorBufferInfo dbi{};
		if(d.hasUbo)
		{
			VkBuffer ub;
			mkBuffer(blobSize(d.uboBlob), VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT, ub, p);
			memcpy(p, blobPtr(d.uboBlob), blobSize(d.uboBlob));
			dbi = { ub, 0, VK_WHOLE_SIZE };
			lbs.push_back({ d.uboBinding, VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, 1, VK_SHADER_STAGE_VERTEX_BIT, nullptr });
		}
		if(d.hasTexture)
		{
			VkBuffer sb;
			mkBuffer(blobSize(d.texBlob), VK_BUFFER_USAGE_TRANSFER_SRC_BIT, sb, p);
			memcpy(p, blobPtr(d.texBlob), blobSize(d.texBlob));
			VkImage timg;
			VkImageView tview;
			mkImage(d.texWidth, d.texHeight, d.texLevels, VK_SAMPLE_COUNT_1_BIT, d.hasTexture == 2 ? VK_FORMAT_R8G8B8A8_SRGB : VK_FORMAT_R8G8B8A8_UNORM, // hasTexture: 1 = UNORM, 2 = SRGB
			        VK_IMAGE_USAGE_SAMPLED_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT, VK_IMAGE_ASPECT_COLOR_BIT, timg, tview);
			VkImageMemoryBarrier b{ VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER };
			b.oldLayout = VK_IMAGE_LAYOUT_UNDEFINED;
			b.newLayout = VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL;
			b.srcQueueFamilyIndex = b.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
			b.image = timg;
			b.subresourceRange = { VK_IMAGE_ASPECT_COLOR_BIT, 0, d.texLevels, 0, 1 };
			b.dstAccessMask = VK_ACCESS_TRANSFER_WRITE_BIT;
			vkCmdPipelineBarrier(cmd[2], VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, VK_PIPELINE_STAGE_TRANSFER_BIT, 0, 0, nullptr, 0, nullptr, 1, &b);
			size_t off = 0;
			for(uint32_t l = 0; l < d.texLevels; l++)
			{
				uint32_t lw = std::max(1u, d.texWidth >> l), lh = std::max(1u, d.texHeight >> l);
				VkBufferImageCopy r{};
				r.bufferOffset = off;
				r.imageSubresource = { VK_IMAGE_ASPECT_COLOR_BIT, l, 0, 1 };
				r.imageExtent = { lw, lh, 1 };
				vkCmdCopyBufferToImage(cmd[2], sb, timg, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &r);
				off += (size_t)lw * lh * 4;
			}
			b.oldLayout = VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL;
			b.newLayout = VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL;
			b.srcAccessMask = VK_ACCESS_TRANSFER_WRITE_BIT;
			b.dstAccessMask = VK_ACCESS_SHADER_READ_BIT;
			vkCmdPipelineBarrier(cmd[2], VK_PIPELINE_STAGE_TRANSFER_BIT, VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT, 0, 0, nullptr, 0, nullptr, 1, &b);
			VkSamplerCreateInfo sci{ VK_STRUCTURE_TYPE_SAMPLER_CREATE_INFO };
			sci.magFilter = (VkFilter)d.magFilter;
			sci.minFilter = (VkFilter)d.minFilter;
			sci.mipmapMode = (VkSamplerMipmapMode)d.mipmapMode;
			sci.addressModeU = (VkSamplerAddressMode)d.addressModeU;
			sci.addressModeV = (VkSamplerAddressMode)d.addressModeV;
			sci.addressModeW = VK_SAMPLER_ADDRESS_MODE_REPEAT;
			sci.mipLodBias = d.mipLodBias;
			sci.minLod = d.minLod;
			sci.maxLod = d.maxLod;
			VkSampler smp;
			CHECK(vkCreateSampler(dev, &sci, nullptr, &smp));
			lbs.push_back({ d.texBinding, VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, 1, VK_SHADER_STAGE_FRAGMENT_BIT, nullptr });
			dii = { smp, tview, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL };
		}
		if(!lbs.empty())
		{
			VkDescriptorSetLayoutCreateInfo li{ VK_STRUCTURE_TYPE_DESCRIPTOR_SET_LAYOUT_CREATE_INFO };
			li.bindingCount = (uint32_t)lbs.size();
			li.pBindings = lbs.data();
			CHECK(vkCreateDescriptorSetLayout(dev, &li, nullptr, &dsl));
			VkDescriptorSetAllocateInfo dai{ VK_STRUCTURE_TYPE_DESCRIPTOR_SET_ALLOCATE_INFO };
			dai.descriptorPool = dpool;
			dai.descriptorSetCount = 1;
			dai.pSetLayouts = &dsl;
			CHECK(vkAllocateDescriptorSets(dev, &dai, &o.dset));
			for(const VkDescriptorSetLayoutBinding &lb : lbs)
			{
				VkWriteDescriptorSet wd{ VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET };
				wd.dstSet = o.dset;
				wd.dstBinding = lb.binding;
				wd.descriptorCount = 1;
				wd.descriptorType = lb.descriptorType;
				if(lb.descriptorType == VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER) wd.pBufferInfo = &dbi;
				else wd.pImageInfo = &dii;
				vkUpdateDescriptorSets(dev, 1, &wd, 0, nullptr);
			}
		}
		VkPipelineLayoutCreateInfo pli{ VK_STRUCTURE_TYPE_PIPELINE_LAYOUT_CREATE_INFO };
		if(dsl) { pli.setLayoutCount = 1; pli.pSetLayouts = &dsl; }
		VkPushConstantRange pcr{ VK_SHADER_STAGE_VERTEX_BIT, 0, d.pushConstantBytes };
		if(d.pushConstantBytes) { pli.pushConstantRangeCount = 1; pli.pPushConstantRanges = &pcr; }
		CHECK(vkCreatePipelineLayout(dev, &pli, nullptr, &o.layout));

		VkShaderModule vsm, fsm;
		VkShaderModuleCreateInfo smi{ VK_STRUCTURE_TYPE_SHADER_MODULE_CREATE_INFO };
		smi.codeSize = blobSize(d.vsBlob);
		smi.pCode = (const uint32_t *)blobPtr(d.vsBlob);
		CHECK(vkCreateShaderModule(dev, &smi, nullptr, &vsm));
		smi.codeSize = blobSize(d.fsBlob);
		smi.pCode = (const uint32_t *)blobPtr(d.fsBlob);
		CHECK(vkCreateShaderModule(dev, &smi, nullptr, &fsm));
		VkPipelineShaderStageCreateInfo st[2]{};
		st[0].sType = st[1].sType = VK_STRUCTURE_TYPE_PIPELINE_SHADER_STAGE_CREATE_INFO;
		st[0].stage = VK_SHADER_STAGE_VERTEX_BIT; st[0].module = vsm; st[0].pName = "main";
		st[1].stage = VK_SHADER_STAGE_FRAGMENT_BIT; st[1].module = fsm; st[1].pName = "main";
		VkVertexInputBindingDescription vbd[2] = { { 0, d.stride, VK_VERTEX_INPUT_RATE_VERTEX }, { 1, d.instanceStride, VK_VERTEX_INPUT_RATE_INSTANCE } };
		VkVertexInputAttributeDescription vad[SCENE_MAX_ATTRIBS + 4];
		uint32_t nvad = 0;
		for(uint32_t a = 0; a < d.numAttribs; a++) vad[nvad++] = { d.attrib[a].location, 0, (VkFormat)d.attrib[a].format, d.attrib[a].offset };
		for(uint32_t a = 0; a < d.numInstanceAttribs && a < 4; a++) vad[nvad++] = { d.instanceAttrib[a].location, 1, (VkFormat)d.instanceAttrib[a].format, d.instanceAttrib[a].offset };
		VkPipelineVertexInputStateCreateInfo vis{ VK_STRUCTURE_TYPE_PIPELINE_VERTEX_INPUT_STATE_CREATE_INFO };
		vis.vertexBindingDescriptionCount = d.numInstanceAttribs ? 2 : 1;
		vis.pVertexBindingDescriptions = vbd;
		vis.vertexAttributeDescriptionCount = nvad;
		vis.pVertexAttributeDescriptions = vad;
		VkPipelineInputAssemblyStateCreateInfo ias{ VK_STRUCTURE_TYPE_PIPELINE_INPUT_ASSEMBLY_STATE_CREATE_INFO };
		ias.topology = (VkPrimitiveTopology)d.topology;
		VkViewport vp{ d.viewport[0], d.viewport[1], d.viewport[2], d.viewport[3], d.viewport[4], d.viewport[5] };
		VkRect2D sc{ { d.scissor[0], d.scissor[1] }, { (uint32_t)d.scissor[2], (uint32_t)d.scissor[3] } };
		VkPipelineViewportStateCreateInfo vps{ VK_STRUCTURE_TYPE_PIPELINE_VIEWPORT_STATE_CREATE_INFO };
		vps.viewportCount = 1; vps.pViewports = &vp; vps.scissorCount = 1; vps.pScissors = &sc;
		VkPipelineRasterizationStateCreateInfo rs{ VK_STRUCTURE_TYPE_PIPELINE_RASTERIZATION_STATE_CREATE_INFO };
		rs.polygonMode = VK_POLYGON_MODE_FILL;
		rs.cullMode = d.cullMode;
		rs.frontFace = (VkFrontFace)d.frontFace;
		rs.lineWidth = d.lineWidth != 0.0f ? d.lineWidth : 1.0f;
		rs.depthClampEnable = d.depthClampEnable ? VK_TRUE : VK_FALSE;
		rs.depthBiasEnable = d.depthBiasEnable;
		rs.depthBiasConstantFactor = d.depthBiasConstant;
		rs.depthBiasClamp = d.depthBiasClamp;
		rs.depthBiasSlopeFactor = d.depthBiasSlope;
		VkPipelineMultisampleStateCreateInfo mss{ VK_STRUCTURE_TYPE_PIPELINE_MULTISAMPLE_STATE_CREATE_INFO };
		mss.rasterizationSamples = S;
		VkSampleMask smask = d.sampleMask;
		mss.pSampleMask = &smask;
		mss.alphaToCoverageEnable = d.alphaToCoverageEnable;
		VkPipelineDepthStencilStateCreateInfo ds{ VK_STRUCTURE_TYPE_PIPELINE_DEPTH_STENCIL_STATE_CREATE_INFO };
		ds.depthTestEnable = d.depthTestEnable;
		ds.depthWriteEnable = d.depthWriteEnable;
		ds.depthCompareOp = (VkCompareOp)d.depthCompareOp;
		ds.stencilTestEnable = d.stencilTestEnable;
		ds.depthBoundsTestEnable = d.depthBoundsTestEnable;
		ds.minDepthBounds = d.minDepthBounds;
		ds.maxDepthBounds = d.maxDepthBounds;
		auto face = [](const SceneStencilFace &f) {
			VkStencilOpState s{};
			s.failOp = (VkStencilOp)f.failOp; s.passOp = (VkStencilOp)f.passOp; s.depthFailOp = (VkStencilOp)f.depthFailOp;
			s.compareOp = (VkCompareOp)f.compareOp; s.compareMask = f.compareMask; s.writeMask = f.writeMask; s.reference = f.reference;
			return s;
		};
		ds.front = face(d.front);
		ds.back = face(d.back);
		VkPipelineColorBlendAttachmentState ba{};
		ba.colorWriteMask = d.colorWriteMask;
		ba.blendEnable = d.blendEnable;
		ba.srcColorBlendFactor = (VkBlendFactor)d.srcColor; ba.dstColorBlendFactor = (VkBlendFactor)d.dstColor; ba.colorBlendOp = (VkBlendOp)d.colorOp;
		ba.srcAlphaBlendFactor = (VkBlendFactor)d.srcAlpha; ba.dstAlphaBlendFactor = (VkBlendFactor)d.dstAlpha; ba.alphaBlendOp = (VkBlendOp)d.alphaOp;
		VkPipelineColorBlendStateCreateInfo cb{ VK_STRUCTURE_TYPE_PIPELINE_COLOR_BLEND_STATE_CREATE_INFO };
		cb.attachmentCount = 1;
		cb.pAttachments = &ba;
		memcpy(cb.blendConstants, d.blendConstants, 16);
		VkGraphicsPipelineCreateInfo gp{ VK_STRUCTURE_TYPE_GRAPHICS_PIPELINE_CREATE_INFO };
		gp.stageCount = 2; gp.pStages = st; gp.pVertexInputState = &vis; gp.pInputAssemblyState = &ias; gp.pViewportState = &vps;
		gp.pRasterizationState = &rs; gp.pMultisampleState = &mss; gp.pDepthStencilState = &ds; gp.pColorBlendState = &cb; gp.layout = o.layout;
		for(int pass = 0; pass < 2; pass++)
		{
			gp.renderPass = rp[pass];
			CHECK(vkCreateGraphicsPipelines(dev, VK_NULL_HANDLE, 1, &gp, nullptr, &o.pipeline[pass]));
		}
	}
	CHECK(vkEndCommandBuffer(cmd[2]));
	VkSubmitInfo si{ VK_STRUCTURE_TYPE_SUBMIT_INFO };
	si.commandBufferCount = 1;
	si.pCommandBuffers = &cmd[2];
	CHECK(vkQueueSubmit(queue, 1, &si, VK_NULL_HANDLE));
	CHECK(vkQueueWaitIdle(queue));

	// ---- readback buffers ----
	VkBuffer rbC, rbD = VK_NULL_HANDLE, rbS = VK_NULL_HANDLE;
	void *pC, *pD = nullptr, *pS = nullptr;
	const size_t colorBpp = cfmt == VK_FORMAT_R32G32B32A32_SFLOAT ? 16 : (cfmt == VK_FORMAT_R16G16B16A16_SFLOAT ? 8 : 4);
	mkBuffer((size_t)W * H * colorBpp, VK_BUFFER_USAGE_TRANSFER_DST_BIT, rbC, pC);
	const bool readDS = hasDS && !ms;
	if(readDS && hdr->hasDepth) mkBuffer((size_t)W * H * depthBpp, VK_BUFFER_USAGE_TRANSFER_DST_BIT, rbD, pD);
	if(readDS && hdr->hasStencil) mkBuffer((size_t)W * H, VK_BUFFER_USAGE_TRANSFER_DST_BIT, rbS, pS);

	auto record = [&](VkCommandBuffer c, int pass, bool copy) {
		CHECK(vkBeginCommandBuffer(c, &cbi));
		VkClearValue cv[3]{};
		memcpy(cv[0].color.float32, hdr->clearColor, 16);
		cv[1].depthStencil = { hdr->clearDepth, hdr->clearStencil };
		VkRenderPassBeginInfo rbi{ VK_STRUCTURE_TYPE_RENDER_PASS_BEGIN_INFO };
		rbi.renderPass = rp[pass];
		rbi.framebuffer = fb[pass];
		rbi.renderArea = { { 0, 0 }, { W, H } };
		rbi.clearValueCount = 3;
		rbi.pClearValues = cv;
		vkCmdBeginRenderPass(c, &rbi, VK_SUBPASS_CONTENTS_INLINE);
		for(uint32_t di = 0; di < hdr->numDraws; di++)
		{
			const SceneDraw &d = draws[di];
			DrawObjects &o = objs[di];
			vkCmdBindPipeline(c, VK_PIPELINE_BIND_POINT_GRAPHICS, o.pipeline[pass]);
			if(o.dset) vkCmdBindDescriptorSets(c, VK_PIPELINE_BIND_POINT_GRAPHICS, o.layout, d.hasTexture ? d.texSet : d.uboSet, 1, &o.dset, 0, nullptr);
			if(d.pushConstantBytes) vkCmdPushConstants(c, o.layout, VK_SHADER_STAGE_VERTEX_BIT, 0, d.pushConstantBytes, d.pushConstants);
			VkDeviceSize off = 0;
			vkCmdBindVertexBuffers(c, 0, 1, &o.vb, &off);
			if(d.numInstanceAttribs) vkCmdBindVertexBuffers(c, 1, 1, &o.instb, &off);
			const uint32_t instances = d.instanceCount ? d.instanceCount : 1;
			if(d.indexType)
			{
				vkCmdBindIndexBuffer(c, o.ib, 0, d.indexType == 2 ? VK_INDEX_TYPE_UINT16 : VK_INDEX_TYPE_UINT32);
				vkCmdDrawIndexed(c, d.count, instances, d.firstIndex, d.vertexOffset, 0);
			}
			else vkCmdDraw(c, d.count, instances, d.firstIndex, 0);
		}
		vkCmdEndRenderPass(c);
		if(copy)
		{
			VkBufferImageCopy r{};
			r.imageSubresource = { VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1 };
			r.imageExtent = { W, H, 1 };
			vkCmdCopyImageToBuffer(c, ms ? rimg : cimg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, rbC, 1, &r);
			if(rbD) { r.imageSubresource.aspectMask = VK_IMAGE_ASPECT_DEPTH_BIT; vkCmdCopyImageToBuffer(c, dimg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, rbD, 1, &r); }
			if(rbS) { r.imageSubresource.aspectMask = VK_IMAGE_ASPECT_STENCIL_BIT; vkCmdCopyImageToBuffer(c, dimg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, rbS, 1, &r); }
		}
		CHECK(vkEndCommandBuffer(c));
	};

	record(cmd[0], 0, true);
	si.pCommandBuffers = &cmd[0];
	CHECK(vkQueueSubmit(queue, 1, &si, VK_NULL_HANDLE));
	CHECK(vkQueueWaitIdle(queue));

	FILE *fo = fopen(argv[3], "wb");
	if(!fo) { perror("out"); return 1; }
	uint32_t oh[6] = { 0x4F525753u /* SWRO */, W, H, (uint32_t)(rbD != VK_NULL_HANDLE ? (d16 ? 2 : 1) : 0), (uint32_t)(rbS != VK_NULL_HANDLE), hdr->samples };
	fwrite(oh, 4, 6, fo);
	fwrite(pC, 1, (size_t)W * H * colorBpp, fo);
	if(rbD) fwrite(pD, 1, (size_t)W * H * depthBpp, fo);
	if(rbS) fwrite(pS, 1, (size_t)W * H, fo);
	fclose(fo);

	if(timing > 0)
	{
		record(cmd[1], 1, false);
		si.pCommandBuffers = &cmd[1];
		for(int i = 0; i < (warmup < 1 ? 1 : warmup); i++) // warm-up: JIT of the LOAD-pass routines
		{
			CHECK(vkQueueSubmit(queue, 1, &si, VK_NULL_HANDLE));
			CHECK(vkQueueWaitIdle(queue));
		}
		std::vector<double> ts;
		for(int i = 0; i < timing; i++)
		{
			auto t0 = std::chrono::steady_clock::now();
			CHECK(vkQueueSubmit(queue, 1, &si, VK_NULL_HANDLE));
			CHECK(vkQueueWaitIdle(queue));
			auto t1 = std::chrono::steady_clock::now();
			ts.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
		}
		std::vector<double> sorted = ts;
		std::sort(sorted.begin(), sorted.end());
		double sum = 0;
		for(double t : ts) sum += t;
		VkPhysicalDeviceProperties props;
		vkGetPhysicalDeviceProperties(pd, &props);
		printf("{\"frames\": %d, \"median_ms\": %.6f, \"min_ms\": %.6f, \"mean_ms\": %.6f, \"device\": \"%s\"}\n", timing,
		       sorted[sorted.size() / 2], sorted[0], sum / timing, props.deviceName);
	}
	return 0;
}
