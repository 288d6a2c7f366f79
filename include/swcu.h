/*
 * swcu.h — C-ABI of the B200 draw path that stands in for SwiftShader's
 * DrawCall::run (src/Device/Renderer.cpp:551-598) and everything below it.
 *
 * Boundary: the reference hands each draw from
 * sw::Renderer::draw (src/Device/Renderer.hpp:211-213, Renderer.cpp:183-490) to
 * three JIT'd routines per batch/cluster.  That is too fine for a GPU, so the
 * cut is one level up: ONE call per draw (swcu_draw), plus the equivalents of
 * Renderer::synchronize (Renderer.cpp:664-671) and of the memory the draw
 * reads/writes (vk::DeviceMemory is host malloc memory in the reference; here
 * each registered host range gets a device-resident shadow in HBM).
 *
 * Every pointer inside swcu_draw_desc is a HOST address, exactly the address
 * the reference stores in sw::DrawData (Renderer.hpp:58-113), sw::Stream
 * (src/Device/Stream.hpp:22-32) and sw::Mipmap::buffer
 * (src/Device/Sampler.hpp:25-39).  The library translates it to the device
 * shadow of the registered range that contains it.
 *
 * Plain C, POD only, no C++/torch types.  Enum-valued fields carry the Vulkan
 * enum value verbatim (VkFormat, VkCompareOp, VkBlendFactor, ...), so the host
 * shim can copy pipeline state through without translation.
 *
 * Unsupported state is a hard error (SWCU_E_UNSUPPORTED); there is no CPU
 * fallback (north_star).
 */
#ifndef SWCU_H
#define SWCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWCU_OK 0
#define SWCU_E_UNSUPPORTED (-1) /* state outside the implemented subset (reference: UNSUPPORTED(), Debug.hpp:139) */
#define SWCU_E_INVALID (-2)     /* malformed argument / unregistered pointer */
#define SWCU_E_CUDA (-3)        /* CUDA runtime error, see swcu_last_error */
#define SWCU_E_NOMEM (-4)

#define SWCU_MAX_INPUTS 16         /* vertex input locations (reference: MAX_INTERFACE_COMPONENTS/4 = 32) */
#define SWCU_MAX_SAMPLED_IMAGES 4  /* combined image samplers visible to the fragment shader */
#define SWCU_MAX_UNIFORM_BUFFERS 4 /* uniform-buffer descriptors the vertex stage reads */
#define SWCU_MIPMAP_LEVELS 15      /* sw::MIPMAP_LEVELS, src/Device/Config.hpp */
#define SWCU_MAX_VARYING_COMPONENTS 16
#define SWCU_MAX_GROUP 8           /* GPUs of one box that share a frame (swcu_group_*) */

typedef struct swcu_ctx swcu_ctx;

/* sw::Stream (src/Device/Stream.hpp:22-32) + DrawData::input/stride/robustnessSize (Renderer.hpp:63-65). */
typedef struct swcu_vertex_input
{
	const void *buffer;      /* host address of the first vertex's attribute (binding base + binding offset + attribute offset) */
	uint32_t robustnessSize; /* bytes available from `buffer` (robustBufferAccess clamp, VertexRoutine.cpp:192-207); 0 = unchecked */
	uint32_t vertexStride;
	uint32_t format;         /* VkFormat; VK_FORMAT_UNDEFINED (0) = unused location */
	uint32_t reserved;
} swcu_vertex_input;

/* One level of sw::Mipmap (src/Device/Sampler.hpp:25-39). */
typedef struct swcu_mip_level
{
	const void *buffer; /* host address of texel (0,0) */
	uint32_t width;
	uint32_t height;
	uint32_t pitchP;    /* row pitch in texels */
	uint32_t reserved;
} swcu_mip_level;

/* sw::Texture + vk::SamplerState as written by vkUpdateDescriptorSets
 * (src/Vulkan/VkDescriptorSetLayout.cpp:299-336,466-503; src/Vulkan/VkSampler.hpp:29-63). */
typedef struct swcu_sampled_image
{
	uint32_t set;
	uint32_t binding;
	uint32_t format;     /* VkFormat of the image view */
	uint32_t levelCount; /* levels in the view; level[] beyond it replicate the last (VkDescriptorSetLayout.cpp:470) */
	swcu_mip_level level[SWCU_MIPMAP_LEVELS];
	uint32_t magFilter;  /* VkFilter */
	uint32_t minFilter;
	uint32_t mipmapMode; /* VkSamplerMipmapMode */
	uint32_t addressModeU; /* VkSamplerAddressMode */
	uint32_t addressModeV;
	float mipLodBias;
	float minLod;
	float maxLod;
	uint32_t anisotropyEnable;
	uint32_t compareEnable;
	uint32_t unnormalizedCoordinates;
	uint32_t reserved;
} swcu_sampled_image;

/* VkStencilOpState as consumed by PixelProcessor::Stencil::set (PixelProcessor.hpp:122-147). */
typedef struct swcu_stencil_face
{
	uint32_t failOp; /* VkStencilOp */
	uint32_t passOp;
	uint32_t depthFailOp;
	uint32_t compareOp; /* VkCompareOp */
	uint32_t compareMask;
	uint32_t writeMask;
	uint32_t reference;
} swcu_stencil_face;

/* DrawData::{color,depth,stencil}Buffer/PitchB/SliceB (Renderer.hpp:90-98); layout per SURVEY §8a-R15:
 * linear rows, MSAA sample q is a whole slice at +q*sliceB. */
typedef struct swcu_attachment
{
	void *buffer;    /* host address of pixel (0,0) of sample 0; NULL = no attachment */
	uint32_t format; /* VkFormat */
	int32_t pitchB;
	int32_t sliceB;
	uint32_t width;  /* extent of the attachment in pixels (bounds for tile staging) */
	uint32_t height;
	uint32_t reserved;
} swcu_attachment;

typedef struct swcu_rect
{
	int32_t x, y;
	uint32_t width, height;
} swcu_rect;

/* Flattened arguments + gathered state of sw::Renderer::draw (Renderer.cpp:183-490). */
typedef struct swcu_uniform_buffer
{
	uint32_t set, binding;
	const void *data; /* host pointer: BufferDescriptor::ptr (offset of the descriptor and, for a dynamic one, the dynamic offset applied) */
	uint32_t bytes;   /* BufferDescriptor::sizeInBytes; words past it read 0 */
	uint32_t reserved0;
} swcu_uniform_buffer;

typedef struct swcu_draw_desc
{
	uint32_t structSize; /* sizeof(swcu_draw_desc), ABI check */

	/* --- input assembly: Renderer::draw(count, baseVertex, indexBuffer), DrawCall::processVertices (Renderer.cpp:600-628) */
	uint32_t topology;            /* VkPrimitiveTopology: POINT_LIST 0, LINE_LIST 1, LINE_STRIP 2, TRIANGLE_LIST 3, TRIANGLE_STRIP 4, TRIANGLE_FAN 5 */
	uint32_t provokingVertexMode; /* VkProvokingVertexModeEXT (0 = FIRST, reference default Context.hpp:329) */
	uint32_t indexType;           /* 0 = non-indexed, 2 = uint16, 4 = uint32 (bytes per index) */
	const void *indexBuffer;      /* host address of the first index of this draw (NULL when non-indexed) */
	uint32_t primitiveCount;      /* number of primitives: triangles, lines or points (Renderer::draw `count`) */
	int32_t baseVertex;           /* added to every index (firstVertex for non-indexed draws) */
	swcu_vertex_input input[SWCU_MAX_INPUTS];

	/* --- shaders: SpirvShader::insns of the two stages (narrow translator) */
	const uint32_t *vertexShader;
	uint32_t vertexShaderWords;
	uint32_t fragmentShaderWords;
	const uint32_t *fragmentShader;

	/* --- pre-rasterization state (Renderer.cpp:300-345, SetupProcessor.cpp:58-104) */
	float viewportX, viewportY, viewportWidth, viewportHeight, viewportMinDepth, viewportMaxDepth;
	swcu_rect scissor;
	swcu_rect renderArea;
	uint32_t cullMode;  /* VkCullModeFlags */
	uint32_t frontFace; /* VkFrontFace */
	uint32_t depthClipEnable; /* reference default true */
	uint32_t depthClampEnable; /* VkPipelineRasterizationStateCreateInfo::depthClampEnable: fragment depth is clamped to the viewport's depth range
	                              instead of [0, 1] (PixelProcessor.cpp:121-136); the API ties depthClipEnable = !depthClampEnable (Context.cpp:647-648) */
	float depthBiasConstant, depthBiasSlope, depthBiasClamp;

	/* --- multisampling */
	uint32_t sampleCount; /* 1 or 4 */
	uint32_t sampleMask;
	uint32_t alphaToCoverageEnable; /* VkPipelineMultisampleStateCreateInfo (Context.cpp:526); thresholds of Renderer.cpp:391-410 */

	/* --- depth / stencil (PixelProcessor.cpp:74-140) */
	uint32_t depthTestEnable;
	uint32_t depthWriteEnable;
	uint32_t depthCompareOp; /* VkCompareOp */
	uint32_t stencilTestEnable;
	swcu_stencil_face front;
	swcu_stencil_face back;
	uint32_t depthBoundsTestEnable; /* Context.cpp:960-962; active only with a depth attachment (Context.cpp:946-949) */
	float minDepthBounds, maxDepthBounds;

	/* --- colour output: blend state BEFORE folding; the library folds it like
	 *     FragmentOutputInterfaceState::getBlendState (src/Device/Context.cpp:1090-1270). */
	uint32_t blendEnable;
	uint32_t srcColorBlendFactor, dstColorBlendFactor, colorBlendOp; /* VkBlendFactor / VkBlendOp */
	uint32_t srcAlphaBlendFactor, dstAlphaBlendFactor, alphaBlendOp;
	uint32_t colorWriteMask; /* VkColorComponentFlags */
	float blendConstants[4];

	/* --- attachments */
	swcu_attachment color;
	swcu_attachment depth;
	swcu_attachment stencil;

	/* --- push constants of the draw (DrawData::pushConstants, Renderer.cpp:484-486): copied by swcu_draw */
	const void *pushConstants;
	uint32_t pushConstantBytes; /* <= 4 * SWCU_MAX_PUSH_WORDS */
	/* --- lines (DrawData::lineWidth, Renderer.cpp:276; rectangle lines of DrawCall::setupLine, Renderer.cpp:920-1000 — the reference's
	 *     default line rasterization mode; 0 is read as 1.0).  Points take their size from the vertex shader (gl_PointSize). */
	float lineWidth;

	/* --- descriptors visible to the fragment shader */
	uint32_t sampledImageCount;
	/* --- uniform buffers the vertex stage reads (VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER: the vk::BufferDescriptor { ptr, sizeInBytes } at
	 *     the binding's offset in DrawData::descriptorSets[set], VkDescriptorSetLayout.cpp:574-600; what the shader's OpLoad through
	 *     an access chain into the Block reads, SpirvShaderMemory.cpp).  HOST memory: swcu_draw reads the words the vertex program
	 *     uses and folds them into the program like the push constants; nothing of it goes to the device. */
	uint32_t uniformBufferCount;
	swcu_sampled_image sampledImage[SWCU_MAX_SAMPLED_IMAGES];
	swcu_uniform_buffer uniformBuffer[SWCU_MAX_UNIFORM_BUFFERS];
} swcu_draw_desc;

/* What the narrow SPIR-V translator extracted from one module (for tests and caching). */
#define SWCU_SRC_INPUT 0 /* value = location*4 + component of a stage input */
#define SWCU_SRC_CONST 1 /* value = IEEE-754 bits of a float constant */
#define SWCU_SRC_TEXEL 2 /* fragment only: value = component of the OpImageSampleImplicitLod result */
#define SWCU_SRC_PUSH 3  /* vertex only: value = 32-bit word of the push-constant block (sw::DrawData::pushConstants, Renderer.hpp:110) */
#define SWCU_SRC_TEMP 4  /* vertex only: value = index of the program step that computes it */
#define SWCU_SRC_UNIFORM 5 /* vertex only: value = slot << 16 | 32-bit word of the uniform block `slot` (uniformSet / uniformBinding of the shader info) */
/* The arithmetic a vertex shader does on its way from the inputs to gl_Position / the varyings (an MVP from push constants), lowered
 * to straight-line steps with the reference's rounding: MUL / ADD / SUB / NEG are single IEEE operations (SpirvShaderArithmetic.cpp),
 * FMA is Reactor's MulAdd — what OpMatrixTimesVector accumulates with (fused on every host with FMA, LLVMReactor.cpp:3082-3092). */
#define SWCU_OP_MUL 0
#define SWCU_OP_ADD 1
#define SWCU_OP_SUB 2
#define SWCU_OP_FMA 3 /* a * b + c, one rounding */
#define SWCU_OP_NEG 4
#define SWCU_MAX_PROGRAM 48
#define SWCU_MAX_PUSH_WORDS 32 /* 128 bytes: vk::MAX_PUSH_CONSTANT_SIZE (src/Vulkan/VkConfig.hpp) */
typedef struct swcu_shader_operand
{
	uint32_t kind;
	uint32_t value;
} swcu_shader_operand;

typedef struct swcu_shader_op
{
	uint32_t op; /* SWCU_OP_* */
	swcu_shader_operand a, b, c;
} swcu_shader_op;

typedef struct swcu_shader_info
{
	uint32_t stage;                /* 0 = vertex, 4 = fragment (SpvExecutionModel) */
	uint32_t outputMask;           /* vertex: bit i set if varying component i is written; fragment: components of location 0 written */
	swcu_shader_operand position[4];                           /* vertex: gl_Position.xyzw */
	swcu_shader_operand output[SWCU_MAX_VARYING_COMPONENTS];   /* vertex: varyings (location*4+component); fragment: output[0..3] = colour location 0 */
	uint32_t inputMask;            /* stage inputs read (bit = location*4+component) */
	uint32_t flatMask;             /* fragment: inputs decorated Flat */
	uint32_t noPerspectiveMask;    /* fragment: inputs decorated NoPerspective */
	uint32_t usesTexture;          /* fragment: 1 if a combined image sampler is sampled */
	uint32_t textureSet, textureBinding;
	swcu_shader_operand texCoord[2]; /* fragment: u, v operands of the sample */
	uint32_t programLength;          /* vertex: steps of the arithmetic program; step i defines SWCU_SRC_TEMP i */
	swcu_shader_op program[SWCU_MAX_PROGRAM];
	uint32_t writesPointSize;        /* vertex: gl_PointSize is stored (VertexRoutine.cpp:641-650); read by point draws only */
	swcu_shader_operand pointSize;
	uint32_t uniformCount;           /* vertex: uniform blocks (Uniform storage class, Block-decorated) the program reads */
	uint32_t uniformSet[SWCU_MAX_UNIFORM_BUFFERS], uniformBinding[SWCU_MAX_UNIFORM_BUFFERS];
} swcu_shader_info;

/* Counters for bench / tests. */
typedef struct swcu_stats
{
	uint64_t draws;
	uint64_t kernelLaunches;   /* kernels of this library launched since create/reset */
	uint64_t primitives;
	uint64_t h2dBytes;
	uint64_t d2hBytes;
	uint64_t pairs;            /* (region, triangle) pairs binned by the draws seen by swcu_sync since create/reset */
} swcu_stats;

/* ---- lifetime ---- */
int swcu_create(swcu_ctx **out, int device_ordinal);
void swcu_destroy(swcu_ctx *ctx);
const char *swcu_last_error(swcu_ctx *ctx); /* ctx may be NULL: last error of a failed swcu_create / translate */

/* ---- memory: device shadows of host ranges (stands in for vk::DeviceMemory, src/Vulkan/VkDeviceMemory.cpp) ---- */
int swcu_mem_register(swcu_ctx *ctx, const void *host_base, size_t bytes);
/* Adopt an existing DEVICE allocation (e.g. an NCCL buffer owned by the caller): pointers into it are used as they are,
 * nothing is allocated, copied or freed by the library.  Unregister with swcu_mem_unregister(ctx, device_base). */
int swcu_mem_register_device(swcu_ctx *ctx, void *device_base, size_t bytes);
int swcu_mem_unregister(swcu_ctx *ctx, const void *host_base);
/* Copies are asynchronous and run on the library's two copy streams (one per DMA direction), ordered against the kernels by
 * events: an upload is visible to every call issued after it; it waits for kernels issued earlier that touch the same shadow
 * and for a download of it still in flight.  A download sees every call issued before it; calls issued later that overwrite
 * the shadow wait for it.  So frame n+1's inputs go up and frame n-1's pixels come down while frame n renders.  The host
 * may read a downloaded range after swcu_sync or after waiting for a fence signalled after the download.
 * (option "copy_streams" = 0 puts the copies back on the context stream.) */
int swcu_mem_upload(swcu_ctx *ctx, const void *host_ptr, size_t bytes);   /* host -> shadow */
int swcu_mem_download(swcu_ctx *ctx, void *host_ptr, size_t bytes);       /* shadow -> host */
void *swcu_mem_device_ptr(swcu_ctx *ctx, const void *host_ptr);           /* device address of the shadow (for NCCL plumbing); NULL if unregistered */

/* ---- the hot path ---- */
int swcu_draw(swcu_ctx *ctx, const swcu_draw_desc *desc); /* replaces DrawCall::run; asynchronous */
int swcu_sync(swcu_ctx *ctx);                             /* replaces Renderer::synchronize (Renderer.cpp:664-671) */
/* Completion of a submission without draining the device: the CountedEvent argument of sw::Renderer::draw (Renderer.cpp:184,
 * counted up at :499-501 and done at :507-510) that vk::Fence waits on.  signal: "everything issued so far, copies included";
 * wait: block the host on the last signal of that slot (returns at once if the slot was never signalled).  With two slots a
 * render loop keeps two frames in flight. */
#define SWCU_MAX_FENCES 8
int swcu_fence_signal(swcu_ctx *ctx, uint32_t slot);
int swcu_fence_wait(swcu_ctx *ctx, uint32_t slot);

/* ---- the steps either side of the draw, on the resident shadows ---- */
/* Blitter::fastClear (src/Device/Blitter.cpp:170-325) for RGBA8 / D32F / S8 (and the 2 / 8 / 16-byte fills of D16, RGBA16F, RGBA32F clears): fills `samples` slices. */
int swcu_clear(swcu_ctx *ctx, const swcu_attachment *att, uint32_t samples, const swcu_rect *area, const void *value /* 4 bytes (1 for S8) */);
/* Blitter::fastResolve (Blitter.cpp:2079-2205): 4x RGBA8 -> 1x, avg(avg(s0,s1),avg(s2,s3)) with (a+b+1)>>1. */
int swcu_resolve(swcu_ctx *ctx, const swcu_attachment *src, uint32_t samples, const swcu_attachment *dst);

/* ---- multi-GPU: one process per GPU; finished bands reach the presenting GPU by stores over NVLink ----
 * The presenting rank exports the CUDA IPC handle of its frame's shadow (and of a small flag array), the other ranks open
 * it, adopt the mapping with swcu_mem_register_device and use addresses inside it as the destination of swcu_resolve /
 * swcu_copy_image; swcu_signal / swcu_wait_flags order frames between the GPUs (counters that only grow). */
int swcu_ipc_export(swcu_ctx *ctx, const void *ptr, void *handle64 /* 64 bytes out */, uint64_t *offset /* of ptr inside the exported allocation */);
int swcu_ipc_open(swcu_ctx *ctx, const void *handle64, void **device_base);
int swcu_ipc_close(swcu_ctx *ctx, void *device_base);
/* rows of an RGBA8 image -> another image of the same extent (either side may be adopted peer memory) */
int swcu_copy_image(swcu_ctx *ctx, const swcu_attachment *src, const swcu_attachment *dst);
/* *flag = value once everything enqueued before on this context is visible system-wide (flag: registered / adopted memory) */
int swcu_signal(swcu_ctx *ctx, void *flag, uint32_t value);
/* the stream waits until flags[first .. first+count) >= value (wrap-around compare), count <= 64 */
int swcu_wait_flags(swcu_ctx *ctx, const void *flags, uint32_t first, uint32_t count, uint32_t value);
/* The hand-over stream: between swcu_side_begin and swcu_side_end the calls that pass a finished frame on — swcu_wait_flags,
 * swcu_copy_image, swcu_signal, swcu_mem_download — are issued on a second stream that has waited for everything issued before
 * swcu_side_begin; the main stream goes on with the next frame's draws meanwhile (the reference's counterpart: presenting a frame is
 * not part of the render loop either — vkQueuePresentKHR waits on the frame's semaphore, the next submission does not wait for it).
 * swcu_side_end(slot) marks the end of that work; swcu_side_wait(slot) makes the MAIN stream wait for the mark of `slot` — to be
 * called before something the hand-over read is overwritten (a band buffer of a two-deep ring: slot = frame & 1). */
#define SWCU_SIDE_SLOTS 4
int swcu_side_begin(swcu_ctx *ctx);
int swcu_side_end(swcu_ctx *ctx, uint32_t slot);
int swcu_side_wait(swcu_ctx *ctx, uint32_t slot);

/* ---- groups: the GPUs of one box render ONE frame — the setup of every draw is sharded by triangle range, the pixels by screen band ----
 * Without a group, a rank that renders a band (renderArea = band) still fetches and projects every triangle of the draw.  In a group
 * rank r sets up triangles [r n / world, (r + 1) n / world) for the WHOLE frame and stores each record, its bin counts and its big-list
 * entry straight into the memory of the rank(s) whose band the triangle touches (CUDA IPC mappings of the peers' work buffers, NVLink
 * stores and atomics); one flag barrier later every rank bins and rasterises what has arrived for its band.  All O(triangles) work
 * is divided by the number of GPUs, and the only collective step of a draw is that barrier.
 * Every rank: swcu_group_reserve (sizes its work buffers once, returns an IPC handle), exchange the handles (torch.distributed /
 * MPI / a pipe: not this library's business), swcu_group_attach with all of them, in rank order.  From then on every BINNED swcu_draw
 * must be issued by all ranks, in the same order, with renderArea = the rank's band of fbHeight / world rows and otherwise equal
 * state; swcu_group_detach (after swcu_sync) ends it. */
typedef struct swcu_group_desc
{
	uint32_t structSize;
	uint32_t rank, world;        /* world <= SWCU_MAX_GROUP */
	uint32_t maxPrimitives;      /* largest primitiveCount a group draw will have */
	uint32_t maxSlots;           /* interpolated scalars the fragment shaders consume at most (0..6); sizes the triangle records */
	uint32_t maxSamples;         /* 1 or 4 */
	uint32_t fbWidth, fbHeight;  /* framebuffer extent of the group draws (the bins are laid out for it) */
} swcu_group_desc;
int swcu_group_reserve(swcu_ctx *ctx, const swcu_group_desc *desc, void *handle64 /* 64 bytes out */);
int swcu_group_attach(swcu_ctx *ctx, const void *handles /* world * 64 bytes, rank order */);
int swcu_group_detach(swcu_ctx *ctx);

/* The caller's stream (swcu_set_stream) is about to read or write the shadow of host_ptr itself — a collective over
 * swcu_mem_device_ptr, a kernel of its own: the stream waits for the last upload into the shadow (and for a download of it still in
 * flight), and uploads issued later wait for what the stream has been given up to their call. */
int swcu_mem_acquire(swcu_ctx *ctx, const void *host_ptr);
/* ... and once that work has been enqueued on the stream: whatever reads the shadow from now on (the setup phase of a draw runs on
 * a stream of its own) waits for it, as it would for an upload. */
int swcu_mem_release(swcu_ctx *ctx, const void *host_ptr);
/* The same pair for work the caller runs on a stream of ITS OWN (a collective that assembles the shadow from the ranks' slices while
 * the library's stream renders the previous frame): `cuda_stream` waits for the last upload into the shadow; after release whatever
 * reads the shadow waits for what that stream had been given.  The library's own stream is not held up. */
int swcu_mem_acquire_on(swcu_ctx *ctx, const void *host_ptr, void *cuda_stream);
int swcu_mem_release_on(swcu_ctx *ctx, const void *host_ptr, void *cuda_stream);

/* ---- narrow SPIR-V translator (host-only, no GPU needed) ---- */
int swcu_shader_translate(const uint32_t *code, uint32_t words, swcu_shader_info *out, char *err, size_t errlen);

/* ---- plumbing for measurement and multi-GPU ---- */
int swcu_set_stream(swcu_ctx *ctx, void *cuda_stream); /* use an external cudaStream_t (e.g. torch's current stream); NULL = own stream */
int swcu_timer_begin(swcu_ctx *ctx);                   /* cudaEventRecord on the context stream */
int swcu_timer_end(swcu_ctx *ctx, float *elapsed_ms);  /* records, synchronises, returns elapsed ms */
int swcu_get_stats(swcu_ctx *ctx, swcu_stats *out);
int swcu_reset_stats(swcu_ctx *ctx);
/* Per-kernel device time of the LAST swcu_draw when profiling is on (events around every launch).
 * names/ms arrays of capacity n; returns number of kernels written. */
int swcu_set_profiling(swcu_ctx *ctx, int enable); /* 1: per-kernel times of the last draw (the draw runs on one stream); 2: timeline (below) */
/* Diagnostics: begin / end (ms since the first one) of every kernel of the library issued since swcu_set_profiling(ctx, 2), measured by
 * events on the streams the kernels really run on (the draws stay pipelined); returns the count, at most n, and starts a new timeline. */
int swcu_timeline(swcu_ctx *ctx, const char **names, float *begin_ms, float *end_ms, int n);
int swcu_last_draw_kernels(swcu_ctx *ctx, const char **names, float *ms, int n);
/* Tuning knob for tests: force the binned path even for tiny draws (default: direct mode below a threshold). */
int swcu_set_option(swcu_ctx *ctx, const char *name, int value);
const char *swcu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SWCU_H */
