/*
 * scene_format.h — on-disk scene consumed by oracle/refrender.cpp (the harness that drives the
 * REFERENCE ICD, oracle/_ref/libvk_swiftshader.so).  TEST INFRASTRUCTURE ONLY.
 *
 * Layout: SceneHeader, numDraws x SceneDraw, numBlobs x SceneBlob, then blob bytes.
 * All little-endian, 4-byte packed.  Written by swiftshader_b200/scene.py (Scene.write_ref_scene).
 * The inputs mirror what tests/VulkanWrapper/DrawTester.cpp:62-102,205-406 sets up for the
 * reference's own benchmarks (tests/VulkanBenchmarks/TriangleBenchmarks.cpp).
 */
#ifndef SCENE_FORMAT_H
#define SCENE_FORMAT_H
#include <stdint.h>

#define SCENE_MAGIC 0x43535753u /* "SWSC" */
#define SCENE_VERSION 7u
#define SCENE_MAX_ATTRIBS 8

#pragma pack(push, 4)
typedef struct SceneHeader
{
	uint32_t magic, version;
	uint32_t width, height, samples;
	uint32_t colorFormat;     /* VkFormat */
	uint32_t hasDepth, hasStencil;
	float clearColor[4];
	float clearDepth;
	uint32_t clearStencil;
	uint32_t numDraws, numBlobs;
} SceneHeader;

typedef struct SceneStencilFace { uint32_t failOp, passOp, depthFailOp, compareOp, compareMask, writeMask, reference; } SceneStencilFace;

typedef struct SceneDraw
{
	uint32_t topology;
	uint32_t indexType;   /* 0 none, 2 u16, 4 u32 */
	uint32_t count;       /* vertexCount / indexCount */
	uint32_t firstIndex;  /* firstVertex for non-indexed */
	int32_t vertexOffset;
	uint32_t vsBlob, fsBlob, vertexBlob, indexBlob; /* blob ids; indexBlob ignored if indexType==0 */
	uint32_t stride;
	uint32_t numAttribs;
	struct { uint32_t location, format, offset; } attrib[SCENE_MAX_ATTRIBS];
	float viewport[6];   /* x y w h minDepth maxDepth */
	int32_t scissor[4];  /* x y w h */
	uint32_t cullMode, frontFace;
	uint32_t depthTestEnable, depthWriteEnable, depthCompareOp;
	uint32_t stencilTestEnable;
	SceneStencilFace front, back;
	uint32_t depthBiasEnable;
	float depthBiasConstant, depthBiasClamp, depthBiasSlope;
	uint32_t blendEnable, srcColor, dstColor, colorOp, srcAlpha, dstAlpha, alphaOp, colorWriteMask;
	float blendConstants[4];
	uint32_t sampleMask;
	uint32_t hasTexture, texBlob, texWidth, texHeight, texLevels;
	uint32_t magFilter, minFilter, mipmapMode, addressModeU, addressModeV;
	float mipLodBias, minLod, maxLod;
	uint32_t texSet, texBinding;
	uint32_t alphaToCoverageEnable, depthBoundsTestEnable;
	float minDepthBounds, maxDepthBounds;
	float lineWidth;            /* VkPipelineRasterizationStateCreateInfo::lineWidth (0 is read as 1) */
	uint32_t pushConstantBytes; /* vkCmdPushConstants(VERTEX, 0, bytes) ahead of the draw; 0 = no push-constant range */
	uint32_t pushConstants[32];
	uint32_t depthClampEnable;  /* VkPipelineRasterizationStateCreateInfo::depthClampEnable */
	/* instancing: vkCmdDraw*(…, instanceCount, …); attributes of vertex binding 1 (VK_VERTEX_INPUT_RATE_INSTANCE) */
	uint32_t instanceCount;     /* >= 1 */
	uint32_t instanceBlob, instanceStride, numInstanceAttribs;
	struct { uint32_t location, format, offset; } instanceAttrib[4];
	/* a uniform buffer for the vertex stage (VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER at (uboSet, uboBinding)); set 0 only */
	uint32_t hasUbo, uboBlob, uboSet, uboBinding;
} SceneDraw;

typedef struct SceneBlob { uint64_t offset, size; } SceneBlob;
#pragma pack(pop)

#endif
