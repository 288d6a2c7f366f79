// swcu_internal.h — structures shared by the host side (draw.cu, context.cpp) and the kernels (kernels.cuh).
//
// The draw path replaces DrawCall::run and everything below it (reference: src/Device/Renderer.cpp:551-662).
// Per draw the host folds the pipeline state exactly like sw::Renderer::draw (Renderer.cpp:183-490) into one
// DrawConst that is passed BY VALUE as a __grid_constant__ kernel parameter; all pointers inside are DEVICE
// addresses of the registered shadows.
#pragma once

#include "../../include/swcu.h"

#include <stdint.h>

// ---- Vulkan enum values used on this path (vulkan_core.h) ----
enum
{
	VKF_UNDEFINED = 0,
	VKF_R8G8B8A8_UNORM = 37,
	VKF_R8G8B8A8_SRGB = 43,
	VKF_B8G8R8A8_UNORM = 44,
	VKF_B8G8R8A8_SRGB = 50,
	VKF_R16G16B16A16_SFLOAT = 97,
	VKF_R32_SFLOAT = 100,
	VKF_R32G32_SFLOAT = 103,
	VKF_R32G32B32_SFLOAT = 106,
	VKF_R32G32B32A32_SFLOAT = 109,
	VKF_D16_UNORM = 124,
	VKF_D32_SFLOAT = 126,
	VKF_S8_UINT = 127,
};
enum { TOPO_POINT_LIST = 0, TOPO_LINE_LIST = 1, TOPO_LINE_STRIP = 2, TOPO_TRIANGLE_LIST = 3, TOPO_TRIANGLE_STRIP = 4, TOPO_TRIANGLE_FAN = 5 };
enum { PRIM_TRIANGLE = 0, PRIM_LINE = 1, PRIM_POINT = 2 }; // SetupProcessor.cpp:71-73
enum { CMP_NEVER, CMP_LESS, CMP_EQUAL, CMP_LESS_OR_EQUAL, CMP_GREATER, CMP_NOT_EQUAL, CMP_GREATER_OR_EQUAL, CMP_ALWAYS };
enum { SOP_KEEP, SOP_ZERO, SOP_REPLACE, SOP_INC_CLAMP, SOP_DEC_CLAMP, SOP_INVERT, SOP_INC_WRAP, SOP_DEC_WRAP };
enum
{
	BF_ZERO, BF_ONE, BF_SRC_COLOR, BF_ONE_MINUS_SRC_COLOR, BF_DST_COLOR, BF_ONE_MINUS_DST_COLOR,
	BF_SRC_ALPHA, BF_ONE_MINUS_SRC_ALPHA, BF_DST_ALPHA, BF_ONE_MINUS_DST_ALPHA,
	BF_CONSTANT_COLOR, BF_ONE_MINUS_CONSTANT_COLOR, BF_CONSTANT_ALPHA, BF_ONE_MINUS_CONSTANT_ALPHA,
	BF_SRC_ALPHA_SATURATE
};
enum { BOP_ADD, BOP_SUBTRACT, BOP_REVERSE_SUBTRACT, BOP_MIN, BOP_MAX,
	   BOP_ZERO_EXT = 1000148000, BOP_SRC_EXT = 1000148001, BOP_DST_EXT = 1000148002 };
// compact blend-op codes used inside the kernels
enum { KOP_ADD, KOP_SUB, KOP_RSUB, KOP_MIN, KOP_MAX, KOP_ZERO, KOP_SRC, KOP_DST };
enum { CULL_FRONT = 1, CULL_BACK = 2 };
enum { FRONT_FACE_CCW = 0, FRONT_FACE_CW = 1 };
enum { FILTER_NEAREST = 0, FILTER_LINEAR = 1 };
enum { MIPMAP_MODE_NEAREST = 0, MIPMAP_MODE_LINEAR = 1 };
enum { ADDR_REPEAT = 0, ADDR_MIRRORED_REPEAT = 1, ADDR_CLAMP_TO_EDGE = 2 };

// Device/Clipper.hpp:28-41
enum { CLIP_RIGHT = 1, CLIP_TOP = 2, CLIP_FAR = 4, CLIP_LEFT = 8, CLIP_BOTTOM = 16, CLIP_NEAR = 32, CLIP_FINITE = 128 };
#define CLIP_FRUSTUM (CLIP_RIGHT | CLIP_TOP | CLIP_FAR | CLIP_LEFT | CLIP_BOTTOM | CLIP_NEAR)
#define CLIP_SIDES (CLIP_LEFT | CLIP_RIGHT | CLIP_TOP | CLIP_BOTTOM)

#define SWCU_MAXSLOTS 6      // plane-equation slots per triangle: 4 colour channels + (u, v)
#define SWCU_TILE_W 32       // screen tile staged in shared memory by one CTA
#define SWCU_TILE_H 16
#define SWCU_REGION_W 16     // sub-tile owned by one warp; also the granularity of the triangle bins
#define SWCU_REGION_H 8
#define SWCU_TILE_WARPS ((SWCU_TILE_W / SWCU_REGION_W) * (SWCU_TILE_H / SWCU_REGION_H))
#define SWCU_SMALL_ROWS 8    // a SMALL triangle fits an 8 x 8 pixel frame: its coverage travels as bit masks inside its record
#define SWCU_SMALL_COLS 8
#define SWCU_POLY_MAX 10     // 3 + 6 clip planes (+1 wrap slot)
#define SWCU_SORT_CAP 256    // bins up to this many entries are put in triangle order in shared memory (longer ones in place, in global memory)

// operand kinds after routing (device side)
enum { OPK_CONST = 0, OPK_INPUT = 1, OPK_TEXEL = 2 };
// fragment shader classes the tile kernel is specialised on
enum { SH_CONST = 0, SH_VARY = 1, SH_TEX = 2, SH_GENERIC = 3 };
// interpolation mode of a plane slot (SpirvShader.hpp:761-782: Flat / NoPerspective; default = perspective)
enum { IM_PERSP = 0, IM_NOPERSP = 1, IM_FLAT = 2 };
enum { CK_CONST = 0, CK_SLOT = 1, CK_TEXEL = 2 };
// blend classes
enum { BL_OFF = 0, BL_SRC_ALPHA = 1, BL_GENERIC = 2 };

struct KOperand
{
	uint32_t kind;  // OPK_*
	uint32_t value; // CONST: float bits; INPUT: vertex stage = location*4+component, fragment stage = packed interpolant index; TEXEL: channel
};

// One scalar the vertex stage produces, resolved on the host to "constant" or "component c of attribute stream l"
// (VertexRoutine::readStream, VertexRoutine.cpp:173-245, for R32..R32G32B32A32_SFLOAT streams).
struct KVSrc
{
	const unsigned char *ptr; // attribute base + 4*component; nullptr => constant
	uint32_t stride;
	uint32_t limit;           // robustBufferAccess: fetch yields 0 when byte offset > limit (0xFFFFFFFF = unchecked)
	float constant;           // value when ptr == nullptr (shader constant, or the (0,0,0,1) default of a short format)
	uint32_t pad;
};

// One scalar step of the vertex stage's arithmetic (an MVP transform).  The translator lowers OpMatrixTimesVector & co.
// to SWCU_OP_* steps; the host resolves their operands per draw: push-constant words become constants (DrawData::pushConstants is
// fixed for the draw), inputs become entries of DrawConst::vsIn.
enum { VK_CONST = 0, VK_INPUT = 1, VK_TEMP = 2 };
struct KVsOperand
{
	uint32_t kind;  // VK_*
	uint32_t value; // CONST: float bits; INPUT: index into DrawConst::vsIn; TEMP: index of the step that computed it
};
struct KVsStep
{
	uint32_t op; // SWCU_OP_*
	KVsOperand a, b, c;
};
#define SWCU_VS_INPUTS 16 // distinct attribute components a vertex program may read

struct KMip
{
	const unsigned char *buffer;
	uint32_t width, height, pitchP;
	uint32_t half; // (0x8000 / width) | (0x8000 / height) << 16: the half-texel offsets of Mipmap::uHalf/vHalf (Sampler.hpp:25-39)
};

struct KStencilFace
{
	uint32_t failOp, passOp, depthFailOp, compareOp, compareMask, writeMask, reference, pad;
};

// One BIG triangle (anything that does not fit the 8 x 8 pixel frame of a small one, clipped polygons, garbage coordinates):
// the polygon in 24.8 fixed point.  Its spans are evaluated in closed form by the region warps of the tile kernel, its
// (region, triangle) pairs by k_bigcount / k_fill.
struct BigTri
{
	uint32_t tri;
	uint32_t walk;    // 1: its pairs did not fit the pair budget (see DrawConst::bigBudget): not binned
	int32_t n, dir;
	int32_t yMin, yMax;
	int32_t pxMin, pxMax;
	int32_t X[SWCU_POLY_MAX], Y[SWCU_POLY_MAX];
};

// Per-draw device counters (one struct in device memory, zeroed per draw)
struct DrawCounters
{
	unsigned long long bigSlots;    // entries appended to the big list
	unsigned long long bigReserved; // (region, triangle) pairs reserved by big triangles (bounding-box count)
	unsigned long long pairTotal;   // number of (region, triangle) pairs binned, written by the scan
	uint32_t overflow;              // bit1: big list full; bit2: a big triangle did not fit the pair budget
	uint32_t visible;               // triangles that survived setup
	uint32_t scanTicket;            // k_binscan: block tickets
	uint32_t bigDone;               // k_binscan: blocks that have finished counting the big triangles' regions
};

struct DrawConst
{
	// ---- DrawData scalars (Renderer.hpp:58-113, Renderer.cpp:300-345) ----
	float WxF, HxF, X0xF, Y0xF, depthRange, depthNear;
	int32_t scX0, scX1, scY0, scY1; // scissor ∩ render area: what the tile phase renders (a rank of a group: its band)
	int32_t suY0, suY1;             // rows the SETUP clamps against: the same, except in a group, where every rank sets its share of
	                                // the triangles up for the whole frame (scissor ∩ framebuffer) and the owners clip to their bands
	int32_t ms; // 1 or 4
	uint32_t sampleMask;

	// ---- input assembly ----
	const void *indexBuffer;
	uint32_t indexType, topology, provokingFirst, primCount;
	int32_t baseVertex;

	// ---- shader routing (output of the SPIR-V subset translator), resolved to per-triangle plane slots ----
	// The fragment shaders of the subset are pure routing, so k_setup evaluates the vertex-stage operand of every value
	// the fragment stage consumes and writes one plane equation per SLOT: slots [0,4) are the colour channels of
	// output 0 (shaderClass VARY/GENERIC), the last two are the texture coordinate (TEX/GENERIC).
	KVSrc vsPos[4];
	uint32_t shaderClass;        // SH_*
	int32_t nslots;
	KVSrc slotSrc[SWCU_MAXSLOTS];     // vertex-stage scalar feeding the slot
	uint32_t slotMode[SWCU_MAXSLOTS]; // IM_*
	uint32_t chanKind[4];        // CK_* : where colour channel ch comes from
	uint32_t chanValue[4];       // CK_CONST: float bits; CK_SLOT: slot index; CK_TEXEL: texel component
	uint32_t usesTexture;
	int32_t uvSlot;              // first of the two texcoord slots
	// lines and points (DrawCall::setupLine / setupPoint, Renderer.cpp:920-1185): the primitive becomes a clipped quad in k_setup_prog
	uint32_t primKind;            // PRIM_*
	float lineWidth, halfPixelX, halfPixelY; // Renderer.cpp:276,317-318
	KVSrc pointSizeSrc;           // gl_PointSize (constant 1 when the shader does not store it)
	int32_t pointSizeTemp;
	// vertex-stage arithmetic: when vsProgLen > 0, k_setup_prog runs the steps for each of the three vertices first; a position
	// component / slot source with a non-negative *Temp index takes the result of that step instead of its KVSrc
	uint32_t vsProgLen;
	int32_t posTemp[4];
	int32_t slotTemp[SWCU_MAXSLOTS];
	KVSrc vsIn[SWCU_VS_INPUTS];
	KVsStep vsProg[SWCU_MAX_PROGRAM];

	// ---- setup state ----
	uint32_t cullMode, frontFace, depthClipEnable;
	float depthBiasConstant, depthBiasSlope, depthBiasClamp;
	uint32_t depthBiasEnable;

	// ---- pixel state (PixelProcessor.cpp:74-140, Context.cpp:1090-1312 folded) ----
	uint32_t depthTestActive, depthWriteEnable, depthCompareOp;
	float minDepthClamp, maxDepthClamp; // clampDepth (PixelRoutine.cpp:484-492): [0, 1], or the viewport's depth range with depthClampEnable
	uint32_t depth16; // D16_UNORM depth buffer: quantised compare / saturating write (PixelRoutine.cpp:466-482,508-511,687-711)
	uint32_t stencilActive, stencilWrite;
	uint32_t alphaToCoverage; // c[0].w against the per-sample thresholds of Renderer.cpp:391-410 (PixelRoutine.cpp:643-658)
	uint32_t depthBounds;     // 0 off; 1: failing samples leave the depth mask (:638-641); 2: no depth test in the reference's state, failing
	                          //    samples leave the coverage mask (:634-637) - the depth plane is still staged, compare op forced to ALWAYS
	float minDepthBounds, maxDepthBounds;
	KStencilFace front, back;
	uint32_t blendEnable, srcF, dstF, op, srcFA, dstFA, opA; // op/opA are KOP_*
	uint32_t colorWriteMask;
	float blendConstant[4]; // clamped to [0,1] (UNORM target)
	uint32_t bgr;
	uint32_t colorEpp; // 32-bit words per colour pixel: 1 (RGBA8 family), 2 (R16G16B16A16_SFLOAT), 4 (R32G32B32A32_SFLOAT); > 1 = floating-point target:
	                   // no clamping of shader output / blend constants, Half conversions of Reactor.cpp:3744-3815, masked bit-exact stores
	uint32_t srgb; // sRGB colour target: encode before the pack, decode the destination when blending (PixelRoutine.cpp:1821-1826,1965-1970)

	// ---- attachments (device addresses) ----
	unsigned char *colorBuf, *depthBuf, *stencilBuf;
	int32_t colorPitchB, colorSliceB, depthPitchB, depthSliceB, stencilPitchB, stencilSliceB;
	int32_t fbWidth, fbHeight;

	// ---- sampled image (SamplerCore state, SpirvShaderSampling.cpp:49-128) ----
	KMip mip[SWCU_MIPMAP_LEVELS];
	uint32_t texLevels;
	uint32_t magFilter, minFilter, mipmapMode, addressU, addressV;
	float mipLodBias, minLod, maxLod;
	uint32_t texSrgb; // R8G8B8A8_SRGB image: RGB texels go through the reference's sRGBtoLinearFF_FF00 table (SamplerCore.cpp:1966-1977)
	uint32_t texFast; // REPEAT/REPEAT, LINEAR/LINEAR, MIPMAP_LINEAR: the benchmark sampler, state tests folded away

	// ---- work buffers ----
	unsigned char *triRecords; // primCount records of triStride bytes
	uint32_t triStride;
	uint32_t planeOffset;      // bytes from the start of a record to its plane equations (16, or 48 with the 4x masks)
	BigTri *bigList;
	uint32_t bigCapacity;
	uint32_t *triRect;         // per triangle: packed region rectangle of a small triangle / BIG | big-list slot / NONE
	uint32_t *binCount;        // per region bin: pairs counted by k_setup / k_bigcount, counted back down to zero by k_fill
	uint32_t *binStart;        // per region bin (+1): exclusive scan of binCount
	uint32_t *pairs;           // triangle ids, bin after bin
	unsigned long long bigBudget; // pairs the big triangles of this draw may add (capacity of `pairs` minus 4 per triangle)
	uint32_t inputsExternal;        // an index / vertex stream lives in caller-owned device memory (host side only)
	const unsigned char *cullFlags; // band mode: 0 = rows outside the band (k_cull), nullptr when the pass is not run
	DrawCounters *counters;
	DrawCounters *hostCounters; // mapped pinned page: the tile kernel leaves the counters of the draw there (nobody waits for them)
	const void *zeroPage;      // 256 readable bytes: target of the discarded loads of branch-free attribute fetches
	int32_t tilesX, tilesY;    // tile grid of the framebuffer
	int32_t tileX0, tileY0, tileX1, tileY1; // tile range touched by the scissor (exclusive upper)
	uint32_t numBins;          // tilesX * tilesY * 4: bin = tile * 4 + (region row & 1) * 2 + (region column & 1)
	// ---- group (multi-GPU): this rank sets up triangles [triLo, triHi) and delivers record, rectangle, big-list entry and bin
	//      counts of each to the rank(s) whose band of bandRows rows the triangle touches, over NVLink.  world == 1: everything local ----
	uint32_t triLo, triHi;
	uint32_t world, rank;
	int32_t bandRows;
	unsigned char *peerRecords[SWCU_MAX_GROUP];
	uint32_t *peerRect[SWCU_MAX_GROUP];
	uint32_t *peerBinCount[SWCU_MAX_GROUP];
	BigTri *peerBig[SWCU_MAX_GROUP];
	DrawCounters *peerCounters[SWCU_MAX_GROUP];
	uint32_t direct;           // 1: no binning, every region warp walks all triangles
	uint32_t blendClass;       // BL_*
	uint32_t useTma;           // attachments satisfy the tensor-map alignment rules: stage the tile with TMA
	uint32_t writeOnly;        // fast state, no blending, no depth / stencil: no attachment is read, the colour goes straight to the framebuffer
};

// TriRecord layout (triStride bytes, 16-byte aligned):
//   small triangle (fits an 8 x 8 pixel frame at (fx, fy)):
//     uint32 fx | fy << 16;  uint32 flags;  uint32 mask[2]        1x: bit (8 * row + column) = pixel (fx + column, fy + row) covered
//     uint32 mask[8]                                             4x only: word = row, byte = sample, bit = column
//   big triangle:
//     uint32 pxMin | pxMax << 16;  uint32 flags;  uint32 yMin | yMax << 16;  uint32 slot in the big list
//   flags: bit0 = clockwiseMask (front facing, Primitive.hpp:60); bit1 = big
//   then the plane equations (Primitive::{x0,y0,w,V,z,zBias}), 128-bit aligned:
//     float x0, y0, wA, wB, wC, rhwConst;  float V[nslots][3];  (pad to 16 bytes)  float zBias, zA, zB, zC (only with a depth test)
//   rhwConst = 1/w when the w plane is constant (0 otherwise).
#define TRI_HEADER_BYTES 16
#define TRI_FLAG_FRONT 1u
#define TRI_FLAG_BIG 2u
#define TRI_FLOATS_FRONT 6

// Per-triangle word k_setup leaves for the binning steps.  Small: region rectangle rx0 (9 bits) | ry0 << 9 (10 bits) |
// (rx1 - rx0) << 19 (1 bit) | (ry1 - ry0) << 20 (1 bit); big: TRI_RECT_BIG | slot in the big list; invisible: TRI_RECT_NONE.
#define TRI_RECT_BIG 0x80000000u
#define TRI_RECT_NONE 0xFFFFFFFFu
static inline uint32_t swcu_front_f4(int nslots) { return (uint32_t)((TRI_FLOATS_FRONT + 3 * nslots + 3) / 4); }
static inline uint32_t swcu_plane_offset(int ms) { return ms > 1 ? TRI_HEADER_BYTES + 32u : TRI_HEADER_BYTES; }
static inline uint32_t swcu_tri_stride(int nslots, int ms, int depth) { return swcu_plane_offset(ms) + 16u * (swcu_front_f4(nslots) + (depth ? 1u : 0u)); }
