/*
 * swref.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of SwiftShader's draw hot path, written to follow the
 * reference's own structure (per-primitive span table filled by an integer DDA,
 * scanline-pair traversal, 2x2 quads) so that it is an INDEPENDENT check of the
 * CUDA path in swiftshader_b200/csrc (which uses closed-form spans, screen tiles
 * and per-quad-position ownership).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call this.
 *
 * Parity pinning: this oracle is checked against renders of the reference ICD
 * itself (oracle/_ref/libvk_swiftshader.so driven by oracle/refrender.cpp);
 * the resulting golden fixtures live in tests/golden/ (see tests/golden/README.md).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/src).  Float discipline: every '*' '+' '-' '/'
 * is a single IEEE-754 binary32 op (compile with -ffp-contract=off), fmaf() only
 * where the reference writes MulAdd()/mulAdd(), workers run FTZ+DAZ
 * (System/SwiftConfig.cpp:136-139) which we set in MXCSR for the duration of a call.
 */
#include "../include/swcu.h"
#include "../include/swcu_srgb_lut.h"

#include <immintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- Vulkan enum values used (vulkan_core.h) ---- */
enum
{
	FMT_UNDEFINED = 0,
	FMT_R8G8B8A8_UNORM = 37,
	FMT_R8G8B8A8_SRGB = 43,
	FMT_B8G8R8A8_UNORM = 44,
	FMT_B8G8R8A8_SRGB = 50,
	FMT_R16G16B16A16_SFLOAT = 97,
	FMT_R32_SFLOAT = 100,
	FMT_R32G32_SFLOAT = 103,
	FMT_R32G32B32_SFLOAT = 106,
	FMT_R32G32B32A32_SFLOAT = 109,
	FMT_D16_UNORM = 124,
	FMT_D32_SFLOAT = 126,
	FMT_S8_UINT = 127,
};
enum { TOPO_POINT_LIST = 0, TOPO_LINE_LIST = 1, TOPO_LINE_STRIP = 2, TOPO_TRIANGLE_LIST = 3, TOPO_TRIANGLE_STRIP = 4, TOPO_TRIANGLE_FAN = 5 };
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
enum { CULL_FRONT = 1, CULL_BACK = 2 };
enum { FRONT_FACE_CCW = 0, FRONT_FACE_CW = 1 };
enum { FILTER_NEAREST = 0, FILTER_LINEAR = 1 };
enum { MIPMAP_MODE_NEAREST = 0, MIPMAP_MODE_LINEAR = 1 };
enum { ADDR_REPEAT = 0, ADDR_MIRRORED_REPEAT = 1, ADDR_CLAMP_TO_EDGE = 2 };

/* Device/Clipper.hpp:28-41 */
enum { CLIP_RIGHT = 1, CLIP_TOP = 2, CLIP_FAR = 4, CLIP_LEFT = 8, CLIP_BOTTOM = 16, CLIP_NEAR = 32, CLIP_FINITE = 128 };
#define CLIP_FRUSTUM (CLIP_RIGHT | CLIP_TOP | CLIP_FAR | CLIP_LEFT | CLIP_BOTTOM | CLIP_NEAR)
#define CLIP_SIDES (CLIP_LEFT | CLIP_RIGHT | CLIP_TOP | CLIP_BOTTOM) /* Clipper.hpp:41 */
/* `lineWidth * 0.5f / sqrt(dx * dx + dy * dy)` of Renderer.cpp:965.  <cmath>'s unqualified sqrt on a float argument: which overload the
 * reference build picks decides whether the quotient is formed in float or in double; pinned against the reference ICD's line goldens. */
#ifndef SWREF_LINE_SQRT_DOUBLE
#define SWREF_LINE_SQRT_DOUBLE 0
#endif
#if SWREF_LINE_SQRT_DOUBLE
#define LINE_SCALE(lw, dx, dy) ((float)((double)((lw) * 0.5f) / sqrt((double)((dx) * (dx) + (dy) * (dy)))))
#else
#define LINE_SCALE(lw, dx, dy) ((lw) * 0.5f / sqrtf((dx) * (dx) + (dy) * (dy)))
#endif

#define OUTLINE_RESOLUTION 8192 /* Device/Config.hpp:20 */
#define SUBPIX_B 8              /* Vulkan/VkConfig.hpp:101 (CMake build) */
#define SUBPIX_M 255

/* Pipeline/Constants.hpp:26-52, Constants.cpp:291-297 */
static const float SampleLocationsX[4] = { 0.375f - 0.5f, 0.875f - 0.5f, 0.125f - 0.5f, 0.625f - 0.5f };
static const float SampleLocationsY[4] = { 0.125f - 0.5f, 0.375f - 0.5f, 0.625f - 0.5f, 0.875f - 0.5f };
static const int Xf[4] = { -32, 96, -96, 32 };
static const int Yf[4] = { -96, -32, 32, 96 };
#define YMIN_MS_OFFSET (256 - 96 - 1) /* 159 */
#define YMAX_MS_OFFSET (256 + 96 - 1) /* 351 */

typedef struct { float x, y, z, w; } f4;
typedef struct { float A, B, C; } Plane;
typedef struct { short left, right; } Span;

/* Device/Vertex.hpp:23-54 */
typedef struct
{
	f4 position;
	int clipFlags;
	int X, Y;     /* projected.x/.y, 24.8 fixed point */
	float pz, pw; /* projected.z, projected.w (= rhw) */
	float pointSize; /* gl_PointSize (VertexRoutine.cpp:641-650); 1.0 when the shader does not store it (the reference leaves it unset) */
	float v[SWCU_MAX_VARYING_COMPONENTS];
} Vertex;

/* Device/Primitive.hpp:37-70 — one per sample */
typedef struct
{
	int yMin, yMax;
	float x0, y0;
	Plane z, w, V[SWCU_MAX_VARYING_COMPONENTS];
	float zBias;
	int clockwise; /* clockwiseMask != 0 */
	Span outlineStore[OUTLINE_RESOLUTION + 4];
} Primitive;
#define OUTLINE(p) ((p)->outlineStore + 2)

typedef struct
{
	const swcu_draw_desc *d;
	const swcu_shader_info *vs, *fs;
	/* DrawData (Device/Renderer.hpp:58-113) */
	float WxF, HxF, X0xF, Y0xF, depthRange, depthNear;
	int scissorX0, scissorX1, scissorY0, scissorY1;
	int ms; /* multiSampleCount */
	int enableMultiSampling;
	/* folded blend state (Device/Context.cpp:1090-1117) */
	int blendEnable, srcF, dstF, op, srcFA, dstFA, opA;
	int colorWriteMask;
	int depthTestActive, depthWriteEnable, stencilActive, depthBoundsActive;
	int floatTarget, colorBpp; /* R32G32B32A32_SFLOAT / R16G16B16A16_SFLOAT colour attachment; bytes per pixel */
	int numVaryings; /* packed interpolants = set bits of fs->inputMask */
	int interpolateZ, interpolateW;
	int prim; /* 0 = triangles, 1 = lines, 2 = points (SetupProcessor.cpp:71-73 isDrawTriangle / isDrawLine / isDrawPoint) */
	float lineWidth, halfPixelX, halfPixelY; /* Renderer.cpp:276,317-318 */
	float minDepthClamp, maxDepthClamp;      /* PixelProcessor.cpp:121-136 */
} Draw;

/* x86 conversions used by Reactor: RoundInt = cvtps2dq, Int(float) = cvttps2dq (Reactor/LLVMReactor.cpp:135-138,2694-2703) */
static inline int round_int(float x) { return _mm_cvtss_si32(_mm_set_ss(x)); }
static inline int trunc_int(float x) { return _mm_cvttss_si32(_mm_set_ss(x)); }
/* RoundIntClamped: cvtps2dq(Min(x, 0x7FFFFF80)) */
static inline int round_int_clamped(float x) { float c = x < 2147483520.0f ? x : 2147483520.0f; return round_int(c); }
/* SSE maxps/minps operand semantics: result is the second operand unless the compare is strictly true */
static inline float sse_max(float a, float b) { return a > b ? a : b; }
static inline float sse_min(float a, float b) { return a < b ? a : b; }
static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* ------------------------------------------------------------------------------------------------
 * Vertex stage: Pipeline/VertexRoutine.cpp:173-245 (readStream), :116-154 (computeClipFlags),
 * :570-610 (projection in writeCache).  The shader body is the hand-specified swcu_shader_info.
 * ---------------------------------------------------------------------------------------------- */
static void read_stream(const swcu_vertex_input *in, uint32_t index, int baseVertex, float out[4])
{
	out[0] = out[1] = out[2] = 0.0f;
	out[3] = 1.0f;
	int n;
	switch(in->format)
	{
	case FMT_R32_SFLOAT: n = 1; break;
	case FMT_R32G32_SFLOAT: n = 2; break;
	case FMT_R32G32B32_SFLOAT: n = 3; break;
	case FMT_R32G32B32A32_SFLOAT: n = 4; break;
	default: return; /* null stream: defaults */
	}
	uint32_t offset = (index + (uint32_t)baseVertex) * in->vertexStride;
	float zero[4] = { 0, 0, 0, 0 };
	const float *src = (const float *)((const char *)in->buffer + offset);
	if(in->robustnessSize)
	{
		uint32_t o = offset < in->robustnessSize ? offset : in->robustnessSize;
		if(o + (uint32_t)n * 4 > in->robustnessSize) src = zero;
	}
	for(int i = 0; i < n; i++) out[i] = src[i];
}

/* One operand of the vertex stage: constant, input component, word of the push-constant block (DrawData::pushConstants, words the
 * application never pushed read 0 here) or the result of an earlier step of the arithmetic program. */
static float operand(const Draw *dr, const swcu_shader_operand *op, float inputs[SWCU_MAX_INPUTS][4], const float *temps)
{
	if(op->kind == SWCU_SRC_CONST) return as_float(op->value);
	if(op->kind == SWCU_SRC_TEMP) return temps[op->value];
	if(op->kind == SWCU_SRC_PUSH)
	{
		uint32_t w = 0;
		if(4 * op->value + 4 <= dr->d->pushConstantBytes) memcpy(&w, (const char *)dr->d->pushConstants + 4 * op->value, 4);
		return as_float(w);
	}
	if(op->kind == SWCU_SRC_UNIFORM)
	{
		/* word of the uniform block behind the BufferDescriptor bound at the block's (set, binding): the Load through the access chain
		 * into the Block (SpirvShaderMemory.cpp), the descriptor's ptr / sizeInBytes (VkDescriptorSetLayout.cpp:574-600) */
		const uint32_t slot = op->value >> 16, wi = op->value & 0xFFFFu;
		uint32_t w = 0;
		if(slot < dr->vs->uniformCount)
			for(uint32_t i = 0; i < dr->d->uniformBufferCount && i < SWCU_MAX_UNIFORM_BUFFERS; i++)
			{
				const swcu_uniform_buffer *u = &dr->d->uniformBuffer[i];
				if(u->set != dr->vs->uniformSet[slot] || u->binding != dr->vs->uniformBinding[slot]) continue;
				if(u->data && (size_t)4 * wi + 4 <= u->bytes) memcpy(&w, (const char *)u->data + 4 * wi, 4);
				break;
			}
		return as_float(w);
	}
	return inputs[op->value >> 2][op->value & 3];
}

/* The straight-line arithmetic of the vertex stage (an MVP transform): SpirvShaderArithmetic.cpp:39-75,449-457,611-621.  MUL / ADD /
 * SUB are single operations; FMA is Reactor's MulAdd = llvm.fmuladd, one rounding on a host with FMA units (LLVMReactor.cpp:4425-4429;
 * this file is compiled with -mfma -ffp-contract=off, so fmaf() is the fused instruction and nothing else is contracted). */
static void run_program(const Draw *dr, float inputs[SWCU_MAX_INPUTS][4], float *temps)
{
	const swcu_shader_info *vs = dr->vs;
	for(uint32_t i = 0; i < vs->programLength && i < SWCU_MAX_PROGRAM; i++)
	{
		const swcu_shader_op *st = &vs->program[i];
		float a = operand(dr, &st->a, inputs, temps);
		switch(st->op)
		{
		case SWCU_OP_MUL: temps[i] = a * operand(dr, &st->b, inputs, temps); break;
		case SWCU_OP_ADD: temps[i] = a + operand(dr, &st->b, inputs, temps); break;
		case SWCU_OP_SUB: temps[i] = a - operand(dr, &st->b, inputs, temps); break;
		case SWCU_OP_FMA: temps[i] = fmaf(a, operand(dr, &st->b, inputs, temps), operand(dr, &st->c, inputs, temps)); break;
		default: temps[i] = as_float(as_uint(a) ^ 0x80000000u); break; /* SWCU_OP_NEG: LLVM fneg */
		}
	}
}

static void process_vertex(const Draw *dr, uint32_t index, Vertex *v)
{
	const swcu_draw_desc *d = dr->d;
	float inputs[SWCU_MAX_INPUTS][4];
	for(int l = 0; l < SWCU_MAX_INPUTS; l++)
	{
		if(dr->vs->inputMask & (0xFu << (4 * l)) ) read_stream(&d->input[l], index, d->baseVertex, inputs[l]);
		else { inputs[l][0] = inputs[l][1] = inputs[l][2] = 0; inputs[l][3] = 1; }
	}
	float temps[SWCU_MAX_PROGRAM];
	run_program(dr, inputs, temps);
	float px = operand(dr, &dr->vs->position[0], inputs, temps), py = operand(dr, &dr->vs->position[1], inputs, temps);
	float pz = operand(dr, &dr->vs->position[2], inputs, temps), pw = operand(dr, &dr->vs->position[3], inputs, temps);
	v->position.x = px; v->position.y = py; v->position.z = pz; v->position.w = pw;
	for(int i = 0; i < SWCU_MAX_VARYING_COMPONENTS; i++)
		v->v[i] = (dr->vs->outputMask >> i) & 1 ? operand(dr, &dr->vs->output[i], inputs, temps) : 0.0f;
	v->pointSize = dr->vs->writesPointSize ? operand(dr, &dr->vs->pointSize, inputs, temps) : 1.0f;

	/* computeClipFlags, VertexRoutine.cpp:128-152.  Reactor's CmpNLE / CmpNLT / CmpNEQ are ORDERED compares in the LLVM backend
	 * (FCmpOGT / FCmpOGE / FCmpONE, LLVMReactor.cpp:3161-3180,4479-4495), not the x86 cmpnleps they are named after: a NaN w sets
	 * no flag, so the triangle goes to setup unclipped with the clamped projection of VertexRoutine.cpp:609-610 */
	int f = 0;
	if(pw < px) f |= CLIP_RIGHT;
	if(pw < py) f |= CLIP_TOP;
	if(-pw > px) f |= CLIP_LEFT;
	if(-pw > py) f |= CLIP_BOTTOM;
	if(d->depthClipEnable)
	{
		if(pw < pz) f |= CLIP_FAR;
		if(0.0f > pz) f |= CLIP_NEAR;
	}
	if(fabsf(px) <= 3.40282347e38f && fabsf(py) <= 3.40282347e38f && fabsf(pz) <= 3.40282347e38f) f |= CLIP_FINITE;
	v->clipFlags = f;

	/* VertexRoutine.cpp:599-606: w = pos.w | (pos.w == 0 ? bits(1.0f) : 0); rhw = 1/w */
	uint32_t wb = as_uint(pw);
	if(pw == 0.0f) wb |= 0x3F800000u;
	float w = as_float(wb);
	float rhw = 1.0f / w;
	v->X = round_int_clamped(dr->X0xF + px * rhw * dr->WxF);
	v->Y = round_int_clamped(dr->Y0xF + py * rhw * dr->HxF);
	v->pz = pz * rhw;
	v->pw = rhw;
}

/* ------------------------------------------------------------------------------------------------
 * Clipper: Device/Clipper.cpp:22-30 (clipEdge), :32-265 (six planes), :271-299 (Clip)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { f4 P[16]; int n; int i; } Polygon; /* i = number of clip stages run (Polygon.hpp:20-51) */

static f4 clip_edge(f4 Vi, f4 Vj, float di, float dj)
{
	float D = 1.0f / (dj - di);
	f4 o;
	o.x = (dj * Vi.x - di * Vj.x) * D;
	o.y = (dj * Vi.y - di * Vj.y) * D;
	o.z = (dj * Vi.z - di * Vj.z) * D;
	o.w = (dj * Vi.w - di * Vj.w) * D;
	return o;
}

static float plane_dist(int plane, f4 v)
{
	switch(plane)
	{
	case CLIP_NEAR: return v.z;
	case CLIP_FAR: return v.w - v.z;
	case CLIP_LEFT: return v.w + v.x;
	case CLIP_RIGHT: return v.w - v.x;
	case CLIP_TOP: return v.w - v.y;
	default: return v.w + v.y; /* CLIP_BOTTOM */
	}
}

static void clip_plane(Polygon *p, int plane)
{
	f4 T[16];
	int t = 0;
	for(int i = 0; i < p->n; i++)
	{
		int j = i == p->n - 1 ? 0 : i + 1;
		float di = plane_dist(plane, p->P[i]);
		float dj = plane_dist(plane, p->P[j]);
		if(di >= 0)
		{
			T[t++] = p->P[i];
			if(dj < 0) T[t++] = clip_edge(p->P[i], p->P[j], di, dj);
		}
		else
		{
			if(dj > 0) T[t++] = clip_edge(p->P[j], p->P[i], dj, di);
		}
	}
	memcpy(p->P, T, sizeof(f4) * (size_t)t);
	p->n = t;
	p->i += 1;
}

static int clip_polygon(Polygon *p, int flagsOr)
{
	static const int order[6] = { CLIP_NEAR, CLIP_FAR, CLIP_LEFT, CLIP_RIGHT, CLIP_TOP, CLIP_BOTTOM };
	if(flagsOr & CLIP_FRUSTUM)
	{
		for(int k = 0; k < 6; k++)
		{
			if(p->n < 3) break;
			if(flagsOr & order[k]) clip_plane(p, order[k]);
		}
	}
	return p->n >= 3;
}

/* ------------------------------------------------------------------------------------------------
 * Setup: Pipeline/SetupRoutine.cpp:36-512 (generate), :514-548 (setupGradient), :550-621 (edge)
 * ---------------------------------------------------------------------------------------------- */
static void edge(const Draw *dr, Primitive *prim, int Xa, int Ya, int Xb, int Yb)
{
	if(Ya == Yb) return;
	int swap = Yb < Ya;
	int X1 = swap ? Xb : Xa, X2 = swap ? Xa : Xb;
	int Y1 = swap ? Yb : Ya, Y2 = swap ? Ya : Yb;

	int y1 = (Y1 + SUBPIX_M) >> SUBPIX_B;
	int y2 = (Y2 + SUBPIX_M) >> SUBPIX_B;
	int yMin = y1 > dr->scissorY0 ? y1 : dr->scissorY0;
	int yMax = y2 < dr->scissorY1 ? y2 : dr->scissorY1;
	if(!(yMin < yMax)) return;

	int xMin = dr->scissorX0, xMax = dr->scissorX1;
	Span *outline = OUTLINE(prim);

	int DX12 = X2 - X1, DY12 = Y2 - Y1;
	int FDX12 = DX12 << SUBPIX_B, FDY12 = DY12 << SUBPIX_B;

	int X = DX12 * ((y1 << SUBPIX_B) - Y1) + (X1 & SUBPIX_M) * DY12;
	int x = (X1 >> SUBPIX_B) + X / FDY12;
	int d = X % FDY12;
	int ceil = -d >> 31; /* ceiling division: remainder <= 0 */
	x -= ceil;
	d -= ceil & FDY12;

	int Q = FDX12 / FDY12;
	int R = FDX12 % FDY12;
	int floor = R >> 31; /* flooring division: remainder >= 0 */
	Q += floor;
	R += floor & FDY12;

	int D = FDY12;
	int y = y1;
	do
	{
		if(y >= yMin)
		{
			int c = x < xMin ? xMin : (x > xMax ? xMax : x);
			if(swap) outline[y].right = (short)c; else outline[y].left = (short)c;
		}
		x += Q;
		d += R;
		int overflow = -d >> 31;
		d -= D & overflow;
		x -= overflow;
		y++;
	} while(y < yMax);
}

static void rotate1(int c, const Vertex **v0, const Vertex **v1, const Vertex **v2)
{ if(c) { const Vertex *t = *v0; *v0 = *v1; *v1 = *v2; *v2 = t; } }
static void rotate2(int c, const Vertex **v0, const Vertex **v1, const Vertex **v2)
{ if(c) { const Vertex *t = *v2; *v2 = *v1; *v1 = *v0; *v0 = t; } }

/* returns 1 if visible; fills prim[0..ms) */
static int setup_triangle(const Draw *dr, Primitive *prim, const Vertex *tv0, const Vertex *tv1, const Vertex *tv2, const Polygon *poly)
{
	const swcu_draw_desc *d = dr->d;
	int X[16], Y[16];
	X[0] = tv0->X; X[1] = tv1->X; X[2] = tv2->X;
	Y[0] = tv0->Y; Y[1] = tv1->Y; Y[2] = tv2->Y;

	int dir = 1;
	int clockwise = 1; /* lines and points: clockwiseMask = 0xFF, no culling, winding as given (SetupRoutine.cpp:110-114) */
	if(dr->prim == 0)
	{ /* SetupRoutine.cpp:73-115 culling */
		float x0 = (float)X[0], x1 = (float)X[1], x2 = (float)X[2];
		float y0 = (float)Y[0], y1 = (float)Y[1], y2 = (float)Y[2];
		float A = (y0 - y2) * x1 + (y2 - y1) * x0 + (y1 - y0) * x2;
		int w0w1w2 = (int)(as_uint(tv0->position.w) ^ as_uint(tv1->position.w) ^ as_uint(tv2->position.w));
		if(w0w1w2 < 0) A = -A;
		int frontFacing = d->frontFace == FRONT_FACE_CCW ? (A >= 0.0f) : (A <= 0.0f);
		if((d->cullMode & CULL_FRONT) && frontFacing) return 0;
		if((d->cullMode & CULL_BACK) && !frontFacing) return 0;
		if(!(A > 0.0f)) dir = 0;
		clockwise = frontFacing;
	}

	int n = poly->n;
	if(poly->i != 0 || dr->prim != 0) /* clipped, or the quad of a line / point: reproject, SetupRoutine.cpp:120-145 */
	{
		for(int i = 0; i < n; i++)
		{
			f4 v = poly->P[i];
			float rhw = (v.w < 0.0f || v.w > 0.0f) ? 1.0f / v.w : 1.0f; /* Float != is FCmpONE (Reactor.cpp:3965-3968): false for NaN */
			X[i] = round_int(dr->X0xF + v.x * rhw * dr->WxF);
			Y[i] = round_int(dr->Y0xF + v.y * rhw * dr->HxF);
		}
	}

	int yMin = Y[0], yMax = Y[0];
	for(int i = 1; i < n; i++) { if(Y[i] < yMin) yMin = Y[i]; if(Y[i] > yMax) yMax = Y[i]; }
	if(dr->enableMultiSampling) { yMin = (yMin + YMIN_MS_OFFSET) >> SUBPIX_B; yMax = (yMax + YMAX_MS_OFFSET) >> SUBPIX_B; }
	else { yMin = (yMin + SUBPIX_M) >> SUBPIX_B; yMax = (yMax + SUBPIX_M) >> SUBPIX_B; }
	if(yMin < dr->scissorY0) yMin = dr->scissorY0;
	if(yMax > dr->scissorY1) yMax = dr->scissorY1;
	if(yMin >= yMax) return 0;

	for(int q = 0; q < dr->ms; q++)
	{
		int Xq[17], Yq[17];
		for(int i = 0; i < n; i++)
		{
			Xq[i] = X[i]; Yq[i] = Y[i];
			if(dr->enableMultiSampling) { Xq[i] -= Xf[q]; Yq[i] -= Yf[q]; }
		}
		Primitive *p = prim + q;
		Span *outline = OUTLINE(p);
		if(dr->enableMultiSampling)
		{
			int x = (X[0] + SUBPIX_M) >> SUBPIX_B;
			x = x < dr->scissorX0 ? dr->scissorX0 : (x > dr->scissorX1 ? dr->scissorX1 : x);
			for(int y = yMin - 1; y < yMax + 1; y++) { outline[y].left = (short)x; outline[y].right = (short)x; }
		}
		Xq[n] = Xq[0]; Yq[n] = Yq[0];
		for(int i = 0; i < n; i++) edge(dr, p, Xq[i + 1 - dir], Yq[i + 1 - dir], Xq[i + dir], Yq[i + dir]);

		if(!dr->enableMultiSampling)
		{
			for(; yMin < yMax && outline[yMin].left == outline[yMin].right; yMin++) {}
			for(; yMax > yMin && outline[yMax - 1].left == outline[yMax - 1].right; yMax--) {}
			if(yMin == yMax) return 0;
			outline[yMin - 1].left = outline[yMin].left;
			outline[yMin - 1].right = outline[yMin].left;
			outline[yMax].left = outline[yMax - 1].left;
			outline[yMax].right = outline[yMax - 1].left;
		}
	}
	prim->yMin = yMin;
	prim->yMax = yMax;
	prim->clockwise = clockwise;

	/* vertex sort, SetupRoutine.cpp:271-294 (triangles only) */
	const Vertex *v0 = tv0, *v1 = tv1, *v2 = tv2;
	if(dr->prim == 0)
	{
		float y0 = v0->position.y, y1 = v1->position.y, y2 = v2->position.y;
		float ym = sse_min(sse_min(y0, y1), y2);
		rotate1(ym == y1, &v0, &v1, &v2);
		rotate2(ym == y2, &v0, &v1, &v2);
	}
	if(dr->prim == 0)
	{
		float w0 = v0->position.w, w1 = v1->position.w, w2 = v2->position.w;
		float wm = sse_max(sse_max(w0, w1), w2);
		rotate1(wm == w1, &v0, &v1, &v2);
		rotate2(wm == w2, &v0, &v1, &v2);
	}

	float w0 = v0->position.w, w1 = v1->position.w, w2 = v2->position.w;
	float w012[3] = { w0, w1, w2 };
	float rhw0 = v0->pw;
	int X0 = v0->X, X1 = v1->X, X2 = v2->X, Y0 = v0->Y, Y1 = v1->Y, Y2 = v2->Y;
	if(dr->prim == 1) /* line: the third point of the plane equations is the second end point turned by 90 degrees, SetupRoutine.cpp:317-321 */
	{
		X2 = (int)((uint32_t)X1 + (uint32_t)Y1 - (uint32_t)Y0);
		Y2 = (int)((uint32_t)Y1 + (uint32_t)X0 - (uint32_t)X1);
	}
	const float rsub = 1.0f / 256.0f;
	float x0 = (float)X0 * rsub, y0 = (float)Y0 * rsub;
	prim->x0 = x0; prim->y0 = y0;
	X1 -= X0; Y1 -= Y0; X2 -= X0; Y2 -= Y0;
	float x1 = w1 * rsub * (float)X1, y1 = w1 * rsub * (float)Y1;
	float x2 = w2 * rsub * (float)X2, y2 = w2 * rsub * (float)Y2;
	float a = x1 * y2 - x2 * y1;
	float M[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } }; /* [row][x,y,z] */
	M[0][2] = rhw0;
	if(a < 0.0f || a > 0.0f) /* If(a != 0.0f), FCmpONE: a NaN area leaves the zero matrix */
	{
		float A = 1.0f / a;
		float D = A * rhw0;
		M[0][0] = (y1 * w2 - y2 * w1) * D;
		M[0][1] = (x2 * w1 - x1 * w2) * D;
		M[1][0] = y2 * A;
		M[1][1] = -x2 * A;
		M[2][0] = -y1 * A;
		M[2][1] = x1 * A;
	}
	if(dr->interpolateW)
	{
		prim->w.A = M[0][0] + M[1][0] + M[2][0];
		prim->w.B = M[0][1] + M[1][1] + M[2][1];
		prim->w.C = M[0][2] + M[1][2] + M[2][2];
	}
	prim->zBias = 0.0f;
	if(dr->interpolateZ)
	{
		float z0 = v0->pz, z1 = v1->pz, z2 = v2->pz;
		z1 -= z0; z2 -= z0;
		float px1 = (float)X1 * rsub, py1 = (float)Y1 * rsub, px2 = (float)X2 * rsub, py2 = (float)Y2 * rsub;
		float D = dr->depthRange / (px1 * py2 - px2 * py1);
		float A = (py2 * z1 - py1 * z2) * D;
		float B = (px1 * z2 - px2 * z1) * D;
		if(dr->prim == 2) { A = 0.0f; B = 0.0f; } /* point: constant depth, SetupRoutine.cpp:405-409 */
		float C = z0 * dr->depthRange + dr->depthNear;
		prim->z.A = A; prim->z.B = B; prim->z.C = C;
		/* depth bias, SetupRoutine.cpp:417-475 (floating-point depth buffer branch) */
		float bias = 0.0f;
		int applyConst = d->depthBiasConstant != 0.0f, applySlope = d->depthBiasSlope != 0.0f;
		if(dr->prim != 0) applyConst = applySlope = 0; /* SetupProcessor.cpp:75-77: isDrawTriangle(false, polygonMode) */
		if(applyConst)
		{
			float r;
			if(d->depth.format == FMT_D16_UNORM)
				r = 1.01f / 0xFFFF; /* fixed-point depth buffer: DrawData::minimumResolvableDepthDifference, Renderer.cpp:430 */
			else
			{
				float Z0 = C;
				float Z1 = z1 * dr->depthRange + dr->depthNear;
				float Z2 = z2 * dr->depthRange + dr->depthNear;
				int e0 = (int)(as_uint(Z0) & 0x7F800000u), e1 = (int)(as_uint(Z1) & 0x7F800000u), e2 = (int)(as_uint(Z2) & 0x7F800000u);
				int e = e0 > e1 ? e0 : e1; e = e > e2 ? e : e2;
				r = as_float((uint32_t)e) * (1.0f / (1 << 23));
			}
			bias = r * d->depthBiasConstant;
		}
		if(applySlope)
		{
			float m = sse_max(fabsf(A), fabsf(B));
			bias += m * d->depthBiasSlope;
		}
		if(applyConst || applySlope)
		{
			if(d->depthBiasClamp != 0.0f)
			{
				float c = d->depthBiasClamp;
				bias = c > 0.0f ? sse_min(bias, c) : sse_max(bias, c);
			}
			prim->zBias = bias;
		}
	}
	/* setupGradient, SetupRoutine.cpp:514-548 */
	int packed = 0;
	for(int k = 0; k < SWCU_MAX_VARYING_COMPONENTS; k++)
	{
		if(!((dr->fs->inputMask >> k) & 1)) continue;
		Plane *P = &prim->V[packed++];
		if((dr->fs->flatMask >> k) & 1)
		{
			P->A = 0; P->B = 0; P->C = tv0->v[k]; /* provoking vertex = Triangle.v0 */
			continue;
		}
		float iv[3] = { v0->v[k], v1->v[k], v2->v[k] };
		if((dr->fs->noPerspectiveMask >> k) & 1) { iv[0] *= w012[0]; iv[1] *= w012[1]; iv[2] *= w012[2]; }
		P->A = iv[0] * M[0][0] + iv[1] * M[1][0] + iv[2] * M[2][0];
		P->B = iv[0] * M[0][1] + iv[1] * M[1][1] + iv[2] * M[2][1];
		P->C = iv[0] * M[0][2] + iv[1] * M[1][2] + iv[2] * M[2][2];
	}
	return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Sampler: Pipeline/SamplerCore.cpp (16-bit fixed-point path :173-198), RGBA8 2D, normalised coords.
 * ---------------------------------------------------------------------------------------------- */
static inline uint16_t mulhi_u16(uint16_t a, uint16_t b) { return (uint16_t)(((uint32_t)a * b) >> 16); }

/* SamplerCore::address, :2399-2435 */
static uint16_t address(float u, uint32_t mode)
{
	if(mode == ADDR_CLAMP_TO_EDGE)
	{
		float c = sse_min(sse_max(u, 0.0f), 65535.0f / 65536.0f);
		return (uint16_t)trunc_int(c * 65536.0f);
	}
	if(mode == ADDR_MIRRORED_REPEAT)
	{
		int convert = trunc_int(u * 65536.0f);
		int mirror = (int)((uint32_t)convert << 15) >> 31;
		convert ^= mirror;
		return (uint16_t)convert;
	}
	return (uint16_t)trunc_int(u * 65536.0f); /* wrap */
}

/* offsetSample, :278-313 */
static uint16_t offset_sample(uint16_t uvw, uint16_t half, int wrap, int count)
{
	if(wrap) return (uint16_t)(count < 0 ? uvw - half : uvw + half);
	if(count < 0) return uvw < half ? 0 : (uint16_t)(uvw - half);            /* SubSat */
	uint32_t s = (uint32_t)uvw + half; return s > 0xFFFF ? 0xFFFF : (uint16_t)s; /* AddSat */
}

/* one bilinear (or point) tap of one mip level: sampleQuad2D :668-909, computeIndices :1593-1608,
 * sampleTexel :1781-1788 (byte b -> b<<8), bilinearInterpolate :516-575 */
/* One 8-bit texel channel in the 16-bit sampler path: b << 8, or — RGB of an sRGB texture — the reference's start-up table
 * sRGBtoLinearFF_FF00 (SamplerCore.cpp:1966-1977, :2670-2680; include/swcu_srgb_lut.h) */
static const uint16_t srgb_lut[256] = { SWCU_SRGB_LUT_VALUES };
static inline uint16_t texel16(const swcu_sampled_image *t, uint8_t b, int ch)
{
	return (t->format == FMT_R8G8B8A8_SRGB && ch < 3) ? srgb_lut[b] : (uint16_t)(b << 8);
}

static void sample_level(const swcu_sampled_image *t, int level, float u, float v, int linear, uint16_t out[4])
{
	if(level < 0) level = 0;
	if(level > SWCU_MIPMAP_LEVELS - 1) level = SWCU_MIPMAP_LEVELS - 1;
	int l = level < (int)t->levelCount ? level : (int)t->levelCount - 1; /* VkDescriptorSetLayout.cpp:470 */
	const swcu_mip_level *m = &t->level[l];
	const uint8_t *buf = (const uint8_t *)m->buffer;
	uint16_t W = (uint16_t)m->width, H = (uint16_t)m->height;
	uint16_t uuuu = address(u, t->addressModeU), vvvv = address(v, t->addressModeV);
	if(!linear)
	{
		uint32_t x = mulhi_u16(uuuu, W), y = mulhi_u16(vvvv, H);
		const uint8_t *p = buf + 4 * (size_t)(x + y * m->pitchP);
		for(int c = 0; c < 4; c++) out[c] = texel16(t, p[c], c);
		return;
	}
	uint16_t uHalf = (uint16_t)(0x8000 / m->width), vHalf = (uint16_t)(0x8000 / m->height);
	int wrapU = t->addressModeU == ADDR_REPEAT, wrapV = t->addressModeV == ADDR_REPEAT;
	uint16_t u0 = offset_sample(uuuu, uHalf, wrapU, -1), u1 = offset_sample(uuuu, uHalf, wrapU, +1);
	uint16_t v0 = offset_sample(vvvv, vHalf, wrapV, -1), v1 = offset_sample(vvvv, vHalf, wrapV, +1);
	uint32_t x0 = mulhi_u16(u0, W), x1 = mulhi_u16(u1, W), y0 = mulhi_u16(v0, H), y1 = mulhi_u16(v1, H);
	const uint8_t *p00 = buf + 4 * (size_t)(x0 + y0 * m->pitchP);
	const uint8_t *p10 = buf + 4 * (size_t)(x1 + y0 * m->pitchP);
	const uint8_t *p01 = buf + 4 * (size_t)(x0 + y1 * m->pitchP);
	const uint8_t *p11 = buf + 4 * (size_t)(x1 + y1 * m->pitchP);
	uint16_t f0u = (uint16_t)(u0 * W), f0v = (uint16_t)(v0 * H); /* low 16 bits */
	uint16_t f1u = (uint16_t)~f0u, f1v = (uint16_t)~f0v;
	uint16_t f0u0v = mulhi_u16(f0u, f0v), f1u0v = mulhi_u16(f1u, f0v), f0u1v = mulhi_u16(f0u, f1v), f1u1v = mulhi_u16(f1u, f1v);
	for(int c = 0; c < 4; c++)
	{
		uint16_t c00 = mulhi_u16(texel16(t, p00[c], c), f1u1v);
		uint16_t c10 = mulhi_u16(texel16(t, p10[c], c), f0u1v);
		uint16_t c01 = mulhi_u16(texel16(t, p01[c], c), f1u0v);
		uint16_t c11 = mulhi_u16(texel16(t, p11[c], c), f0u0v);
		out[c] = (uint16_t)((uint16_t)(c00 + c10) + (uint16_t)(c01 + c11));
	}
}

/* Min and mag filters differ and the point one is selected: offsetSample masks the half-texel offset to 0 (:282-289), the four
 * taps coincide and the 16-bit blend weights (fractions of the un-offset coordinate) still apply - they sum to less than 1 */
static void sample_level_split_point(const swcu_sampled_image *t, int ilod, float u, float v, uint16_t c[4])
{
	const swcu_mip_level *m = &t->level[ilod < (int)t->levelCount ? (ilod < 0 ? 0 : ilod) : (int)t->levelCount - 1];
	uint16_t uu = address(u, t->addressModeU), vv = address(v, t->addressModeV);
	uint16_t W = (uint16_t)m->width, H = (uint16_t)m->height;
	uint32_t x = mulhi_u16(uu, W), y = mulhi_u16(vv, H);
	const uint8_t *p = (const uint8_t *)m->buffer + 4 * (size_t)(x + y * m->pitchP);
	uint16_t f0u = (uint16_t)(uu * W), f0v = (uint16_t)(vv * H), f1u = (uint16_t)~f0u, f1v = (uint16_t)~f0v;
	uint16_t w00 = mulhi_u16(f1u, f1v), w10 = mulhi_u16(f0u, f1v), w01 = mulhi_u16(f1u, f0v), w11 = mulhi_u16(f0u, f0v);
	for(int ch = 0; ch < 4; ch++)
	{
		uint16_t tx = texel16(t, p[ch], ch);
		c[ch] = (uint16_t)((uint16_t)(mulhi_u16(tx, w00) + mulhi_u16(tx, w10)) + (uint16_t)(mulhi_u16(tx, w01) + mulhi_u16(tx, w11)));
	}
}

/* sampleTexture128 :59-253 for function == Implicit; u[4], v[4] are the quad's lanes */
static void sample_quad(const swcu_sampled_image *t, const float u[4], const float v[4], float out[4][4])
{
	/* state derivation: SpirvShaderSampling.cpp:49-128 */
	int filterLinear;
	int minMagSplit = 0; /* FILTER_MIN_POINT_MAG_LINEAR / MIN_LINEAR_MAG_POINT need the lod sign */
	if(t->magFilter == t->minFilter) filterLinear = t->magFilter == FILTER_LINEAR;
	else { filterLinear = 0; minMagSplit = 1; }
	float minLod = t->minLod, maxLod = t->maxLod;
	if(t->levelCount == 1 && !minMagSplit) { minLod = 0.0f; maxLod = 0.0f; }

	float lod;
	if(minLod == maxLod) lod = minLod; /* skipLodComputation :89-96 */
	else
	{
		/* computeLod2D :1376-1422, log2sqrt :1333-1341 */
		float Wf = (float)t->level[0].width, Hf = (float)t->level[0].height;
		float dUdx = (u[1] - u[0]) * Wf, dUdy = (u[2] - u[0]) * Wf;
		float dVdx = (v[1] - v[0]) * Hf, dVdy = (v[2] - v[0]) * Hf;
		float sx = dUdx * dUdx + dVdx * dVdx, sy = dUdy * dUdy + dVdy * dVdy;
		lod = sse_max(sx, sy);
		lod *= lod;
		lod = (float)(int)as_uint(lod) - (float)0x3F800000;
		lod *= as_float(0x33000000u);
		lod += t->mipLodBias;
		lod = sse_max(lod, minLod);
		lod = sse_min(lod, maxLod);
	}
	int linear = filterLinear;
	if(minMagSplit)
	{
		/* offsetSample masks the half-texel offset by the lod sign (:282-289): MIN_LINEAR_MAG_POINT: linear iff lod > 0 */
		int minLinear = t->minFilter == FILTER_LINEAR;
		linear = minLinear ? (lod > 0.0f) : (lod <= 0.0f); /* CmpNLE is FCmpOGT */
		/* with the offset masked to 0 the four taps coincide and the blend weights still apply (fractions of u0 == u) */
	}
	int ilod;
	if(t->mipmapMode == MIPMAP_MODE_NEAREST) ilod = round_int(lod); else ilod = trunc_int(lod); /* selectMipmap :2357-2379 */

	for(int k = 0; k < 4; k++)
	{
		uint16_t c[4];
		if(minMagSplit && !linear) sample_level_split_point(t, ilod, u[k], v[k], c);
		else sample_level(t, ilod, u[k], v[k], linear, c);
		if(t->mipmapMode == MIPMAP_MODE_LINEAR) /* sampleFilter :324-373 */
		{
			uint16_t cc[4];
			if(minMagSplit && !linear) sample_level_split_point(t, ilod + 1, u[k], v[k], cc); /* both levels of a trilinear fetch */
			else sample_level(t, ilod + 1, u[k], v[k], linear, cc);
			uint16_t utri = (uint16_t)trunc_int(lod * 65536.0f);
			uint16_t inv = (uint16_t)~utri;
			for(int ch = 0; ch < 4; ch++) c[ch] = (uint16_t)(mulhi_u16(c[ch], inv) + mulhi_u16(cc[ch], utri));
		}
		for(int ch = 0; ch < 4; ch++) out[k][ch] = (float)c[ch] * (1.0f / 0xFF00); /* :187-208 scale */
	}
}

/* ------------------------------------------------------------------------------------------------
 * Pixel stage: Device/QuadRasterizer.cpp:71-250, Pipeline/PixelRoutine.cpp:97-358 (+ helpers)
 * ---------------------------------------------------------------------------------------------- */
static int stencil_compare(int op, uint8_t value, uint8_t refMasked)
{
	/* PixelRoutine.cpp:406-449: "(ref & mask) OP (value & mask)" */
	switch(op)
	{
	case CMP_ALWAYS: return 1;
	case CMP_NEVER: return 0;
	case CMP_LESS: return refMasked < value;
	case CMP_EQUAL: return refMasked == value;
	case CMP_NOT_EQUAL: return refMasked != value;
	case CMP_LESS_OR_EQUAL: return refMasked <= value;
	case CMP_GREATER: return refMasked > value;
	default: return refMasked >= value; /* GREATER_OR_EQUAL */
	}
}

static uint8_t stencil_op(int op, uint8_t v, uint8_t ref)
{
	/* PixelRoutine.cpp:870-902 */
	switch(op)
	{
	case SOP_KEEP: return v;
	case SOP_ZERO: return 0;
	case SOP_REPLACE: return ref;
	case SOP_INC_CLAMP: return v == 0xFF ? 0xFF : (uint8_t)(v + 1);
	case SOP_DEC_CLAMP: return v == 0 ? 0 : (uint8_t)(v - 1);
	case SOP_INVERT: return (uint8_t)~v;
	case SOP_INC_WRAP: return (uint8_t)(v + 1);
	default: return (uint8_t)(v - 1);
	}
}

/* Pow<Mediump> = Exp2(y * Log2(x)) with the relaxed-precision polynomials of ShaderCore.cpp:352-382 (Exp2), :412-436 (Log2),
 * :472-477 (Pow); MulAdd is an FMA on every AVX2 host, Float(Int) is cvtdq2ps, Int(Float) truncates. */
static float log2_mediump(float x)
{
	int32_t im = (int32_t)as_uint(x);
	float y = fmaf((float)im, 1.0f / (1 << 23), -127.0f);
	if(im == 0x7F800000) y = as_float(as_uint(y) | 0x7F800000u);
	float m = (float)(im & 0x007FFFFF);
	const float a = 2.8017103e-22f, b = -8.373131e-15f, c = 5.0615534e-8f;
	float f = fmaf(fmaf(a, m, b), m, c);
	return fmaf(f, m, y);
}
static float exp2_mediump(float x)
{
	float x0 = sse_min(x, 128.0f);
	x0 = sse_max(x0, as_float(0xC2FDFFFFu));
	float xi = floorf(x0);
	float f = x0 - xi;
	const float a = 7.8145574e-2f, b = 2.2617357e-1f, c = -3.0444314e-1f;
	float r = fmaf(fmaf(a, f, b), f, c);
	float y = fmaf(r, f, x0);
	int32_t i = trunc_int(fmaf((float)(1 << 23), y, (float)(127 << 23)));
	return as_float((uint32_t)i);
}
static float pow_mediump(float x, float y) { return exp2_mediump(log2_mediump(x) * y); }
/* ShaderCore.cpp:673-689 */
static float linear_to_srgb(float c)
{
	float lc = c * 12.92f;
	float ec = fmaf(1.055f, pow_mediump(c, 1.0f / 2.4f), -0.055f);
	return c < 0.0031308f ? lc : ec;
}
static float srgb_to_linear(float c)
{
	float lc = c * (1.0f / 12.92f);
	float ec = pow_mediump(fmaf(c, 1.0f / 1.055f, 0.055f / 1.055f), 2.4f);
	return c < 0.04045f ? lc : ec;
}

/* Reactor's scalar Half <-> Float conversions (Reactor.cpp:3744-3770, :3787-3815): round-to-nearest-even on the way down with
 * everything above 0x47FFEFFF (incl. NaN) becoming 0x7FFF; on the way up exponent 31 is NOT special (it decodes to 2^16 * 1.m). */
static uint16_t float_to_half(float f)
{
	uint32_t fp32i = as_uint(f), a = fp32i & 0x7FFFFFFFu;
	uint16_t h = (uint16_t)((fp32i & 0x80000000u) >> 16);
	if(a > 0x47FFEFFFu) h |= 0x7FFF;
	else if(a < 0x38800000u)
	{
		int32_t mantissa = (int32_t)((a & 0x007FFFFFu) | 0x00800000u);
		int32_t e = 113 - (int32_t)(a >> 23);
		a = e < 24 ? (uint32_t)(mantissa >> e) : 0u;
		h |= (uint16_t)((a + 0x00000FFFu + ((a >> 13) & 1u)) >> 13);
	}
	else h |= (uint16_t)((a + 0xC8000000u + 0x00000FFFu + ((a >> 13) & 1u)) >> 13);
	return h;
}
static float half_to_float(uint16_t h)
{
	int32_t sgn = (h >> 15) & 1, e = (h >> 10) & 0x1F, m = h & 0x3FF;
	uint32_t fp32i = (uint32_t)sgn << 31;
	if(e == 0)
	{
		if(m != 0)
		{
			while((m & 0x400) == 0) { m <<= 1; e -= 1; }
			fp32i |= (uint32_t)(((e + (127 - 15) + 1) << 23) | ((m & ~0x400) << 13));
		}
	}
	else fp32i |= (uint32_t)(((e + (127 - 15)) << 23) | (m << 13));
	return as_float(fp32i);
}

static float blend_factor(const Draw *dr, int f, int ch, const float s[4], const float dst[4])
{
	/* PixelRoutine.cpp:1225-1393; ch 0..2 = RGB path, 3 = alpha path */
	/* blend constants: PixelProcessor::setBlendConstant keeps a [0,1]-clamped copy for UNORM targets (blendConstantU) and the
	 * raw one for floating-point targets (blendConstantF), PixelRoutine.cpp:1203-1223 */
	float bc[4];
	for(int k = 0; k < 4; k++) bc[k] = dr->floatTarget ? dr->d->blendConstants[k] : sse_min(sse_max(dr->d->blendConstants[k], 0.0f), 1.0f);
	float r;
	if(ch < 3)
	{
		switch(f)
		{
		case BF_ZERO: return 0.0f;
		case BF_ONE: return 1.0f;
		case BF_SRC_COLOR: return s[ch];
		case BF_ONE_MINUS_SRC_COLOR: return 1.0f - s[ch];
		case BF_DST_COLOR: return dst[ch];
		case BF_ONE_MINUS_DST_COLOR: return 1.0f - dst[ch];
		case BF_SRC_ALPHA: return s[3];
		case BF_ONE_MINUS_SRC_ALPHA: return 1.0f - s[3];
		case BF_DST_ALPHA: return dst[3];
		case BF_ONE_MINUS_DST_ALPHA: return 1.0f - dst[3];
		case BF_SRC_ALPHA_SATURATE: r = 1.0f - dst[3]; return sse_min(r, s[3]);
		case BF_CONSTANT_COLOR: return bc[ch];
		case BF_CONSTANT_ALPHA: return bc[3];
		case BF_ONE_MINUS_CONSTANT_COLOR: return 1.0f - bc[ch];
		case BF_ONE_MINUS_CONSTANT_ALPHA: return 1.0f - bc[3];
		}
		return 0.0f;
	}
	switch(f)
	{
	case BF_ZERO: return 0.0f;
	case BF_ONE: return 1.0f;
	case BF_SRC_COLOR: case BF_SRC_ALPHA: return s[3];
	case BF_ONE_MINUS_SRC_COLOR: case BF_ONE_MINUS_SRC_ALPHA: return 1.0f - s[3];
	case BF_DST_COLOR: case BF_DST_ALPHA: return dst[3];
	case BF_ONE_MINUS_DST_COLOR: case BF_ONE_MINUS_DST_ALPHA: return 1.0f - dst[3];
	case BF_SRC_ALPHA_SATURATE: return 1.0f;
	case BF_CONSTANT_COLOR: case BF_CONSTANT_ALPHA: return bc[3];
	case BF_ONE_MINUS_CONSTANT_COLOR: case BF_ONE_MINUS_CONSTANT_ALPHA: return 1.0f - bc[3];
	}
	return 0.0f;
}

static float blend_op(int op, float s, float sf, float dd, float df)
{
	switch(op) /* PixelRoutine.cpp:1849-1958 */
	{
	case BOP_ADD: return s * sf + dd * df;
	case BOP_SUBTRACT: return s * sf - dd * df;
	case BOP_REVERSE_SUBTRACT: return dd * df - s * sf;
	case BOP_MIN: return sse_min(s, dd);
	case BOP_MAX: return sse_max(s, dd);
	case BOP_SRC_EXT: return s;
	case BOP_DST_EXT: return dd;
	default: return 0.0f; /* ZERO_EXT */
	}
}

/* Context.cpp:1165-1270 */
static int fold_blend_op(int op, int sf, int df, int unorm)
{
	switch(op)
	{
	case BOP_ADD:
		if(sf == BF_ZERO) { if(df == BF_ZERO) return BOP_ZERO_EXT; if(df == BF_ONE) return BOP_DST_EXT; }
		else if(sf == BF_ONE) { if(df == BF_ZERO) return BOP_SRC_EXT; }
		break;
	case BOP_SUBTRACT:
		if(sf == BF_ZERO) { if(df == BF_ZERO || unorm) return BOP_ZERO_EXT; } /* ZERO,ZERO; or negative, clamped to zero (UNORM only) */
		else if(sf == BF_ONE) { if(df == BF_ZERO) return BOP_SRC_EXT; }
		break;
	case BOP_REVERSE_SUBTRACT:
		if(sf == BF_ZERO) { if(df == BF_ZERO) return BOP_ZERO_EXT; if(df == BF_ONE) return BOP_DST_EXT; }
		else { if(df == BF_ZERO && unorm) return BOP_ZERO_EXT; }
		break;
	}
	return op;
}
static int fold_blend_factor(int op, int f) { return (op == BOP_MIN || op == BOP_MAX) ? BF_ONE : f; }

static void rasterize(const Draw *dr, const Primitive *prim)
{
	const swcu_draw_desc *d = dr->d;
	const int ms = dr->ms;
	int yMin = prim->yMin & ~1; /* clusterCount = 1: QuadRasterizer.cpp:46-49 */
	const int yMax = prim->yMax;
	const int bgr = d->color.format == FMT_B8G8R8A8_UNORM || d->color.format == FMT_B8G8R8A8_SRGB;
	const int srgb = d->color.format == FMT_R8G8B8A8_SRGB || d->color.format == FMT_B8G8R8A8_SRGB;

	for(int y = yMin; y < yMax; y += 2)
	{
		/* QuadRasterizer.cpp:101-123 */
		int x0 = 0x7FFFFFFF, x1 = -0x7FFFFFFF;
		for(int q = 0; q < ms; q++)
		{
			const Span *o = OUTLINE(prim + q);
			int a = o[y].left < o[y + 1].left ? o[y].left : o[y + 1].left;
			int b = o[y].right > o[y + 1].right ? o[y].right : o[y + 1].right;
			if(a < x0) x0 = a;
			if(b > x1) x1 = b;
		}
		x0 &= ~1;
		float yFragment[4];
		for(int i = 0; i < 4; i++) yFragment[i] = (float)y + (float)(i >> 1) - prim->y0;
		float Dz[4][4];
		if(dr->interpolateZ)
			for(int q = 0; q < ms; q++)
				for(int i = 0; i < 4; i++)
				{
					float yy = yFragment[i];
					if(dr->enableMultiSampling) yy += SampleLocationsY[q];
					Dz[q][i] = prim->z.C + yy * prim->z.B;
				}
		if(!(x0 < x1)) continue;
		float Dw[4], Dv[SWCU_MAX_VARYING_COMPONENTS][4];
		for(int i = 0; i < 4; i++) Dw[i] = prim->w.C + yFragment[i] * prim->w.B;
		{
			int packed = 0;
			for(int k = 0; k < SWCU_MAX_VARYING_COMPONENTS; k++)
			{
				if(!((dr->fs->inputMask >> k) & 1)) continue;
				for(int i = 0; i < 4; i++)
				{
					Dv[k][i] = prim->V[packed].C;
					if(!((dr->fs->flatMask >> k) & 1)) Dv[k][i] += yFragment[i] * prim->V[packed].B;
				}
				packed++;
			}
		}

		for(int x = x0; x < x1; x += 2)
		{
			int cMask[4] = { 0, 0, 0, 0 };
			for(int q = 0; q < ms; q++)
			{
				if(!(d->sampleMask & (1u << q))) continue;
				const Span *o = OUTLINE(prim + (dr->enableMultiSampling ? q : 0));
				for(int i = 0; i < 4; i++)
				{
					int px = x + (i & 1), py = y + (i >> 1);
					if((short)px >= o[py].left && (short)px < o[py].right) cMask[q] |= 1 << i;
				}
			}
			/* ---- PixelRoutine::quad ---- */
			int sMask[4], zMask[4];
			for(int q = 0; q < ms; q++) { sMask[q] = cMask[q]; zMask[q] = cMask[q]; }

			/* stencilTest :360-404 */
			if(dr->stencilActive)
			{
				const swcu_stencil_face *sf = prim->clockwise ? &d->front : &d->back;
				for(int q = 0; q < ms; q++)
				{
					if(!(d->sampleMask & (1u << q))) continue;
					const uint8_t *sb = (const uint8_t *)d->stencil.buffer + (size_t)q * d->stencil.sliceB;
					int pass = 0;
					for(int i = 0; i < 4; i++)
					{
						uint8_t value = sb[(size_t)(y + (i >> 1)) * d->stencil.pitchB + x + (i & 1)];
						if(stencil_compare(sf->compareOp, value & sf->compareMask, sf->reference & sf->compareMask)) pass |= 1 << i;
					}
					sMask[q] &= pass;
				}
			}

			float xFragment[4];
			for(int i = 0; i < 4; i++) xFragment[i] = (float)x + (float)(i & 1) - prim->x0;

			float z[4][4];
			if(dr->interpolateZ)
				for(int q = 0; q < ms; q++)
					for(int i = 0; i < 4; i++)
					{
						float xx = xFragment[i];
						if(dr->enableMultiSampling) xx -= SampleLocationsX[q];
						z[q][i] = fmaf(xx, prim->z.A, Dz[q][i]);
						if(d->depthBiasConstant != 0.0f || d->depthBiasSlope != 0.0f) z[q][i] += prim->zBias;
					}

			/* late-test path (our shaders do not declare EarlyFragmentTests, :99): shade first */
			float w[4], rhw[4];
			for(int i = 0; i < 4; i++) { w[i] = fmaf(xFragment[i], prim->w.A, Dw[i]); rhw[i] = 1.0f / w[i]; }
			float in[SWCU_MAX_VARYING_COMPONENTS][4];
			{
				int packed = 0;
				for(int k = 0; k < SWCU_MAX_VARYING_COMPONENTS; k++)
				{
					if(!((dr->fs->inputMask >> k) & 1)) continue;
					for(int i = 0; i < 4; i++)
					{
						if((dr->fs->flatMask >> k) & 1) in[k][i] = Dv[k][i];
						else
						{
							float t = fmaf(xFragment[i], prim->V[packed].A, Dv[k][i]);
							if(!((dr->fs->noPerspectiveMask >> k) & 1)) t *= rhw[i];
							in[k][i] = t;
						}
					}
					packed++;
				}
			}
			/* executeShader: the hand-specified fragment body */
			float texel[4][4];
			if(dr->fs->usesTexture)
			{
				const swcu_sampled_image *t = NULL;
				for(uint32_t s = 0; s < d->sampledImageCount; s++)
					if(d->sampledImage[s].set == dr->fs->textureSet && d->sampledImage[s].binding == dr->fs->textureBinding) t = &d->sampledImage[s];
				float uu[4], vv[4];
				for(int i = 0; i < 4; i++)
				{
					const swcu_shader_operand *ou = &dr->fs->texCoord[0], *ov = &dr->fs->texCoord[1];
					uu[i] = ou->kind == SWCU_SRC_CONST ? as_float(ou->value) : in[ou->value][i];
					vv[i] = ov->kind == SWCU_SRC_CONST ? as_float(ov->value) : in[ov->value][i];
				}
				if(t) sample_quad(t, uu, vv, texel); else memset(texel, 0, sizeof(texel));
			}
			float c[4][4]; /* [lane][channel] */
			for(int i = 0; i < 4; i++)
				for(int ch = 0; ch < 4; ch++)
				{
					const swcu_shader_operand *o = &dr->fs->output[ch];
					float val;
					if(!((dr->fs->outputMask >> ch) & 1)) val = 0.0f;
					else if(o->kind == SWCU_SRC_CONST) val = as_float(o->value);
					else if(o->kind == SWCU_SRC_TEXEL) val = texel[i][o->value];
					else val = in[o->value][i];
					/* PixelProgram::clampColor :286-364: UNORM targets only, "if the color attachment is floating-point, no clamping occurs" */
					c[i][ch] = dr->floatTarget ? val : sse_min(sse_max(val, 0.0f), 1.0f);
				}

			/* alphaTest (PixelProgram.cpp:243-259) -> alphaToCoverage (PixelRoutine.cpp:643-658): the clamped c[0].w against the
			 * per-sample thresholds of Renderer.cpp:391-410; then :319-326 */
			if(d->alphaToCoverageEnable)
			{
				static const float a2c4[4] = { 0.2f, 0.4f, 0.6f, 0.8f };
				for(int q = 0; q < ms; q++)
				{
					if(!(d->sampleMask & (1u << q))) continue;
					const float thr = ms == 4 ? a2c4[q] : 0.5f;
					int aMask = 0;
					for(int i = 0; i < 4; i++)
						if(c[i][3] >= thr) aMask |= 1 << i; /* CmpNLT is FCmpOGE: a NaN alpha loses its coverage */
					cMask[q] &= aMask;
					zMask[q] &= cMask[q];
					sMask[q] &= cMask[q];
				}
			}

			/* depth test :494-574 (late) */
			int depthPass = !dr->depthTestActive;
			for(int q = 0; q < ms; q++)
			{
				if(!(d->sampleMask & (1u << q))) continue;
				/* depthBoundsTest :576-641 reads the stored depth BEFORE this fragment's write (both run before writeDepth) */
				int bTest = 0xF;
				if(dr->depthBoundsActive)
				{
					const char *zb = (const char *)d->depth.buffer + (size_t)q * d->depth.sliceB;
					bTest = 0;
					for(int i = 0; i < 4; i++)
					{
						float zValue;
						if(d->depth.format == FMT_D16_UNORM)
							zValue = (float)*(const uint16_t *)(zb + (size_t)(y + (i >> 1)) * d->depth.pitchB + 2 * (size_t)(x + (i & 1))) * (1.0f / 0xFFFF);
						else
							zValue = *(const float *)(zb + (size_t)(y + (i >> 1)) * d->depth.pitchB + 4 * (size_t)(x + (i & 1)));
						if(d->minDepthBounds <= zValue && zValue <= d->maxDepthBounds) bTest |= 1 << i;
					}
					if(!dr->depthTestActive) cMask[q] &= zMask[q] & bTest; /* :634-637 */
				}
				if(dr->depthTestActive)
				{
					char *zb = (char *)d->depth.buffer + (size_t)q * d->depth.sliceB;
					const int d16 = d->depth.format == FMT_D16_UNORM;
					int zTest = 0;
					for(int i = 0; i < 4; i++)
					{
						/* clampDepth :484-492: fixed point, or D32F without VK_EXT_depth_range_unrestricted => [0,1] (PixelProcessor.cpp:121-136) */
						z[q][i] = sse_min(sse_max(z[q][i], dr->minDepthClamp), dr->maxDepthClamp);
						float Z = z[q][i];
						float zValue;
						if(d16)
						{
							/* depthTest :508-511: Z = Min(Max(Round(z * 0xFFFF), 0), 0xFFFF), compared as floats with Float(UShort) (:466-482) */
							Z = sse_min(sse_max(rintf(Z * 65535.0f), 0.0f), 65535.0f);
							zValue = (float)*(uint16_t *)(zb + (size_t)(y + (i >> 1)) * d->depth.pitchB + 2 * (size_t)(x + (i & 1)));
						}
						else
							zValue = *(float *)(zb + (size_t)(y + (i >> 1)) * d->depth.pitchB + 4 * (size_t)(x + (i & 1)));
						int t;
						switch(d->depthCompareOp)
						{
						case CMP_ALWAYS: t = 1; break;
						case CMP_NEVER: t = 0; break;
						case CMP_EQUAL: t = zValue == Z; break;
						case CMP_NOT_EQUAL: t = zValue < Z || zValue > Z; break; /* CmpNEQ is FCmpONE */
						case CMP_LESS: t = zValue > Z; break;            /* CmpNLE is FCmpOGT */
						case CMP_GREATER_OR_EQUAL: t = zValue <= Z; break;
						case CMP_LESS_OR_EQUAL: t = zValue >= Z; break;  /* CmpNLT is FCmpOGE */
						default: t = zValue < Z; break; /* GREATER */
						}
						if(t) zTest |= 1 << i;
					}
					zMask[q] = zTest & cMask[q];
					if(dr->stencilActive) zMask[q] &= sMask[q];
					if(zMask[q]) depthPass = 1;
					zMask[q] &= cMask[q] & bTest; /* :638-641, after depthPass was taken */
				}
			}
			if(depthPass)
			{
				/* writeDepth :660-685 */
				if(dr->depthTestActive && dr->depthWriteEnable)
					for(int q = 0; q < ms; q++)
					{
						if(!(d->sampleMask & (1u << q))) continue;
						char *zb = (char *)d->depth.buffer + (size_t)q * d->depth.sliceB;
						for(int i = 0; i < 4; i++)
							if(zMask[q] & (1 << i))
							{
								if(d->depth.format == FMT_D16_UNORM)
								{
									/* writeDepth16 :687-711: UShort4(Round(z * 0xFFFF), saturate) */
									float r = rintf(z[q][i] * 65535.0f);
									r = r < 0.0f ? 0.0f : (r > 65535.0f ? 65535.0f : r);
									*(uint16_t *)(zb + (size_t)(y + (i >> 1)) * d->depth.pitchB + 2 * (size_t)(x + (i & 1))) = (uint16_t)r;
								}
								else
									*(float *)(zb + (size_t)(y + (i >> 1)) * d->depth.pitchB + 4 * (size_t)(x + (i & 1))) = z[q][i];
							}
					}
				/* blendColor (PixelProgram.cpp:261-284) -> alphaBlend (PixelRoutine.cpp:1653-1961) -> writeColor (:1963-2655) */
				if(dr->colorWriteMask && d->color.buffer)
					for(int q = 0; q < ms; q++)
					{
						if(!(d->sampleMask & (1u << q))) continue;
						uint8_t *cb = (uint8_t *)d->color.buffer + (size_t)q * d->color.sliceB;
						int xMask = dr->depthTestActive ? zMask[q] : cMask[q];
						if(dr->stencilActive) xMask &= sMask[q];
						for(int i = 0; i < 4; i++)
						{
							if(!(xMask & (1 << i))) continue;
							uint8_t *px = cb + (size_t)(y + (i >> 1)) * d->color.pitchB + (size_t)dr->colorBpp * (size_t)(x + (i & 1));
							float out[4];
							if(dr->blendEnable)
							{
								/* readPixel :1111-1130: byte b -> b*257 -> float * (1/65535) */
								float dst[4];
								for(int ch = 0; ch < 4; ch++)
								{
									if(d->color.format == FMT_R32G32B32A32_SFLOAT) { dst[ch] = ((const float *)px)[ch]; continue; }  /* :1700-1710 */
									if(d->color.format == FMT_R16G16B16A16_SFLOAT) { dst[ch] = half_to_float(((const uint16_t *)px)[ch]); continue; } /* :1782-1801 */
									uint8_t b = px[bgr && ch < 3 ? 2 - ch : ch];
									dst[ch] = (float)(uint16_t)(b * 257) * (1.0f / 0xFFFF);
									if(srgb && ch < 3) dst[ch] = srgb_to_linear(dst[ch]); /* :1821-1826 */
								}
								for(int ch = 0; ch < 3; ch++)
								{
									float sF = blend_factor(dr, dr->srcF, ch, c[i], dst);
									float dF = blend_factor(dr, dr->dstF, ch, c[i], dst);
									out[ch] = blend_op(dr->op, c[i][ch], sF, dst[ch], dF);
								}
								{
									float sF = blend_factor(dr, dr->srcFA, 3, c[i], dst);
									float dF = blend_factor(dr, dr->dstFA, 3, c[i], dst);
									out[3] = blend_op(dr->opA, c[i][3], sF, dst[3], dF);
								}
							}
							else memcpy(out, c[i], sizeof(out));
							if(srgb) /* writeColor :1965-1970 */
								for(int ch = 0; ch < 3; ch++) out[ch] = linear_to_srgb(out[ch]);
							for(int ch = 0; ch < 4; ch++)
							{
								if(!((dr->colorWriteMask >> ch) & 1)) continue;
								if(d->color.format == FMT_R32G32B32A32_SFLOAT) { ((float *)px)[ch] = out[ch]; continue; }                 /* :2429-2447: bits stored as they are */
								if(d->color.format == FMT_R16G16B16A16_SFLOAT) { ((uint16_t *)px)[ch] = float_to_half(out[ch]); continue; } /* :2504-2540 */
								float cl = sse_min(sse_max(out[ch], 0.0f), 1.0f);
								int v = round_int(cl * 255.0f);
								v = v < 0 ? 0 : (v > 255 ? 255 : v); /* PackUnsigned saturation */
								px[bgr && ch < 3 ? 2 - ch : ch] = (uint8_t)v;
							}
						}
					}
			}
			/* writeStencil :754-817 (late) */
			if(dr->stencilActive)
			{
				const swcu_stencil_face *sf = prim->clockwise ? &d->front : &d->back;
				int allKeep = d->front.passOp == SOP_KEEP && d->front.depthFailOp == SOP_KEEP && d->front.failOp == SOP_KEEP &&
				              d->back.passOp == SOP_KEEP && d->back.depthFailOp == SOP_KEEP && d->back.failOp == SOP_KEEP;
				int writeEnabled = (d->front.writeMask & 0xFF) != 0 || (d->back.writeMask & 0xFF) != 0;
				if(!allKeep && writeEnabled)
					for(int q = 0; q < ms; q++)
					{
						if(!(d->sampleMask & (1u << q))) continue;
						uint8_t *sb = (uint8_t *)d->stencil.buffer + (size_t)q * d->stencil.sliceB;
						for(int i = 0; i < 4; i++)
						{
							if(!(cMask[q] & (1 << i))) continue;
							uint8_t *p = sb + (size_t)(y + (i >> 1)) * d->stencil.pitchB + x + (i & 1);
							uint8_t v = *p, nv;
							uint8_t ref = (uint8_t)sf->reference;
							/* stencilOperation :819-842; when the depth test is inactive zMask == cMask here */
							int zOk = dr->depthTestActive ? (zMask[q] >> i) & 1 : 1;
							if(!((sMask[q] >> i) & 1)) nv = stencil_op(sf->failOp, v, ref);
							else if(!zOk) nv = stencil_op(sf->depthFailOp, v, ref);
							else nv = stencil_op(sf->passOp, v, ref);
							uint8_t wm = (uint8_t)sf->writeMask;
							*p = (uint8_t)((nv & wm) | (v & ~wm));
						}
					}
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------------
 * Entry: Renderer::draw state gathering (Device/Renderer.cpp:183-490) + DrawCall::run/processVertices/
 * processPrimitives/processPixels (:551-662) + setupSolidTriangles (:733-776), one primitive at a time
 * (batching and clusters only change scheduling: SURVEY §3.1 ordering contract).
 * ---------------------------------------------------------------------------------------------- */
static uint32_t fetch_index(const swcu_draw_desc *d, uint32_t i)
{
	if(d->indexType == 2) return ((const uint16_t *)d->indexBuffer)[i];
	if(d->indexType == 4) return ((const uint32_t *)d->indexBuffer)[i];
	return i;
}

/* setBatchIndices, Renderer.cpp:50-145 */
static void triangle_indices(const swcu_draw_desc *d, uint32_t i, uint32_t idx[3])
{
	int pf = d->provokingVertexMode == 0;
	switch(d->topology)
	{
	case TOPO_TRIANGLE_STRIP:
		idx[0] = fetch_index(d, i + (pf ? 0 : 2));
		idx[1] = fetch_index(d, i + (i & 1) + (pf ? 1 : 0));
		idx[2] = fetch_index(d, i + (~i & 1) + (pf ? 1 : 0));
		break;
	case TOPO_POINT_LIST: /* Renderer.cpp:57-73 + VertexRoutine.cpp:74: the one vertex stands for all three */
		idx[0] = idx[1] = idx[2] = fetch_index(d, i);
		break;
	case TOPO_LINE_LIST: /* Renderer.cpp:74-86 */
		idx[0] = fetch_index(d, 2 * i + (pf ? 0 : 1));
		idx[1] = fetch_index(d, 2 * i + (pf ? 1 : 0));
		idx[2] = fetch_index(d, 2 * i + 1);
		break;
	case TOPO_LINE_STRIP: /* Renderer.cpp:87-99 */
		idx[0] = fetch_index(d, i + (pf ? 0 : 1));
		idx[1] = fetch_index(d, i + (pf ? 1 : 0));
		idx[2] = fetch_index(d, i + 1);
		break;
	case TOPO_TRIANGLE_FAN:
		idx[pf ? 0 : 2] = fetch_index(d, i + 1);
		idx[pf ? 1 : 0] = fetch_index(d, i + 2);
		idx[pf ? 2 : 1] = fetch_index(d, 0);
		break;
	default:
		idx[0] = fetch_index(d, 3 * i + (pf ? 0 : 2));
		idx[1] = fetch_index(d, 3 * i + (pf ? 1 : 0));
		idx[2] = fetch_index(d, 3 * i + (pf ? 2 : 1));
		break;
	}
}

int swref_draw(const swcu_draw_desc *d, const swcu_shader_info *vs, const swcu_shader_info *fs)
{
	if(!d || d->structSize != sizeof(swcu_draw_desc) || !vs || !fs) return SWCU_E_INVALID;
	if(d->sampleCount != 1 && d->sampleCount != 4) return SWCU_E_UNSUPPORTED;
	const int floatTarget = d->color.buffer && (d->color.format == FMT_R32G32B32A32_SFLOAT || d->color.format == FMT_R16G16B16A16_SFLOAT);
	if(d->color.buffer && !floatTarget && d->color.format != FMT_R8G8B8A8_UNORM && d->color.format != FMT_B8G8R8A8_UNORM &&
	   d->color.format != FMT_R8G8B8A8_SRGB && d->color.format != FMT_B8G8R8A8_SRGB) return SWCU_E_UNSUPPORTED;
	if(d->depth.buffer && d->depth.format != FMT_D32_SFLOAT && d->depth.format != FMT_D16_UNORM) return SWCU_E_UNSUPPORTED;
	if(d->depth.buffer && d->depth.format == FMT_D16_UNORM && d->stencil.buffer) return SWCU_E_UNSUPPORTED; /* D16_UNORM_S8_UINT: not in the subset */

	unsigned int csr = _mm_getcsr();
	_mm_setcsr(csr | 0x8040); /* FTZ | DAZ, System/SwiftConfig.cpp:136-139 */

	Draw dr;
	memset(&dr, 0, sizeof(dr));
	dr.d = d; dr.vs = vs; dr.fs = fs;
	{ /* Renderer.cpp:300-331 */
		float W = 0.5f * d->viewportWidth, H = 0.5f * d->viewportHeight;
		float X0 = d->viewportX + W, Y0 = d->viewportY + H;
		dr.WxF = W * 256.0f; dr.HxF = H * 256.0f;
		dr.X0xF = X0 * 256.0f - 128.0f; dr.Y0xF = Y0 * 256.0f - 128.0f;
		dr.depthRange = d->viewportMaxDepth - d->viewportMinDepth;
		dr.depthNear = d->viewportMinDepth;
	}
	{ /* Renderer.cpp:333-345 */
		int x0 = d->renderArea.x, y0 = d->renderArea.y;
		int x1 = x0 + (int)d->renderArea.width, y1 = y0 + (int)d->renderArea.height;
#define CLAMPI(v, lo, hi) ((v) < (lo) ? (lo) : ((v) > (hi) ? (hi) : (v)))
		dr.scissorX0 = CLAMPI(d->scissor.x, x0, x1);
		dr.scissorX1 = CLAMPI(d->scissor.x + (int)d->scissor.width, x0, x1);
		dr.scissorY0 = CLAMPI(d->scissor.y, y0, y1);
		dr.scissorY1 = CLAMPI(d->scissor.y + (int)d->scissor.height, y0, y1);
	}
	dr.prim = d->topology == TOPO_POINT_LIST ? 2 : ((d->topology == TOPO_LINE_LIST || d->topology == TOPO_LINE_STRIP) ? 1 : 0);
	dr.lineWidth = d->lineWidth == 0.0f ? 1.0f : d->lineWidth;
	dr.halfPixelX = 0.5f / (0.5f * d->viewportWidth); dr.halfPixelY = 0.5f / (0.5f * d->viewportHeight);
	dr.minDepthClamp = 0.0f; dr.maxDepthClamp = 1.0f; /* no VK_EXT_depth_range_unrestricted: always clamped */
	if(d->depthClampEnable)
	{
		dr.minDepthClamp = d->viewportMinDepth < d->viewportMaxDepth ? d->viewportMinDepth : d->viewportMaxDepth;
		dr.maxDepthClamp = d->viewportMinDepth < d->viewportMaxDepth ? d->viewportMaxDepth : d->viewportMinDepth;
	}
	dr.ms = (int)d->sampleCount;
	dr.enableMultiSampling = dr.ms > 1;
	dr.depthTestActive = d->depthTestEnable && d->depth.buffer;
	dr.depthBoundsActive = d->depthBoundsTestEnable && d->depth.buffer; /* Context.cpp:946-949 */
	dr.floatTarget = floatTarget;
	dr.colorBpp = d->color.format == FMT_R32G32B32A32_SFLOAT ? 16 : (d->color.format == FMT_R16G16B16A16_SFLOAT ? 8 : 4);
	dr.depthWriteEnable = dr.depthTestActive && d->depthWriteEnable; /* FragmentState::depthWriteActive */
	dr.stencilActive = d->stencilTestEnable && d->stencil.buffer;
	dr.interpolateZ = dr.depthTestActive;
	dr.interpolateW = 1;
	{ /* Context.cpp:1090-1147, :1272-1300 */
		int cop = fold_blend_op((int)d->colorBlendOp, (int)d->srcColorBlendFactor, (int)d->dstColorBlendFactor, !dr.floatTarget);
		int aop = fold_blend_op((int)d->alphaBlendOp, (int)d->srcAlphaBlendFactor, (int)d->dstAlphaBlendFactor, !dr.floatTarget);
		dr.colorWriteMask = d->color.buffer ? (int)(d->colorWriteMask & 0xF) : 0;
		if(cop == BOP_DST_EXT && aop == BOP_DST_EXT) dr.colorWriteMask = 0; /* colorWriteActive, Context.cpp:1304-1308: the stored factors, whether or not blending is enabled */
		dr.blendEnable = d->blendEnable && dr.colorWriteMask && (cop != BOP_SRC_EXT || aop != BOP_SRC_EXT);
		dr.srcF = fold_blend_factor((int)d->colorBlendOp, (int)d->srcColorBlendFactor);
		dr.dstF = fold_blend_factor((int)d->colorBlendOp, (int)d->dstColorBlendFactor);
		dr.srcFA = fold_blend_factor((int)d->alphaBlendOp, (int)d->srcAlphaBlendFactor);
		dr.dstFA = fold_blend_factor((int)d->alphaBlendOp, (int)d->dstAlphaBlendFactor);
		dr.op = cop; dr.opA = aop;
	}

	Primitive *prim = (Primitive *)malloc(sizeof(Primitive) * (size_t)dr.ms);
	if(!prim) { _mm_setcsr(csr); return SWCU_E_NOMEM; }
	memset(prim, 0, sizeof(Primitive) * (size_t)dr.ms);

	for(uint32_t i = 0; i < d->primitiveCount; i++)
	{
		uint32_t idx[3];
		triangle_indices(d, i, idx);
		Vertex v[3];
		for(int k = 0; k < 3; k++) process_vertex(&dr, idx[k], &v[k]);

		Polygon poly;
		if(dr.prim == 1)
		{
			/* DrawCall::setupLine, Renderer.cpp:920-1000: rectangle centred on the segment (the default line rasterization mode);
			 * host C++ there, plain float operations here in the same order */
			const f4 P0 = v[0].position, P1 = v[1].position;
			if(P0.w <= 0 && P1.w <= 0) continue;
			const float W = dr.WxF * (1.0f / 256.0f), H = dr.HxF * (1.0f / 256.0f);
			float dx = W * (P1.x / P1.w - P0.x / P0.w);
			float dy = H * (P1.y / P1.w - P0.y / P0.w);
			if(dx == 0 && dy == 0) continue;
			float scale = LINE_SCALE(dr.lineWidth, dx, dy);
			dx *= scale; dy *= scale;
			float dx0h = dx * P0.w / H, dy0w = dy * P0.w / W;
			float dx1h = dx * P1.w / H, dy1w = dy * P1.w / W;
			poly.P[0] = P0; poly.P[1] = P1; poly.P[2] = P1; poly.P[3] = P0;
			poly.P[0].x += -dy0w; poly.P[0].y += +dx0h;
			poly.P[1].x += -dy1w; poly.P[1].y += +dx1h;
			poly.P[2].x += +dy1w; poly.P[2].y += -dx1h;
			poly.P[3].x += +dy0w; poly.P[3].y += -dx0h;
			poly.n = 4; poly.i = 0;
			if(!clip_polygon(&poly, d->depthClipEnable ? CLIP_FRUSTUM : CLIP_SIDES)) continue;
		}
		else if(dr.prim == 2)
		{
			/* DrawCall::setupPoint, Renderer.cpp:1137-1185: a square of gl_PointSize pixels around the vertex */
			const float pSize = v[0].pointSize < 1.0f ? 1.0f : (v[0].pointSize > 1023.0f ? 1023.0f : v[0].pointSize); /* clamp(): NaN passes through */
			const float X = pSize * v[0].position.w * dr.halfPixelX, Y = pSize * v[0].position.w * dr.halfPixelY;
			for(int k = 0; k < 4; k++) poly.P[k] = v[0].position;
			poly.P[0].x -= X; poly.P[0].y += Y;
			poly.P[1].x += X; poly.P[1].y += Y;
			poly.P[2].x += X; poly.P[2].y -= Y;
			poly.P[3].x -= X; poly.P[3].y -= Y;
			poly.n = 4; poly.i = 0;
			if(!clip_polygon(&poly, d->depthClipEnable ? CLIP_FRUSTUM : CLIP_SIDES)) continue;
		}
		else
		{
		/* setupSolidTriangles, Renderer.cpp:733-776 */
		if((v[0].clipFlags & v[1].clipFlags & v[2].clipFlags) != CLIP_FINITE) continue;
		poly.P[0] = v[0].position; poly.P[1] = v[1].position; poly.P[2] = v[2].position;
		poly.n = 3; poly.i = 0;
		int flagsOr = v[0].clipFlags | v[1].clipFlags | v[2].clipFlags;
		if(flagsOr != CLIP_FINITE)
			if(!clip_polygon(&poly, flagsOr)) continue;
		}
		if(!setup_triangle(&dr, prim, &v[0], &v[1], &v[2], &poly)) continue;
		for(int q = 1; q < dr.ms; q++) /* planes live in the first Primitive only; copy what rasterize() reads */
		{
			prim[q].yMin = prim[0].yMin; prim[q].yMax = prim[0].yMax;
		}
		rasterize(&dr, prim);
	}
	free(prim);
	_mm_setcsr(csr);
	return SWCU_OK;
}

/* Blitter::fastClear, Device/Blitter.cpp:170-325 — rectangle fill of every sample slice */
int swref_clear(const swcu_attachment *att, uint32_t samples, const swcu_rect *area, const void *value)
{
	int bpp = att->format == FMT_S8_UINT ? 1 : (att->format == FMT_D16_UNORM ? 2 : (att->format == FMT_R32G32B32A32_SFLOAT ? 16 : (att->format == FMT_R16G16B16A16_SFLOAT ? 8 : 4)));
	for(uint32_t q = 0; q < samples; q++)
		for(uint32_t y = 0; y < area->height; y++)
		{
			uint8_t *row = (uint8_t *)att->buffer + (size_t)q * att->sliceB + (size_t)(area->y + (int)y) * att->pitchB + (size_t)area->x * bpp;
			for(uint32_t x = 0; x < area->width; x++) memcpy(row + (size_t)x * bpp, value, (size_t)bpp);
		}
	return SWCU_OK;
}

/* Blitter::fastResolve, Device/Blitter.cpp:2079-2205 — RGBA8 4x: avg(avg(s0,s1),avg(s2,s3)), avg = pavgb = (a+b+1)>>1 */
int swref_resolve(const swcu_attachment *src, uint32_t samples, const swcu_attachment *dst)
{
	if(samples != 4) return SWCU_E_UNSUPPORTED;
	if(src->format == FMT_R8G8B8A8_SRGB || src->format == FMT_B8G8R8A8_SRGB || src->format == FMT_R16G16B16A16_SFLOAT || src->format == FMT_R32G32B32A32_SFLOAT)
	{
		/* Not Blitter::fastResolve (its format switch, Blitter.cpp:2142-2200, knows RGBA8 / BGRA8 UNORM only): Blitter::resolve falls
		 * back to the generic blit (:2053-2071), whose routine reads every sample as floats (readFloat4 :368-452), brings sRGB samples
		 * to linear light (ApplyScaleAndClamp :1459-1465: * 1/255, sRGBtoLinear on rgb, * 255), adds them up in sample order, scales by
		 * 1 / samples (:1524-1545), takes the result back through the same function (pre-scaled: * 1/255, linearToSRGB, * 255) and
		 * writes it (:640-745: RoundShort4 + unsigned saturation for the 8-bit formats, Reactor's Half() for R16G16B16A16_SFLOAT). */
		unsigned int csr = _mm_getcsr();
		_mm_setcsr(csr | 0x8040);
		const int bpp = src->format == FMT_R32G32B32A32_SFLOAT ? 16 : (src->format == FMT_R16G16B16A16_SFLOAT ? 8 : 4);
		const int srgb = bpp == 4;
		for(uint32_t y = 0; y < dst->height; y++)
			for(uint32_t x = 0; x < dst->width; x++)
			{
				float accum[4] = { 0, 0, 0, 0 };
				for(int q = 0; q < 4; q++)
				{
					const uint8_t *s = (const uint8_t *)src->buffer + (size_t)q * src->sliceB + (size_t)y * src->pitchB + (size_t)x * bpp;
					for(int ch = 0; ch < 4; ch++)
					{
						float c;
						if(bpp == 16) c = ((const float *)s)[ch];
						else if(bpp == 8) c = half_to_float(((const uint16_t *)s)[ch]);
						else
						{
							c = (float)s[ch] * (1.0f / 255.0f);
							if(ch < 3) c = srgb_to_linear(c); /* (byte order does not matter: the three colour bytes are treated alike) */
							c = c * 255.0f;
						}
						accum[ch] = q == 0 ? c : accum[ch] + c;
					}
				}
				uint8_t *t = (uint8_t *)dst->buffer + (size_t)y * dst->pitchB + (size_t)x * bpp;
				for(int ch = 0; ch < 4; ch++)
				{
					float c = accum[ch] * 0.25f;
					if(bpp == 16) ((float *)t)[ch] = c;
					else if(bpp == 8) ((uint16_t *)t)[ch] = float_to_half(c);
					else
					{
						c = c * (1.0f / 255.0f);
						if(ch < 3) c = linear_to_srgb(c);
						c = c * 255.0f;
						int i = round_int(c); /* RoundShort4: cvtps2dq, packssdw; then packuswb */
						i = i < -32768 ? -32768 : (i > 32767 ? 32767 : i);
						t[ch] = (uint8_t)(i < 0 ? 0 : (i > 255 ? 255 : i));
					}
				}
			}
		_mm_setcsr(csr);
		(void)srgb;
		return SWCU_OK;
	}
	for(uint32_t y = 0; y < dst->height; y++)
		for(uint32_t x = 0; x < dst->width * 4; x++)
		{
			const uint8_t *s = (const uint8_t *)src->buffer + (size_t)y * src->pitchB + x;
			unsigned a = (s[0] + s[(size_t)src->sliceB] + 1u) >> 1;
			unsigned b = (s[2 * (size_t)src->sliceB] + s[3 * (size_t)src->sliceB] + 1u) >> 1;
			((uint8_t *)dst->buffer)[(size_t)y * dst->pitchB + x] = (uint8_t)((a + b + 1u) >> 1);
		}
	return SWCU_OK;
}

const char *swref_version(void) { return "swref oracle 1 (restates google/swiftshader @7868bf37 draw path; test infrastructure only)"; }
