// kernels.cuh — the sm_100a kernels of the draw path.
//
//   k_cull    band mode only: marks the triangles whose rows miss the render area (a rank of a multi-GPU frame).
//   k_setup   one thread per triangle: index fetch (Renderer.cpp:50-145), vertex stage + clip flags + projection
//             (VertexRoutine.cpp:116-154,570-610), trivial reject / Sutherland-Hodgman clip (Renderer.cpp:733-776,
//             Clipper.cpp), cull / row range / plane equations (SetupRoutine.cpp:36-548) and — for small triangles —
//             the per-row spans (SetupRoutine.cpp:550-621); writes one record + one packed tile rectangle per triangle.
//   k_big     one CTA per large triangle: spans in closed form, one (row, sample) per thread, and its tile pairs.
//   k_emit    (tile, triangle) pairs of the small triangles; sorted by tile with a stable radix sort => per-tile
//             triangle lists in API order (the ordering contract of Renderer.cpp:573-576,652-661).
//   k_tile    one CTA per 32x16 screen tile, one warp per 16x8 region: stages colour/depth/stencil of the tile in shared
//             memory (TMA), walks the tile's triangle list in order, turns the spans that cross the region into
//             (candidate, row, sample) runs of covered pixels (QuadRasterizer coverage) and consumes the covered samples 32 at
//             a time, one per lane, through PixelRoutine::quad (interpolation, shader routing, sampler, depth/stencil
//             test, blend, format write); the tile goes back with TMA stores.
//   k_clear, k_resolve4, k_copy_rows, k_signal, k_wait_flags   the steps either side of the draw and the multi-GPU delivery.
//
// Float discipline: compiled with -fmad=false -ftz=true -prec-div=true -prec-sqrt=true; every
// product/sum is a single rounded op and __fmaf_rn appears only where the reference writes MulAdd().
#pragma once

#include "swcu_internal.h"
#include "../../include/swcu_srgb_lut.h"

#include <cuda.h>
#include <cuda_runtime.h>

#define DEVI __device__ __forceinline__

// ------------------------------------------------------------------------------------------------------------------
// small helpers (Reactor semantics on x86: LLVMReactor.cpp:135-138,2694-2703)
// ------------------------------------------------------------------------------------------------------------------
DEVI int round_int(float x) { return !(x < 2147483648.0f) ? (int)0x80000000 : __float2int_rn(x); } // cvtps2dq: NaN / overflow -> "integer indefinite"
DEVI int trunc_int(float x) { return !(x < 2147483648.0f) ? (int)0x80000000 : __float2int_rz(x); } // cvttps2dq
DEVI int round_int_clamped(float x) { float c = x < 2147483520.0f ? x : 2147483520.0f; return round_int(c); }
DEVI float sse_max(float a, float b) { return a > b ? a : b; } // maxps: second operand unless strictly greater
DEVI float sse_min(float a, float b) { return a < b ? a : b; }
DEVI float fmul(float a, float b) { return __fmul_rn(a, b); }
DEVI float fadd(float a, float b) { return __fadd_rn(a, b); }
DEVI float fsub(float a, float b) { return __fsub_rn(a, b); }
DEVI float fdiv(float a, float b) { return __fdiv_rn(a, b); }
DEVI float frcp(float a) { return __frcp_rn(a); } // 1.0f / a, correctly rounded: the same value as fdiv(1.0f, a) in fewer instructions
DEVI int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__constant__ float c_SampleX[4] = { 0.375f - 0.5f, 0.875f - 0.5f, 0.125f - 0.5f, 0.625f - 0.5f }; // Constants.cpp:291-297
__constant__ float c_SampleY[4] = { 0.125f - 0.5f, 0.375f - 0.5f, 0.625f - 0.5f, 0.875f - 0.5f };
__constant__ int c_Xf[4] = { -32, 96, -96, 32 }; // Constants.hpp:26-52 (8-bit sub-pixel precision)
__constant__ int c_Yf[4] = { -96, -32, 32, 96 };

// ------------------------------------------------------------------------------------------------------------------
// vertex stage
// ------------------------------------------------------------------------------------------------------------------
DEVI uint32_t fetch_index(const DrawConst &d, uint32_t i)
{
	if(d.indexType == 2) return ((const uint16_t *)d.indexBuffer)[i];
	if(d.indexType == 4) return ((const uint32_t *)d.indexBuffer)[i];
	return i;
}

// setBatchIndices, Renderer.cpp:50-145 (triangle list / strip / fan, provoking-vertex rotation)
DEVI void triangle_indices(const DrawConst &d, uint32_t i, uint32_t idx[3])
{
	const uint32_t pf = d.provokingFirst;
	if(d.topology == TOPO_TRIANGLE_STRIP)
	{
		idx[0] = fetch_index(d, i + (pf ? 0 : 2));
		idx[1] = fetch_index(d, i + (i & 1) + (pf ? 1 : 0));
		idx[2] = fetch_index(d, i + (~i & 1) + (pf ? 1 : 0));
	}
	else if(d.topology == TOPO_POINT_LIST) // Renderer.cpp:57-73; the one vertex stands for all three (VertexRoutine.cpp:74)
	{
		idx[0] = idx[1] = idx[2] = fetch_index(d, i);
	}
	else if(d.topology == TOPO_LINE_LIST || d.topology == TOPO_LINE_STRIP) // Renderer.cpp:74-99
	{
		const uint32_t base = d.topology == TOPO_LINE_LIST ? 2 * i : i;
		idx[0] = fetch_index(d, base + (pf ? 0 : 1));
		idx[1] = fetch_index(d, base + (pf ? 1 : 0));
		idx[2] = fetch_index(d, base + 1);
	}
	else if(d.topology == TOPO_TRIANGLE_FAN)
	{
		uint32_t a = fetch_index(d, i + 1), b = fetch_index(d, i + 2), c = fetch_index(d, 0);
		if(pf) { idx[0] = a; idx[1] = b; idx[2] = c; }
		else { idx[2] = a; idx[0] = b; idx[1] = c; }
	}
	else
	{
		idx[0] = fetch_index(d, 3 * i + (pf ? 0 : 2));
		idx[1] = fetch_index(d, 3 * i + (pf ? 1 : 0));
		idx[2] = fetch_index(d, 3 * i + (pf ? 2 : 1));
	}
}

// one scalar of the vertex stage: shader constant / format default, or a component of an attribute stream with the
// robustBufferAccess clamp (VertexRoutine::readStream, VertexRoutine.cpp:173-245; offsets wrap in 32 bits like the reference)
DEVI float vs_operand(const DrawConst &d, const KVSrc &src, uint32_t index)
{
	// constant or stream is the same for every thread (a uniform branch); the robustness check does not branch around the
	// load: the offset is clamped into the buffer and an out-of-range fetch is zeroed afterwards, so all attribute fetches of
	// a triangle are issued back to back
	if(src.ptr == nullptr) return src.constant;
	const uint32_t offset = (index + (uint32_t)d.baseVertex) * src.stride;
	const float v = __ldg((const float *)(src.ptr + min(offset, src.limit)));
	return offset <= src.limit ? v : 0.0f;
}

// The vertex stage's arithmetic (an MVP transform): the translator's scalar steps, run once per vertex.  MUL / ADD / SUB
// are single IEEE operations, FMA is Reactor's MulAdd — llvm.fmuladd, fused on every host the reference runs on that has FMA units
// (SpirvShaderArithmetic.cpp:39-55, LLVMReactor.cpp:4425-4429) — NEG flips the sign bit (LLVM fneg).
DEVI void vs_run(const DrawConst &d, uint32_t index, float *t)
{
	for(uint32_t i = 0; i < d.vsProgLen; i++)
	{
		const KVsStep &st = d.vsProg[i];
		auto get = [&](const KVsOperand &o) -> float {
			if(o.kind == VK_CONST) return __uint_as_float(o.value);
			if(o.kind == VK_TEMP) return t[o.value];
			return vs_operand(d, d.vsIn[o.value], index);
		};
		const float a = get(st.a);
		float r;
		switch(st.op)
		{
		case SWCU_OP_MUL: r = fmul(a, get(st.b)); break;
		case SWCU_OP_ADD: r = fadd(a, get(st.b)); break;
		case SWCU_OP_SUB: r = fsub(a, get(st.b)); break;
		case SWCU_OP_FMA: r = __fmaf_rn(a, get(st.b), get(st.c)); break;
		default: r = __uint_as_float(__float_as_uint(a) ^ 0x80000000u); break;
		}
		t[i] = r;
	}
}

struct VOut
{
	float px, py, pz, pw; // clip-space position
	int flags;
	int X, Y;      // projected, 24.8
	float zp, rhw; // projected.z, projected.w
};

DEVI void process_vertex(const DrawConst &d, float px, float py, float pz, float pw, VOut &v)
{
	v.px = px; v.py = py; v.pz = pz; v.pw = pw;
	// computeClipFlags, VertexRoutine.cpp:128-152.  Reactor's CmpNLE is an ORDERED greater-than (FCmpOGT, LLVMReactor.cpp:4491-4495):
	// a NaN w sets no flag
	int f = 0;
	if(pw < px) f |= CLIP_RIGHT;
	if(pw < py) f |= CLIP_TOP;
	if(-pw > px) f |= CLIP_LEFT;
	if(-pw > py) f |= CLIP_BOTTOM;
	if(d.depthClipEnable)
	{
		if(pw < pz) f |= CLIP_FAR;
		if(0.0f > pz) f |= CLIP_NEAR;
	}
	if(fabsf(px) <= 3.40282347e38f && fabsf(py) <= 3.40282347e38f && fabsf(pz) <= 3.40282347e38f) f |= CLIP_FINITE;
	v.flags = f;
	uint32_t wb = __float_as_uint(pw); // VertexRoutine.cpp:599-606
	if(pw == 0.0f) wb |= 0x3F800000u;
	const float w = __uint_as_float(wb);
	const float rhw = fdiv(1.0f, w);
	v.X = round_int_clamped(fadd(d.X0xF, fmul(fmul(px, rhw), d.WxF)));
	v.Y = round_int_clamped(fadd(d.Y0xF, fmul(fmul(py, rhw), d.HxF)));
	v.zp = fmul(pz, rhw);
	v.rhw = rhw;
}

// ------------------------------------------------------------------------------------------------------------------
// clipper (Clipper.cpp:22-30 clipEdge, :32-265 planes, :271-299 Clip)
// ------------------------------------------------------------------------------------------------------------------
DEVI float4 clip_edge(float4 Vi, float4 Vj, float di, float dj)
{
	const float D = fdiv(1.0f, fsub(dj, di));
	float4 o;
	o.x = fmul(fsub(fmul(dj, Vi.x), fmul(di, Vj.x)), D);
	o.y = fmul(fsub(fmul(dj, Vi.y), fmul(di, Vj.y)), D);
	o.z = fmul(fsub(fmul(dj, Vi.z), fmul(di, Vj.z)), D);
	o.w = fmul(fsub(fmul(dj, Vi.w), fmul(di, Vj.w)), D);
	return o;
}

DEVI float plane_dist(int plane, float4 v)
{
	switch(plane)
	{
	case CLIP_NEAR: return v.z;
	case CLIP_FAR: return fsub(v.w, v.z);
	case CLIP_LEFT: return fadd(v.w, v.x);
	case CLIP_RIGHT: return fsub(v.w, v.x);
	case CLIP_TOP: return fsub(v.w, v.y);
	default: return fadd(v.w, v.y);
	}
}

// returns the vertex count (0 if clipped away); P has room for 16
__device__ __noinline__ int clip_polygon(float4 *P, int n, int flagsOr)
{
	const int order[6] = { CLIP_NEAR, CLIP_FAR, CLIP_LEFT, CLIP_RIGHT, CLIP_TOP, CLIP_BOTTOM };
	float4 T[16];
	for(int k = 0; k < 6; k++)
	{
		if(n < 3) break;
		if(!(flagsOr & order[k])) continue;
		int t = 0;
		for(int i = 0; i < n; i++)
		{
			int j = i == n - 1 ? 0 : i + 1;
			float di = plane_dist(order[k], P[i]);
			float dj = plane_dist(order[k], P[j]);
			if(di >= 0)
			{
				T[t++] = P[i];
				if(dj < 0) T[t++] = clip_edge(P[i], P[j], di, dj);
			}
			else if(dj > 0) T[t++] = clip_edge(P[j], P[i], dj, di);
		}
		for(int i = 0; i < t; i++) P[i] = T[i];
		n = t;
	}
	return n >= 3 ? n : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// spans
// ------------------------------------------------------------------------------------------------------------------
// SetupRoutine::edge (SetupRoutine.cpp:550-621), row-stepping form for triangles of a few rows.
//
// The reference runs the DDA once per (edge, sample) with two integer divisions each.  Here the per-edge step
// (Q, R) = floor-divmod(DX, DY) is computed once for all samples (the sample offset moves both end points, so DX, DY do
// not change), and the divisions are done as a float estimate fixed up with the exact integer remainder — the result is
// the exact quotient for every int32 input (the fix-up loops run 0 or 1 times for screen-sized operands).
#define SETUP_THREADS 128
#ifndef SETUP_BLOCKS_1X
#define SETUP_BLOCKS_1X 7 // resident CTAs per SM asked of the 1x instantiation of k_setup (register cap 72)
#endif
DEVI int floor_div(int n, int d, int &rem) // d > 0; returns floor(n / d), rem = n - q*d in [0, d)
{
	int q = __float2int_rd(__fdividef((float)n, (float)d));
	int r = n - q * d;
	while(r < 0) { r += d; q--; }
	while(r >= d) { r -= d; q++; }
	rem = r;
	return q;
}

// One edge a->b of a small triangle / clipped polygon, all samples.  Writes the clamped x of every row of the edge inside
// [rowMin, rowMin + SWCU_SMALL_ROWS) into the left or right half of the span entries; `rows` is this thread's column of
// the shared scratch: entry e lives at rows[e * SETUP_THREADS].
// Operands of an edge for which the closed form of edge_at_row equals the reference's 32-bit arithmetic: screen-sized coordinates.
// Anything else (a vertex that projected to INT_MIN because its w was -Inf or NaN, ...) goes through edge_wrapped(); k_setup
// sends every polygon with such a coordinate to k_big.
DEVI bool polygon_insane(int minX, int maxX, int minY, int maxY)
{
	const int lim = (1 << 21) + 4096;
	return minX < -lim || maxX > lim || minY < -lim || maxY > lim;
}
DEVI bool edge_is_sane(int DX, int DY) { return DY > 0 && DY < (1 << 22) && DX > -(1 << 22) && DX < (1 << 22); }

// x of the edge at row y exactly as SetupRoutine::edge (SetupRoutine.cpp:550-621) computes it in wrapping 32-bit integers,
// for ANY operands: the reference's set-up values (x0, d0, Q, R) are formed with the same wrapping operations, then its
// row-by-row stepping (d += R; carry into x when d > 0) is closed over k = y - y1 rows in 64 bits — the step count itself can
// be millions when a coordinate is garbage.  Returns false if the edge does not own the row.
DEVI bool edge_wrapped(const DrawConst &d, int Xa, int Ya, int Xb, int Yb, int y, bool &right, int &xo)
{
	if(Ya == Yb) return false;
	const bool swap = Yb < Ya;
	const int X1 = swap ? Xb : Xa, X2 = swap ? Xa : Xb;
	const int Y1 = swap ? Yb : Ya, Y2 = swap ? Ya : Yb;
	const int y1 = (int)((uint32_t)Y1 + 255u) >> 8, y2 = (int)((uint32_t)Y2 + 255u) >> 8;
	if(y < max(y1, d.scY0) || y >= min(y2, d.scY1)) return false;
	const uint32_t DX12 = (uint32_t)X2 - (uint32_t)X1, DY12 = (uint32_t)Y2 - (uint32_t)Y1;
	const int FDX12 = (int)(DX12 << 8), FDY12 = (int)(DY12 << 8);
	if(FDY12 <= 0) return false; // the reference divides by a non-positive value here (undefined); nothing is drawn for the edge
	const int X = (int)(DX12 * (((uint32_t)y1 << 8) - (uint32_t)Y1) + (uint32_t)(X1 & 255) * DY12);
	int x0 = (int)((uint32_t)(X1 >> 8) + (uint32_t)(X / FDY12));
	int d0 = X % FDY12;
	if(d0 > 0) { x0 = (int)((uint32_t)x0 + 1u); d0 -= FDY12; } // ceiling: remainder in (-D, 0]
	int Q = FDX12 / FDY12, R = FDX12 % FDY12;
	if(R < 0) { Q -= 1; R += FDY12; }                          // flooring: remainder in [0, D)
	const long long k = (long long)y - (long long)y1;
	const long long total = (long long)d0 + k * (long long)R;
	const long long carries = total > 0 ? (total + FDY12 - 1) / FDY12 : 0;
	const int x = (int)(uint32_t)((unsigned long long)(long long)x0 + (unsigned long long)(k * (long long)Q) + (unsigned long long)carries);
	xo = clampi(x, d.scX0, d.scX1);
	right = swap;
	return true;
}

template<int MS>
DEVI void edge_small(const DrawConst &d, uint32_t *rows, int rowMin, int Xa, int Ya, int Xb, int Yb)
{
	if(Ya == Yb) return;
	const bool swap = Yb < Ya;
	const int X1 = swap ? Xb : Xa, X2 = swap ? Xa : Xb;
	const int Y1 = swap ? Yb : Ya, Y2 = swap ? Ya : Yb;
	const int DX = (int)((uint32_t)X2 - (uint32_t)X1), DY = (int)((uint32_t)Y2 - (uint32_t)Y1), FDY = DY << 8;
	int R;
	const int Q = floor_div(DX, DY, R); // == floor-divmod(DX << 8, DY << 8) with the remainder scaled by 256
	R <<= 8;
	unsigned short *half = (unsigned short *)rows + (swap ? 1 : 0);
#pragma unroll
	for(int q = 0; q < MS; q++)
	{
		const int X1q = X1 - (MS > 1 ? c_Xf[q] : 0), Y1q = Y1 - (MS > 1 ? c_Yf[q] : 0), Y2q = Y2 - (MS > 1 ? c_Yf[q] : 0);
		const int y1 = (Y1q + 255) >> 8, y2 = (Y2q + 255) >> 8;
		const int yMin = max(y1, d.suY0), yMax = min(y2, d.suY1);
		if(!(yMin < yMax)) continue;
		// x(y1) = (X1 >> 8) + ceil(N / FDY), dd = N - ceil * FDY in (-FDY, 0]
		const int N = DX * ((y1 << 8) - Y1q) + (X1q & 255) * DY;
		int dd;
		int x = (X1q >> 8) + floor_div(N, FDY, dd);
		if(dd > 0) { x++; dd -= FDY; }
		for(int y = y1; y < yMax; y++)
		{
			if(y >= yMin) half[2 * (((y - rowMin) * MS + q) * SETUP_THREADS)] = (unsigned short)clampi(x, d.scX0, d.scX1);
			x += Q;
			dd += R;
			if(dd > 0) { dd -= FDY; x++; }
		}
	}
}

// The same edge in closed form (SURVEY §9.1 "edges"): x(y) = (X1>>8) + ceil((DX*(256y - Y1) + (X1&255)*DY) / (256*DY)).
// Returns true if the edge owns row y; *right tells which half it writes.
DEVI bool edge_at_row(const DrawConst &d, int Xa, int Ya, int Xb, int Yb, int y, bool &right, int &xo)
{
	if(Ya == Yb) return false;
	if(!edge_is_sane((int)((uint32_t)Xb - (uint32_t)Xa), Ya < Yb ? (int)((uint32_t)Yb - (uint32_t)Ya) : (int)((uint32_t)Ya - (uint32_t)Yb)))
		return edge_wrapped(d, Xa, Ya, Xb, Yb, y, right, xo);
	const bool swap = Yb < Ya;
	const int X1 = swap ? Xb : Xa, X2 = swap ? Xa : Xb;
	const int Y1 = swap ? Yb : Ya, Y2 = swap ? Ya : Yb;
	const int y1 = (Y1 + 255) >> 8, y2 = (Y2 + 255) >> 8;
	if(y < max(y1, d.scY0) || y >= min(y2, d.scY1)) return false;
	const long long DX = X2 - X1, DY = Y2 - Y1;
	const long long N = DX * (256ll * y - Y1) + (long long)(X1 & 255) * DY;
	const long long D = 256ll * DY;
	// ceil(N / D), exactly: a double-precision estimate of the quotient (|N| < 2^46, 0 < D < 2^31, so it is off by one at most),
	// put right with the integer remainder — a 64-bit integer division costs several times as many instructions
	long long qv = __double2ll_rd(__ddiv_rn((double)N, (double)D));
	long long rem = N - qv * D;
	while(rem < 0) { rem += D; qv -= 1; }
	while(rem >= D) { rem -= D; qv += 1; }
	if(rem > 0) qv += 1; // ceiling
	xo = clampi((X1 >> 8) + (int)qv, d.scX0, d.scX1);
	right = swap;
	return true;
}

// `lineWidth * 0.5f / sqrt(dx * dx + dy * dy)` of DrawCall::setupLine (Renderer.cpp:965)
#ifndef SWCU_LINE_SQRT_DOUBLE
#define SWCU_LINE_SQRT_DOUBLE 0
#endif
DEVI float line_scale(float lineWidth, float dx, float dy)
{
	const float len2 = fadd(fmul(dx, dx), fmul(dy, dy));
#if SWCU_LINE_SQRT_DOUBLE
	return __double2float_rn(__ddiv_rn((double)fmul(lineWidth, 0.5f), __dsqrt_rn((double)len2)));
#else
	return fdiv(fmul(lineWidth, 0.5f), __fsqrt_rn(len2));
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// k_setup
// ------------------------------------------------------------------------------------------------------------------
DEVI void rot1(bool c, int &a, int &b, int &e) { if(c) { int t = a; a = b; b = e; e = t; } }
DEVI void rot2(bool c, int &a, int &b, int &e) { if(c) { int t = e; e = b; b = a; a = t; } }

// warp-aggregated bump allocation: one atomic per warp instead of one per triangle (all 32 lanes must call)
DEVI unsigned long long warp_alloc(unsigned long long *cursor, uint32_t count)
{
	const int lane = threadIdx.x & 31;
	uint32_t incl = count;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1)
	{
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if(lane >= o) incl += t;
	}
	const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
	unsigned long long base = 0;
	if(lane == 31 && total) base = atomicAdd(cursor, (unsigned long long)total);
	base = __shfl_sync(0xFFFFFFFFu, base, 31);
	return base + (incl - count);
}

// Approximate band / scissor-row reject shared by k_cull and k_setup: projected y with a fast reciprocal and a 4-pixel margin.
// With all three w > 0 the clipped polygon stays inside the hull of the projected vertices, so a triangle whose hull misses
// the scissor rows cannot produce a span.  Returns true if the triangle is certainly invisible.
DEVI bool rows_missed(const DrawConst &d, float y0, float w0, float y1, float w1, float y2, float w2)
{
	if(!(w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) return false;
	const float y0a = __fmaf_rn(__fdividef(y0, w0), d.HxF, d.Y0xF);
	const float y1a = __fmaf_rn(__fdividef(y1, w1), d.HxF, d.Y0xF);
	const float y2a = __fmaf_rn(__fdividef(y2, w2), d.HxF, d.Y0xF);
	const float ylo = fminf(fminf(y0a, y1a), y2a), yhi = fmaxf(fmaxf(y0a, y1a), y2a);
	return yhi + 1024.0f < (float)(d.suY0 << 8) || ylo - 1024.0f > (float)(d.suY1 << 8);
}

// Band mode (the render area covers only part of the framebuffer rows — a rank of a multi-GPU frame): a light first pass,
// one thread per triangle at full occupancy, that only fetches y and w of the three vertices and marks the triangles whose
// rows miss the band; k_setup then drops them on one coalesced byte load instead of a dependent index -> vertex fetch chain.
__global__ void __launch_bounds__(256) k_cull(const __grid_constant__ DrawConst d, unsigned char *flags)
{
	const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
	if(tri >= d.primCount) return;
	uint32_t idx[3];
	triangle_indices(d, tri, idx);
	float y[3], w[3];
#pragma unroll
	for(int a = 0; a < 3; a++)
	{
		y[a] = vs_operand(d, d.vsPos[1], idx[a]);
		w[a] = vs_operand(d, d.vsPos[3], idx[a]);
	}
	// (a draw whose vertex stage has arithmetic does not run this pass: the host leaves cullFlags null, k_setup_prog rejects on its own)
	flags[tri] = rows_missed(d, y[0], w[0], y[1], w[1], y[2], w[2]) ? 0 : 1;
}

DEVI uint32_t region_bin(const DrawConst &d, int rx, int ry) { return (uint32_t)(((ry >> 1) * d.tilesX + (rx >> 1)) * 4 + (ry & 1) * 2 + (rx & 1)); }

// One add per distinct bin of the warp instead of one per lane: neighbouring triangles of a mesh land in the same few region bins,
// and the L2 atomic unit serialises same-address operations (all 32 lanes must call).
DEVI void warp_bin_count(uint32_t *binCount, bool has, uint32_t bin)
{
	const int lane = threadIdx.x & 31;
	const uint32_t peers = __match_any_sync(0xFFFFFFFFu, has ? bin : (0x80000000u | (uint32_t)lane));
	if(has && lane == __ffs(peers) - 1) atomicAdd(binCount + bin, (uint32_t)__popc(peers));
}
// group: the rank whose band holds row y, and the count of bin `bin` AT that rank (a peer's memory over NVLink, or my own)
DEVI int band_of(const DrawConst &d, int y) { return min(y / d.bandRows, (int)d.world - 1); }
DEVI void warp_bin_count_at(const DrawConst &d, bool has, int owner, uint32_t bin)
{
	const int lane = threadIdx.x & 31;
	const uint32_t peers = __match_any_sync(0xFFFFFFFFu, has ? (bin | ((uint32_t)owner << 24)) : (0x80000000u | (uint32_t)lane));
	if(has && lane == __ffs(peers) - 1) atomicAdd(d.peerBinCount[owner] + bin, (uint32_t)__popc(peers));
}
// ... and the slots of the fill pass: the bin's count is taken down by the number of peers, every peer gets one of the slots
DEVI uint32_t warp_bin_slot(uint32_t *binCount, bool has, uint32_t bin)
{
	const int lane = threadIdx.x & 31;
	const uint32_t peers = __match_any_sync(0xFFFFFFFFu, has ? bin : (0x80000000u | (uint32_t)lane));
	const int leader = __ffs(peers) - 1;
	const uint32_t cnt = (uint32_t)__popc(peers);
	uint32_t old = 0;
	if(has && lane == leader) old = atomicSub(binCount + bin, cnt);
	old = __shfl_sync(0xFFFFFFFFu, old, leader);
	return old - cnt + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

DEVI int sel3(int i, int a0, int a1, int a2) { return i == 0 ? a0 : (i == 1 ? a1 : a2); }
DEVI float sel3(int i, float a0, float a1, float a2) { return i == 0 ? a0 : (i == 1 ? a1 : a2); }

// The unclipped triangle lives in registers only (three named vertices, selects instead of indexed arrays); the clipped
// polygon — rare — goes through local-memory arrays.
// MSC = 1: the 1x instantiation (sample count folded, 7 CTAs / SM); MSC = 0: sample count read at run time (used for 4x — measured
// faster than a folded 4x instantiation, whose natural register allocation costs a resident CTA)
// PROG: the vertex stage has arithmetic (DrawConst::vsProg): its steps run per vertex ahead of everything else; position components
// and slot sources may then come from a step's result.  A separate instantiation, so that the routing-only draws keep their registers.
template<int MSC, bool PROG = false>
DEVI void setup_triangle(const DrawConst &d)
{
	extern __shared__ uint32_t s_rows[]; // [SWCU_SMALL_ROWS * ms][SETUP_THREADS]: one scratch column of span rows per thread
	const uint32_t tri = d.triLo + blockIdx.x * blockDim.x + threadIdx.x; // my share of the draw; the grid is padded to whole warps
	const bool live = tri < d.triHi;
	const int MS = MSC ? MSC : d.ms;
	const bool msaa = MS > 1;
	bool visible = false;
	uint32_t idx[3] = { 0, 0, 0 };
	const bool precull = live && d.cullFlags != nullptr && d.cullFlags[tri] == 0; // marked by k_cull: rows outside the band
	VOut va, vb, vc;
	float sv[3][SWCU_MAXSLOTS]; // slot sources at the three vertices
	float vt[PROG ? 3 : 1][PROG ? SWCU_MAX_PROGRAM : 1]; // PROG: the results of the program's steps at the three vertices
	int PX[SWCU_POLY_MAX], PY[SWCU_POLY_MAX]; // clipped polygons only
	int n = 3, dir = 1;
	bool clipped = false;
	bool frontFacing = false;
	int yMin = 0, yMax = 0, pxMin = 0, pxMax = 0;
	int minXs = 0, maxXs = 0, minYs = 0, maxYs = 0; // 24.8 bounds of the (clipped) polygon
	if(live && !precull)
	{
		do
		{
			triangle_indices(d, tri, idx);
			// the positions of the three vertices, fetched together; the other attributes wait until the triangle is known to be
			// visible (a rank of a multi-GPU frame rejects everything outside its band here)
			float pos[3][4];
			if(PROG)
			{
				for(int a = 0; a < 3; a++)
				{
					vs_run(d, idx[a], vt[a]);
					for(int c = 0; c < 4; c++) pos[a][c] = d.posTemp[c] >= 0 ? vt[a][d.posTemp[c]] : vs_operand(d, d.vsPos[c], idx[a]);
				}
			}
			else
			{
#pragma unroll
			for(int a = 0; a < 3; a++)
#pragma unroll
				for(int c = 0; c < 4; c++) pos[a][c] = vs_operand(d, d.vsPos[c], idx[a]);
			}
			// Early band / scissor reject on an approximate projected y (fast reciprocal, 4-pixel margin): with all three w > 0 the
			// clipped polygon stays inside the hull of the projected vertices, so a triangle whose hull misses the scissor rows
			// cannot produce a span.  This is what a rank of a multi-GPU frame pays for a triangle outside its band.
			// (not for lines and points: their quads reach beyond the hull of the vertices, by up to half of MAX_POINT_SIZE)
			if(!(PROG && d.primKind != PRIM_TRIANGLE) && rows_missed(d, pos[0][1], pos[0][3], pos[1][1], pos[1][3], pos[2][1], pos[2][3])) break;
			process_vertex(d, pos[0][0], pos[0][1], pos[0][2], pos[0][3], va);
			process_vertex(d, pos[1][0], pos[1][1], pos[1][2], pos[1][3], vb);
			process_vertex(d, pos[2][0], pos[2][1], pos[2][2], pos[2][3], vc);

			const bool poly4 = PROG && d.primKind != PRIM_TRIANGLE; // a line or a point: a quad, always clipped and re-projected
			// setupSolidTriangles, Renderer.cpp:749-757
			if(!poly4 && (va.flags & vb.flags & vc.flags) != CLIP_FINITE) break;
			const int flagsOr = va.flags | vb.flags | vc.flags;

			// culling, SetupRoutine.cpp:73-115 (on the original three vertices; triangles only: lines and points keep dir = 1, front facing)
			if(poly4) frontFacing = true;
			else
			{
				const float x0 = (float)va.X, x1 = (float)vb.X, x2 = (float)vc.X;
				const float y0 = (float)va.Y, y1 = (float)vb.Y, y2 = (float)vc.Y;
				float A = fadd(fadd(fmul(fsub(y0, y2), x1), fmul(fsub(y2, y1), x0)), fmul(fsub(y1, y0), x2));
				const int s = (int)(__float_as_uint(va.pw) ^ __float_as_uint(vb.pw) ^ __float_as_uint(vc.pw));
				if(s < 0) A = -A;
				frontFacing = d.frontFace == FRONT_FACE_CCW ? (A >= 0.0f) : (A <= 0.0f);
				if((d.cullMode & CULL_FRONT) && frontFacing) break;
				if((d.cullMode & CULL_BACK) && !frontFacing) break;
				if(!(A > 0.0f)) dir = 0;
			}

			int minY = min(min(va.Y, vb.Y), vc.Y), maxY = max(max(va.Y, vb.Y), vc.Y);
			int minX = min(min(va.X, vb.X), vc.X), maxX = max(max(va.X, vb.X), vc.X);
			if(poly4 || (flagsOr != CLIP_FINITE && (flagsOr & CLIP_FRUSTUM)))
			{
				float4 P[16];
				P[0] = make_float4(va.px, va.py, va.pz, va.pw);
				P[1] = make_float4(vb.px, vb.py, vb.pz, vb.pw);
				P[2] = make_float4(vc.px, vc.py, vc.pz, vc.pw);
				if(PROG && d.primKind == PRIM_LINE)
				{
					// DrawCall::setupLine, Renderer.cpp:920-1000: the rectangle centred on the segment (host C++ in the reference: plain
					// float operations in the same order; its sqrt() on a float argument is the double overload, the quotient a double)
					const float4 P0 = P[0], P1 = P[1];
					if(P0.w <= 0.0f && P1.w <= 0.0f) break;
					const float W = fmul(d.WxF, 1.0f / 256.0f), H = fmul(d.HxF, 1.0f / 256.0f);
					float dx = fmul(W, fsub(fdiv(P1.x, P1.w), fdiv(P0.x, P0.w)));
					float dy = fmul(H, fsub(fdiv(P1.y, P1.w), fdiv(P0.y, P0.w)));
					if(dx == 0.0f && dy == 0.0f) break;
					const float scale = line_scale(d.lineWidth, dx, dy);
					dx = fmul(dx, scale); dy = fmul(dy, scale);
					const float dx0h = fdiv(fmul(dx, P0.w), H), dy0w = fdiv(fmul(dy, P0.w), W);
					const float dx1h = fdiv(fmul(dx, P1.w), H), dy1w = fdiv(fmul(dy, P1.w), W);
					P[2] = P1; P[3] = P0;
					P[0].x = fadd(P[0].x, -dy0w); P[0].y = fadd(P[0].y, dx0h);
					P[1].x = fadd(P[1].x, -dy1w); P[1].y = fadd(P[1].y, dx1h);
					P[2].x = fadd(P[2].x, dy1w); P[2].y = fadd(P[2].y, -dx1h);
					P[3].x = fadd(P[3].x, dy0w); P[3].y = fadd(P[3].y, -dx0h);
				}
				else if(PROG && d.primKind == PRIM_POINT)
				{
					// DrawCall::setupPoint, Renderer.cpp:1137-1185: a square of gl_PointSize pixels around the vertex
					float ps = d.pointSizeTemp >= 0 ? vt[0][d.pointSizeTemp] : vs_operand(d, d.pointSizeSrc, idx[0]);
					ps = ps < 1.0f ? 1.0f : (ps > 1023.0f ? 1023.0f : ps); // clamp(v.pointSize, 1.0f, MAX_POINT_SIZE): a NaN passes through
					const float X = fmul(fmul(ps, va.pw), d.halfPixelX), Y = fmul(fmul(ps, va.pw), d.halfPixelY);
					P[1] = P[0]; P[2] = P[0]; P[3] = P[0];
					P[0].x = fsub(P[0].x, X); P[0].y = fadd(P[0].y, Y);
					P[1].x = fadd(P[1].x, X); P[1].y = fadd(P[1].y, Y);
					P[2].x = fadd(P[2].x, X); P[2].y = fsub(P[2].y, Y);
					P[3].x = fsub(P[3].x, X); P[3].y = fsub(P[3].y, Y);
				}
				n = clip_polygon(P, poly4 ? 4 : 3, poly4 ? (d.depthClipEnable ? CLIP_FRUSTUM : CLIP_SIDES) : flagsOr);
				if(n == 0) break;
				clipped = true;
				for(int i = 0; i < n; i++) // re-projection, SetupRoutine.cpp:125-145
				{
					const float rhw = (P[i].w < 0.0f || P[i].w > 0.0f) ? fdiv(1.0f, P[i].w) : 1.0f; // Float != is FCmpONE: false for NaN
					PX[i] = round_int(fadd(d.X0xF, fmul(fmul(P[i].x, rhw), d.WxF)));
					PY[i] = round_int(fadd(d.Y0xF, fmul(fmul(P[i].y, rhw), d.HxF)));
				}
				minY = maxY = PY[0]; minX = maxX = PX[0];
				for(int i = 1; i < n; i++)
				{
					minY = min(minY, PY[i]); maxY = max(maxY, PY[i]);
					minX = min(minX, PX[i]); maxX = max(maxX, PX[i]);
				}
			}
			minXs = minX; maxXs = maxX; minYs = minY; maxYs = maxY;
			// SetupRoutine.cpp:147-186, in the reference's wrapping 32-bit arithmetic: a vertex with a NaN w projects to the clamp
			// value 2147483520 (RoundIntClamped), the sum wraps negative and the triangle ends here, as it does in the reference
			yMin = (int)((uint32_t)minY + (msaa ? 159u : 255u)) >> 8;
			yMax = (int)((uint32_t)maxY + (msaa ? 351u : 255u)) >> 8;
			yMin = max(yMin, d.suY0);
			yMax = min(yMax, d.suY1);
			if(yMin >= yMax) break;
			// conservative pixel-x bounds of the spans (left = ceil of an edge x >= minX; right <= ceil(maxX))
			const int margin = msaa ? 96 : 0;
			pxMin = clampi((int)(((long long)minX - margin + 255) >> 8), d.scX0, d.scX1);
			pxMax = clampi((int)(((long long)maxX + margin + 255) >> 8), d.scX0, d.scX1);
			// A polygon with a coordinate far outside the screen range (a vertex that projected to INT_MIN or to the clamp value
			// because its w was -Inf or NaN, ...): the reference's edge walk wraps for it, so its spans are not bounded by the
			// polygon's x extent - every tile column of the scissor is a candidate, and k_big must not cull tiles geometrically
			if(polygon_insane(minX, maxX, minY, maxY)) { pxMin = d.scX0; pxMax = d.scX1; }
			if(pxMin >= pxMax) break;
			visible = true;
		} while(0);
	}

	// ---- classification.  SMALL: the clamped pixel bounds fit an 8 x 8 frame — the spans become bit masks stored in the record,
	//      and the triangle touches at most 2 x 2 region bins.  BIG: everything else (also clipped-to-garbage polygons whose edges the
	//      reference walks in wrapped arithmetic): the polygon goes to the big list, the region warps of the tile kernel evaluate
	//      its spans in closed form (edge_at_row) ----
	const int rows = visible ? yMax - yMin : 0;
	bool big = false;
	if(visible) big = rows > SWCU_SMALL_ROWS || (pxMax - pxMin) > SWCU_SMALL_COLS || polygon_insane(minXs, maxXs, minYs, maxYs);
	// ---- owners: the ranks whose bands hold rows of the triangle's regions (one rank, everything local, outside a group).  The rule
	//      only looks at the region rows, so that the owner's k_fill can apply it to the rectangle alone ----
	const int ry0 = visible ? yMin / SWCU_REGION_H : 0, ry1 = visible ? (yMax - 1) / SWCU_REGION_H : 0;
	int o0 = 0, o1 = 0;
	if(visible && d.world > 1)
	{
		o0 = band_of(d, max(ry0 * SWCU_REGION_H, d.suY0));
		o1 = band_of(d, min(ry1 * SWCU_REGION_H + SWCU_REGION_H, d.suY1) - 1);
	}
	unsigned long long slot = 0;
	if(d.world == 1)
	{
		if(__any_sync(0xFFFFFFFFu, big)) slot = warp_alloc(&d.counters->bigSlots, big ? 1u : 0u);
	}
	else if(big) slot = atomicAdd(&d.peerCounters[o0]->bigSlots, 1ull);
	const uint32_t nvis = __popc(__ballot_sync(0xFFFFFFFFu, visible));
	if((threadIdx.x & 31) == 0 && nvis) atomicAdd(&d.counters->visible, nvis);
	// ---- region bins of a small triangle's frame: at most 2 x 2, counted here where the warp is still converged; a region row that
	//      straddles two bands counts at both owners ----
	uint32_t smallRect = TRI_RECT_NONE;
	{
		const bool small = visible && !big;
		const int rx0 = pxMin / SWCU_REGION_W, rx1 = (pxMax - 1) / SWCU_REGION_W;
		if(small) smallRect = (uint32_t)rx0 | ((uint32_t)ry0 << 9) | ((uint32_t)(rx1 - rx0) << 19) | ((uint32_t)(ry1 - ry0) << 20);
		if(!d.direct)
		{
#pragma unroll
			for(int cell = 0; cell < 4; cell++)
			{
				const int cx = (cell & 1) ? rx1 : rx0, cy = (cell & 2) ? ry1 : ry0;
				const bool has = small && (!(cell & 1) || rx1 > rx0) && (!(cell & 2) || ry1 > ry0);
				if(cell && !__any_sync(0xFFFFFFFFu, has)) continue;
				const uint32_t bin = has ? region_bin(d, cx, cy) : 0u;
				int oA = 0, oB = 0;
				if(has && d.world > 1)
				{
					oA = band_of(d, max(cy * SWCU_REGION_H, d.suY0));
					oB = band_of(d, min(cy * SWCU_REGION_H + SWCU_REGION_H, d.suY1) - 1);
				}
				warp_bin_count_at(d, has, oA, bin);
				if(__any_sync(0xFFFFFFFFu, has && oB != oA)) warp_bin_count_at(d, has && oB != oA, oB, bin);
			}
		}
	}
	if(!live) return;
	unsigned char *rec = d.peerRecords[o0] + (size_t)tri * d.triStride;
	if(big && slot >= d.bigCapacity) { atomicOr(&d.counters->overflow, 2u); visible = false; }
	if(!visible)
	{
		// Only the direct mode reads the header of an invisible triangle (every region warp walks the whole list): a small triangle
		// whose frame lies outside every region.  A binned draw never puts it in a bin, so the 32-byte sector is not written at all.
		if(d.direct) *(uint4 *)rec = make_uint4(0xFFFFFFFFu, 0, 0, 0);
		if(d.world == 1) d.triRect[tri] = TRI_RECT_NONE; // (the rectangles of a group member are cleared ahead of the draw: most come from peers)
		return;
	}
	// the attributes behind the plane slots (usually the same cache lines as the positions); in flight during the span work
#pragma unroll
	for(int a = 0; a < 3; a++)
#pragma unroll
		for(int k = 0; k < SWCU_MAXSLOTS; k++)
		{
			sv[a][k] = k < d.nslots ? vs_operand(d, d.slotSrc[k], idx[a]) : 0.0f;
			if(PROG && k < d.nslots && d.slotTemp[k] >= 0) sv[a][k] = vt[a][d.slotTemp[k]];
		}

	uint4 hdr;
	if(big)
	{
		BigTri &b = d.peerBig[o0][slot];
		b.tri = tri; b.walk = 0; b.n = n; b.dir = dir;
		b.yMin = yMin; b.yMax = yMax; b.pxMin = pxMin; b.pxMax = pxMax;
		if(clipped)
			for(int i = 0; i < n; i++) { b.X[i] = PX[i]; b.Y[i] = PY[i]; }
		else
		{
			b.X[0] = va.X; b.X[1] = vb.X; b.X[2] = vc.X;
			b.Y[0] = va.Y; b.Y[1] = vb.Y; b.Y[2] = vc.Y;
		}
		d.peerRect[o0][tri] = TRI_RECT_BIG | (uint32_t)slot;
		hdr = make_uint4((uint32_t)pxMin | ((uint32_t)pxMax << 16), (frontFacing ? TRI_FLAG_FRONT : 0u) | TRI_FLAG_BIG, (uint32_t)yMin | ((uint32_t)yMax << 16), (uint32_t)slot);
	}
	else
	{
		// span rows of the small triangle, built in a shared scratch column (SetupRoutine::edge writes the left and the right
		// half of a row from different edges, the last edge that owns a row wins)
		uint32_t *col = s_rows + threadIdx.x;
		// MSAA pre-fill (SetupRoutine.cpp:214-225): left = right = the clamped pixel of the polygon's first vertex — an empty span,
		// but also what a half keeps when only the other half of a row gets written (degenerate / garbage edges)
		uint32_t fill = 0;
		if(msaa)
		{
			const uint32_t x = (uint32_t)clampi((int)((uint32_t)(clipped ? PX[0] : va.X) + 255u) >> 8, d.scX0, d.scX1);
			fill = x | (x << 16);
		}
		for(int i = 0; i < SWCU_SMALL_ROWS * MS; i++) col[i * SETUP_THREADS] = fill;
		if(clipped) { PX[n] = PX[0]; PY[n] = PY[0]; }
		for(int i = 0; i < n; i++)
		{
			// edge i runs from vertex i to vertex i + 1 (reversed when the winding is reversed)
			int Xs, Ys, Xe, Ye;
			if(clipped) { Xs = PX[i]; Ys = PY[i]; Xe = PX[i + 1]; Ye = PY[i + 1]; }
			else
			{
				Xs = sel3(i, va.X, vb.X, vc.X); Ys = sel3(i, va.Y, vb.Y, vc.Y);
				Xe = sel3(i, vb.X, vc.X, va.X); Ye = sel3(i, vb.Y, vc.Y, va.Y);
			}
			const int Xa = dir ? Xs : Xe, Ya = dir ? Ys : Ye, Xb = dir ? Xe : Xs, Yb = dir ? Ye : Ys;
			if(MSC == 1) edge_small<1>(d, col, yMin, Xa, Ya, Xb, Yb);
			else if(msaa) edge_small<4>(d, col, yMin, Xa, Ya, Xb, Yb);
			else edge_small<1>(d, col, yMin, Xa, Ya, Xb, Yb);
		}
		// ---- spans -> coverage masks in the 8 x 8 frame at (pxMin, yMin): pixel x of row y is covered <=> left <= x < right ----
		auto span_bits = [&](uint32_t v) -> uint32_t {
			const int a = clampi((int)(v & 0xFFFFu) - pxMin, 0, SWCU_SMALL_COLS), e = clampi((int)(v >> 16) - pxMin, 0, SWCU_SMALL_COLS);
			return e > a ? ((1u << e) - 1u) & ~((1u << a) - 1u) : 0u;
		};
		uint32_t m0 = 0, m1 = 0;
		if(!msaa)
		{
#pragma unroll
			for(int r = 0; r < SWCU_SMALL_ROWS; r++)
			{
				const uint32_t bits = span_bits(col[r * SETUP_THREADS]) << (8 * (r & 3));
				if(r < 4) m0 |= bits; else m1 |= bits;
			}
		}
		else
		{
			uint32_t w[SWCU_SMALL_ROWS];
#pragma unroll
			for(int r = 0; r < SWCU_SMALL_ROWS; r++)
			{
				w[r] = 0;
#pragma unroll
				for(int q = 0; q < 4; q++) w[r] |= span_bits(col[(r * 4 + q) * SETUP_THREADS]) << (8 * q);
			}
			((uint4 *)(rec + TRI_HEADER_BYTES))[0] = make_uint4(w[0], w[1], w[2], w[3]);
			((uint4 *)(rec + TRI_HEADER_BYTES))[1] = make_uint4(w[4], w[5], w[6], w[7]);
		}
		hdr = make_uint4((uint32_t)pxMin | ((uint32_t)yMin << 16), frontFacing ? TRI_FLAG_FRONT : 0u, m0, m1);
		d.peerRect[o0][tri] = smallRect;
	}

	// ---- vertex sort (SetupRoutine.cpp:271-294): only changes float rounding of the planes ----
	int i0 = 0, i1 = 1, i2 = 2;
	const bool isTri = !PROG || d.primKind == PRIM_TRIANGLE;
	if(isTri)
	{
		const float y0 = va.py, y1 = vb.py, y2 = vc.py;
		const float ym = sse_min(sse_min(y0, y1), y2);
		rot1(ym == y1, i0, i1, i2);
		rot2(ym == y2, i0, i1, i2);
	}
	if(isTri)
	{
		const float w0 = sel3(i0, va.pw, vb.pw, vc.pw), w1 = sel3(i1, va.pw, vb.pw, vc.pw), w2 = sel3(i2, va.pw, vb.pw, vc.pw);
		const float wm = sse_max(sse_max(w0, w1), w2);
		rot1(wm == w1, i0, i1, i2);
		rot2(wm == w2, i0, i1, i2);
	}
	const float w0 = sel3(i0, va.pw, vb.pw, vc.pw), w1 = sel3(i1, va.pw, vb.pw, vc.pw), w2 = sel3(i2, va.pw, vb.pw, vc.pw);
	const int X0 = sel3(i0, va.X, vb.X, vc.X), X1 = sel3(i1, va.X, vb.X, vc.X);
	const int Y0 = sel3(i0, va.Y, vb.Y, vc.Y), Y1 = sel3(i1, va.Y, vb.Y, vc.Y);
	int X2 = sel3(i2, va.X, vb.X, vc.X), Y2 = sel3(i2, va.Y, vb.Y, vc.Y);
	if(PROG && d.primKind == PRIM_LINE) // the third point of a line's plane equations: the second end point turned by 90 degrees (SetupRoutine.cpp:317-321)
	{
		X2 = (int)((uint32_t)X1 + (uint32_t)Y1 - (uint32_t)Y0);
		Y2 = (int)((uint32_t)Y1 + (uint32_t)X0 - (uint32_t)X1);
	}
	const float rhw0 = sel3(i0, va.rhw, vb.rhw, vc.rhw);
	const float rsub = 1.0f / 256.0f;
	const float x0 = fmul((float)X0, rsub), y0 = fmul((float)Y0, rsub);
	const int dX1 = X1 - X0, dY1 = Y1 - Y0, dX2 = X2 - X0, dY2 = Y2 - Y0;
	const float x1 = fmul(fmul(w1, rsub), (float)dX1), y1 = fmul(fmul(w1, rsub), (float)dY1);
	const float x2 = fmul(fmul(w2, rsub), (float)dX2), y2 = fmul(fmul(w2, rsub), (float)dY2);
	const float a = fsub(fmul(x1, y2), fmul(x2, y1));
	float M00 = 0, M01 = 0, M02 = rhw0, M10 = 0, M11 = 0, M20 = 0, M21 = 0;
	if(a < 0.0f || a > 0.0f) // If(a != 0.0f) is FCmpONE: a NaN area leaves the zero matrix
	{
		const float A = fdiv(1.0f, a);
		const float D = fmul(A, rhw0);
		M00 = fmul(fsub(fmul(y1, w2), fmul(y2, w1)), D);
		M01 = fmul(fsub(fmul(x2, w1), fmul(x1, w2)), D);
		M10 = fmul(y2, A);
		M11 = fmul(-x2, A);
		M20 = fmul(-y1, A);
		M21 = fmul(x1, A);
	}
	float *f = (float *)(rec + d.planeOffset);
	float F[TRI_FLOATS_FRONT + 3 * SWCU_MAXSLOTS + 4]; // the front block, stored with 128-bit writes below
#pragma unroll
	for(int i = 0; i < TRI_FLOATS_FRONT + 3 * SWCU_MAXSLOTS + 4; i++) F[i] = 0.0f;
	F[0] = x0; F[1] = y0;
	F[2] = fadd(fadd(M00, M10), M20);
	F[3] = fadd(fadd(M01, M11), M21);
	F[4] = fadd(fadd(M02, 0.0f), 0.0f);
	// 1/w when the w plane is constant (wA == wB == 0: MulAdd(x, 0, wC + y * 0) == wC at every pixel, bit for bit), so the tile
	// kernel can skip the per-fragment division (PixelRoutine.cpp:196-199); 0 = "not constant".
	if(F[2] == 0.0f && F[3] == 0.0f && F[4] != 0.0f)
	{
		const float r = fdiv(1.0f, F[4]);
		if(r != 0.0f && fabsf(r) <= 3.40282347e38f) F[5] = r;
	}
	const int nf4 = (TRI_FLOATS_FRONT + 3 * d.nslots + 3) >> 2;
	if(d.depthTestActive)
	{
		const float zp0 = sel3(i0, va.zp, vb.zp, vc.zp), zp1 = sel3(i1, va.zp, vb.zp, vc.zp), zp2 = sel3(i2, va.zp, vb.zp, vc.zp);
		const float z0 = zp0;
		const float z1 = fsub(zp1, z0), z2 = fsub(zp2, z0);
		const float px1 = fmul((float)dX1, rsub), py1 = fmul((float)dY1, rsub), px2 = fmul((float)dX2, rsub), py2 = fmul((float)dY2, rsub);
		const float D = fdiv(d.depthRange, fsub(fmul(px1, py2), fmul(px2, py1)));
		float A = fmul(fsub(fmul(py2, z1), fmul(py1, z2)), D);
		float B = fmul(fsub(fmul(px1, z2), fmul(px2, z1)), D);
		if(PROG && d.primKind == PRIM_POINT) { A = 0.0f; B = 0.0f; } // constant depth over a point (SetupRoutine.cpp:405-409)
		const float C = fadd(fmul(z0, d.depthRange), d.depthNear);
		const bool applyConst = d.depthBiasConstant != 0.0f, applySlope = d.depthBiasSlope != 0.0f;
		float zBias = 0.0f;
		float bias = 0.0f; // SetupRoutine.cpp:417-475, floating-point depth buffer branch
		if(applyConst)
		{
			const float Z1 = fadd(fmul(z1, d.depthRange), d.depthNear);
			const float Z2 = fadd(fmul(z2, d.depthRange), d.depthNear);
			const int e0 = (int)(__float_as_uint(C) & 0x7F800000u), e1 = (int)(__float_as_uint(Z1) & 0x7F800000u), e2 = (int)(__float_as_uint(Z2) & 0x7F800000u);
			const int e = max(max(e0, e1), e2);
			// fixed-point depth buffer: the constant minimum resolvable difference of Renderer.cpp:430
			const float r = d.depth16 ? 1.01f / 0xFFFF : fmul(__uint_as_float((uint32_t)e), 1.0f / (1 << 23));
			bias = fmul(r, d.depthBiasConstant);
		}
		if(applySlope) bias = fadd(bias, fmul(sse_max(fabsf(A), fabsf(B)), d.depthBiasSlope));
		if(applyConst || applySlope)
		{
			if(d.depthBiasClamp != 0.0f)
			{
				const float c = d.depthBiasClamp;
				bias = c > 0.0f ? sse_min(bias, c) : sse_max(bias, c);
			}
			zBias = bias;
		}
		((float4 *)f)[nf4] = make_float4(zBias, A, B, C); // the depth block follows the front block
	}
	auto put = [&](int j) { ((float4 *)f)[j] = make_float4(F[4 * j], F[4 * j + 1], F[4 * j + 2], F[4 * j + 3]); };
	put(0);
	// setupGradient, SetupRoutine.cpp:514-548
#pragma unroll
	for(int k = 0; k < SWCU_MAXSLOTS; k++)
	{
		if(k >= d.nslots) break;
		float *P = F + TRI_FLOATS_FRONT + 3 * k;
		const uint32_t mode = d.slotMode[k];
		if(mode == IM_FLAT)
		{
			P[0] = 0; P[1] = 0; P[2] = sv[0][k]; // provoking vertex = Triangle.v0 (or a constant)
		}
		else
		{
			float a0 = sel3(i0, sv[0][k], sv[1][k], sv[2][k]);
			float a1 = sel3(i1, sv[0][k], sv[1][k], sv[2][k]);
			float a2 = sel3(i2, sv[0][k], sv[1][k], sv[2][k]);
			if(mode == IM_NOPERSP) { a0 = fmul(a0, w0); a1 = fmul(a1, w1); a2 = fmul(a2, w2); }
			P[0] = fadd(fadd(fmul(a0, M00), fmul(a1, M10)), fmul(a2, M20));
			P[1] = fadd(fadd(fmul(a0, M01), fmul(a1, M11)), fmul(a2, M21));
			P[2] = fadd(fadd(fmul(a0, M02), fmul(a1, 0.0f)), fmul(a2, 0.0f));
		}
		// floats up to index 8 + 3k are final: store the float4s this slot completed
#pragma unroll
		for(int j = 1; j < (TRI_FLOATS_FRONT + 3 * SWCU_MAXSLOTS + 3) / 4; j++)
			if(4 * j + 3 <= 8 + 3 * k && 4 * j + 3 > 5 + 3 * k) put(j);
	}
	// the float4 that holds the padding is still open
#pragma unroll
	for(int j = 1; j < (TRI_FLOATS_FRONT + 3 * SWCU_MAXSLOTS + 3) / 4; j++)
		if(j < nf4 && 4 * j + 3 > 5 + 3 * d.nslots) put(j);
	*(uint4 *)rec = hdr;
	// ---- a triangle whose regions reach into further bands: the same record (read back) for each further owner; a big triangle
	//      gets a slot in that owner's big list ----
	for(int o = o0 + 1; o <= o1; o++)
	{
		unsigned char *rec2 = d.peerRecords[o] + (size_t)tri * d.triStride;
		for(uint32_t i = 16; i < d.triStride; i += 16) *(uint4 *)(rec2 + i) = *(const uint4 *)(rec + i);
		uint4 h2 = hdr;
		if(big)
		{
			const unsigned long long slot2 = atomicAdd(&d.peerCounters[o]->bigSlots, 1ull);
			if(slot2 >= d.bigCapacity) { atomicOr(&d.peerCounters[o]->overflow, 2u); continue; }
			d.peerBig[o][slot2] = d.peerBig[o0][slot];
			h2.w = (uint32_t)slot2;
			d.peerRect[o][tri] = TRI_RECT_BIG | (uint32_t)slot2;
		}
		else d.peerRect[o][tri] = smallRect;
		*(uint4 *)rec2 = h2;
	}
}

__global__ void __launch_bounds__(SETUP_THREADS, SETUP_BLOCKS_1X) k_setup_1x(const __grid_constant__ DrawConst d) { setup_triangle<1>(d); }
#ifndef SETUP_BLOCKS_4X
#define SETUP_BLOCKS_4X 6 // 80 registers, no spills: C4 set-up 0.183 -> 0.175 ms (5 CTAs of 96 registers before; 7 CTAs of 72 spill and take 0.186 ms)
#endif
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_BLOCKS_4X) k_setup(const __grid_constant__ DrawConst d) { setup_triangle<0>(d); }
// The same kernel capped at 72 registers (7 CTAs per SM; it spills a little and a wave of it takes ~1.24x as long): launched when the
// draw — a rank's share of a group draw — then fits in fewer waves.  Every thread is one triangle's serial chain, so a wave costs its
// full latency however empty it is: 125 k triangles are 1.1 waves of k_setup (two wave times) but one wave of this one.
#define SETUP_BLOCKS_WIDE 7
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_BLOCKS_WIDE) k_setup_wide(const __grid_constant__ DrawConst d) { setup_triangle<0>(d); }
__global__ void __launch_bounds__(SETUP_THREADS) k_setup_prog(const __grid_constant__ DrawConst d) { setup_triangle<0, true>(d); }

// ------------------------------------------------------------------------------------------------------------------
// binning: (region, triangle) pairs without a sort and without a host round trip
//
//   k_setup      counts the (at most 2 x 2) region bins of every small triangle           binCount[bin]++
//   k_binscan    its first blocks count the regions of the big triangles (one warp per triangle: the regions of its bounding box that
//                an edge does not exclude; a big triangle whose bounding box does not fit the remaining pair budget is marked `walk`
//                and not binned), the others then scan the counts (single pass, decoupled look-back)         binStart[bin]
//   k_fill       every pair takes a slot of its bin: binStart[bin] + (--binCount[bin]); the counts end at zero again
//
// The slots of a bin are handed out by atomics, i.e. in no particular order; the API order of the triangles (the reference's cluster
// tickets, Renderer.cpp:573-576) is restored by sorting each bin's few entries by triangle id.  The pair buffer holds 4 pairs per
// triangle of the draw plus a budget for the big triangles, so no count has to reach the host before the next launch.
// ------------------------------------------------------------------------------------------------------------------

// Can a fragment of big triangle b land in region (rx, ry)?  Conservative — the bounding box, minus the regions all of whose pixel
// centres (widened by the sample offsets) lie strictly outside one edge — and deterministic: the count (k_binscan) and the fill (k_fill) must agree.
// sgn: orientation of the polygon, 0 = do not cull geometrically (garbage coordinates the reference walks in wrapped arithmetic).
DEVI int big_orientation(const BigTri &b)
{
	long long area2 = 0;
	int loX = b.X[0], hiX = b.X[0], loY = b.Y[0], hiY = b.Y[0];
	for(int i = 0; i < b.n; i++)
	{
		const int j = i + 1 == b.n ? 0 : i + 1;
		area2 += (long long)b.X[i] * b.Y[j] - (long long)b.X[j] * b.Y[i];
		loX = min(loX, b.X[i]); hiX = max(hiX, b.X[i]); loY = min(loY, b.Y[i]); hiY = max(hiY, b.Y[i]);
	}
	return polygon_insane(loX, hiX, loY, hiY) ? 0 : (area2 > 0 ? 1 : (area2 < 0 ? -1 : 0));
}
DEVI bool big_touches_region(const DrawConst &d, const BigTri &b, int rx, int ry, int sgn)
{
	if(sgn == 0) return true;
	const int m = d.ms > 1 ? 96 : 0;
	// pixel centres of the region in 24.8 (centre of pixel x is X = 256x), widened by the sample offsets
	const long long cx0 = 256ll * max(rx * SWCU_REGION_W, b.pxMin) - m, cx1 = 256ll * (min(rx * SWCU_REGION_W + SWCU_REGION_W, b.pxMax) - 1) + m;
	const long long cy0 = 256ll * max(ry * SWCU_REGION_H, b.yMin) - m, cy1 = 256ll * (min(ry * SWCU_REGION_H + SWCU_REGION_H, b.yMax) - 1) + m;
	for(int i = 0; i < b.n; i++)
	{
		const int j = i + 1 == b.n ? 0 : i + 1;
		const long long ex = (long long)b.X[j] - b.X[i], ey = (long long)b.Y[j] - b.Y[i];
		const long long slack = 4 * (llabs(ex) + llabs(ey)) + 1024; // rounding of re-projected clip vertices
		// E(P) = ex*(Py - Yi) - ey*(Px - Xi); inside when sgn*E >= 0
		const long long e00 = sgn * (ex * (cy0 - b.Y[i]) - ey * (cx0 - b.X[i]));
		const long long e10 = sgn * (ex * (cy0 - b.Y[i]) - ey * (cx1 - b.X[i]));
		const long long e01 = sgn * (ex * (cy1 - b.Y[i]) - ey * (cx0 - b.X[i]));
		const long long e11 = sgn * (ex * (cy1 - b.Y[i]) - ey * (cx1 - b.X[i]));
		if(e00 < -slack && e10 < -slack && e01 < -slack && e11 < -slack) return false;
	}
	return true;
}

// FILL = false: count the regions of every big triangle (first blocks of k_binscan); FILL = true: hand out their slots (second part of k_fill)
template<bool FILL>
DEVI void big_regions(const DrawConst &d, uint32_t warpIndex, uint32_t warpCount)
{
	const int lane = threadIdx.x & 31;
	const uint32_t nbig = (uint32_t)min(d.counters->bigSlots, (unsigned long long)d.bigCapacity);
	for(uint32_t e = warpIndex; e < nbig; e += warpCount)
	{
		BigTri &b = d.bigList[e];
		const int rx0 = b.pxMin / SWCU_REGION_W, rx1 = (b.pxMax - 1) / SWCU_REGION_W;
		// the region rows of the triangle that hold rows of MY part of the frame (a member of a group: its band)
		const int ry0 = max(b.yMin, d.scY0) / SWCU_REGION_H, ry1 = (min(b.yMax, d.scY1) - 1) / SWCU_REGION_H;
		if(ry1 < ry0) continue;
		const int w = rx1 - rx0 + 1, total = w * (ry1 - ry0 + 1);
		uint32_t walk = 0;
		if(!FILL)
		{
			// the bounding box must fit what is left of the pair budget; otherwise the triangle is not binned at all
			if(lane == 0)
			{
				const unsigned long long before = atomicAdd(&d.counters->bigReserved, (unsigned long long)total);
				if(before + (unsigned long long)total > d.bigBudget) { walk = 1; b.walk = 1; atomicOr(&d.counters->overflow, 4u); }
			}
			walk = __shfl_sync(0xFFFFFFFFu, walk, 0);
		}
		else walk = b.walk;
		if(walk) continue;
		const int sgn = big_orientation(b);
		for(int t = lane; t < total; t += 32)
		{
			const int rx = rx0 + t % w, ry = ry0 + t / w;
			if(!big_touches_region(d, b, rx, ry, sgn)) continue;
			const uint32_t bin = region_bin(d, rx, ry);
			if(!FILL) atomicAdd(d.binCount + bin, 1u);
			else d.pairs[d.binStart[bin] + (atomicSub(d.binCount + bin, 1u) - 1u)] = b.tri;
		}
	}
}

#define BIG_WARPS_PER_BLOCK 8
// Exclusive scan of the bin counts in one pass (decoupled look-back: a block publishes its sum, then adds up the sums of its
// predecessors until it meets one that already knows its prefix).  state[] and the ticket start at zero.
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
// The first bigBlocks blocks of the grid count the regions of the big triangles (k_setup has counted the small ones) and check in;
// the scan blocks wait for them — all blocks of this small grid are resident together — unless the draw has no big triangle at all.
__global__ void __launch_bounds__(SCAN_THREADS) k_binscan(const __grid_constant__ DrawConst d, uint32_t bigBlocks)
{
	__shared__ uint32_t s_block, s_warp[SCAN_THREADS / 32], s_prefix;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	DrawCounters *c = d.counters;
	const uint32_t *count = d.binCount;
	uint32_t *start = d.binStart;
	const uint32_t n = d.numBins;
	volatile uint32_t *state = (volatile uint32_t *)(d.counters + 1);
	if(blockIdx.x < bigBlocks)
	{
		big_regions<false>(d, blockIdx.x * BIG_WARPS_PER_BLOCK + (threadIdx.x >> 5), bigBlocks * BIG_WARPS_PER_BLOCK);
		__syncthreads();
		if(tid == 0) { __threadfence(); atomicAdd(&c->bigDone, 1u); }
		return;
	}
	if(tid == 0)
	{
		if(c->bigSlots != 0)
			while(*(volatile uint32_t *)&c->bigDone < bigBlocks) __nanosleep(32);
		__threadfence();
		s_block = atomicAdd(&c->scanTicket, 1u); // blocks take their part in the order they start: a predecessor is always running
	}
	__syncthreads();
	const uint32_t blk = s_block;
	const uint32_t base = (blk * SCAN_THREADS + tid) * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
#pragma unroll
	for(int i = 0; i < SCAN_ITEMS; i++) v[i] = base + i < n ? count[base + i] : 0u;
	uint32_t sum = 0;
#pragma unroll
	for(int i = 0; i < SCAN_ITEMS; i++) sum += v[i];
	uint32_t incl = sum;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1)
	{
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if(lane >= o) incl += t;
	}
	if(lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint32_t warpBase = 0, blockTotal = 0;
#pragma unroll
	for(int w = 0; w < SCAN_THREADS / 32; w++)
	{
		if(w < warp) warpBase += s_warp[w];
		blockTotal += s_warp[w];
	}
	if(warp == 0)
	{
		// look-back by the whole warp: lane l reads the state of predecessor blk - 1 - l (32 at a time, one trip to L2 per window
		// instead of one per predecessor), the lanes up to the nearest predecessor that already knows its prefix add up
		uint32_t prefix = 0;
		if(blk == 0) { if(lane == 0) state[0] = 0x80000000u | blockTotal; }
		else
		{
			if(lane == 0) { state[blk] = 0x40000000u | blockTotal; __threadfence(); }
			for(int hi = (int)blk - 1; hi >= 0; hi -= 32)
			{
				const int j = hi - lane;
				uint32_t sv = 0x80000000u; // (lanes past block 0 contribute nothing and end the walk)
				if(j >= 0)
					do { sv = state[j]; } while((sv & 0xC0000000u) == 0);
				const uint32_t done = __ballot_sync(0xFFFFFFFFu, (sv & 0x80000000u) != 0);
				const int last = done ? __ffs(done) - 1 : 31; // nearest predecessor with an inclusive prefix
				prefix += __reduce_add_sync(0xFFFFFFFFu, (lane <= last && j >= 0) ? (sv & 0x3FFFFFFFu) : 0u);
				if(done) break;
			}
			if(lane == 0) state[blk] = 0x80000000u | (prefix + blockTotal);
		}
		if(lane == 0)
		{
			s_prefix = prefix;
			if((blk + 1) * SCAN_THREADS * SCAN_ITEMS >= n) // the block that holds the last bin
			{
				start[n] = prefix + blockTotal;
				c->pairTotal = prefix + blockTotal;
			}
		}
	}
	__syncthreads();
	uint32_t run = s_prefix + warpBase + incl - sum;
#pragma unroll
	for(int i = 0; i < SCAN_ITEMS; i++)
	{
		if(base + i < n) start[base + i] = run;
		run += v[i];
	}
}

// every (region, triangle) pair takes a slot of its bin; blocks [0, smallBlocks) walk the small triangles (one thread each),
// the blocks after them the big list (one warp per triangle)
#define FILL_TRIS 4
__global__ void __launch_bounds__(256) k_fill(const __grid_constant__ DrawConst d, uint32_t smallBlocks)
{
	if(blockIdx.x >= smallBlocks)
	{
		big_regions<true>(d, (blockIdx.x - smallBlocks) * BIG_WARPS_PER_BLOCK + (threadIdx.x >> 5), (gridDim.x - smallBlocks) * BIG_WARPS_PER_BLOCK);
		return;
	}
	// FILL_TRIS triangles per thread — a warp walks 32 * FILL_TRIS consecutive triangles, 32 consecutive ones per step (neighbours of
	// a mesh share their bins: one atomic per distinct bin of a step), all rectangles requested up front.  A block of one-triangle
	// threads lived for one trip to memory, and the pass over a big mesh — of which a rank of a group owns only its band — was bound
	// by the turnover of such blocks (10 M triangles: 54 us however few of them were the rank's).
	const uint32_t warpBase = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32u * FILL_TRIS) + (threadIdx.x & 31);
	uint32_t rr[FILL_TRIS];
#pragma unroll
	for(int k = 0; k < FILL_TRIS; k++) rr[k] = warpBase + 32u * k < d.primCount ? d.triRect[warpBase + 32u * k] : TRI_RECT_NONE;
	// (a thread of invisible / big / foreign triangles has nothing to do; a warp of them leaves here)
	if(!__any_sync(0xFFFFFFFFu, !(rr[0] & rr[1] & rr[2] & rr[3] & TRI_RECT_BIG))) return;
#pragma unroll
	for(int k = 0; k < FILL_TRIS; k++)
	{
		const uint32_t tri = warpBase + 32u * k, r = rr[k];
		// a member of a group: the rectangles come from the peers (only those of visible triangles are stored), so every entry that has
		// been read goes back to "none" for the draw that uses this set next — no pass over the whole array in between
		if(d.world > 1 && r != TRI_RECT_NONE) d.triRect[tri] = TRI_RECT_NONE;
		const bool small = !(r & TRI_RECT_BIG); // big: the blocks behind the small ones; invisible: nothing to do
		if(!__any_sync(0xFFFFFFFFu, small)) continue;
		const int rx0 = r & 0x1FF, ry0 = (r >> 9) & 0x3FF, rx1 = rx0 + ((r >> 19) & 1), ry1 = ry0 + ((r >> 20) & 1);
		// warp-aggregated: one atomic per distinct bin of the warp's triangles.  A cell counts here if its region row holds rows of MY
		// part of the frame — the rule k_setup (possibly on another rank) counted it by
#pragma unroll
		for(int cell = 0; cell < 4; cell++)
		{
			const int cx = (cell & 1) ? rx1 : rx0, cy = (cell & 2) ? ry1 : ry0;
			const bool has = small && (!(cell & 1) || rx1 > rx0) && (!(cell & 2) || ry1 > ry0) &&
			                 max(cy * SWCU_REGION_H, d.scY0) < min(cy * SWCU_REGION_H + SWCU_REGION_H, d.scY1);
			if(!__any_sync(0xFFFFFFFFu, has)) continue;
			const uint32_t bin = has ? region_bin(d, cx, cy) : 0u;
			const uint32_t slot = warp_bin_slot(d.binCount, has, bin);
			if(has) d.pairs[d.binStart[bin] + slot] = tri;
		}
	}
}

// Ascending sort of n values with the bitonic network in its "all comparisons point the same way" form (the first step of
// every merge mirrors the block, the following ones are the usual half-cleaners), so indices >= n simply behave like +infinity and
// n need not be a power of two.  `sync` separates the steps (warp or CTA barrier), the threads stride over the comparators.
template<typename Sync>
DEVI void bitonic_sort_any(uint32_t *a, uint32_t n, uint32_t tid, uint32_t nthreads, Sync sync)
{
	for(uint32_t k = 2; (k >> 1) < n; k <<= 1)
	{
		for(uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			const bool mirror = j == (k >> 1);
			for(uint32_t t = tid; t < (n + 1) / 2 + j; t += nthreads) // comparator t of this step: low index lo, partner hi > lo
			{
				const uint32_t lo = ((t / j) * 2 * j) + (t % j);
				const uint32_t hi = mirror ? (lo ^ (k - 1)) : (lo + j);
				if(lo < n && hi < n && hi > lo)
				{
					const uint32_t x = a[lo], y = a[hi];
					if(x > y) { a[lo] = y; a[hi] = x; }
				}
			}
			sync();
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// sampler (SamplerCore.cpp 16-bit fixed-point path :173-198) — RGBA8, 2D, normalised coordinates
// ------------------------------------------------------------------------------------------------------------------
DEVI uint32_t mulhi16(uint32_t a, uint32_t b) { return (a * b) >> 16; } // on values < 65536

DEVI uint32_t tex_address(float u, uint32_t mode) // SamplerCore::address :2399-2435
{
	if(mode == ADDR_CLAMP_TO_EDGE)
	{
		const float c = sse_min(sse_max(u, 0.0f), 65535.0f / 65536.0f);
		return (uint32_t)trunc_int(fmul(c, 65536.0f)) & 0xFFFF;
	}
	int convert = trunc_int(fmul(u, 65536.0f));
	if(mode == ADDR_MIRRORED_REPEAT)
	{
		const int mirror = (int)((uint32_t)convert << 15) >> 31;
		convert ^= mirror;
	}
	return (uint32_t)convert & 0xFFFF;
}

DEVI uint32_t offset_sample(uint32_t uvw, uint32_t half, bool wrap, int count) // :278-313
{
	if(wrap) return (count < 0 ? uvw - half : uvw + half) & 0xFFFF;
	if(count < 0) return uvw < half ? 0u : uvw - half;
	const uint32_t s = uvw + half;
	return s > 0xFFFF ? 0xFFFFu : s;
}

// One 8-bit texel channel in the 16-bit sampler path: b << 8, or — RGB of an sRGB image — the reference's start-up table
// sRGBtoLinearFF_FF00 (SamplerCore.cpp:1966-1977, :2670-2680).  The table sits in global memory and is read through the
// read-only path: 512 bytes, resident in L1, and unlike constant memory not serialised when the lanes' texels differ.
__device__ const unsigned short g_srgbLut[256] = { SWCU_SRGB_LUT_VALUES };
DEVI uint32_t texel16(bool srgb, uint32_t t, int c)
{
	const uint32_t b = __byte_perm(t, 0, 0x4440 + c);
	return (srgb && c < 3) ? (uint32_t)__ldg(g_srgbLut + b) : b << 8;
}

DEVI uint32_t load_texel(const KMip &m, uint32_t x, uint32_t y) { return __ldg((const uint32_t *)m.buffer + (x + y * m.pitchP)); }

// one tap set of one mip level; out[c] are 16-bit channel values.  FAST = REPEAT/REPEAT addressing + linear filter (the
// benchmark sampler): same arithmetic with the state tests folded away.
template<bool FAST>
DEVI void sample_level(const DrawConst &d, int level, float u, float v, bool linear, uint32_t out[4])
{
	level = clampi(level, 0, SWCU_MIPMAP_LEVELS - 1);
	const int l = level < (int)d.texLevels ? level : (int)d.texLevels - 1; // VkDescriptorSetLayout.cpp:470
	const KMip &m = d.mip[l];
	const uint32_t W = m.width & 0xFFFF, H = m.height & 0xFFFF;
	const uint32_t uu = tex_address(u, FAST ? (uint32_t)ADDR_REPEAT : d.addressU), vv = tex_address(v, FAST ? (uint32_t)ADDR_REPEAT : d.addressV);
	if(!FAST && !linear)
	{
		const uint32_t t = load_texel(m, mulhi16(uu, W), mulhi16(vv, H));
#pragma unroll
		for(int c = 0; c < 4; c++) out[c] = texel16(d.texSrgb != 0, t, c); // (!FAST only)
		return;
	}
	const uint32_t uHalf = m.half & 0xFFFF, vHalf = m.half >> 16; // 0x8000 / extent, VkDescriptorSetLayout.cpp:315
	const bool wrapU = FAST || d.addressU == ADDR_REPEAT, wrapV = FAST || d.addressV == ADDR_REPEAT;
	const uint32_t u0 = offset_sample(uu, uHalf, wrapU, -1), u1 = offset_sample(uu, uHalf, wrapU, +1);
	const uint32_t v0 = offset_sample(vv, vHalf, wrapV, -1), v1 = offset_sample(vv, vHalf, wrapV, +1);
	const uint32_t x0 = mulhi16(u0, W), x1 = mulhi16(u1, W), y0 = mulhi16(v0, H), y1 = mulhi16(v1, H);
	const uint32_t *base = (const uint32_t *)m.buffer;
	const uint32_t r0 = y0 * m.pitchP, r1 = y1 * m.pitchP;
	const uint32_t t00 = __ldg(base + r0 + x0), t10 = __ldg(base + r0 + x1), t01 = __ldg(base + r1 + x0), t11 = __ldg(base + r1 + x1);
	const uint32_t f0u = (u0 * W) & 0xFFFF, f0v = (v0 * H) & 0xFFFF;
	const uint32_t f1u = ~f0u & 0xFFFF, f1v = ~f0v & 0xFFFF;
	const uint32_t f0u0v = mulhi16(f0u, f0v), f1u0v = mulhi16(f1u, f0v), f0u1v = mulhi16(f0u, f1v), f1u1v = mulhi16(f1u, f1v);
#pragma unroll
	for(int c = 0; c < 4; c++)
	{
		uint32_t c00, c10, c01, c11;
		if(!FAST && d.texSrgb) // the benchmark sampler (FAST) is only selected for UNORM images
		{
			c00 = mulhi16(texel16(true, t00, c), f1u1v); c10 = mulhi16(texel16(true, t10, c), f0u1v);
			c01 = mulhi16(texel16(true, t01, c), f1u0v); c11 = mulhi16(texel16(true, t11, c), f0u0v);
		}
		else
		{
			// mulhi(b << 8, w) == (b * w) >> 8 for a byte b and a 16-bit weight w
			c00 = (__byte_perm(t00, 0, 0x4440 + c) * f1u1v) >> 8;
			c10 = (__byte_perm(t10, 0, 0x4440 + c) * f0u1v) >> 8;
			c01 = (__byte_perm(t01, 0, 0x4440 + c) * f1u0v) >> 8;
			c11 = (__byte_perm(t11, 0, 0x4440 + c) * f0u0v) >> 8;
		}
		out[c] = (((c00 + c10) & 0xFFFF) + ((c01 + c11) & 0xFFFF)) & 0xFFFF;
	}
}

// the zero-offset 4-tap variant the reference runs when min and mag filters differ and the point filter is selected
DEVI void sample_level_split_point(const DrawConst &d, int ilod, float u, float v, uint32_t out[4])
{
	const int l = ilod < (int)d.texLevels ? (ilod < 0 ? 0 : ilod) : (int)d.texLevels - 1;
	const KMip &m = d.mip[l];
	const uint32_t W = m.width & 0xFFFF, H = m.height & 0xFFFF;
	const uint32_t uu = tex_address(u, d.addressU), vv = tex_address(v, d.addressV);
	const uint32_t t = load_texel(m, mulhi16(uu, W), mulhi16(vv, H));
	const uint32_t f0u = (uu * W) & 0xFFFF, f0v = (vv * H) & 0xFFFF, f1u = ~f0u & 0xFFFF, f1v = ~f0v & 0xFFFF;
	const uint32_t w00 = mulhi16(f1u, f1v), w10 = mulhi16(f0u, f1v), w01 = mulhi16(f1u, f0v), w11 = mulhi16(f0u, f0v);
#pragma unroll
	for(int c = 0; c < 4; c++)
	{
		const uint32_t tx = texel16(d.texSrgb != 0, t, c);
		out[c] = (((mulhi16(tx, w00) + mulhi16(tx, w10)) & 0xFFFF) + ((mulhi16(tx, w01) + mulhi16(tx, w11)) & 0xFFFF)) & 0xFFFF;
	}
}

struct LodState
{
	float lod;
	int ilod;
	bool linear, split;
};

// computeLod2D :1376-1422 + log2sqrt :1333-1341 + selectMipmap :2357-2379; u/v of quad lanes 0,1,2
template<bool FAST>
DEVI LodState compute_lod(const DrawConst &d, float u0, float u1, float u2, float v0, float v1, float v2)
{
	LodState s;
	s.split = FAST ? false : d.magFilter != d.minFilter;
	bool filterLinear = FAST ? true : (s.split ? false : d.magFilter == FILTER_LINEAR);
	float minLod = d.minLod, maxLod = d.maxLod;
	if(d.texLevels == 1 && !s.split) { minLod = 0.0f; maxLod = 0.0f; }
	float lod;
	if(minLod == maxLod) lod = minLod;
	else
	{
		const float Wf = (float)d.mip[0].width, Hf = (float)d.mip[0].height;
		const float dUdx = fmul(fsub(u1, u0), Wf), dUdy = fmul(fsub(u2, u0), Wf);
		const float dVdx = fmul(fsub(v1, v0), Hf), dVdy = fmul(fsub(v2, v0), Hf);
		const float sx = fadd(fmul(dUdx, dUdx), fmul(dVdx, dVdx)), sy = fadd(fmul(dUdy, dUdy), fmul(dVdy, dVdy));
		lod = sse_max(sx, sy);
		lod = fmul(lod, lod);
		lod = fsub((float)(int)__float_as_uint(lod), (float)0x3F800000);
		lod = fmul(lod, __uint_as_float(0x33000000u));
		lod = fadd(lod, d.mipLodBias);
		lod = sse_max(lod, minLod);
		lod = sse_min(lod, maxLod);
	}
	s.linear = filterLinear;
	if(s.split)
	{
		const bool minLinear = d.minFilter == FILTER_LINEAR;
		s.linear = minLinear ? (lod > 0.0f) : (lod <= 0.0f); // CmpNLE is FCmpOGT
	}
	s.lod = lod;
	s.ilod = (!FAST && d.mipmapMode == MIPMAP_MODE_NEAREST) ? round_int(lod) : trunc_int(lod);
	return s;
}

template<bool FAST>
DEVI void sample_texture(const DrawConst &d, const LodState &s, float u, float v, float out[4])
{
	uint32_t c[4];
	if(!FAST && s.split && !s.linear) sample_level_split_point(d, s.ilod, u, v, c);
	else sample_level<FAST>(d, s.ilod, u, v, s.linear, c);
	if(FAST || d.mipmapMode == MIPMAP_MODE_LINEAR) // sampleFilter :324-373
	{
		const uint32_t utri = (uint32_t)trunc_int(fmul(s.lod, 65536.0f)) & 0xFFFF;
		const uint32_t inv = ~utri & 0xFFFF;
		// the reference always fetches level ilod + 1; with a zero weight (magnification, integer LOD) its term mulhi(cc, 0) is 0
		// whatever the texels are, so the fetch is skipped: c = mulhi(c, 0xFFFF) + 0.  (Measured: a per-lane branch costs the
		// minified 10 M-triangle scene 2.6 % of its tile kernel and saves the magnified 4K triangle 30 %; warp-voted variants that
		// keep both levels' loads together were slower on both.)
		uint32_t cc[4] = { 0, 0, 0, 0 };
		if(utri != 0)
		{
			if(!FAST && s.split && !s.linear) sample_level_split_point(d, s.ilod + 1, u, v, cc); // both levels of a trilinear fetch
			else sample_level<FAST>(d, s.ilod + 1, u, v, s.linear, cc);
		}
#pragma unroll
		for(int ch = 0; ch < 4; ch++) c[ch] = (mulhi16(c[ch], inv) + mulhi16(cc[ch], utri)) & 0xFFFF;
	}
#pragma unroll
	for(int ch = 0; ch < 4; ch++) out[ch] = fmul((float)c[ch], 1.0f / 0xFF00);
}

// ------------------------------------------------------------------------------------------------------------------
// pixel helpers (PixelRoutine.cpp)
// ------------------------------------------------------------------------------------------------------------------
DEVI bool stencil_compare(uint32_t op, uint32_t value, uint32_t refMasked) // :406-449
{
	switch(op)
	{
	case CMP_ALWAYS: return true;
	case CMP_NEVER: return false;
	case CMP_LESS: return refMasked < value;
	case CMP_EQUAL: return refMasked == value;
	case CMP_NOT_EQUAL: return refMasked != value;
	case CMP_LESS_OR_EQUAL: return refMasked <= value;
	case CMP_GREATER: return refMasked > value;
	default: return refMasked >= value;
	}
}

DEVI uint32_t stencil_op(uint32_t op, uint32_t v, uint32_t ref) // :870-902
{
	switch(op)
	{
	case SOP_KEEP: return v;
	case SOP_ZERO: return 0;
	case SOP_REPLACE: return ref;
	case SOP_INC_CLAMP: return v == 0xFF ? 0xFF : v + 1;
	case SOP_DEC_CLAMP: return v == 0 ? 0 : v - 1;
	case SOP_INVERT: return ~v & 0xFF;
	case SOP_INC_WRAP: return (v + 1) & 0xFF;
	default: return (v - 1) & 0xFF;
	}
}

DEVI bool depth_compare(uint32_t op, float zValue, float Z) // :533-553; CmpNEQ / CmpNLE / CmpNLT are the ORDERED FCmpONE / OGT / OGE
{
	switch(op)
	{
	case CMP_ALWAYS: return true;
	case CMP_NEVER: return false;
	case CMP_EQUAL: return zValue == Z;
	case CMP_NOT_EQUAL: return zValue < Z || zValue > Z;
	case CMP_LESS: return zValue > Z;
	case CMP_GREATER_OR_EQUAL: return zValue <= Z;
	case CMP_LESS_OR_EQUAL: return zValue >= Z;
	default: return zValue < Z;
	}
}

DEVI float blend_factor_rgb(const DrawConst &d, uint32_t f, int ch, const float s[4], const float dst[4]) // :1225-1393
{
	switch(f)
	{
	case BF_ZERO: return 0.0f;
	case BF_ONE: return 1.0f;
	case BF_SRC_COLOR: return s[ch];
	case BF_ONE_MINUS_SRC_COLOR: return fsub(1.0f, s[ch]);
	case BF_DST_COLOR: return dst[ch];
	case BF_ONE_MINUS_DST_COLOR: return fsub(1.0f, dst[ch]);
	case BF_SRC_ALPHA: return s[3];
	case BF_ONE_MINUS_SRC_ALPHA: return fsub(1.0f, s[3]);
	case BF_DST_ALPHA: return dst[3];
	case BF_ONE_MINUS_DST_ALPHA: return fsub(1.0f, dst[3]);
	case BF_SRC_ALPHA_SATURATE: return sse_min(fsub(1.0f, dst[3]), s[3]);
	case BF_CONSTANT_COLOR: return d.blendConstant[ch];
	case BF_CONSTANT_ALPHA: return d.blendConstant[3];
	case BF_ONE_MINUS_CONSTANT_COLOR: return fsub(1.0f, d.blendConstant[ch]);
	case BF_ONE_MINUS_CONSTANT_ALPHA: return fsub(1.0f, d.blendConstant[3]);
	}
	return 0.0f;
}

DEVI float blend_factor_a(const DrawConst &d, uint32_t f, const float s[4], const float dst[4])
{
	switch(f)
	{
	case BF_ZERO: return 0.0f;
	case BF_ONE: return 1.0f;
	case BF_SRC_COLOR: case BF_SRC_ALPHA: return s[3];
	case BF_ONE_MINUS_SRC_COLOR: case BF_ONE_MINUS_SRC_ALPHA: return fsub(1.0f, s[3]);
	case BF_DST_COLOR: case BF_DST_ALPHA: return dst[3];
	case BF_ONE_MINUS_DST_COLOR: case BF_ONE_MINUS_DST_ALPHA: return fsub(1.0f, dst[3]);
	case BF_SRC_ALPHA_SATURATE: return 1.0f;
	case BF_CONSTANT_COLOR: case BF_CONSTANT_ALPHA: return d.blendConstant[3];
	case BF_ONE_MINUS_CONSTANT_COLOR: case BF_ONE_MINUS_CONSTANT_ALPHA: return fsub(1.0f, d.blendConstant[3]);
	}
	return 0.0f;
}

// sRGB colour targets.  Pow<Mediump> = Exp2(y * Log2(x)) with the relaxed-precision polynomials of ShaderCore.cpp:352-382 (Exp2),
// :412-436 (Log2), :472-477 (Pow); MulAdd is an FMA, Float(Int) rounds to nearest, Int(Float) truncates.
DEVI float log2_mediump(float x)
{
	const int im = (int)__float_as_uint(x);
	float y = __fmaf_rn((float)im, 1.0f / (1 << 23), -127.0f);
	if(im == 0x7F800000) y = __uint_as_float(__float_as_uint(y) | 0x7F800000u);
	const float m = (float)(im & 0x007FFFFF);
	const float f = __fmaf_rn(__fmaf_rn(2.8017103e-22f, m, -8.373131e-15f), m, 5.0615534e-8f);
	return __fmaf_rn(f, m, y);
}
DEVI float exp2_mediump(float x)
{
	float x0 = sse_min(x, 128.0f);
	x0 = sse_max(x0, __uint_as_float(0xC2FDFFFFu));
	const float f = fsub(x0, floorf(x0));
	const float r = __fmaf_rn(__fmaf_rn(7.8145574e-2f, f, 2.2617357e-1f), f, -3.0444314e-1f);
	const float y = __fmaf_rn(r, f, x0);
	return __uint_as_float((uint32_t)trunc_int(__fmaf_rn((float)(1 << 23), y, (float)(127 << 23))));
}
DEVI float pow_mediump(float x, float y) { return exp2_mediump(fmul(log2_mediump(x), y)); }
DEVI float linear_to_srgb(float c) // ShaderCore.cpp:673-680
{
	const float lc = fmul(c, 12.92f);
	const float ec = __fmaf_rn(1.055f, pow_mediump(c, 1.0f / 2.4f), -0.055f);
	return c < 0.0031308f ? lc : ec;
}
DEVI float srgb_to_linear(float c) // ShaderCore.cpp:682-689
{
	const float lc = fmul(c, 1.0f / 12.92f);
	const float ec = pow_mediump(__fmaf_rn(c, 1.0f / 1.055f, 0.055f / 1.055f), 2.4f);
	return c < 0.04045f ? lc : ec;
}

// Reactor's scalar Half <-> Float conversions (Reactor.cpp:3744-3770, :3787-3815): round-to-nearest-even on the way down, everything
// above 0x47FFEFFF (incl. NaN) becomes 0x7FFF; on the way up exponent 31 is NOT special (it decodes to 2^16 * 1.m).
DEVI uint32_t float_to_half(float f)
{
	const uint32_t fp32i = __float_as_uint(f);
	uint32_t a = fp32i & 0x7FFFFFFFu;
	uint32_t h = (fp32i & 0x80000000u) >> 16;
	if(a > 0x47FFEFFFu) h |= 0x7FFFu;
	else if(a < 0x38800000u)
	{
		const int mantissa = (int)((a & 0x007FFFFFu) | 0x00800000u);
		const int e = 113 - (int)(a >> 23);
		a = e < 24 ? (uint32_t)(mantissa >> e) : 0u;
		h |= ((a + 0x00000FFFu + ((a >> 13) & 1u)) >> 13) & 0xFFFFu;
	}
	else h |= ((a + 0xC8000000u + 0x00000FFFu + ((a >> 13) & 1u)) >> 13) & 0xFFFFu;
	return h & 0xFFFFu;
}
DEVI float half_to_float(uint32_t h)
{
	int e = (int)(h >> 10) & 0x1F, m = (int)(h & 0x3FFu);
	uint32_t fp32i = (h & 0x8000u) << 16;
	if(e == 0)
	{
		if(m != 0)
		{
			while((m & 0x400) == 0) { m <<= 1; e -= 1; }
			fp32i |= (uint32_t)(((e + (127 - 15) + 1) << 23) | ((m & ~0x400) << 13));
		}
	}
	else fp32i |= (uint32_t)(((e + (127 - 15)) << 23) | (m << 13));
	return __uint_as_float(fp32i);
}

DEVI float blend_apply(uint32_t op, float s, float sf, float dd, float df) // :1849-1958
{
	switch(op)
	{
	case KOP_ADD: return fadd(fmul(s, sf), fmul(dd, df));
	case KOP_SUB: return fsub(fmul(s, sf), fmul(dd, df));
	case KOP_RSUB: return fsub(fmul(dd, df), fmul(s, sf));
	case KOP_MIN: return sse_min(s, dd);
	case KOP_MAX: return sse_max(s, dd);
	case KOP_SRC: return s;
	case KOP_DST: return dd;
	default: return 0.0f;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// k_tile — one CTA per 32x16 screen tile = four INDEPENDENT warps, one per 16x8 region (= one triangle bin) of it
//
//   * a warp stages the colour / depth / stencil planes of ITS region in shared memory with one TMA load per attachment
//     (cp.async.bulk.tensor.3d: x, y, sample plane) behind its own mbarrier, keeps them there for the whole bin and writes them
//     back with one TMA store per attachment; there is no CTA barrier anywhere, a warp with an empty bin leaves at once
//     (fallback: lane copies when an attachment does not satisfy the tensor-map alignment rules, or when the scissor rows cut
//     through the region — a band of a multi-GPU frame must not store rows that belong to another rank);
//   * the bin's entries come in no particular order (k_fill hands the slots out with atomics): the warp first puts them in
//     triangle order — the API order of the fragments — in registers (<= 128 entries), in shared memory (<= SWCU_SORT_CAP) or, a long
//     bin, in place in global memory;
//   * coverage, 32 bin entries at a time, one per lane.  SMALL triangle: the lane clips the coverage masks of the record to
//     the region (a few logic ops per row) and counts the bits.  BIG triangle: one (row, sample) per lane evaluates the
//     reference's span in closed form (edge_at_row) from the polygon in the big list;
//   * a warp scan places every lane's items; the lanes then WRITE their covered samples (`lane | sample | row | x`, 16 bits) into
//     the warp's item queue in list order — producer-side expansion: no pair lists, no start marks, no search by the consumers;
//   * the queue is consumed 32 items per round, one covered sample per lane, so lane utilisation does not depend on triangle
//     size.  Two fragments of one round on the same sample (overlapping triangles) are detected with an owner byte per
//     sample; only then are the items of the round serialised by __match_any_sync rank (list order);
//   * specialised on <samples, fragment shader class, blend class, fast state>; everything else is warp-uniform run-time state.
// ------------------------------------------------------------------------------------------------------------------
#define TILE_THREADS (SWCU_TILE_WARPS * 32)
#define REGION_PX (SWCU_REGION_W * SWCU_REGION_H)

template<int SH> struct ShaderSlots { static constexpr int N = SH == SH_CONST ? 0 : SH == SH_VARY ? 4 : SH == SH_TEX ? 2 : 6; };

// ---- TMA / mbarrier primitives (PTX ISA 8.x, sm_90+) ----
DEVI uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVI void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
DEVI void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
DEVI bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
DEVI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DEVI void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	             ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
DEVI void tma_store_3d(const CUtensorMap *map, const void *src, int x, int y, int z)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z) : "memory");
}
DEVI void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DEVI void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// shared-memory layout of one region warp (dynamic shared memory; the host computes the same size)
template<int MS, int SH>
struct TileLayout
{
	static constexpr int FRONT4 = (TRI_FLOATS_FRONT + 3 * ShaderSlots<SH>::N + 3) / 4; // float4s of the front plane block
	static constexpr int PLANE_B = REGION_PX * 4 * MS;                                 // one 32-bit-per-sample plane of the region
	static constexpr int STENCIL_B = REGION_PX * MS;
	static constexpr int ICAP = MS == 4 ? 512 : 1024;   // items (covered samples) per queue fill; one region holds at most 128 * MS
	static constexpr int Q_B = 2 * ICAP;                // uint16 queue[ICAP]
	static constexpr int OWNER_B = REGION_PX * MS;      // uint8 owner[sample of the region]
	static constexpr int SORT_B = 4 * SWCU_SORT_CAP;    // uint32 sorted[SWCU_SORT_CAP]
	static constexpr int HEAD_B = 128;                  // the warp's mbarrier
	// Per warp, at compile-time offsets: mbarrier | queue | owner | sorted | colour plane (one word per sample).  What depends on the
	// draw's state — the further colour words of a floating-point target, the depth plane, the stencil plane — lies behind the four
	// fixed parts, so the hot pointers need no run-time arithmetic (the compiler re-derived them inside the item loop otherwise).
	static constexpr int FIXED_B = HEAD_B + Q_B + OWNER_B + SORT_B + PLANE_B;
	__host__ __device__ static int var_bytes(bool depth, bool stencil, int colorEpp)
	{
		return (colorEpp > 1 ? PLANE_B * colorEpp : 0) + (depth ? PLANE_B : 0) + (stencil ? ((STENCIL_B + 127) & ~127) : 0);
	}
	__host__ __device__ static int total(bool depth, bool stencil, int colorEpp = 1) { return SWCU_TILE_WARPS * (FIXED_B + var_bytes(depth, stencil, colorEpp)); }
};

// writeColor for the RGBA8 family (PixelRoutine.cpp:1981-1992, :2603-2655): RoundInt(clamp(c, 0, 1) * 255) per channel,
// packed with saturation.  The float clamp is folded into the integer saturation: values above 1 round to >= 255, negative
// ones to <= 0 and NaN converts to 0, exactly what min(max(c, 0), 1) gives before the conversion.
DEVI uint32_t pack_unorm8(float b0, float b1, float b2, float b3)
{
	const int i0 = __float2int_rn(fmul(b0, 255.0f)), i1 = __float2int_rn(fmul(b1, 255.0f));
	const int i2 = __float2int_rn(fmul(b2, 255.0f)), i3 = __float2int_rn(fmul(b3, 255.0f));
	uint32_t hi, pk;
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(i3), "r"(i2), "r"(0));
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(pk) : "r"(i1), "r"(i0), "r"(hi));
	return pk;
}

// region <-> framebuffer copy by the lanes of one warp (TMA-ineligible attachments, regions cut by the scissor rows):
// rows [y0, y1) of the region only
template<int MS, typename T, bool STORE>
DEVI void region_copy(T *sm, unsigned char *base, int pitchB, int sliceB, int rx, int ry, int fbW, int y0, int y1, int lane)
{
	for(int i = lane; i < MS * REGION_PX; i += 32)
	{
		const int q = i / REGION_PX, r = (i / SWCU_REGION_W) % SWCU_REGION_H, c = i % SWCU_REGION_W;
		const int y = ry + r, x = rx + c;
		if(y < y0 || y >= y1 || x >= fbW) continue;
		T *g = (T *)(base + (size_t)q * sliceB + (size_t)y * pitchB) + x;
		if(STORE) *g = sm[i]; else sm[i] = *g;
	}
}

// P(x, y) of one plane slot: QuadRasterizer::interpolate (QuadRasterizer.cpp:235-250) on top of the row constant
// D = C + y*B (:158-163, unfused) — MulAdd(x, A, D) is the one fused op — then * rhw for perspective-correct slots.
DEVI float interp_slot(float A, float B, float C, uint32_t mode, float xf, float yf, float rhw)
{
	if(mode == IM_FLAT) return C;
	float t = __fmaf_rn(xf, A, fadd(C, fmul(yf, B)));
	if(mode == IM_PERSP) t = fmul(t, rhw);
	return t;
}

struct TileMaps
{
	CUtensorMap color, depth, stencil; // boxes of one region: (16 pixels, 8 rows, all sample planes)
};

DEVI uint32_t byte_range(int a, int b) { return b > a ? (0xFFFFFFFFu >> (32 - 8 * (b - a))) << (8 * a) : 0u; } // bytes [a, b) of a word, 0 <= a, b <= 4

// FS ("fast state"): the host has checked the common fixed-function state — no stencil, full colour write mask, RGBA byte
// order, no depth bias, full sample mask, depth test off or LESS / LESS_OR_EQUAL, perspective slots routed one to one —
// so none of it is decoded per fragment.  FS == false is the same code with every state read at run time.
#ifndef TILE_CTAS_4X
#define TILE_CTAS_4X 7 // 72 registers: measured 0.357 ms on C4 against 0.385 ms with 8 CTAs of 64 (spills and re-derived addresses in the item loop)
#endif
#ifndef TILE_CTAS_1X
#define TILE_CTAS_1X 7
#endif
#ifndef TILE_MAP_1X
#define TILE_MAP_1X 0
#endif
#ifndef TILE_MAP_4X
#define TILE_MAP_4X 1
#endif

template<int MS, int SH, int BL, bool FS>
__global__ void __launch_bounds__(TILE_THREADS, MS == 4 ? TILE_CTAS_4X : TILE_CTAS_1X) k_tile(const __grid_constant__ DrawConst d, const __grid_constant__ TileMaps maps)
{
	using L = TileLayout<MS, SH>;
	constexpr int FRONT4 = L::FRONT4, ICAP = L::ICAP;
	constexpr bool TEX = SH == SH_TEX || SH == SH_GENERIC;
	constexpr int UV = SH == SH_TEX ? 0 : 4;
	constexpr int BIG_GROUP = MS == 4 ? 1 : 4; // big triangles evaluated together: 8 * MS (row, sample) lanes each
	extern __shared__ __align__(128) unsigned char smem[];

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t laneLt = (1u << lane) - 1u;
	if(blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && d.hostCounters)
	{
		// the counters of the draw's setup phase, for swcu_sync to look at (through the mapped page: a cudaMemcpyAsync would queue in the
		// device-to-host copy engine behind a frame download)
		*d.hostCounters = *d.counters;
		__threadfence_system();
	}
	const int tx = d.tileX0 + blockIdx.x, ty = d.tileY0 + blockIdx.y;
	const uint32_t bin = (uint32_t)(ty * d.tilesX + tx) * 4u + (uint32_t)warp;
	uint32_t begin = 0, n = d.primCount;
	if(!d.direct)
	{
		begin = d.binStart[bin];
		n = d.binStart[bin + 1] - begin;
	}
	if(n == 0) return;
	const int rx = tx * SWCU_TILE_W + (warp & 1) * SWCU_REGION_W; // this warp's region
	const int ry = ty * SWCU_TILE_H + (warp >> 1) * SWCU_REGION_H;
	const int ryLo = max(ry, d.scY0), ryHi = min(ry + SWCU_REGION_H, d.scY1); // rows of the region inside the scissor
	if(ryLo >= ryHi || rx >= d.scX1 || rx + SWCU_REGION_W <= d.scX0) return;

	const bool colorOn = FS ? true : (d.colorWriteMask != 0 && d.colorBuf != nullptr);
	// WRITE-ONLY draws (fast state, no blending, no depth test — the host has checked): no attachment is read, so the region is not
	// staged at all; a fragment's colour goes straight to the framebuffer, in the order the rounds apply them (the lanes of a warp
	// are ordered by the __syncwarp between the rounds), and nothing is written back at the end
	const bool wo = (FS && BL == BL_OFF) ? d.writeOnly != 0 : false;
	const int colorEpp = FS ? 1 : (int)d.colorEpp; // 32-bit words per colour pixel (floating-point targets: 2 or 4)
	unsigned char *wa = smem + warp * L::FIXED_B;
	uint64_t *bar = (uint64_t *)wa;
	unsigned short *wQueue = (unsigned short *)(wa + L::HEAD_B);
	unsigned char *wOwner = wa + L::HEAD_B + L::Q_B;
	// The first words of the owner array double as the COVERAGE MAP of a queue fill: one bit per sample of the region (word 2 * row:
	// samples 0 | 1 << 16, word 2 * row + 1: samples 2 | 3 << 16, a bit per column; 1x: word row / 2, half row & 1).  Every producer ORs
	// its clipped masks in; a bit found set already means two fragments of the fill meet on a sample.  Only then do the rounds of the
	// fill run the exact per-round test (owner bytes, below) — a mesh without overdraw never does.
	uint32_t *wBits = (uint32_t *)wOwner;
	// (1x: measured slower than signing the owner bytes round by round — 32 lanes meet on the 4 words of the map — so one sample per
	// pixel keeps the per-round test, and its optimistic pass the signatures)
	constexpr bool USE_MAP = MS == 4 ? TILE_MAP_4X : TILE_MAP_1X;
	uint32_t *wSort = (uint32_t *)(wa + L::HEAD_B + L::Q_B + L::OWNER_B);
	// the colour plane; a floating-point target (2 or 4 words per pixel, never with FS) lives in the variable part instead
	unsigned char *wv = smem + SWCU_TILE_WARPS * L::FIXED_B + warp * L::var_bytes(d.depthTestActive != 0, d.stencilActive != 0, colorEpp);
	uint32_t *smColor = (FS || colorEpp == 1) ? (uint32_t *)(wa + L::HEAD_B + L::Q_B + L::OWNER_B + L::SORT_B) : (uint32_t *)wv;
	float *smDepth = (float *)(wv + ((FS || colorEpp == 1) ? 0 : L::PLANE_B * colorEpp));
	unsigned char *smStencil = (unsigned char *)smDepth + (d.depthTestActive ? L::PLANE_B : 0);

	// ---- stage the region: TMA when the attachments allow it.  (A second time if the optimistic pass below has to be redone:
	//      nothing has been stored by then, the attachments still hold what they held.) ----
	bool tileReady = true;
	uint32_t tmaPhase = 0;
	auto stage_region = [&](bool first) {
		if(wo) return;
		if(d.useTma)
		{
			if(lane == 0)
			{
				if(first)
				{
					mbar_init(bar, 1);
					asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
				}
				const uint32_t bytes = (colorOn ? L::PLANE_B * colorEpp : 0) + (d.depthTestActive ? (d.depth16 ? L::PLANE_B / 2 : L::PLANE_B) : 0) + (d.stencilActive ? L::STENCIL_B : 0);
				mbar_expect_tx(bar, bytes);
				if(colorOn) tma_load_3d(smColor, &maps.color, bar, rx * colorEpp, ry, 0); // the map counts 32-bit words along x
				if(d.depthTestActive) tma_load_3d(smDepth, &maps.depth, bar, rx, ry, 0);
				if(d.stencilActive) tma_load_3d(smStencil, &maps.stencil, bar, rx, ry, 0);
			}
			__syncwarp();
			tileReady = false;
		}
		else
		{
			const int yEnd = min(ry + SWCU_REGION_H, d.fbHeight);
			if(colorOn && colorEpp == 4) region_copy<MS, uint4, false>((uint4 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, rx, ry, d.fbWidth, ry, yEnd, lane);
			else if(colorOn && colorEpp == 2) region_copy<MS, uint2, false>((uint2 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, rx, ry, d.fbWidth, ry, yEnd, lane);
			else if(colorOn) region_copy<MS, uint32_t, false>(smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, rx, ry, d.fbWidth, ry, yEnd, lane);
			if(d.depthTestActive && d.depth16) region_copy<MS, unsigned short, false>((unsigned short *)smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, rx, ry, d.fbWidth, ry, yEnd, lane);
			else if(d.depthTestActive) region_copy<MS, float, false>(smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, rx, ry, d.fbWidth, ry, yEnd, lane);
			if(d.stencilActive) region_copy<MS, unsigned char, false>(smStencil, d.stencilBuf, d.stencilPitchB, d.stencilSliceB, rx, ry, d.fbWidth, ry, yEnd, lane);
			__syncwarp();
		}
	};
	stage_region(true);

	bool dirty = false;
	uint32_t constChan = 0; // colour channels that are shader constants
#pragma unroll
	for(int ch = 0; ch < 4; ch++)
		if(d.chanKind[ch] == CK_CONST) constChan |= 1u << ch;
	const bool biasOn = FS ? false : d.depthBiasEnable != 0;
	const bool bgr = FS ? false : d.bgr != 0;
	uint32_t wmask32 = FS ? 0xFFFFFFFFu : 0u; // byte lanes of the packed pixel the draw may write
	uint32_t sampleBytes = 0xFFFFFFFFu;       // 4x: byte q of a mask word = sample q enabled
	if(!FS)
	{
#pragma unroll
		for(int ch = 0; ch < 4; ch++)
			if((d.colorWriteMask >> ch) & 1) wmask32 |= 0xFFu << (8 * ((bgr && ch < 3) ? 2 - ch : ch));
		if(MS == 4)
		{
			sampleBytes = 0;
#pragma unroll
			for(int q = 0; q < 4; q++)
				if((d.sampleMask >> q) & 1) sampleBytes |= 0xFFu << (8 * q);
		}
	}

	// ---- the bin in triangle order — or, 1x, optimistically NOT: the order of the fragments only matters where two of them land on
	//      the same sample.  A bin of a mesh without overdraw is walked as k_fill left it, every sample signing the owner array
	//      as it is written; the first sample found signed already (or signed twice in one round) abandons the pass: the region is
	//      staged again — nothing has been stored — and the bin is walked once more, sorted.  Order-free bins skip the sort; bins
	//      with overdraw pay for the part of the first pass they got through ----
	const uint32_t *list = d.pairs + begin;
	uint32_t regId = 0xFFFFFFFFu;
	bool sortedInSmem = false, listInPlace = false;
	const bool optimisticBin = MS == 1 && !d.direct && n > 32 && n <= SWCU_SORT_CAP;
	bool redo = false;
	for(int pass = 0; pass < 2; pass++)
	{
	const bool optimistic = optimisticBin && pass == 0;
	if(pass == 1)
	{
		if(!tileReady)
			while(!mbar_try_wait(bar, tmaPhase)) {}
		tmaPhase ^= 1u;
		stage_region(false);
		dirty = false;
		redo = false;
	}
	if(optimistic)
	{
		if(USE_MAP) { if(lane < 4) wBits[lane] = 0u; } // the pass keeps ONE map of the samples written so far: 8 rows x 16 columns
		else ((uint32_t *)wOwner)[lane] = 0xFFFFFFFFu; // 128 samples, nobody has signed yet
		__syncwarp();
	}
	if(!d.direct && !optimistic)
	{
		if(n <= 32)
		{
			if((uint32_t)lane < n) regId = __ldg(list + lane);
#pragma unroll
			for(int k = 2; k <= 32; k <<= 1)
#pragma unroll
				for(int j = k >> 1; j > 0; j >>= 1)
				{
					const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, regId, j);
					const bool keepMin = ((lane & j) == 0) == ((lane & k) == 0);
					regId = keepMin ? min(regId, o) : max(regId, o);
				}
		}
		else if(n <= 128)
		{
			// four entries per lane (element 4 * lane + i), bitonic network in registers: partners closer than 4 are in the same lane,
			// the others one shuffle away; the sorted entries go to the warp's shared-memory list
			uint32_t v[4];
#pragma unroll
			for(int i = 0; i < 4; i++) v[i] = (uint32_t)(4 * lane + i) < n ? __ldg(list + 4 * lane + i) : 0xFFFFFFFFu;
#pragma unroll
			for(int k = 2; k <= 128; k <<= 1)
#pragma unroll
				for(int j = k >> 1; j > 0; j >>= 1)
				{
					if(j >= 4)
					{
#pragma unroll
						for(int i = 0; i < 4; i++)
						{
							const int e = 4 * lane + i;
							const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v[i], j >> 2);
							const bool keepMin = ((e & j) == 0) == ((e & k) == 0);
							v[i] = keepMin ? min(v[i], o) : max(v[i], o);
						}
					}
					else
					{
#pragma unroll
						for(int i = 0; i < 4; i++)
						{
							if(i & j) continue; // (i, i | j) is a comparator inside the lane
							const int e = 4 * lane + i;
							const bool up = (e & k) == 0;
							const uint32_t lo = min(v[i], v[i | j]), hi = max(v[i], v[i | j]);
							v[i] = up ? lo : hi;
							v[i | j] = up ? hi : lo;
						}
					}
				}
			*(uint4 *)(wSort + 4 * lane) = make_uint4(v[0], v[1], v[2], v[3]);
			__syncwarp();
			sortedInSmem = true;
		}
		else if(n <= SWCU_SORT_CAP)
		{
			for(uint32_t i = lane; i < n; i += 32) wSort[i] = __ldg(list + i);
			__syncwarp();
			bitonic_sort_any(wSort, n, (uint32_t)lane, 32u, [] { __syncwarp(); });
			sortedInSmem = true;
		}
		else
		{
			// a long bin (heavy overdraw): ordered in place in global memory — the segment belongs to this warp alone
			bitonic_sort_any(d.pairs + begin, n, (uint32_t)lane, 32u, [] { __syncwarp(); });
			listInPlace = true;
		}
	}

	// ---- 32 bin entries at a time, one per lane; the ids and record headers of the next 32 are requested before the current ones
	//      are used, and the plane equations of every entry are pulled into L2 ahead of the items that will read them ----
	auto fetch_block = [&](uint32_t pos0, uint32_t &idOut, uint4 &hOut) {
		const uint32_t li = pos0 + lane;
		idOut = 0;
		hOut = make_uint4(0, 0, 0, 0);
		if(li < n)
		{
			idOut = d.direct ? li : (n <= 32 ? regId : (sortedInSmem ? wSort[li] : (listInPlace ? d.pairs[begin + li] : __ldg(list + li))));
			const unsigned char *r = d.triRecords + (size_t)idOut * d.triStride;
			hOut = __ldg((const uint4 *)r);
			asm volatile("prefetch.global.L2 [%0];" ::"l"(r + d.planeOffset));
			asm volatile("prefetch.global.L2 [%0];" ::"l"(r + d.triStride - 16));
		}
	};
	uint32_t idN;
	uint4 hN;
	fetch_block(0, idN, hN);
	for(uint32_t pos0 = 0; pos0 < n && !redo; pos0 += 32)
	{
		const uint32_t li = pos0 + lane;
		bool valid = li < n;
		const uint32_t id = idN;
		const uint4 h = hN;
		if(pos0 + 32 < n) fetch_block(pos0 + 32, idN, hN);
		const unsigned char *rec = d.triRecords + (size_t)id * d.triStride;
		bool isBig = valid && (h.y & TRI_FLAG_BIG);
		if(valid)
		{
			// does the entry touch my region at all?  (always asked in direct mode; bins are conservative too)
			if(isBig)
			{
				const int pxMin = h.x & 0xFFFF, pxMax = h.x >> 16, yMin = h.z & 0xFFFF, yMax = h.z >> 16;
				if(!(pxMin < rx + SWCU_REGION_W && pxMax > rx && yMin < ryHi && yMax > ryLo)) { valid = false; isBig = false; }
			}
			else
			{
				const int fx = h.x & 0xFFFF, fy = h.x >> 16;
				if(!(fx < rx + SWCU_REGION_W && fx + SWCU_SMALL_COLS > rx && fy < ryHi && fy + SWCU_SMALL_ROWS > ryLo)) valid = false;
			}
		}
		const uint32_t bigMask = __ballot_sync(0xFFFFFFFFu, isBig);
		const int cnt = (int)min(32u, n - pos0);
		int p = 0;
		while(p < cnt && !redo)
		{
			uint32_t total = 0;
			bool oneCandidate = false; // no two items of this queue fill can lie on the same sample (one triangle, or the coverage map says so)
			bool mapHit = false;       // optimistic pass: an item of this fill lands on a sample an earlier fragment of the pass has written
			const uint32_t rest = bigMask >> p;
			if(rest & 1u)
			{
				if(MS == 1)
				{
					// ---- BIG, 1x: one entry; lane -> (row, edge slot): the four lanes of a row evaluate the polygon's edges 4 at a time in
					//      closed form (SetupRoutine::edge as edge_at_row) and keep, for the left and the right end of the span, the hit
					//      of the LAST edge that owns the row (the reference's span table is written edge after edge) ----
					const int row = lane >> 2, es = lane & 3;
					const uint32_t bslot = __shfl_sync(0xFFFFFFFFu, h.w, p);
					const int y = ry + row;
					int Lx = 0, Li = -1, Rx = 0, Ri = -1;
					if(y >= ryLo && y < ryHi)
					{
						const BigTri &b = d.bigList[bslot];
						if(y >= b.yMin && y < b.yMax)
						{
							const int bn = b.n, bdir = b.dir;
							for(int i = es; i < bn; i += 4)
							{
								const int ia = i + 1 - bdir, ib = i + bdir;
								const int va = ia == bn ? 0 : ia, vb = ib == bn ? 0 : ib;
								bool right; int x;
								if(edge_at_row(d, b.X[va], b.Y[va], b.X[vb], b.Y[vb], y, right, x)) { if(right) { Rx = x; Ri = i; } else { Lx = x; Li = i; } }
							}
						}
					}
#pragma unroll
					for(int o = 1; o <= 2; o <<= 1)
					{
						const int oLx = __shfl_xor_sync(0xFFFFFFFFu, Lx, o), oLi = __shfl_xor_sync(0xFFFFFFFFu, Li, o);
						const int oRx = __shfl_xor_sync(0xFFFFFFFFu, Rx, o), oRi = __shfl_xor_sync(0xFFFFFFFFu, Ri, o);
						if(oLi > Li) { Li = oLi; Lx = oLx; }
						if(oRi > Ri) { Ri = oRi; Rx = oRx; }
					}
					// (a side no edge owns keeps 0, the value the reference's cleared span entry holds)
					const int a = clampi(Lx - rx, 0, SWCU_REGION_W), e = clampi(Rx - rx, 0, SWCU_REGION_W);
					const int nrow = e > a ? e - a : 0;       // the same in the four lanes of the row
					uint32_t incl = es == 0 ? (uint32_t)nrow : 0u;
#pragma unroll
					for(int o = 1; o < 32; o <<= 1)
					{
						const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
						if(lane >= o) incl += t;
					}
					total = __shfl_sync(0xFFFFFFFFu, incl, 31); // <= 128
					const uint32_t start = __shfl_sync(0xFFFFFFFFu, incl, lane & ~3) - (uint32_t)nrow; // first item of my row
					const uint32_t item = ((uint32_t)p << 11) | (1u << 7) | ((uint32_t)row << 4);
					for(int i = es; i < nrow; i += 4) wQueue[start + i] = (unsigned short)(item | (uint32_t)(a + i)); // every fourth pixel of the run
					if(USE_MAP && optimistic && es == 0 && nrow)
					{
						const uint32_t v = ((0xFFFFu >> (16 - nrow)) << a) << (16 * (row & 1));
						if(atomicOr(wBits + (row >> 1), v) & v) mapHit = true;
					}
					p += 1;
					oneCandidate = true;
				}
				else
				{
					// ---- BIG: up to BIG_GROUP consecutive big entries; lane -> (entry, row, sample); the span of the row in closed form
					//      (SetupRoutine::edge as edge_at_row: the last edge that owns the row wins; 4x: pre-filled, SetupRoutine.cpp:214-225) ----
					int group = 1;
					if(BIG_GROUP > 1) group = min(BIG_GROUP, rest == 0xFFFFFFFFu ? 32 : (int)__ffs(~rest) - 1);
					const int c = lane / (SWCU_REGION_H * MS), row = (lane / MS) % SWCU_REGION_H, q = lane % MS;
					const int src = min(p + c, 31);
					const uint32_t bslot = __shfl_sync(0xFFFFFFFFu, h.w, src);
					int a = 0, e = 0;
					const int y = ry + row;
					if(c < group && y >= ryLo && y < ryHi && (MS == 1 || FS || ((d.sampleMask >> q) & 1)))
					{
						const BigTri &b = d.bigList[bslot];
						if(y >= b.yMin && y < b.yMax)
						{
							const int ox = MS > 1 ? c_Xf[q] : 0, oy = MS > 1 ? c_Yf[q] : 0;
							int Ls = 0, Rs = 0;
							if(MS > 1) Ls = Rs = clampi((int)((uint32_t)b.X[0] + 255u) >> 8, d.scX0, d.scX1);
							const int bn = b.n, bdir = b.dir;
							for(int i = 0; i < bn; i++)
							{
								const int ia = i + 1 - bdir, ib = i + bdir;
								const int va = ia == bn ? 0 : ia, vb = ib == bn ? 0 : ib;
								bool right; int x;
								if(edge_at_row(d, b.X[va] - ox, b.Y[va] - oy, b.X[vb] - ox, b.Y[vb] - oy, y, right, x)) { if(right) Rs = x; else Ls = x; }
							}
							a = clampi(Ls - rx, 0, SWCU_REGION_W); e = clampi(Rs - rx, 0, SWCU_REGION_W);
						}
					}
					// pixel items: at 4x the four sample lanes of a row exchange their runs; lane q of the row takes the columns [4q, 4q + 4)
					const uint32_t run = e > a ? ((1u << e) - 1u) & ~((1u << a) - 1u) : 0u; // columns of the region my (row, sample) covers
					uint32_t cols, sm0 = run, sm1 = 0, sm2 = 0, sm3 = 0;
					if(MS == 4)
					{
						const uint32_t r1 = __shfl_xor_sync(0xFFFFFFFFu, run, 1), r2 = __shfl_xor_sync(0xFFFFFFFFu, run, 2), r3 = __shfl_xor_sync(0xFFFFFFFFu, run, 3);
						// runs by SAMPLE number: lane q holds sample q, its partner under xor j holds sample q ^ j
						sm0 = q == 0 ? run : q == 1 ? r1 : q == 2 ? r2 : r3;
						sm1 = q == 1 ? run : q == 0 ? r1 : q == 3 ? r2 : r3;
						sm2 = q == 2 ? run : q == 3 ? r1 : q == 0 ? r2 : r3;
						sm3 = q == 3 ? run : q == 2 ? r1 : q == 1 ? r2 : r3;
						cols = (sm0 | sm1 | sm2 | sm3) & (0xFu << (4 * q));
					}
					else cols = run;
					const int nn = __popc(cols);
					uint32_t incl = (uint32_t)nn;
	#pragma unroll
					for(int o = 1; o < 32; o <<= 1)
					{
						const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
						if(lane >= o) incl += t;
					}
					total = __shfl_sync(0xFFFFFFFFu, incl, 31); // <= 128 pixels (4x) / 32 runs of <= 16 pixels (1x): never more than ICAP
					uint32_t w = incl - (uint32_t)nn;
					const uint32_t item = ((uint32_t)(p + c) << 11) | ((uint32_t)row << 4);
					while(cols)
					{
						const uint32_t x = (uint32_t)__ffs(cols) - 1u;
						cols &= cols - 1u;
						const uint32_t sm = MS == 4 ? (((sm0 >> x) & 1u) | (((sm1 >> x) & 1u) << 1) | (((sm2 >> x) & 1u) << 2) | (((sm3 >> x) & 1u) << 3)) : 1u;
						wQueue[w++] = (unsigned short)(item | (sm << 7) | x);
					}
				p += group;
				oneCandidate = group == 1;
				}
			}
			else
			{
				// ---- SMALL: lanes [p, hi) clip the coverage masks of their records to the region and count the bits ----
				const int hi = rest ? p + (int)__ffs(rest) - 1 : cnt;
				const bool mine = valid && lane >= p && lane < hi;
				const int fx = h.x & 0xFFFF, fy = h.x >> 16;
				uint32_t m[MS == 4 ? 8 : 2];
				uint32_t count = 0;
				if(mine)
				{
					const int rlo = clampi(ryLo - fy, 0, SWCU_SMALL_ROWS), rhi = clampi(ryHi - fy, 0, SWCU_SMALL_ROWS);
					const int clo = clampi(rx - fx, 0, SWCU_SMALL_COLS), chi = clampi(rx + SWCU_REGION_W - fx, 0, SWCU_SMALL_COLS);
					const uint32_t cm4 = (chi > clo ? (((1u << chi) - 1u) & ~((1u << clo) - 1u)) : 0u) * 0x01010101u;
					if(MS == 4)
					{
						const uint4 wa4 = __ldg((const uint4 *)(rec + TRI_HEADER_BYTES)), wb4 = __ldg((const uint4 *)(rec + TRI_HEADER_BYTES + 16));
						m[0] = wa4.x; m[1] = wa4.y; m[MS == 4 ? 2 : 0] = wa4.z; m[MS == 4 ? 3 : 0] = wa4.w;
						m[MS == 4 ? 4 : 0] = wb4.x; m[MS == 4 ? 5 : 0] = wb4.y; m[MS == 4 ? 6 : 0] = wb4.z; m[MS == 4 ? 7 : 0] = wb4.w;
						const uint32_t keep = cm4 & sampleBytes;
#pragma unroll
						for(int r = 0; r < (MS == 4 ? 8 : 0); r++)
						{
							m[r] = (r >= rlo && r < rhi) ? (m[r] & keep) : 0u;
							// items are PIXELS: a column counts once however many of its samples are covered
							count += __popc((m[r] | (m[r] >> 8) | (m[r] >> 16) | (m[r] >> 24)) & 0xFFu);
						}
					}
					else
					{
						m[0] = h.z & cm4 & byte_range(min(rlo, 4), min(rhi, 4));
						m[1] = h.w & cm4 & byte_range(max(rlo - 4, 0), max(rhi - 4, 0));
						count = __popc(m[0]) + __popc(m[1]);
					}
				}
				uint32_t incl = count;
#pragma unroll
				for(int o = 1; o < 32; o <<= 1)
				{
					const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
					if(lane >= o) incl += t;
				}
				// the lanes whose items still fit the queue: a prefix of [p, hi) (one small triangle always fits)
				const uint32_t fitMask = __ballot_sync(0xFFFFFFFFu, lane >= p && lane < hi && incl <= (uint32_t)ICAP);
				const int fitEnd = p + __popc(fitMask);
				total = __shfl_sync(0xFFFFFFFFu, incl, fitEnd - 1);
				if(mine && lane < fitEnd && count)
				{
					// producer-side expansion: item = lane << 11 | sample mask << 7 | region row << 4 | region column
					uint32_t w = incl - count;
					const uint32_t item0 = (uint32_t)((lane << 11) + (fy - ry) * 16 + (fx - rx));
					if(MS == 4)
					{
#pragma unroll
						for(int r = 0; r < (MS == 4 ? 8 : 0); r++)
						{
							const uint32_t mr = m[r];
							uint32_t cols = (mr | (mr >> 8) | (mr >> 16) | (mr >> 24)) & 0xFFu;
							while(cols)
							{
								const uint32_t c = (uint32_t)__ffs(cols) - 1u;
								cols &= cols - 1u;
								// bit c of the four sample bytes -> a 4-bit sample mask (the partial products of the multiply do not collide)
								const uint32_t sm = (((mr >> c) & 0x01010101u) * 0x01020408u) >> 24;
								wQueue[w++] = (unsigned short)(item0 + 16u * r + c + (sm << 7));
							}
						}
					}
					else
					{
#pragma unroll
						for(int j = 0; j < 2; j++)
						{
							uint32_t bits = m[j];
							while(bits)
							{
								const uint32_t b = (uint32_t)__ffs(bits) - 1u;
								bits &= bits - 1u;
								wQueue[w++] = (unsigned short)(item0 + (1u << 7) + 64u * j + ((b & 0x38u) << 1) + (b & 7u));
							}
						}
					}
				}
				// ---- coverage map of the fill: do two of its fragments meet on a sample?  (Not asked when the fill holds one triangle.) ----
				if(fitEnd - p == 1 && !optimistic) oneCandidate = true;
				else if(USE_MAP)
				{
					if(!optimistic)
					{
						if(lane < (MS == 4 ? 16 : 4)) wBits[lane] = 0u;
						__syncwarp();
					}
					bool hit = false;
					if(mine && lane < fitEnd && count)
					{
						const int dx = fx - rx, dy = fy - ry;
						const int sl = max(dx, 0), sr = max(-dx, 0); // (columns left of the region are clipped away: nothing is shifted out)
						if(MS == 4)
						{
#pragma unroll
							for(int r = 0; r < (MS == 4 ? 8 : 0); r++)
							{
								const uint32_t mr = m[r];
								if(mr)
								{
									const uint32_t lo = (__byte_perm(mr, 0, 0x4140) << sl) >> sr, hi2 = (__byte_perm(mr, 0, 0x4342) << sl) >> sr;
									uint32_t *w = wBits + 2 * (dy + r);
									if(lo && (atomicOr(w, lo) & lo)) hit = true;
									if(hi2 && (atomicOr(w + 1, hi2) & hi2)) hit = true;
								}
							}
						}
						else
						{
#pragma unroll
							for(int r = 0; r < 8; r++)
							{
								const uint32_t byte = (m[r >> 2] >> (8 * (r & 3))) & 0xFFu;
								if(byte)
								{
									const int R = dy + r;
									const uint32_t v = ((byte << sl) >> sr) << (16 * (R & 1));
									if(atomicOr(wBits + (R >> 1), v) & v) hit = true;
								}
							}
						}
					}
					if(optimistic) mapHit = hit;
					else oneCandidate = !__any_sync(0xFFFFFFFFu, hit);
				}
				p = fitEnd;
			}
			__syncwarp();
			if(USE_MAP && optimistic && __any_sync(0xFFFFFFFFu, mapHit))
			{
				redo = true; // (warp-uniform) nothing of this fill has been applied: the region is staged again and the bin walked in order
				break;
			}
			if(total && !tileReady)
			{
				while(!mbar_try_wait(bar, tmaPhase)) {} // the TMA loads of the region have landed
				tileReady = true;
			}
			// ---- consume the items 32 at a time, one covered PIXEL per lane: interpolation and the routed shader run once per pixel
			//      (PixelRoutine.cpp:196-261), stencil / depth / blend / write once per covered sample of it (:124-148, :262-358) ----
			for(uint32_t base = 0; base < total && !redo; base += 32)
			{
				const uint32_t g = base + lane;
				const bool live = g < total;
				const uint32_t e = live ? (uint32_t)wQueue[g] : 0u;
				const int k = (int)(e >> 11);
				const uint32_t smask = MS == 4 ? (e >> 7) & 15u : 1u; // covered samples of the pixel
				const uint32_t pkey = e & 0x7Fu;                        // pixel of the region: row << 4 | column
				const uint32_t idk = __shfl_sync(0xFFFFFFFFu, id, k);
				const uint32_t flagsk = FS ? 0u : __shfl_sync(0xFFFFFFFFu, h.y, k);
				// ---- plane equations of the item's triangle, requested before the conflict test below (the items of a triangle sit next
				//      to each other in the queue, so a round reads two or three records) ----
				const float4 *pl = (const float4 *)(d.triRecords + (size_t)idk * d.triStride + d.planeOffset);
				float pf[FRONT4 * 4];
				float4 zv = make_float4(0, 0, 0, 0); // zBias, zA, zB, zC
				if(live)
				{
#pragma unroll
					for(int t = 0; t < FRONT4; t++)
					{
						const float4 v = __ldg(pl + t);
						pf[4 * t] = v.x; pf[4 * t + 1] = v.y; pf[4 * t + 2] = v.z; pf[4 * t + 3] = v.w;
					}
					if(d.depthTestActive) zv = __ldg(pl + FRONT4);
				}
				// two fragments of this round on one SAMPLE?  Every lane signs its samples; a lane that reads back another signature has
				// company.  (Two triangles that share an edge pixel with disjoint samples — every edge of a mesh — are not a conflict.)
				bool shared = false;
				// (an optimistic pass with the map has asked it about the whole pass; one without still has to sign the samples for the rounds to come)
				const bool noCheck = USE_MAP ? (oneCandidate || optimistic) : (oneCandidate && !optimistic);
				if(!noCheck)
				{
				if(!USE_MAP && MS == 1 && optimistic)
				{
					// signed in an earlier round of this pass?  Then two fragments meet on the sample and the walk needs the order
					if(live && wOwner[pkey] != 0xFFu) shared = true;
					__syncwarp();
				}
				if(live)
				{
#pragma unroll
					for(int q = 0; q < MS; q++)
						if((smask >> q) & 1u) wOwner[q * REGION_PX + pkey] = (unsigned char)lane;
				}
				__syncwarp();
				if(live)
				{
#pragma unroll
					for(int q = 0; q < MS; q++)
						if(((smask >> q) & 1u) && wOwner[q * REGION_PX + pkey] != (unsigned char)lane) shared = true;
				}
				}
				int prank = 0, maxRank = 0;
				const bool anyShared = !noCheck && __any_sync(0xFFFFFFFFu, shared);
				if(!USE_MAP && MS == 1 && optimistic && anyShared)
				{
					redo = true; // (warp-uniform) nothing of this round has been applied
					break;
				}
				if(anyShared) // overlapping triangles: the items of a pixel run in list order
				{
					const uint32_t peers = __match_any_sync(0xFFFFFFFFu, live ? pkey : (0x200u | (uint32_t)lane));
					prank = __popc(peers & laneLt);
					maxRank = (int)__reduce_max_sync(0xFFFFFFFFu, (uint32_t)(live ? prank : 0));
				}
				for(int rr = 0; rr <= maxRank; rr++)
				{
					if(live && prank == rr)
					{
						const int row = (int)((e >> 4) & 7u), bit = (int)(e & 15u);
						const int x = rx + bit, y = ry + row;
						const int ix = x & 1, iy = y & 1;
						const float x0 = pf[0], y0 = pf[1], wA = pf[2], wB = pf[3], wC = pf[4];
						const float rhwConst = pf[5]; // 1/w of a constant w plane (k_setup), 0 otherwise
						const float *S = pf + TRI_FLOATS_FRONT; // slot s: S[3s], S[3s+1], S[3s+2]
						// ---- interpolate + routed fragment shader (PixelRoutine.cpp:196-261, PixelProgram.cpp:138-241) ----
						const float xf = fsub((float)x, x0), yf = fsub((float)y, y0);
						float rhw = SH != SH_CONST ? rhwConst : 1.0f;
						if(SH != SH_CONST && rhwConst == 0.0f) rhw = frcp(__fmaf_rn(xf, wA, fadd(wC, fmul(yf, wB)))); // (a branch: the division is ~10 instructions)
						float texel[4] = { 0, 0, 0, 0 };
						if(TEX)
						{
							// implicit LOD from lanes 0,1,2 of the pixel's quad (helper pixels included), SamplerCore.cpp:1376-1422
							const uint32_t modeU = FS ? (uint32_t)IM_PERSP : d.slotMode[UV], modeV = FS ? (uint32_t)IM_PERSP : d.slotMode[UV + 1];
							float uu[3], vv[3];
#pragma unroll
							for(int t = 0; t < 3; t++)
							{
								const float xk = fsub((float)(x - ix + (t & 1)), x0), yk = fsub((float)(y - iy + (t >> 1)), y0);
								float rk = rhwConst;
								if(rhwConst == 0.0f) rk = frcp(__fmaf_rn(xk, wA, fadd(wC, fmul(yk, wB))));
								uu[t] = interp_slot(S[3 * UV], S[3 * UV + 1], S[3 * UV + 2], modeU, xk, yk, rk);
								vv[t] = interp_slot(S[3 * UV + 3], S[3 * UV + 4], S[3 * UV + 5], modeV, xk, yk, rk);
							}
							const float u = interp_slot(S[3 * UV], S[3 * UV + 1], S[3 * UV + 2], modeU, xf, yf, rhw);
							const float v = interp_slot(S[3 * UV + 3], S[3 * UV + 4], S[3 * UV + 5], modeV, xf, yf, rhw);
							if(d.texFast) sample_texture<true>(d, compute_lod<true>(d, uu[0], uu[1], uu[2], vv[0], vv[1], vv[2]), u, v, texel);
							else sample_texture<false>(d, compute_lod<false>(d, uu[0], uu[1], uu[2], vv[0], vv[1], vv[2]), u, v, texel);
						}
						float rgba[4];
#pragma unroll
						for(int ch = 0; ch < 4; ch++)
						{
							float val;
							if(FS)
							{
								// routing checked on the host: constant shader -> constants; texture shader -> texel channel ch;
								// varying shader -> slot ch, or a constant (e.g. alpha = 1)
								if(SH == SH_CONST) val = __uint_as_float(d.chanValue[ch]);
								else if(SH == SH_TEX) val = texel[ch];
								else
								{
									if((constChan >> ch) & 1) val = __uint_as_float(d.chanValue[ch]);
									else val = interp_slot(S[3 * ch], S[3 * ch + 1], S[3 * ch + 2], IM_PERSP, xf, yf, rhw);
								}
							}
							else
							{
								const uint32_t kind = d.chanKind[ch];
								if(kind == CK_CONST) val = __uint_as_float(d.chanValue[ch]);
								else if(TEX && kind == CK_TEXEL)
								{
									const uint32_t t = d.chanValue[ch];
									val = t == 0 ? texel[0] : t == 1 ? texel[1] : t == 2 ? texel[2] : texel[3];
								}
								else if(SH == SH_VARY || SH == SH_GENERIC) val = interp_slot(S[3 * ch], S[3 * ch + 1], S[3 * ch + 2], d.slotMode[ch], xf, yf, rhw);
								else val = 0.0f;
							}
							// PixelProgram::clampColor :286-364 — UNORM targets only ("if the color attachment is floating-point, no clamping occurs")
							rgba[ch] = (!FS && colorEpp > 1) ? val : sse_min(sse_max(val, 0.0f), 1.0f);
						}
						const uint32_t frontFacing = FS ? 1u : (flagsk & TRI_FLAG_FRONT);
						float premul[3] = { 0, 0, 0 }; // source colour times source alpha: the first product of the SRC_ALPHA blend, the same for every sample
						if(BL == BL_SRC_ALPHA)
						{
#pragma unroll
							for(int ch = 0; ch < 3; ch++) premul[ch] = fmul(rgba[ch], rgba[3]);
						}

						// ---- per covered sample: stencil test, depth test, depth write, blend + colour write, stencil write ----
#pragma unroll
						for(int q = 0; q < MS; q++) // (unrolled: the sample offsets and plane strides of each copy are constants)
						{
							if(!((smask >> q) & 1u)) continue;
							const int pi = q * REGION_PX + (int)pkey; // index inside the staged planes
							bool sPass = true;
							uint32_t sValue = 0;
							if(!FS && d.stencilActive)
							{
								const KStencilFace &face = frontFacing ? d.front : d.back;
								sValue = smStencil[pi];
								sPass = stencil_compare(face.compareOp, sValue & face.compareMask, face.reference & face.compareMask);
							}
							// alphaToCoverage (PixelRoutine.cpp:643-658, thresholds Renderer.cpp:391-410): CmpNLT = ordered >=, a NaN alpha loses its coverage; a
							// sample that loses its coverage leaves every later stage, stencil write included (:319-326)
							bool alive = true;
							if(!FS && d.alphaToCoverage) alive = rgba[3] >= (MS == 4 ? (q == 0 ? 0.2f : q == 1 ? 0.4f : q == 2 ? 0.6f : 0.8f) : 0.5f);
							bool zPass = true;
							float z = 0.0f;
							if(d.depthTestActive)
							{
								float yy = yf, xx = xf;
								if(MS > 1)
								{
									// sample position relative to the pixel centre (Constants.cpp:291-297), in eighths: Y = 2q - 3, X = {-1, 3, -3, 1}
									const float sy = fmul((float)(2 * q - 3), 0.125f);
									const float sx = fmul((float)(int)(signed char)(0x01FD03FFu >> (8 * q)), 0.125f);
									yy = fadd(yy, sy);
									xx = fsub(xx, sx);
								}
								z = __fmaf_rn(xx, zv.y, fadd(zv.w, fmul(yy, zv.z)));
								if(biasOn) z = fadd(z, zv.x);
								z = FS ? sse_min(sse_max(z, 0.0f), 1.0f) : sse_min(sse_max(z, d.minDepthClamp), d.maxDepthClamp); // clampDepth :484-492
								if(!FS && d.depth16)
								{
									// D16_UNORM (:466-482, :508-511): Z = Min(Max(Round(z * 0xFFFF), 0), 0xFFFF) against Float(UShort), as floats;
									// the value written (:687-711, saturating UShort of Round(z * 0xFFFF)) is that same Z
									z = sse_min(sse_max(rintf(fmul(z, 65535.0f)), 0.0f), 65535.0f);
									zPass = depth_compare(d.depthCompareOp, (float)((const unsigned short *)smDepth)[pi], z);
								}
								else
								{
									const float zValue = smDepth[pi];
									if(FS) zPass = d.depthCompareOp == CMP_LESS ? zValue > z : zValue >= z; // LESS / LESS_OR_EQUAL (:533-553)
									else zPass = depth_compare(d.depthCompareOp, zValue, z);
								}
								if(!FS && d.depthBounds)
								{
									// depthBoundsTest :576-641: the STORED depth (read before this fragment's write) against [min, max]; with a depth
									// test it narrows the depth mask, so the stencil depth-fail op sees it; without one it narrows the coverage
									const float stored = d.depth16 ? fmul((float)((const unsigned short *)smDepth)[pi], 1.0f / 0xFFFF) : smDepth[pi];
									const bool inside = d.minDepthBounds <= stored && stored <= d.maxDepthBounds;
									if(d.depthBounds == 2) alive = alive && inside;
									else zPass = zPass && inside;
								}
							}
							if(alive && zPass && sPass) // zMask (& sMask); without a depth test this is cMask & sMask
							{
								if(d.depthWriteEnable)
								{
									if(!FS && d.depth16) ((unsigned short *)smDepth)[pi] = (unsigned short)z;
									else smDepth[pi] = z;
									dirty = true;
								}
								if(colorOn)
								{
									const bool floatTarget = !FS && colorEpp > 1;
									const uint32_t px = floatTarget ? 0u : smColor[pi];
									float o[4] = { rgba[0], rgba[1], rgba[2], rgba[3] };
									if(BL != BL_OFF)
									{
										float dst[4]; // readPixel :1111-1130: b -> b*257 -> float * (1/65535)
										if(floatTarget)
										{
											// floating-point targets: the stored value itself (:1700-1710), or Reactor's Float(Half) (:1782-1801)
											if(colorEpp == 4)
											{
												const float4 t = ((const float4 *)smColor)[pi];
												dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
											}
											else
											{
												const uint2 t = ((const uint2 *)smColor)[pi];
												dst[0] = half_to_float(t.x & 0xFFFFu); dst[1] = half_to_float(t.x >> 16);
												dst[2] = half_to_float(t.y & 0xFFFFu); dst[3] = half_to_float(t.y >> 16);
											}
										}
										else
										{
#pragma unroll
											for(int ch = 0; ch < 4; ch++)
											{
												const uint32_t byte = (bgr && ch < 3) ? 2 - ch : ch;
												dst[ch] = fmul((float)__byte_perm(px, 0, 0x4400u | byte | (byte << 4)), 1.0f / 0xFFFF); // b * 257 == b << 8 | b
												if(!FS && d.srgb && ch < 3) dst[ch] = srgb_to_linear(dst[ch]);
											}
										}
										if(BL == BL_SRC_ALPHA)
										{
											const float da = fsub(1.0f, rgba[3]);
#pragma unroll
											for(int ch = 0; ch < 3; ch++) o[ch] = fadd(premul[ch], fmul(dst[ch], da));
										}
										else
										{
#pragma unroll
											for(int ch = 0; ch < 3; ch++)
												o[ch] = blend_apply(d.op, rgba[ch], blend_factor_rgb(d, d.srcF, ch, rgba, dst), dst[ch], blend_factor_rgb(d, d.dstF, ch, rgba, dst));
											o[3] = blend_apply(d.opA, rgba[3], blend_factor_a(d, d.srcFA, rgba, dst), dst[3], blend_factor_a(d, d.dstFA, rgba, dst));
										}
									}
									if(!FS && d.srgb)
									{
#pragma unroll
										for(int ch = 0; ch < 3; ch++) o[ch] = linear_to_srgb(o[ch]);
									}
									if(floatTarget)
									{
										// masked store of the bits (:2429-2447), or of Reactor's Half(Float) (:2504-2540); no clamp, no rounding of fp32
										const uint32_t cm = d.colorWriteMask;
										if(colorEpp == 4)
										{
											float *t = (float *)smColor + 4 * pi;
#pragma unroll
											for(int ch = 0; ch < 4; ch++)
												if((cm >> ch) & 1) t[ch] = o[ch];
										}
										else
										{
											unsigned short *t = (unsigned short *)smColor + 4 * pi;
#pragma unroll
											for(int ch = 0; ch < 4; ch++)
												if((cm >> ch) & 1) t[ch] = (unsigned short)float_to_half(o[ch]);
										}
									}
									else
									{
										const uint32_t pk = bgr ? pack_unorm8(o[2], o[1], o[0], o[3]) : pack_unorm8(o[0], o[1], o[2], o[3]);
										if(FS && BL == BL_OFF && wo) *(uint32_t *)(d.colorBuf + (size_t)q * d.colorSliceB + (size_t)y * d.colorPitchB + 4 * x) = pk;
										else smColor[pi] = (px & ~wmask32) | (pk & wmask32);
									}
									dirty = true;
								}
							}
							if(!FS && d.stencilWrite && alive) // writeStencil :754-817
							{
								const KStencilFace &face = frontFacing ? d.front : d.back;
								const uint32_t ref = face.reference & 0xFF;
								uint32_t nv;
								if(!sPass) nv = stencil_op(face.failOp, sValue, ref);
								else if(!zPass) nv = stencil_op(face.depthFailOp, sValue, ref);
								else nv = stencil_op(face.passOp, sValue, ref);
								const uint32_t wm = face.writeMask & 0xFF;
								smStencil[pi] = (unsigned char)((nv & wm) | (sValue & ~wm));
								dirty = true;
							}
						}
					}
					__syncwarp();
				}
			}
			__syncwarp(); // the queue is refilled by the next entries
		}
	}

	if(!redo) break;
	} // pass

	if(!tileReady)
		while(!mbar_try_wait(bar, tmaPhase)) {} // never leave with a bulk copy into this CTA's shared memory still in flight
	if(wo) return; // every fragment is in the framebuffer already
	if(!__any_sync(0xFFFFFFFFu, dirty)) return;
	// Only rows inside the scissor go back: the rows of a region that the scissor cuts off may belong to another rank's band of
	// the same frame (multi-GPU), whose pixels this warp has staged but must not overwrite.
	const bool whole = ryLo == ry && ryHi == min(ry + SWCU_REGION_H, d.fbHeight);
	if(d.useTma && whole)
	{
		fence_proxy_async(); // generic-proxy writes of the region -> visible to the async proxy
		__syncwarp();
		if(lane == 0)
		{
			if(colorOn) tma_store_3d(&maps.color, smColor, rx * colorEpp, ry, 0);
			if(d.depthWriteEnable) tma_store_3d(&maps.depth, smDepth, rx, ry, 0);
			if(d.stencilWrite) tma_store_3d(&maps.stencil, smStencil, rx, ry, 0);
			tma_commit();
			tma_wait_read0();
		}
	}
	else
	{
		__syncwarp();
		if(colorOn && colorEpp == 4) region_copy<MS, uint4, true>((uint4 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, rx, ry, d.fbWidth, ryLo, ryHi, lane);
		else if(colorOn && colorEpp == 2) region_copy<MS, uint2, true>((uint2 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, rx, ry, d.fbWidth, ryLo, ryHi, lane);
		else if(colorOn) region_copy<MS, uint32_t, true>(smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, rx, ry, d.fbWidth, ryLo, ryHi, lane);
		if(d.depthWriteEnable && d.depth16) region_copy<MS, unsigned short, true>((unsigned short *)smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, rx, ry, d.fbWidth, ryLo, ryHi, lane);
		else if(d.depthWriteEnable) region_copy<MS, float, true>(smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, rx, ry, d.fbWidth, ryLo, ryHi, lane);
		if(d.stencilWrite) region_copy<MS, unsigned char, true>(smStencil, d.stencilBuf, d.stencilPitchB, d.stencilSliceB, rx, ry, d.fbWidth, ryLo, ryHi, lane);
	}
}

// ------------------------------------------------------------------------------------------------------------------
// the steps either side of the draw
// ------------------------------------------------------------------------------------------------------------------
// Blitter::fastClear (Blitter.cpp:170-325): rectangle fill of every sample slice; bpp 4 or 1
__global__ void k_clear(unsigned char *base, int pitchB, int sliceB, int bpp, int x0, int y0, int w, int h, int samples, uint4 value4)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if(x >= w || y >= h) return;
	for(int q = 0; q < samples; q++)
	{
		unsigned char *row = base + (size_t)q * sliceB + (size_t)(y0 + y) * pitchB;
		const uint32_t value = value4.x;
		if(bpp == 16) ((uint4 *)row)[x0 + x] = value4;
		else if(bpp == 8) ((uint2 *)row)[x0 + x] = make_uint2(value4.x, value4.y);
		else if(bpp == 4) ((uint32_t *)row)[x0 + x] = value;
		else if(bpp == 2) ((unsigned short *)row)[x0 + x] = (unsigned short)value;
		else row[x0 + x] = (unsigned char)value;
	}
}

// Blitter::fastResolve (Blitter.cpp:2079-2205): RGBA8 4x -> 1x, avg(avg(s0,s1),avg(s2,s3)) with pavgb = (a+b+1)>>1
DEVI uint32_t pavgb4(uint32_t a, uint32_t b) { return __vavgu4(a, b); }
__global__ void k_resolve4(const unsigned char *src, int srcPitchB, int srcSliceB, unsigned char *dst, int dstPitchB, int w, int h)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if(x >= w || y >= h) return;
	const size_t o = (size_t)y * srcPitchB + 4 * (size_t)x;
	const uint32_t s0 = *(const uint32_t *)(src + o), s1 = *(const uint32_t *)(src + o + srcSliceB);
	const uint32_t s2 = *(const uint32_t *)(src + o + 2 * (size_t)srcSliceB), s3 = *(const uint32_t *)(src + o + 3 * (size_t)srcSliceB);
	*(uint32_t *)(dst + (size_t)y * dstPitchB + 4 * (size_t)x) = pavgb4(pavgb4(s0, s1), pavgb4(s2, s3));
}

// Blitter::resolve for the formats fastResolve does not know (its switch, Blitter.cpp:2142-2200, has RGBA8 / BGRA8 UNORM only): the
// generic blit routine (:2053-2071).  Every sample is read as floats (readFloat4 :368-452); sRGB samples go to linear light
// (ApplyScaleAndClamp :1459-1465: * 1/255, sRGBtoLinear on rgb, * 255); the samples are added in order and scaled by 1/4 (:1524-1545);
// the result comes back through the same function (pre-scaled: * 1/255, linearToSRGB, * 255) and is written (:640-745: RoundShort4 +
// unsigned saturation for the 8-bit formats, Reactor's Half() for R16G16B16A16_SFLOAT, the floats themselves for R32G32B32A32_SFLOAT).
// epp = 32-bit words per pixel: 1 (sRGB8), 2 (RGBA16F), 4 (RGBA32F).
__global__ void k_resolve4_generic(const unsigned char *src, int srcPitchB, int srcSliceB, unsigned char *dst, int dstPitchB, int w, int h, int epp)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if(x >= w || y >= h) return;
	const size_t o = (size_t)y * srcPitchB + 4 * (size_t)epp * x;
	float acc[4] = { 0, 0, 0, 0 };
#pragma unroll
	for(int q = 0; q < 4; q++)
	{
		const unsigned char *s = src + o + (size_t)q * srcSliceB;
		float c[4];
		if(epp == 4)
		{
			const float4 v = *(const float4 *)s;
			c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
		}
		else if(epp == 2)
		{
			const uint2 v = *(const uint2 *)s;
			c[0] = half_to_float(v.x & 0xFFFFu); c[1] = half_to_float(v.x >> 16); c[2] = half_to_float(v.y & 0xFFFFu); c[3] = half_to_float(v.y >> 16);
		}
		else
		{
			const uint32_t v = *(const uint32_t *)s;
#pragma unroll
			for(int ch = 0; ch < 4; ch++)
			{
				float t = fmul((float)((v >> (8 * ch)) & 0xFFu), 1.0f / 255.0f);
				if(ch < 3) t = srgb_to_linear(t); // (the three colour bytes are treated alike: RGBA or BGRA order does not matter)
				c[ch] = fmul(t, 255.0f);
			}
		}
#pragma unroll
		for(int ch = 0; ch < 4; ch++) acc[ch] = q == 0 ? c[ch] : fadd(acc[ch], c[ch]);
	}
	unsigned char *t = dst + (size_t)y * dstPitchB + 4 * (size_t)epp * x;
	float r[4];
#pragma unroll
	for(int ch = 0; ch < 4; ch++) r[ch] = fmul(acc[ch], 0.25f);
	if(epp == 4) *(float4 *)t = make_float4(r[0], r[1], r[2], r[3]);
	else if(epp == 2) *(uint2 *)t = make_uint2(float_to_half(r[0]) | (float_to_half(r[1]) << 16), float_to_half(r[2]) | (float_to_half(r[3]) << 16));
	else
	{
		uint32_t pk = 0;
#pragma unroll
		for(int ch = 0; ch < 4; ch++)
		{
			float v = fmul(r[ch], 1.0f / 255.0f);
			if(ch < 3) v = linear_to_srgb(v);
			v = fmul(v, 255.0f);
			const int i = clampi(round_int(v), -32768, 32767); // RoundShort4: cvtps2dq + packssdw; then packuswb
			pk |= (uint32_t)clampi(i, 0, 255) << (8 * ch);
		}
		*(uint32_t *)t = pk;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// multi-GPU: finished bands go to the presenting GPU by stores over NVLink into its (IPC-mapped) frame, not by a collective
// ------------------------------------------------------------------------------------------------------------------
// rows of a 4-byte-per-pixel image -> another image (the destination may be peer memory); 16 bytes per thread when aligned
__global__ void k_copy_rows(const unsigned char *src, int srcPitchB, unsigned char *dst, int dstPitchB, int rowBytes, int h, int vec)
{
	const int y = blockIdx.y;
	if(y >= h) return;
	const unsigned char *s = src + (size_t)y * srcPitchB;
	unsigned char *t = dst + (size_t)y * dstPitchB;
	if(vec)
	{
		const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
		if(i < rowBytes) *(uint4 *)(t + i) = *(const uint4 *)(s + i);
	}
	else
	{
		const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
		if(i < rowBytes) *(uint32_t *)(t + i) = *(const uint32_t *)(s + i);
	}
}

// Group barrier of the setup phase: tell every rank that my share of the draw has been delivered (my stores and atomics into its
// memory are visible system-wide), then wait until every rank has told me the same.  flags[p]: rank p's flag row of this buffer set.
struct GroupFlags
{
	uint32_t *flags[SWCU_MAX_GROUP];
};
__global__ void k_xbarrier(const GroupFlags g, uint32_t world, uint32_t rank, uint32_t epoch)
{
	const uint32_t lane = threadIdx.x;
	__threadfence_system();
	if(lane < world) *(volatile uint32_t *)(g.flags[lane] + rank) = epoch;
	if(lane < world)
		while((int32_t)(*(volatile const uint32_t *)(g.flags[rank] + lane) - epoch) < 0) __nanosleep(64);
	__syncwarp();
	__threadfence_system();
}

// flag <- value, after everything this stream wrote before (the kernels ahead of this one) is visible system-wide
__global__ void k_signal(uint32_t *flag, uint32_t value)
{
	__threadfence_system();
	*(volatile uint32_t *)flag = value;
}

// spin until flags[first .. first + count) have all reached `value` (they only grow), then make the data they announce visible
__global__ void k_wait_flags(const uint32_t *flags, int first, int count, uint32_t value)
{
	const int i = threadIdx.x;
	if(i < count)
		while((int32_t)(*(volatile const uint32_t *)(flags + first + i) - value) < 0) __nanosleep(100);
	__syncthreads();
	__threadfence_system();
}
