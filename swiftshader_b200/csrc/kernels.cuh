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
		const int yMin = max(y1, d.scY0), yMax = min(y2, d.scY1);
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
	long long qv = N / D;
	if(N % D > 0) qv += 1; // ceiling
	xo = clampi((X1 >> 8) + (int)qv, d.scX0, d.scX1);
	right = swap;
	return true;
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
	return yhi + 1024.0f < (float)(d.scY0 << 8) || ylo - 1024.0f > (float)(d.scY1 << 8);
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
	flags[tri] = rows_missed(d, y[0], w[0], y[1], w[1], y[2], w[2]) ? 0 : 1;
}

DEVI int sel3(int i, int a0, int a1, int a2) { return i == 0 ? a0 : (i == 1 ? a1 : a2); }
DEVI float sel3(int i, float a0, float a1, float a2) { return i == 0 ? a0 : (i == 1 ? a1 : a2); }

// The unclipped triangle lives in registers only (three named vertices, selects instead of indexed arrays); the clipped
// polygon — rare — goes through local-memory arrays.
// MSC = 1: the 1x instantiation (sample count folded, 7 CTAs / SM); MSC = 0: sample count read at run time (used for 4x — measured
// faster than a folded 4x instantiation, whose natural register allocation costs a resident CTA)
template<int MSC>
DEVI void setup_triangle(const DrawConst &d)
{
	extern __shared__ uint32_t s_rows[]; // [SWCU_SMALL_ROWS * ms][SETUP_THREADS]: one scratch column of span rows per thread
	const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x; // grid is padded to whole warps; inactive lanes just allocate 0
	const bool live = tri < d.primCount;
	const int MS = MSC ? MSC : d.ms;
	const bool msaa = MS > 1;
	uint32_t nTiles = 0;
	bool visible = false;
	uint32_t idx[3] = { 0, 0, 0 };
	const bool precull = live && d.cullFlags != nullptr && d.cullFlags[tri] == 0; // marked by k_cull: rows outside the band
	VOut va, vb, vc;
	float sv[3][SWCU_MAXSLOTS]; // slot sources at the three vertices
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
#pragma unroll
			for(int a = 0; a < 3; a++)
#pragma unroll
				for(int c = 0; c < 4; c++) pos[a][c] = vs_operand(d, d.vsPos[c], idx[a]);
			// Early band / scissor reject on an approximate projected y (fast reciprocal, 4-pixel margin): with all three w > 0 the
			// clipped polygon stays inside the hull of the projected vertices, so a triangle whose hull misses the scissor rows
			// cannot produce a span.  This is what a rank of a multi-GPU frame pays for a triangle outside its band.
			if(rows_missed(d, pos[0][1], pos[0][3], pos[1][1], pos[1][3], pos[2][1], pos[2][3])) break;
			process_vertex(d, pos[0][0], pos[0][1], pos[0][2], pos[0][3], va);
			process_vertex(d, pos[1][0], pos[1][1], pos[1][2], pos[1][3], vb);
			process_vertex(d, pos[2][0], pos[2][1], pos[2][2], pos[2][3], vc);

			// setupSolidTriangles, Renderer.cpp:749-757
			if((va.flags & vb.flags & vc.flags) != CLIP_FINITE) break;
			const int flagsOr = va.flags | vb.flags | vc.flags;

			// culling, SetupRoutine.cpp:73-115 (on the original three vertices)
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
			if(flagsOr != CLIP_FINITE && (flagsOr & CLIP_FRUSTUM))
			{
				float4 P[16];
				P[0] = make_float4(va.px, va.py, va.pz, va.pw);
				P[1] = make_float4(vb.px, vb.py, vb.pz, vb.pw);
				P[2] = make_float4(vc.px, vc.py, vc.pz, vc.pw);
				n = clip_polygon(P, 3, flagsOr);
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
			yMin = max(yMin, d.scY0);
			yMax = min(yMax, d.scY1);
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

	// ---- span-table and big-list allocation for the big triangles, one atomic per warp ----
	const int rows = visible ? yMax - yMin : 0;
	bool big = false;
	uint32_t tileRect = TILE_RECT_NONE; // what k_emit needs to know about this triangle (see tile_rect_count)
	if(visible)
	{
		const int tx0 = pxMin / SWCU_TILE_W, tx1 = (pxMax - 1) / SWCU_TILE_W;
		const int ty0 = yMin / SWCU_TILE_H, ty1 = (yMax - 1) / SWCU_TILE_H;
		nTiles = (uint32_t)((tx1 - tx0 + 1) * (ty1 - ty0 + 1));
		// A polygon with a coordinate outside the screen range (a vertex that projected to INT_MIN because its w was -Inf or NaN,
		// ...) also goes the big-triangle way: k_big reproduces the reference's wrapping arithmetic for such edges (edge_wrapped),
		// which keeps that case out of the small-triangle DDA below (whose fast division assumes screen-sized operands).
		const bool insane = polygon_insane(minXs, maxXs, minYs, maxYs);
		big = rows > SWCU_SMALL_ROWS || nTiles > SWCU_SMALL_TILES || insane;
		tileRect = big ? (TILE_RECT_BIG | nTiles) : ((uint32_t)tx0 | ((uint32_t)ty0 << 9) | ((uint32_t)(tx1 - tx0) << 19) | ((uint32_t)(ty1 - ty0) << 22));
	}
	const uint32_t count = big ? (uint32_t)(rows * MS) : 0u;
	unsigned long long base = 0, slot = 0;
	if(__any_sync(0xFFFFFFFFu, big))
	{
		base = warp_alloc(&d.counters->spanCursor, count);
		slot = warp_alloc(&d.counters->bigSlots, big ? 1u : 0u);
	}
	const uint32_t nvis = __popc(__ballot_sync(0xFFFFFFFFu, visible));
	if((threadIdx.x & 31) == 0 && nvis) atomicAdd(&d.counters->visible, nvis);
	if(!live) return;
	unsigned char *rec = d.triRecords + (size_t)tri * d.triStride;
	if(big && base + count > d.spanCapacity) { atomicOr(&d.counters->overflow, 1u); visible = false; }
	if(big && slot >= d.bigCapacity) { atomicOr(&d.counters->overflow, 2u); visible = false; }
	if(!visible)
	{
		// empty bounds: never a candidate.  Only the direct mode reads the header of an invisible triangle (every tile CTA walks
		// the whole list); a binned draw never puts it in a tile list, so the 32-byte sector is not written at all.
		if(d.direct) *(uint4 *)rec = make_uint4(0, 0, 0, 0);
		d.tileCount[tri] = TILE_RECT_NONE;
		return;
	}
	d.tileCount[tri] = tileRect;
	// the attributes behind the plane slots (usually the same cache lines as the positions); in flight during the span work
#pragma unroll
	for(int a = 0; a < 3; a++)
#pragma unroll
		for(int k = 0; k < SWCU_MAXSLOTS; k++) sv[a][k] = k < d.nslots ? vs_operand(d, d.slotSrc[k], idx[a]) : 0.0f;

	if(big)
	{
		BigTri &b = d.bigList[slot];
		b.tri = tri; b.spanBase = (uint32_t)base; b.n = n; b.dir = dir;
		b.yMin = yMin; b.yMax = yMax; b.pxMin = pxMin; b.pxMax = pxMax;
		if(clipped)
			for(int i = 0; i < n; i++) { b.X[i] = PX[i]; b.Y[i] = PY[i]; }
		else
		{
			b.X[0] = va.X; b.X[1] = vb.X; b.X[2] = vc.X;
			b.Y[0] = va.Y; b.Y[1] = vb.Y; b.Y[2] = vc.Y;
		}
	}
	else
	{
		// span rows of the small triangle, built in a shared scratch column and stored inline in its record
		uint32_t *col = s_rows + threadIdx.x;
		// every row of the record is written (rows outside the triangle as empty spans): whole 32-byte sectors reach L2, so
		// evicting them needs no fill from DRAM.
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
		uint32_t *out = (uint32_t *)(rec + d.triStride) - SWCU_SMALL_ROWS * MS;
#pragma unroll
		for(int r = 0; r < SWCU_SMALL_ROWS; r++)
			if(r < 2 * MS)
				((uint4 *)out)[r] = make_uint4(col[(4 * r) * SETUP_THREADS], col[(4 * r + 1) * SETUP_THREADS], col[(4 * r + 2) * SETUP_THREADS], col[(4 * r + 3) * SETUP_THREADS]);
	}

	// ---- vertex sort (SetupRoutine.cpp:271-294): only changes float rounding of the planes ----
	int i0 = 0, i1 = 1, i2 = 2;
	{
		const float y0 = va.py, y1 = vb.py, y2 = vc.py;
		const float ym = sse_min(sse_min(y0, y1), y2);
		rot1(ym == y1, i0, i1, i2);
		rot2(ym == y2, i0, i1, i2);
	}
	{
		const float w0 = sel3(i0, va.pw, vb.pw, vc.pw), w1 = sel3(i1, va.pw, vb.pw, vc.pw), w2 = sel3(i2, va.pw, vb.pw, vc.pw);
		const float wm = sse_max(sse_max(w0, w1), w2);
		rot1(wm == w1, i0, i1, i2);
		rot2(wm == w2, i0, i1, i2);
	}
	const float w0 = sel3(i0, va.pw, vb.pw, vc.pw), w1 = sel3(i1, va.pw, vb.pw, vc.pw), w2 = sel3(i2, va.pw, vb.pw, vc.pw);
	const int X0 = sel3(i0, va.X, vb.X, vc.X), X1 = sel3(i1, va.X, vb.X, vc.X), X2 = sel3(i2, va.X, vb.X, vc.X);
	const int Y0 = sel3(i0, va.Y, vb.Y, vc.Y), Y1 = sel3(i1, va.Y, vb.Y, vc.Y), Y2 = sel3(i2, va.Y, vb.Y, vc.Y);
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
	float *f = (float *)(rec + TRI_HEADER_BYTES);
	float F[TRI_FLOATS_FIXED + 3 * SWCU_MAXSLOTS + 3]; // the plane block, stored with 128-bit writes below
#pragma unroll
	for(int i = 0; i < TRI_FLOATS_FIXED + 3 * SWCU_MAXSLOTS + 3; i++) F[i] = 0.0f;
	F[0] = x0; F[1] = y0;
	F[3] = fadd(fadd(M00, M10), M20);
	F[4] = fadd(fadd(M01, M11), M21);
	F[5] = fadd(fadd(M02, 0.0f), 0.0f);
	// The last float of the block is spare for every slot count in use (9 + 3n = 9, 15, 21, 27).  It carries 1/w when the w
	// plane is constant (wA == wB == 0: MulAdd(x, 0, wC + y * 0) == wC at every pixel, bit for bit), so the tile kernel can
	// skip the per-fragment division (PixelRoutine.cpp:196-199); 0 = "not constant".
	float rhwConst = 0.0f;
	if(F[3] == 0.0f && F[4] == 0.0f && F[5] != 0.0f)
	{
		const float r = fdiv(1.0f, F[5]);
		if(r != 0.0f && fabsf(r) <= 3.40282347e38f) rhwConst = r;
	}
	const int nf4 = (TRI_FLOATS_FIXED + 3 * d.nslots + 3) >> 2;
	// 128-bit stores of the block, each issued as soon as its four floats are final (keeps the live range of F short)
	auto put = [&](int j) { ((float4 *)f)[j] = make_float4(F[4 * j], F[4 * j + 1], F[4 * j + 2], j == nf4 - 1 ? rhwConst : F[4 * j + 3]); };
	float zBias = 0.0f;
	if(d.depthTestActive)
	{
		const float zp0 = sel3(i0, va.zp, vb.zp, vc.zp), zp1 = sel3(i1, va.zp, vb.zp, vc.zp), zp2 = sel3(i2, va.zp, vb.zp, vc.zp);
		const float z0 = zp0;
		const float z1 = fsub(zp1, z0), z2 = fsub(zp2, z0);
		const float px1 = fmul((float)dX1, rsub), py1 = fmul((float)dY1, rsub), px2 = fmul((float)dX2, rsub), py2 = fmul((float)dY2, rsub);
		const float D = fdiv(d.depthRange, fsub(fmul(px1, py2), fmul(px2, py1)));
		const float A = fmul(fsub(fmul(py2, z1), fmul(py1, z2)), D);
		const float B = fmul(fsub(fmul(px1, z2), fmul(px2, z1)), D);
		const float C = fadd(fmul(z0, d.depthRange), d.depthNear);
		F[6] = A; F[7] = B; F[8] = C;
		const bool applyConst = d.depthBiasConstant != 0.0f, applySlope = d.depthBiasSlope != 0.0f;
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
	}
	F[2] = zBias;
	put(0);
	put(1);
	// setupGradient, SetupRoutine.cpp:514-548
#pragma unroll
	for(int k = 0; k < SWCU_MAXSLOTS; k++)
	{
		if(k >= d.nslots) break;
		float *P = F + TRI_FLOATS_FIXED + 3 * k;
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
		// floats up to index 11 + 3k are final: store the float4s this slot completed
#pragma unroll
		for(int j = 2; j < (TRI_FLOATS_FIXED + 3 * SWCU_MAXSLOTS + 3) / 4; j++)
			if(4 * j + 3 <= 11 + 3 * k && 4 * j + 3 > 8 + 3 * k) put(j);
	}
	// the float4 that holds the padding (and rhwConst) is still open
#pragma unroll
	for(int j = 2; j < (TRI_FLOATS_FIXED + 3 * SWCU_MAXSLOTS + 3) / 4; j++)
		if(j < nf4 && 4 * j + 3 > 8 + 3 * d.nslots) put(j);
	uint4 hdr;
	hdr.x = (uint32_t)pxMin | ((uint32_t)pxMax << 16);
	hdr.y = (uint32_t)yMin | ((uint32_t)yMax << 16);
	hdr.z = (uint32_t)base;
	hdr.w = (frontFacing ? 1u : 0u) | (big ? 2u : 0u);
	*(uint4 *)rec = hdr;
}

__global__ void __launch_bounds__(SETUP_THREADS, SETUP_BLOCKS_1X) k_setup_1x(const __grid_constant__ DrawConst d) { setup_triangle<1>(d); }
__global__ void __launch_bounds__(SETUP_THREADS) k_setup(const __grid_constant__ DrawConst d) { setup_triangle<0>(d); }

// ------------------------------------------------------------------------------------------------------------------
// k_big: spans (and tile pairs) of the large triangles
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_big(const __grid_constant__ DrawConst d, const uint32_t *pairOffset, uint32_t *keys, uint32_t *vals)
{
	const uint32_t nbig = (uint32_t)min(d.counters->bigSlots, (unsigned long long)d.bigCapacity);
	for(uint32_t e = blockIdx.x; e < nbig; e += gridDim.x)
	{
		const BigTri &b = d.bigList[e];
		const int n = b.n, dir = b.dir;
		const bool msaa = d.ms > 1;
		{
			// (row, sample) entries strided over the y-slices of the grid as well: a full-screen triangle gets one row per thread
			const int total = (b.yMax - b.yMin) * d.ms;
			for(int r = blockIdx.y * blockDim.x + threadIdx.x; r < total; r += blockDim.x * gridDim.y)
			{
				const int y = b.yMin + r / d.ms, q = r % d.ms;
				const int ox = msaa ? c_Xf[q] : 0, oy = msaa ? c_Yf[q] : 0;
				int L = 0, R = 0;
				if(msaa) L = R = clampi((int)((uint32_t)b.X[0] + 255u) >> 8, d.scX0, d.scX1); // MSAA pre-fill, SetupRoutine.cpp:214-225
				for(int i = 0; i < n; i++) // in edge order: the last writer wins, like the reference's span table
				{
					const int ia = i + 1 - dir, ib = i + dir;
					const int a = ia == n ? 0 : ia, bb = ib == n ? 0 : ib;
					bool right; int x;
					if(edge_at_row(d, b.X[a] - ox, b.Y[a] - oy, b.X[bb] - ox, b.Y[bb] - oy, y, right, x)) { if(right) R = x; else L = x; }
				}
				d.spans[b.spanBase + r] = (uint32_t)L | ((uint32_t)R << 16);
			}
		}
		if(keys)
		{
			// tiles of the bounding box; a tile all of whose pixel centres lie strictly outside one edge is dropped
			const int tx0 = b.pxMin / SWCU_TILE_W, tx1 = (b.pxMax - 1) / SWCU_TILE_W;
			const int ty0 = b.yMin / SWCU_TILE_H, ty1 = (b.yMax - 1) / SWCU_TILE_H;
			const int tw = tx1 - tx0 + 1, total = tw * (ty1 - ty0 + 1);
			long long area2 = 0;
			int loX = b.X[0], hiX = b.X[0], loY = b.Y[0], hiY = b.Y[0];
			for(int i = 0; i < n; i++)
			{
				const int j = i + 1 == n ? 0 : i + 1;
				area2 += (long long)b.X[i] * b.Y[j] - (long long)b.X[j] * b.Y[i];
				loX = min(loX, b.X[i]); hiX = max(hiX, b.X[i]); loY = min(loY, b.Y[i]); hiY = max(hiY, b.Y[i]);
			}
			// no geometric culling for a polygon whose edges the reference walks in wrapped arithmetic (see k_setup)
			const long long sgn = polygon_insane(loX, hiX, loY, hiY) ? 0 : (area2 > 0 ? 1 : (area2 < 0 ? -1 : 0));
			const int m = msaa ? 96 : 0;
			const uint32_t off = pairOffset[b.tri];
			for(int t = blockIdx.y * blockDim.x + threadIdx.x; t < total; t += blockDim.x * gridDim.y)
			{
				const int tx = tx0 + t % tw, ty = ty0 + t / tw;
				// pixel centres of the tile in 24.8 (centre of pixel x is X = 256x), widened by the sample offsets
				const long long cx0 = 256ll * max(tx * SWCU_TILE_W, b.pxMin) - m, cx1 = 256ll * (min(tx * SWCU_TILE_W + SWCU_TILE_W, b.pxMax) - 1) + m;
				const long long cy0 = 256ll * max(ty * SWCU_TILE_H, b.yMin) - m, cy1 = 256ll * (min(ty * SWCU_TILE_H + SWCU_TILE_H, b.yMax) - 1) + m;
				bool outside = false;
				if(sgn != 0)
					for(int i = 0; i < n && !outside; i++)
					{
						const int j = i + 1 == n ? 0 : i + 1;
						const long long ex = (long long)b.X[j] - b.X[i], ey = (long long)b.Y[j] - b.Y[i];
						const long long slack = 4 * (llabs(ex) + llabs(ey)) + 1024; // rounding of re-projected clip vertices
						// E(P) = ex*(Py - Yi) - ey*(Px - Xi); inside when sgn*E >= 0
						const long long e00 = sgn * (ex * (cy0 - b.Y[i]) - ey * (cx0 - b.X[i]));
						const long long e10 = sgn * (ex * (cy0 - b.Y[i]) - ey * (cx1 - b.X[i]));
						const long long e01 = sgn * (ex * (cy1 - b.Y[i]) - ey * (cx0 - b.X[i]));
						const long long e11 = sgn * (ex * (cy1 - b.Y[i]) - ey * (cx1 - b.X[i]));
						outside = e00 < -slack && e10 < -slack && e01 < -slack && e11 < -slack;
					}
				keys[off + t] = outside ? (uint32_t)(d.tilesX * d.tilesY) : (uint32_t)(ty * d.tilesX + tx); // numTiles = "no tile", sorts last
				vals[off + t] = b.tri;
			}
		}
	}
}

// (tile, triangle) pairs of the small triangles, from the packed tile rectangle k_setup left for each of them
__global__ void __launch_bounds__(256) k_emit(const __grid_constant__ DrawConst d, const uint32_t *pairOffset, uint32_t *keys, uint32_t *vals)
{
	const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
	if(tri >= d.primCount) return;
	const uint32_t r = d.tileCount[tri];
	if(r & TILE_RECT_BIG) return; // big: k_big emits; invisible: nothing to emit
	const int tx0 = r & 0x1FF, ty0 = (r >> 9) & 0x3FF, tx1 = tx0 + ((r >> 19) & 7), ty1 = ty0 + ((r >> 22) & 7);
	uint32_t o = pairOffset[tri];
	for(int ty = ty0; ty <= ty1; ty++)
		for(int tx = tx0; tx <= tx1; tx++)
		{
			keys[o] = (uint32_t)(ty * d.tilesX + tx);
			vals[o] = tri;
			o++;
		}
}

// start/end of every tile's run in the sorted pair array; four keys per thread
__global__ void k_tile_ranges(const uint32_t *keys, uint32_t n, uint32_t numTiles, uint32_t *tileBegin, uint32_t *tileEnd)
{
	const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if(i0 >= n) return;
	uint32_t k[6]; // keys i0 - 1 .. i0 + 4
	if(i0 + 4 <= n)
	{
		const uint4 v = *(const uint4 *)(keys + i0);
		k[1] = v.x; k[2] = v.y; k[3] = v.z; k[4] = v.w;
	}
	else
	{
#pragma unroll
		for(int j = 0; j < 4; j++) k[1 + j] = i0 + j < n ? keys[i0 + j] : 0xFFFFFFFFu;
	}
	k[0] = i0 > 0 ? keys[i0 - 1] : 0xFFFFFFFFu;
	k[5] = i0 + 4 < n ? keys[i0 + 4] : 0xFFFFFFFFu;
#pragma unroll
	for(int j = 0; j < 4; j++)
	{
		const uint32_t i = i0 + j, key = k[1 + j];
		if(i >= n || key >= numTiles) continue;
		if(i == 0 || k[j] != key) tileBegin[key] = i;
		if(i + 1 == n || k[2 + j] != key) tileEnd[key] = i + 1;
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
// k_tile — one CTA per 32x16 screen tile, one warp per 16x8 region of it
//
//   * the tile's colour / depth / stencil planes are staged in shared memory by ONE TMA load per attachment
//     (cp.async.bulk.tensor.3d: x, y, sample plane) that overlaps the first list scan, stay there for the tile's whole
//     triangle list, and go back with one TMA store per attachment (fallback: cooperative 128-bit copies when the
//     attachment's pitch / base address do not satisfy the tensor-map alignment rules);
//   * the four warps are INDEPENDENT inside the list loop (no CTA barriers): each warp scans the tile's list 32 entries
//     at a time (headers of the next block and list entries of the block after it already in flight), ballots the bounding
//     boxes against its region and collects a batch of candidates; their plane equations come in with cp.async;
//   * coverage: one candidate per lane (1x) or lane pair (4x) reads its span rows, clips them to the region's 16 columns and
//     keeps the non-empty (row, sample) runs in registers; one warp scan places every lane's pairs and first item, the
//     pair words and one start-mark bit per pair go to shared memory.  Batches of up to 4 (1x) / 1 (4x) candidates — big
//     triangles — take a one-(candidate, row, sample)-per-lane path instead;
//   * the covered samples (ITEMS) are consumed 32 at a time, one per lane, so lane utilisation does not depend on triangle
//     size: the round's mark word maps every lane to its pair.  Items of one sample stay in API order: pairs are in list
//     order, and when two fragments of a range hit the same sample (tested once per range) the items of a round are
//     serialised by __match_any_sync rank;
//   * specialised on <samples, fragment shader class, blend class, fast state>; everything else is warp-uniform run-time state.
// ------------------------------------------------------------------------------------------------------------------
#define TILE_THREADS (SWCU_TILE_WARPS * 32)

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

// shared-memory layout of one tile CTA (dynamic shared memory; the host computes the same size)
template<int MS, int SH>
struct TileLayout
{
	static constexpr int NB = MS == 4 ? 16 : 32;                                   // candidates staged per warp batch
	static constexpr int NF4 = (TRI_FLOATS_FIXED + 3 * ShaderSlots<SH>::N + 3) / 4; // float4s of plane data per record
	static constexpr int PLANE_B = SWCU_TILE_W * SWCU_TILE_H * 4 * MS;              // colour or depth tile
	static constexpr int STENCIL_B = SWCU_TILE_W * SWCU_TILE_H * MS;
	// A batch is bounded three ways when it is formed: NB candidates, PCAP (candidate, region row, sample) pairs and ICAP
	// covered samples (both from the bounding boxes, before any span is read), so the per-warp area stays small enough for
	// 8 CTAs per SM with the 4x MSAA colour + depth tile.
	static constexpr int PCAP = MS == 4 ? 208 : 256;
	static constexpr int ICAP = MS == 4 ? 1536 : 2048;
	// per-warp area
	static constexpr int W_HDR = 0;                                  // uint4 hdr[NB]: {span rows pointer (lo, hi), yMin | rows << 14 | flags << 28, triangle}
	static constexpr int W_PLANES = W_HDR + 16 * NB;                 // float4 planes[NB][NF4]
	static constexpr int W_PAIRS = W_PLANES + 16 * NB * NF4;         // uint32 pairs[PCAP]: start << 18 | cand << 13 | code << 8 | x0 << 4 | (n - 1)
	static constexpr int W_BITS = W_PAIRS + 4 * PCAP;                // uint32 bits[ICAP / 32]: bit i set <=> a pair starts at item i
	static constexpr int W_COV = W_BITS + ICAP / 8;                  // uint32 cov[16]: samples of the region covered by the current range (conflict test)
	static constexpr int W_BYTES = (W_COV + 64 + 15) & ~15;
	static constexpr int HEAD_B = 128;                               // mbarrier + dirty flag
	__host__ __device__ static int total(bool depth, bool stencil, int colorEpp = 1)
	{
		return HEAD_B + PLANE_B * colorEpp + (depth ? PLANE_B : 0) + (stencil ? ((STENCIL_B + 127) & ~127) : 0) + SWCU_TILE_WARPS * W_BYTES;
	}
};

DEVI void cp_async16(void *dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory"); }
DEVI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DEVI void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

// cooperative tile <-> framebuffer copy with 128-bit accesses where the layout allows it (TMA-ineligible attachments)
template<int MS, typename T, bool STORE>
DEVI void tile_copy(T *sm, unsigned char *base, int pitchB, int sliceB, int tileX, int tileY, int fbW, int fbH)
{
	constexpr int ROWB = SWCU_TILE_W * (int)sizeof(T);
	constexpr int VEC = 16;
	constexpr int VPR = ROWB / VEC;
	const bool vecOk = ((size_t)base % VEC == 0) && (pitchB % VEC == 0) && (sliceB % VEC == 0) && ((fbW * (int)sizeof(T)) % VEC == 0);
	const int x0B = tileX * (int)sizeof(T);
	if(vecOk)
	{
		for(int i = threadIdx.x; i < MS * SWCU_TILE_H * VPR; i += TILE_THREADS)
		{
			const int q = i / (SWCU_TILE_H * VPR), r = (i / VPR) % SWCU_TILE_H, c = i % VPR;
			const int y = tileY + r, xB = x0B + c * VEC;
			if(y >= fbH || xB >= fbW * (int)sizeof(T)) continue;
			uint4 *g = (uint4 *)(base + (size_t)q * sliceB + (size_t)y * pitchB + xB);
			uint4 *s = (uint4 *)((unsigned char *)(sm + (q * SWCU_TILE_H + r) * SWCU_TILE_W) + c * VEC);
			if(STORE) *g = *s; else *s = *g;
		}
	}
	else
	{
		for(int i = threadIdx.x; i < MS * SWCU_TILE_H * SWCU_TILE_W; i += TILE_THREADS)
		{
			const int q = i / (SWCU_TILE_H * SWCU_TILE_W), r = (i / SWCU_TILE_W) % SWCU_TILE_H, c = i % SWCU_TILE_W;
			const int y = tileY + r, x = tileX + c;
			if(y >= fbH || x >= fbW) continue;
			T *g = (T *)(base + (size_t)q * sliceB + (size_t)y * pitchB) + x;
			T *s = sm + (q * SWCU_TILE_H + r) * SWCU_TILE_W + c;
			if(STORE) *g = *s; else *s = *g;
		}
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
	CUtensorMap color, depth, stencil;
};

// FS ("fast state"): the host has checked the common fixed-function state — no stencil, full colour write mask, RGBA byte
// order, no depth bias, full sample mask, depth test off or LESS / LESS_OR_EQUAL, perspective slots routed one to one —
// so none of it is decoded per fragment.  FS == false is the same code with every state read at run time.
template<int MS, int SH, int BL, bool FS>
__global__ void __launch_bounds__(TILE_THREADS, MS == 4 ? 8 : 7) k_tile(const __grid_constant__ DrawConst d, const __grid_constant__ TileMaps maps,
                                                                         const uint32_t *tileBegin, const uint32_t *tileEnd, const uint32_t *triList)
{
	using L = TileLayout<MS, SH>;
	constexpr int NB = L::NB, NF4 = L::NF4, PCAP = L::PCAP, ICAP = L::ICAP;
	constexpr int LPC = MS == 4 ? 2 : 1;                 // lanes per candidate in the coverage step
	constexpr int RPL = SWCU_REGION_H / LPC;              // region rows per lane
	static_assert(NB * LPC == 32, "one batch fills the warp");
	constexpr bool TEX = SH == SH_TEX || SH == SH_GENERIC;
	constexpr int UV = SH == SH_TEX ? 0 : 4;
	constexpr int TP = SWCU_TILE_W * SWCU_TILE_H; // pixels per sample plane of the tile
	extern __shared__ __align__(128) unsigned char smem[];

	const int tx = d.tileX0 + blockIdx.x, ty = d.tileY0 + blockIdx.y;
	const int tileId = ty * d.tilesX + tx;
	uint32_t begin, end;
	if(d.direct) { begin = 0; end = d.primCount; }
	else { begin = tileBegin[tileId]; end = tileEnd[tileId]; }
	if(begin >= end) return;

	const bool colorOn = FS ? true : (d.colorWriteMask != 0 && d.colorBuf != nullptr);
	uint64_t *bar = (uint64_t *)smem;
	int *dirtyFlag = (int *)(smem + 8);
	uint32_t *smColor = (uint32_t *)(smem + L::HEAD_B);
	const int colorEpp = FS ? 1 : (int)d.colorEpp; // 32-bit words per colour pixel (floating-point targets: 2 or 4)
	float *smDepth = (float *)(smem + L::HEAD_B + L::PLANE_B * colorEpp);
	unsigned char *smStencil = smem + L::HEAD_B + L::PLANE_B * colorEpp + (d.depthTestActive ? L::PLANE_B : 0);
	unsigned char *warpBase = smStencil + (d.stencilActive ? ((L::STENCIL_B + 127) & ~127) : 0);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t laneLt = (1u << lane) - 1u, laneLe = (2u << lane) - 1u;
	const int tileX = tx * SWCU_TILE_W, tileY = ty * SWCU_TILE_H;
	const int rx = tileX + (warp % (SWCU_TILE_W / SWCU_REGION_W)) * SWCU_REGION_W; // this warp's region
	const int ry = tileY + (warp / (SWCU_TILE_W / SWCU_REGION_W)) * SWCU_REGION_H;
	const int regionPi = (ry - tileY) * SWCU_TILE_W + (rx - tileX); // index of the region's first pixel inside a staged plane

	unsigned char *wa = warpBase + warp * L::W_BYTES;
	uint4 *wHdr = (uint4 *)(wa + L::W_HDR);
	float4 *wPlanes = (float4 *)(wa + L::W_PLANES);
	uint32_t *wPairs = (uint32_t *)(wa + L::W_PAIRS);
	uint32_t *wBits = (uint32_t *)(wa + L::W_BITS);
	uint32_t *wCov = (uint32_t *)(wa + L::W_COV);

	// ---- stage the tile: TMA when the attachments allow it ----
	if(threadIdx.x == 0)
	{
		*dirtyFlag = 0;
		if(d.useTma)
		{
			mbar_init(bar, 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	}
	__syncthreads();
	if(d.useTma)
	{
		if(threadIdx.x == 0)
		{
			const uint32_t bytes = (colorOn ? L::PLANE_B * colorEpp : 0) + (d.depthTestActive ? (d.depth16 ? L::PLANE_B / 2 : L::PLANE_B) : 0) + (d.stencilActive ? L::STENCIL_B : 0);
			mbar_expect_tx(bar, bytes);
			if(colorOn) tma_load_3d(smColor, &maps.color, bar, tileX * colorEpp, tileY, 0); // the map counts 32-bit words along x
			if(d.depthTestActive) tma_load_3d(smDepth, &maps.depth, bar, tileX, tileY, 0);
			if(d.stencilActive) tma_load_3d(smStencil, &maps.stencil, bar, tileX, tileY, 0);
		}
	}
	else
	{
		if(colorOn && colorEpp == 4) tile_copy<MS, uint4, false>((uint4 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		else if(colorOn && colorEpp == 2) tile_copy<MS, uint2, false>((uint2 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		else if(colorOn) tile_copy<MS, uint32_t, false>(smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		if(d.depthTestActive && d.depth16) tile_copy<MS, unsigned short, false>((unsigned short *)smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		else if(d.depthTestActive) tile_copy<MS, float, false>(smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		if(d.stencilActive) tile_copy<MS, unsigned char, false>(smStencil, d.stencilBuf, d.stencilPitchB, d.stencilSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		__syncthreads();
	}
	bool tileReady = !d.useTma;

	bool dirty = false;
	const bool biasOn = FS ? false : d.depthBiasEnable != 0;
	const bool bgr = FS ? false : d.bgr != 0;
	uint32_t wmask32 = FS ? 0xFFFFFFFFu : 0u; // byte lanes of the packed pixel the draw may write
	if(!FS)
	{
#pragma unroll
		for(int ch = 0; ch < 4; ch++)
			if((d.colorWriteMask >> ch) & 1) wmask32 |= 0xFFu << (8 * ((bgr && ch < 3) ? 2 - ch : ch));
	}

	// ---- list scan with look-ahead: while block b of 32 list entries is used, the record headers of block b + 1 and the list
	//      entries of block b + 2 are in flight (the header address depends on the list entry) ----
	uint32_t triN = 0, triNN = 0;
	uint4 hN = make_uint4(0, 0, 0, 0);
	{
		const uint32_t li = begin + lane;
		if(li < end)
		{
			triN = d.direct ? li : __ldg(triList + li);
			hN = __ldg((const uint4 *)(d.triRecords + (size_t)triN * d.triStride));
		}
		if(li + 32 < end) triNN = d.direct ? li + 32 : __ldg(triList + li + 32);
	}
	int ns = 0;             // candidates staged so far for the next batch
	uint32_t pend = 0;      // lanes whose scanned hit is not staged yet
	uint32_t tri = 0, hy = 0;
	const unsigned char *hrows = nullptr;
	for(uint32_t pos = begin;;)
	{
		if(!pend && pos < end)
		{
			// ---- which of these 32 list entries touch my region? ----
			bool hit = false;
			if(pos + lane < end)
			{
				const int pxMin = hN.x & 0xFFFF, pxMax = hN.x >> 16, yMin = hN.y & 0xFFFF, yMax = hN.y >> 16;
				hit = pxMin < rx + SWCU_REGION_W && pxMax > rx && yMin < ry + SWCU_REGION_H && yMax > ry;
				if(hit)
				{
					tri = triN;
					hy = (uint32_t)yMin | ((uint32_t)(yMax - yMin) << 14) | ((hN.w & 3u) << 28);
					// span rows of the triangle, rebased so that entry (y * MS + q) is row y: the span table for big triangles,
					// the rows inlined in the record otherwise
					hrows = (hN.w & 2u) ? (const unsigned char *)(d.spans + hN.z) : d.triRecords + (size_t)triN * d.triStride + TRI_HEADER_BYTES + 16 * NF4;
					hrows -= (size_t)yMin * (MS * 4);
				}
			}
			pend = __ballot_sync(0xFFFFFFFFu, hit);
			pos += 32;
			const uint32_t li = pos + lane;
			if(li < end)
			{
				triN = triNN;
				hN = __ldg((const uint4 *)(d.triRecords + (size_t)triN * d.triStride));
			}
			if(li + 32 < end) triNN = d.direct ? li + 32 : __ldg(triList + li + 32);
		}
		if(pend)
		{
			// ---- hits go to the free slots of the batch, in list order; the rest wait for the next batch ----
			const int rank = __popc(pend & laneLt);
			const int take = min(__popc(pend), NB - ns);
			const bool mine = ((pend >> lane) & 1) && rank < take;
			if(mine) wHdr[ns + rank] = make_uint4((uint32_t)(uintptr_t)hrows, (uint32_t)((uintptr_t)hrows >> 32), hy, tri);
			pend &= ~__ballot_sync(0xFFFFFFFFu, mine);
			ns += take;
		}
		const bool listDone = pend == 0 && pos >= end;
		if(ns < NB && !listDone) continue;
		if(ns == 0) break;
		{
			const int nb = ns;
			ns = 0;
			__syncwarp();
			// ---- plane equations of the batch -> shared memory, asynchronously (needed only when the items are consumed) ----
			for(int i = lane; i < nb * NF4; i += 32)
			{
				const int s = i / NF4, j = i - s * NF4;
				cp_async16(wPlanes + i, d.triRecords + (size_t)wHdr[s].w * d.triStride + TRI_HEADER_BYTES + 16 * j);
			}
			cp_async_commit();

			// ---- coverage (QuadRasterizer.cpp:181-206), one candidate per lane (MS == 1) or per lane pair (MS == 4: rows 0-3 and
			//      4-7 of the region): the lane reads its rows' spans, clips [left, right) to the region's 16 columns — a run of
			//      n pixels from x0 — and keeps the non-empty (row, sample) pairs packed in registers.  A warp scan over the lanes
			//      then places the pairs and their first items, so the lanes write the pair words (and the start marks) directly;
			//      no per-candidate warp iteration, no second pass over the pairs.  A range [c0, c1) of the batch whose pairs
			//      and items fit the shared-memory areas is taken at a time (normally the whole batch) ----
			for(int c0 = 0; c0 < nb;)
			{
			int c1;
			uint32_t total;
			bool conflicts = false; // does any sample of the region receive two fragments in this range?
			if(nb * SWCU_REGION_H * MS <= 32)
			{
				// ---- a batch of one (4x) or up to four (1x) candidates — big triangles, direct mode: one (candidate, row, sample)
				//      per lane, pairs compacted with a ballot, first items from one warp scan ----
				const int cand = lane / (SWCU_REGION_H * MS), row = (lane / MS) % SWCU_REGION_H, q = lane % MS;
				uint32_t v = 0;
				if(cand < nb && (MS == 1 || FS || ((d.sampleMask >> q) & 1)))
				{
					const uint4 hh = wHdr[cand];
					const uint32_t y = (uint32_t)(ry + row);
					if(y - (hh.z & 0x3FFFu) < ((hh.z >> 14) & 0x3FFFu))
						v = __ldg((const uint32_t *)(((uintptr_t)hh.y << 32) | hh.x) + y * MS + q);
				}
#pragma unroll
				for(int i = 0; i < (ICAP / 32 + 31) / 32; i++)
					if(lane + 32 * i < ICAP / 32) wBits[lane + 32 * i] = 0;
				const int a = clampi((int)(v & 0xFFFF) - rx, 0, 16), e = clampi((int)(v >> 16) - rx, 0, 16);
				const int n = e - a;
				const bool has = n > 0;
				const uint32_t nz = __ballot_sync(0xFFFFFFFFu, has);
				uint32_t incl = has ? (uint32_t)n : 0u;
#pragma unroll
				for(int o = 1; o < 32; o <<= 1)
				{
					const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
					if(lane >= o) incl += t;
				}
				total = __shfl_sync(0xFFFFFFFFu, incl, 31);
				c1 = nb;
				if(MS == 1 && nb > 1)
				{
					// the other candidates' runs on my row sit 8, 16 and 24 lanes away
					const uint32_t mask = has ? ((1u << n) - 1u) << a : 0u;
					const uint32_t others = __shfl_xor_sync(0xFFFFFFFFu, mask, 8) | __shfl_xor_sync(0xFFFFFFFFu, mask, 16) | __shfl_xor_sync(0xFFFFFFFFu, mask, 24);
					conflicts = __any_sync(0xFFFFFFFFu, (mask & others) != 0);
				}
				__syncwarp(); // start marks zeroed
				if(has)
				{
					const uint32_t start = incl - (uint32_t)n;
					wPairs[__popc(nz & laneLt)] = (start << 18) | ((uint32_t)cand << 13) | ((uint32_t)(row * MS + q) << 8) | ((uint32_t)a << 4) | (uint32_t)(n - 1);
					atomicOr(wBits + (start >> 5), 1u << (start & 31));
				}
				__syncwarp();
			}
			else
			{
				const int cand = c0 + lane / LPC;
				const int rowBase = (lane % LPC) * RPL; // first region row of this lane
				const bool active = cand < nb;
				uint4 hh = make_uint4(0, 0, 0, 0);
				if(active) hh = wHdr[cand];
				const unsigned char *rowsPtr = (const unsigned char *)(((uintptr_t)hh.y << 32) | hh.x);
				const uint32_t yMin = hh.z & 0x3FFFu, nrows = (hh.z >> 14) & 0x3FFFu;
				uint32_t sp[RPL][MS];
#pragma unroll
				for(int r = 0; r < RPL; r++)
				{
					const uint32_t y = (uint32_t)(ry + rowBase + r);
					const bool in = active && (y - yMin) < nrows;
					if(MS == 4)
					{
						uint4 v = make_uint4(0, 0, 0, 0); // empty spans outside the triangle's rows
						if(in) v = __ldg((const uint4 *)(rowsPtr + (size_t)y * 16));
						sp[r][0] = v.x; sp[r][MS > 1 ? 1 : 0] = v.y; sp[r][MS > 1 ? 2 : 0] = v.z; sp[r][MS > 1 ? 3 : 0] = v.w;
					}
					else
					{
						sp[r][0] = 0;
						if(in) sp[r][0] = __ldg((const uint32_t *)(rowsPtr + (size_t)y * 4));
					}
				}
				// zero the start marks while the loads fly (the previous rounds ended with a __syncwarp)
#pragma unroll
				for(int i = 0; i < (ICAP / 32 + 31) / 32; i++)
					if(lane + 32 * i < ICAP / 32) wBits[lane + 32 * i] = 0;
				if(MS == 4 && lane < 16) wCov[lane] = 0;
				uint32_t runs[RPL * MS / 4]; // per pair: x0 << 4 | (n - 1), four pairs per register
				uint32_t valid = 0, items = 0;
				uint32_t cw[MS == 1 ? RPL / 2 : 1]; // MS == 1: my candidate's coverage of the region, two rows per word
				if(MS == 1)
				{
#pragma unroll
					for(int i = 0; i < RPL / 2; i++) cw[i] = 0;
				}
#pragma unroll
				for(int j = 0; j < RPL * MS; j++)
				{
					const int r = j / MS, q = j % MS;
					uint32_t v = sp[r][q];
					if(MS == 4 && !FS && !((d.sampleMask >> q) & 1)) v = 0;
					const int a = clampi((int)(v & 0xFFFF) - rx, 0, 16), e = clampi((int)(v >> 16) - rx, 0, 16);
					const int n = e - a;
					const bool has = n > 0;
					const uint32_t run = has ? (((uint32_t)a << 4) | (uint32_t)(n - 1)) : 0u;
					if(j % 4 == 0) runs[j / 4] = run; else runs[j / 4] |= run << (8 * (j % 4));
					if(has) { valid |= 1u << j; items += (uint32_t)n; }
					if(MS == 1) cw[j / 2 < RPL / 2 ? j / 2 : 0] |= (has ? ((1u << n) - 1u) << a : 0u) << (16 * (j & 1));
				}
				// ---- where do my pairs and items start?  (inclusive scan of pairs << 16 | items over the lanes) ----
				const uint32_t mineU = ((uint32_t)__popc(valid) << 16) | items;
				uint32_t incl = mineU;
#pragma unroll
				for(int o = 1; o < 32; o <<= 1)
				{
					const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
					if(lane >= o) incl += t;
				}
				// a candidate is taken whole: the sums at its last lane must fit
				const uint32_t inclC = LPC == 1 ? incl : __shfl_sync(0xFFFFFFFFu, incl, lane | (LPC - 1));
				const bool fits = active && (inclC >> 16) <= (uint32_t)PCAP && (inclC & 0xFFFFu) <= (uint32_t)ICAP;
				const uint32_t fitMask = __ballot_sync(0xFFFFFFFFu, fits); // a prefix of the lanes, never empty (one candidate always fits)
				const int lastLane = 31 - __clz(fitMask);
				const uint32_t sums = __shfl_sync(0xFFFFFFFFu, incl, lastLane);
				total = sums & 0xFFFFu;
				c1 = c0 + (lastLane + 1) / LPC;
				__syncwarp(); // start marks zeroed
				if(fits)
				{
					uint32_t pairIdx = (incl - mineU) >> 16, start = (incl - mineU) & 0xFFFFu;
					const uint32_t word0 = ((uint32_t)cand << 13) | (MS == 4 ? (uint32_t)rowBase << 10 : 0u);
#pragma unroll
					for(int j = 0; j < RPL * MS; j++)
					{
						if((valid >> j) & 1)
						{
							const uint32_t run = (runs[j / 4] >> (8 * (j % 4))) & 0xFFu;
							// code (bits 8-12) = region row << 2 | sample for MS == 4 (j = r * 4 + q, the lane's first row comes in
							// through word0), region row for MS == 1 (j = r)
							wPairs[pairIdx] = (start << 18) | (word0 + ((uint32_t)j << 8)) | run;
							atomicOr(wBits + (start >> 5), 1u << (start & 31));
							pairIdx++;
							start += (run & 15u) + 1u;
						}
					}
				}
				__syncwarp();
				// ---- two fragments on one sample?  Every pair ORs its run into a bitmap of the region's samples; a bit that was
				//      already set means an earlier (or later) pair covers the sample too.  One candidate alone cannot overlap itself ----
				if(c1 - c0 > 1)
				{
					if(MS == 1)
					{
						// one candidate per lane, all lanes see the same eight rows: the samples covered by the range are the OR of the
						// lanes' words; fewer of them than items means some sample is covered twice
						uint32_t covered = 0;
#pragma unroll
						for(int i = 0; i < RPL / 2; i++) covered += __popc(__reduce_or_sync(0xFFFFFFFFu, fits ? cw[i] : 0u));
						conflicts = covered != total;
					}
					else
					{
						const uint32_t P = sums >> 16;
						uint32_t ov = 0;
#pragma unroll 2
						for(uint32_t p = lane; p < P; p += 32)
						{
							const uint32_t w = wPairs[p];
							const uint32_t code = (w >> 8) & 31u;
							const uint32_t m = ((2u << (w & 15u)) - 1u) << (((w >> 4) & 15u) + 16u * (code & 1u));
							ov |= atomicOr(wCov + (code >> 1), m) & m;
						}
						conflicts = __any_sync(0xFFFFFFFFu, ov != 0);
					}
				}
			}
			cp_async_wait_all();
			__syncwarp();
			if(total && !tileReady)
			{
				while(!mbar_try_wait(bar, 0)) {} // the TMA loads of the tile have landed
				tileReady = true;
			}
			// ---- consume the items 32 at a time, one per lane ----
			{
				int cursor = 0; // pairs that start before this round
				for(uint32_t base = 0; base < total; base += 32)
				{
					const uint32_t g = base + lane;
					const bool valid = g < total;
					const uint32_t starts = wBits[base >> 5];
					const int myPair = cursor + __popc(starts & laneLe) - 1; // the last pair that starts at or before my item
					cursor += __popc(starts);
					const uint32_t e = wPairs[myPair];
					const int bit = (int)((e >> 4) & 15u) + (int)(g - (e >> 18));
					const uint32_t code = (e >> 8) & 31u;
					// items of the same (pixel, sample) in this round run in queue order
					const uint32_t key = valid ? ((code << 4) | (uint32_t)bit) : (0x200u | lane); // one key per sample of the region
					int prank = 0, maxRank = 0;
					if(conflicts) // overlapping triangles in this range: same-sample items of a round run in list order
					{
						const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
						prank = __popc(peers & laneLt);
						maxRank = (int)__reduce_max_sync(0xFFFFFFFFu, (uint32_t)(valid ? prank : 0));
					}
					for(int rr = 0; rr <= maxRank; rr++)
					{
						if(valid && prank == rr)
						{
							const int k = (e >> 13) & 31;
							const int q = MS == 4 ? (int)(code & 3) : 0;
							const int row = MS == 4 ? (int)(code >> 2) : (int)code;
							const int x = rx + bit, y = ry + row;
							const int ix = x & 1, iy = y & 1;
							const int pi = q * TP + regionPi + row * SWCU_TILE_W + bit; // index inside the staged planes
							// ---- plane equations of the triangle ----
							float pf[NF4 * 4];
#pragma unroll
							for(int t = 0; t < NF4; t++)
							{
								const float4 v = wPlanes[k * NF4 + t];
								pf[4 * t] = v.x; pf[4 * t + 1] = v.y; pf[4 * t + 2] = v.z; pf[4 * t + 3] = v.w;
							}
							const float x0 = pf[0], y0 = pf[1], zBias = pf[2], wA = pf[3], wB = pf[4], wC = pf[5], zA = pf[6], zB = pf[7], zC = pf[8];
							const float *S = pf + TRI_FLOATS_FIXED; // slot s: S[3s], S[3s+1], S[3s+2]
							// ---- interpolate + routed fragment shader (PixelRoutine.cpp:196-261, PixelProgram.cpp:138-241) ----
							const float xf = fsub((float)x, x0), yf = fsub((float)y, y0);
							const float rhwConst = pf[NF4 * 4 - 1]; // 1/w of a constant w plane (k_setup), 0 otherwise
							float rhw = 1.0f;
							if(SH != SH_CONST) rhw = rhwConst != 0.0f ? rhwConst : fdiv(1.0f, __fmaf_rn(xf, wA, fadd(wC, fmul(yf, wB))));
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
									const float rk = rhwConst != 0.0f ? rhwConst : fdiv(1.0f, __fmaf_rn(xk, wA, fadd(wC, fmul(yk, wB))));
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
										val = interp_slot(S[3 * ch], S[3 * ch + 1], S[3 * ch + 2], IM_PERSP, xf, yf, rhw);
										if(d.chanKind[ch] == CK_CONST) val = __uint_as_float(d.chanValue[ch]);
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

							// ---- stencil test, depth test, depth write, blend + colour write, stencil write ----
							bool sPass = true;
							uint32_t sValue = 0;
							const uint32_t frontFacing = FS ? 1u : (wHdr[k].z >> 28) & 1u;
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
								z = __fmaf_rn(xx, zA, fadd(zC, fmul(yy, zB)));
								if(biasOn) z = fadd(z, zBias);
								z = sse_min(sse_max(z, 0.0f), 1.0f); // clampDepth :484-492
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
											const float sa = rgba[3], da = fsub(1.0f, rgba[3]);
#pragma unroll
											for(int ch = 0; ch < 3; ch++) o[ch] = fadd(fmul(rgba[ch], sa), fmul(dst[ch], da));
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
										smColor[pi] = (px & ~wmask32) | (pk & wmask32);
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
						__syncwarp();
					}
				}
				__syncwarp();
			}
			__syncwarp(); // the pair / mark areas are reused by the next range
			c0 = c1;
			}
		}
		if(listDone) break;
	}

	if(!tileReady)
		while(!mbar_try_wait(bar, 0)) {} // never leave with a bulk copy into this CTA's shared memory still in flight
	if(__any_sync(0xFFFFFFFFu, dirty) && lane == 0) *dirtyFlag = 1;
	__syncthreads();
	if(!*dirtyFlag) return;
	if(d.useTma)
	{
		fence_proxy_async(); // generic-proxy writes of the tile -> visible to the async proxy
		__syncthreads();
		if(threadIdx.x == 0)
		{
			if(colorOn) tma_store_3d(&maps.color, smColor, tileX * colorEpp, tileY, 0);
			if(d.depthWriteEnable) tma_store_3d(&maps.depth, smDepth, tileX, tileY, 0);
			if(d.stencilWrite) tma_store_3d(&maps.stencil, smStencil, tileX, tileY, 0);
			tma_commit();
			tma_wait_read0();
		}
	}
	else
	{
		if(colorOn && colorEpp == 4) tile_copy<MS, uint4, true>((uint4 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		else if(colorOn && colorEpp == 2) tile_copy<MS, uint2, true>((uint2 *)smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		else if(colorOn) tile_copy<MS, uint32_t, true>(smColor, d.colorBuf, d.colorPitchB, d.colorSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		if(d.depthWriteEnable && d.depth16) tile_copy<MS, unsigned short, true>((unsigned short *)smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		else if(d.depthWriteEnable) tile_copy<MS, float, true>(smDepth, d.depthBuf, d.depthPitchB, d.depthSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
		if(d.stencilWrite) tile_copy<MS, unsigned char, true>(smStencil, d.stencilBuf, d.stencilPitchB, d.stencilSliceB, tileX, tileY, d.fbWidth, d.fbHeight);
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
