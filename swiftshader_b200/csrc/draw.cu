// draw.cu — host side of the C-ABI in include/swcu.h: the context (device shadows of host memory, work buffers,
// stream), the state gathering of sw::Renderer::draw (reference: src/Device/Renderer.cpp:183-490) and the kernel
// sequence that replaces DrawCall::run (Renderer.cpp:551-662).  No torch types, no CPU rendering fallback.
#include "kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

static thread_local std::string g_lastError;

struct DevBuf
{
	void *p = nullptr;
	size_t cap = 0;
};

struct Shadow
{
	uintptr_t host = 0;
	size_t bytes = 0;
	unsigned char *dev = nullptr;
	bool pinned = false;
	bool external = false; // caller-owned device memory (swcu_mem_register_device): identity mapping
	// copy-stream hazards: a download in flight on the download stream must finish before the shadow is overwritten (by a
	// kernel on the main stream or by an upload); a shadow kernels of the main stream touch (attachment, sampled image,
	// clear / resolve / copy operand) must not be overwritten by an upload before those kernels are done
	cudaEvent_t dlEvent = nullptr;
	bool dlPendingMain = false, dlPendingUpload = false;
	bool mainTouched = false;
	// the last upload into this shadow: readers wait for THIS one, not for uploads issued after it (the next frame's inputs
	// already on their way up must not hold back the current frame)
	cudaEvent_t upEvent = nullptr;
	uint64_t upSeq = 0;
};

struct KernelTime
{
	const char *name;
	cudaEvent_t e0, e1;
};

#define SWCU_SETS 3 // buffer sets of the setup phase: in a group the peers write set d + 1 while this rank may still read sets d and d - 1

struct swcu_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr, ownStream = nullptr;
	std::map<uintptr_t, Shadow> mem;
	std::string err;
	// Buffers of the setup phase (k_cull, k_setup, binning) exist twice: the setup of draw i+1 runs on its own stream while the tile
	// kernel of draw i is still reading the other set.  Nothing of a draw has to reach the host before its tile kernel is launched:
	// the pair buffer holds 4 pairs per triangle plus a budget for the big triangles (swcu_set_option "big_pair_budget").
	struct SetupSet
	{
		DevBuf triRecords, bigList, triRect, binCount, binStart, pairs, counters, cullFlags;
		cudaEvent_t tileDone = nullptr;       // recorded on the main stream after the last kernel that reads this set
		bool tileDoneValid = false;
		cudaEvent_t setupDone = nullptr;      // recorded on the setup stream after the binning of a pipelined draw
		DrawCounters *hostCounters = nullptr; // pinned: the counters of the last draw that used this set, copied by its k_tile's successor
		bool countersPending = false;
	} set[SWCU_SETS];
	int cur = 0;
	// The setup streams.  A lone context uses [0] for every pipelined draw.  A group member uses one PER SET: its chain of a draw
	// (set-up -> group barrier -> scan -> fill) is a string of short, latency-bound launches that takes longer than the tile
	// kernel of a band, so on one stream the chains themselves would set the frame rate; on three streams the set-up of draw i+1
	// starts as soon as the barrier of draw i has been passed (evBarrier) and runs beside the scan / fill of draw i.
	cudaStream_t setupStream[SWCU_SETS] = {};
	// The hand-over stream (swcu_side_begin / _end / _wait): what follows a frame's last kernel on its way to whoever presents it —
	// waiting for the flags of a ring slot, the copy of a band into a peer's frame over NVLink, the signal, the download — runs
	// beside the next frame's kernels instead of between them.
	cudaStream_t sideStream = nullptr;
	bool sideActive = false;
	cudaEvent_t evSideBegin = nullptr;
	cudaEvent_t sideDone[SWCU_SIDE_SLOTS] = {};
	bool sideDoneValid[SWCU_SIDE_SLOTS] = {};
	cudaStream_t handover() const { return sideActive ? sideStream : stream; } // where hand-over calls are issued
	cudaEvent_t evBarrier = nullptr;   // group: recorded behind the k_xbarrier of the last grouped draw
	bool evBarrierValid = false;
	cudaEvent_t evSetupMark[SWCU_SETS] = {}; // scratch: "this setup stream up to here" (uploads wait for the readers of the old contents)
	bool setupReadsInputs[SWCU_SETS] = {};   // a setup phase ran on this setup stream since the last upload looked
	// Host<->device copies run on their own two streams (one per DMA direction), so the upload of the next frame's inputs and
	// the download of the previous frame overlap the kernels of the current one; events order them against the kernels:
	cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
	cudaEvent_t evUpload = nullptr;  // after the last swcu_mem_upload
	uint64_t uploadSeq = 0, mainSawUpload = 0, setupSawUpload[SWCU_SETS] = {}, d2hSawUpload = 0;
	cudaEvent_t evMark = nullptr;    // scratch: "the main stream up to here"
	cudaEvent_t evDownload = nullptr; // after the last swcu_mem_download
	uint64_t downloadSeq = 0, mainSawDownload = 0;
	bool mainReadsInputs = false;    // a setup phase ran on the main stream since the last upload looked
	cudaEvent_t fence[SWCU_MAX_FENCES] = {};
	bool fenceValid[SWCU_MAX_FENCES] = {};
	int optCopyStreams = 1;
	int optPipeline = 1;
	DevBuf zeroPage;
	// ---- group (swcu_group_*): the work buffers peers write into live in ONE allocation exported over CUDA IPC ----
	struct Group
	{
		bool reserved = false, attached = false;
		uint32_t rank = 0, world = 1, maxPrims = 0, stride = 0, fbW = 0, fbH = 0, numBins = 0;
		size_t countersBytes = 0;
		unsigned char *arena = nullptr;
		size_t arenaBytes = 0;
		size_t offRecords[SWCU_SETS], offRect[SWCU_SETS], offBig[SWCU_SETS], offBinCount[SWCU_SETS], offCounters[SWCU_SETS], offFlags[SWCU_SETS];
		unsigned char *peer[SWCU_MAX_GROUP] = {}; // mapped arenas, [rank] = my own
		uint32_t epoch[SWCU_SETS] = {};
	} group;
	size_t optBigPairBudget = (size_t)16 << 20; // (region, triangle) pairs the big triangles of one draw may occupy
	int lastOverflow = 0;
	swcu_stats stats{};
	cudaEvent_t t0 = nullptr, t1 = nullptr;
	struct CachedShader { std::vector<uint32_t> words; swcu_shader_info info; };
	std::multimap<uint64_t, CachedShader> shaderCache; // keyed by a hash of the module, entries compared word for word on a hit
	int optForceBinned = 0, optDirectMax = 64, optPinHost = 1;
	int profiling = 0;
	std::vector<KernelTime> lastKernels;
	std::vector<cudaEvent_t> eventPool;
	size_t eventsUsed = 0;
	int optTma = 1;
	int optFastState = 1;
	int optWriteOnly = 1;
	int optSetupWide = 1;
	int smCount = 148;
	void *encodeTiled = nullptr; // cuTensorMapEncodeTiled, resolved through the runtime (no -lcuda link dependency)
	std::map<std::vector<uint64_t>, CUtensorMap> mapCache;
};

static int fail(swcu_ctx *ctx, int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if(ctx) ctx->err = buf;
	g_lastError = buf;
	return code;
}

#define CU(call)                                                                                              \
	do {                                                                                                      \
		cudaError_t _e = (call);                                                                              \
		if(_e != cudaSuccess) return fail(ctx, SWCU_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
	} while(0)

static int ensure(swcu_ctx *ctx, DevBuf &b, size_t bytes)
{
	if(bytes <= b.cap) return SWCU_OK;
	size_t want = std::max(bytes, b.cap + b.cap / 2);
	want = (want + 255) & ~(size_t)255;
	if(b.p)
	{
		CU(cudaStreamSynchronize(ctx->stream)); // earlier launches may still read the old block
		for(cudaStream_t ss : ctx->setupStream)
			if(ss) CU(cudaStreamSynchronize(ss));
		CU(cudaFree(b.p));
		b.p = nullptr;
		b.cap = 0;
	}
	cudaError_t e = cudaMalloc(&b.p, want);
	if(e != cudaSuccess) { cudaGetLastError(); return fail(ctx, SWCU_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
	b.cap = want;
	return SWCU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------------------------
// The setup phase of draw i+1 (set-up, group barrier, scan, fill: mostly short, latency-bound launches) runs beside the tile kernel of
// draw i, which fills the machine.  In a group — where that chain is a third of the frame — its stream gets the highest priority: its
// blocks are placed ahead of the tile kernel's remaining ones instead of behind them, and the chain is over when the tile kernel
// drains (C4 at 8 GPUs: 0.151 -> 0.143 ms per frame; on one GPU the same setting costs 2 %, so a lone context keeps the default).
// SWCU_SETUP_PRIORITY=0 / 1 forces it off / on.
static cudaError_t make_setup_stream(swcu_ctx *ctx, bool grouped)
{
	int least = 0, greatest = 0;
	cudaDeviceGetStreamPriorityRange(&least, &greatest);
	const char *pe = getenv("SWCU_SETUP_PRIORITY");
	const bool high = pe ? pe[0] != '0' : grouped;
	for(int i = 0; i < SWCU_SETS; i++)
	{
		if(ctx->setupStream[i])
		{
			cudaStreamSynchronize(ctx->setupStream[i]);
			cudaStreamDestroy(ctx->setupStream[i]);
			ctx->setupStream[i] = nullptr;
		}
		cudaError_t e = cudaStreamCreateWithPriority(&ctx->setupStream[i], cudaStreamNonBlocking, high ? greatest : least);
		if(e != cudaSuccess) return e;
		ctx->setupSawUpload[i] = 0; // (a new stream has waited for nothing yet)
		ctx->setupReadsInputs[i] = false;
	}
	ctx->evBarrierValid = false;
	return cudaSuccess;
}

extern "C" int swcu_create(swcu_ctx **out, int device_ordinal)
{
	if(!out) return fail(nullptr, SWCU_E_INVALID, "swcu_create: null out");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if(e != cudaSuccess || count == 0)
	{
		cudaGetLastError();
		return fail(nullptr, SWCU_E_CUDA, "no CUDA device available (%s); the draw path has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
	}
	if(device_ordinal < 0 || device_ordinal >= count) return fail(nullptr, SWCU_E_INVALID, "device ordinal %d out of range (0..%d)", device_ordinal, count - 1);
	swcu_ctx *ctx = new swcu_ctx();
	ctx->device = device_ordinal;
	auto bail = [&](const char *what, cudaError_t err) {
		int rc = fail(nullptr, SWCU_E_CUDA, "%s: %s", what, cudaGetErrorString(err));
		delete ctx;
		return rc;
	};
	if((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return bail("cudaSetDevice", e);
	cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device_ordinal);
	if((e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
	ctx->stream = ctx->ownStream;
	if((e = cudaEventCreate(&ctx->t0)) != cudaSuccess) return bail("cudaEventCreate", e);
	if((e = cudaEventCreate(&ctx->t1)) != cudaSuccess) return bail("cudaEventCreate", e);
	if((e = make_setup_stream(ctx, false)) != cudaSuccess) return bail("cudaStreamCreate", e);
	if((e = cudaEventCreateWithFlags(&ctx->evUpload, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
	if((e = cudaEventCreateWithFlags(&ctx->evMark, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
	if((e = cudaEventCreateWithFlags(&ctx->evDownload, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
	if((e = cudaStreamCreateWithFlags(&ctx->h2dStream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
	if((e = cudaStreamCreateWithFlags(&ctx->d2hStream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
	for(auto &f : ctx->fence)
		if((e = cudaEventCreateWithFlags(&f, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
	for(auto &S : ctx->set)
	{
		if((e = cudaMallocHost((void **)&S.hostCounters, sizeof(DrawCounters))) != cudaSuccess) return bail("cudaMallocHost", e);
		memset(S.hostCounters, 0, sizeof(DrawCounters));
		if((e = cudaEventCreateWithFlags(&S.tileDone, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
		if((e = cudaEventCreateWithFlags(&S.setupDone, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
	}
	*out = ctx;
	return SWCU_OK;
}

extern "C" void swcu_destroy(swcu_ctx *ctx)
{
	if(!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if(ctx->h2dStream) cudaStreamSynchronize(ctx->h2dStream);
	if(ctx->d2hStream) cudaStreamSynchronize(ctx->d2hStream);
	for(auto &kv : ctx->mem)
	{
		if(kv.second.pinned) cudaHostUnregister((void *)kv.second.host);
		if(!kv.second.external) cudaFree(kv.second.dev);
		if(kv.second.dlEvent) cudaEventDestroy(kv.second.dlEvent);
		if(kv.second.upEvent) cudaEventDestroy(kv.second.upEvent);
	}
	for(cudaStream_t ss : ctx->setupStream)
		if(ss) cudaStreamSynchronize(ss);
	swcu_group_detach(ctx);
	cudaFree(ctx->zeroPage.p);
	for(auto &S : ctx->set)
	{
		DevBuf *sb[] = { &S.triRecords, &S.bigList, &S.triRect, &S.binCount, &S.binStart, &S.pairs, &S.counters, &S.cullFlags };
		for(DevBuf *b : sb) cudaFree(b->p);
		if(S.hostCounters) cudaFreeHost(S.hostCounters);
		if(S.tileDone) cudaEventDestroy(S.tileDone);
		if(S.setupDone) cudaEventDestroy(S.setupDone);
	}
	if(ctx->evUpload) cudaEventDestroy(ctx->evUpload);
	if(ctx->evMark) cudaEventDestroy(ctx->evMark);
	if(ctx->evDownload) cudaEventDestroy(ctx->evDownload);
	for(auto &f : ctx->fence)
		if(f) cudaEventDestroy(f);
	if(ctx->h2dStream) cudaStreamDestroy(ctx->h2dStream);
	if(ctx->d2hStream) cudaStreamDestroy(ctx->d2hStream);
	for(cudaStream_t ss : ctx->setupStream)
		if(ss) cudaStreamDestroy(ss);
	for(cudaEvent_t ev : ctx->evSetupMark)
		if(ev) cudaEventDestroy(ev);
	if(ctx->evBarrier) cudaEventDestroy(ctx->evBarrier);
	if(ctx->sideStream) { cudaStreamSynchronize(ctx->sideStream); cudaStreamDestroy(ctx->sideStream); }
	if(ctx->evSideBegin) cudaEventDestroy(ctx->evSideBegin);
	for(cudaEvent_t ev : ctx->sideDone)
		if(ev) cudaEventDestroy(ev);
	for(cudaEvent_t ev : ctx->eventPool) cudaEventDestroy(ev);
	if(ctx->t0) cudaEventDestroy(ctx->t0);
	if(ctx->t1) cudaEventDestroy(ctx->t1);
	if(ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
	cudaGetLastError();
	delete ctx;
}

extern "C" const char *swcu_last_error(swcu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_lastError.c_str(); }
extern "C" const char *swcu_version(void) { return "swcuda 0.1 (sm_100a draw path: setup / spans / tile binning / tile raster)"; }

// ------------------------------------------------------------------------------------------------------------------
// memory shadows
// ------------------------------------------------------------------------------------------------------------------
static Shadow *find_shadow(swcu_ctx *ctx, const void *p, size_t bytes)
{
	const uintptr_t a = (uintptr_t)p;
	auto it = ctx->mem.upper_bound(a);
	if(it == ctx->mem.begin()) return nullptr;
	--it;
	Shadow &s = it->second;
	if(a < s.host || a + bytes > s.host + s.bytes) return nullptr;
	return &s;
}

static unsigned char *dev_ptr(swcu_ctx *ctx, const void *p, size_t bytes = 1)
{
	Shadow *s = find_shadow(ctx, p, bytes);
	return s ? s->dev + ((uintptr_t)p - s->host) : nullptr;
}

extern "C" int swcu_mem_register(swcu_ctx *ctx, const void *host_base, size_t bytes)
{
	if(!ctx || !host_base || !bytes) return fail(ctx, SWCU_E_INVALID, "swcu_mem_register: bad arguments");
	CU(cudaSetDevice(ctx->device));
	const uintptr_t a = (uintptr_t)host_base;
	auto it = ctx->mem.upper_bound(a + bytes - 1);
	if(it != ctx->mem.begin())
	{
		--it;
		if(it->second.host + it->second.bytes > a) return fail(ctx, SWCU_E_INVALID, "swcu_mem_register: range overlaps a registered range");
	}
	Shadow s;
	s.host = a;
	s.bytes = bytes;
	cudaError_t e = cudaMalloc((void **)&s.dev, (bytes + 255) & ~(size_t)255);
	if(e != cudaSuccess) { cudaGetLastError(); return fail(ctx, SWCU_E_NOMEM, "cudaMalloc(%zu) for a shadow failed: %s", bytes, cudaGetErrorString(e)); }
	if(ctx->optPinHost)
	{
		// page-lock the host range so uploads/downloads are true async DMA; best effort (fails on read-only mappings)
		e = cudaHostRegister((void *)host_base, bytes, cudaHostRegisterDefault);
		if(e == cudaSuccess) s.pinned = true; else cudaGetLastError();
	}
	ctx->mem[a] = s;
	return SWCU_OK;
}

extern "C" int swcu_mem_register_device(swcu_ctx *ctx, void *device_base, size_t bytes)
{
	if(!ctx || !device_base || !bytes) return fail(ctx, SWCU_E_INVALID, "swcu_mem_register_device: bad arguments");
	const uintptr_t a = (uintptr_t)device_base;
	auto it = ctx->mem.upper_bound(a + bytes - 1);
	if(it != ctx->mem.begin())
	{
		--it;
		if(it->second.host + it->second.bytes > a) return fail(ctx, SWCU_E_INVALID, "swcu_mem_register_device: range overlaps a registered range");
	}
	Shadow s;
	s.host = a;
	s.bytes = bytes;
	s.dev = (unsigned char *)device_base;
	s.external = true;
	ctx->mem[a] = s;
	return SWCU_OK;
}

extern "C" int swcu_mem_unregister(swcu_ctx *ctx, const void *host_base)
{
	if(!ctx) return SWCU_E_INVALID;
	auto it = ctx->mem.find((uintptr_t)host_base);
	if(it == ctx->mem.end()) return fail(ctx, SWCU_E_INVALID, "swcu_mem_unregister: %p is not a registered base", host_base);
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->h2dStream));
	for(cudaStream_t ss : ctx->setupStream) CU(cudaStreamSynchronize(ss));
	CU(cudaStreamSynchronize(ctx->stream));
	if(ctx->sideStream) CU(cudaStreamSynchronize(ctx->sideStream));
	CU(cudaStreamSynchronize(ctx->d2hStream));
	if(it->second.pinned) cudaHostUnregister((void *)it->second.host);
	if(!it->second.external) cudaFree(it->second.dev);
	if(it->second.dlEvent) cudaEventDestroy(it->second.dlEvent);
	if(it->second.upEvent) cudaEventDestroy(it->second.upEvent);
	ctx->mem.erase(it);
	return SWCU_OK;
}

// ---- ordering between the copy streams and the kernel streams ----
// a stream is about to read shadow s: its last upload must have landed (uploads complete in issue order, so a stream that has
// waited for upload number n has seen every earlier one)
static int see_upload(swcu_ctx *ctx, Shadow *s, cudaStream_t st, uint64_t &seen)
{
	if(s && s->upSeq > seen)
	{
		CU(cudaStreamWaitEvent(st, s->upEvent, 0));
		seen = s->upSeq;
	}
	return SWCU_OK;
}
// ... or every upload issued so far
static int see_uploads(swcu_ctx *ctx, cudaStream_t st, uint64_t &seen)
{
	if(seen != ctx->uploadSeq)
	{
		CU(cudaStreamWaitEvent(st, ctx->evUpload, 0));
		seen = ctx->uploadSeq;
	}
	return SWCU_OK;
}
// the main stream is about to read or write [p, ...): later uploads wait for it, and a download of the shadow still in
// flight finishes before a write
static int main_touches(swcu_ctx *ctx, const void *p, bool write)
{
	Shadow *s = p ? find_shadow(ctx, p, 1) : nullptr;
	if(!s) return SWCU_OK;
	s->mainTouched = true;
	int rc = see_upload(ctx, s, ctx->stream, ctx->mainSawUpload);
	if(rc) return rc;
	if(write && s->dlPendingMain)
	{
		CU(cudaStreamWaitEvent(ctx->stream, s->dlEvent, 0));
		s->dlPendingMain = false;
	}
	return SWCU_OK;
}

extern "C" int swcu_mem_upload(swcu_ctx *ctx, const void *host_ptr, size_t bytes)
{
	if(!ctx) return SWCU_E_INVALID;
	unsigned char *d = dev_ptr(ctx, host_ptr, bytes);
	if(!d) return fail(ctx, SWCU_E_INVALID, "swcu_mem_upload: [%p,+%zu) is not inside a registered range", host_ptr, bytes);
	if(d == (const unsigned char *)host_ptr) return fail(ctx, SWCU_E_INVALID, "swcu_mem_upload: %p is caller-owned device memory", host_ptr);
	CU(cudaSetDevice(ctx->device));
	Shadow *s = find_shadow(ctx, host_ptr, bytes);
	const cudaStream_t st = ctx->optCopyStreams ? ctx->h2dStream : ctx->stream;
	if(ctx->optCopyStreams)
	{
		// Readers of the old contents: kernels of the main stream, and the setup phases of pipelined draws on the setup streams
		// (they read the vertex / index streams)
		if(s->mainTouched || ctx->mainReadsInputs)
		{
			CU(cudaEventRecord(ctx->evMark, ctx->stream));
			CU(cudaStreamWaitEvent(st, ctx->evMark, 0));
			ctx->mainReadsInputs = false;
		}
		for(int i = 0; i < SWCU_SETS; i++)
		{
			if(!ctx->setupReadsInputs[i]) continue;
			if(!ctx->evSetupMark[i]) CU(cudaEventCreateWithFlags(&ctx->evSetupMark[i], cudaEventDisableTiming));
			CU(cudaEventRecord(ctx->evSetupMark[i], ctx->setupStream[i]));
			CU(cudaStreamWaitEvent(st, ctx->evSetupMark[i], 0));
			ctx->setupReadsInputs[i] = false;
		}
		if(s->dlPendingUpload)
		{
			CU(cudaStreamWaitEvent(st, s->dlEvent, 0));
			s->dlPendingUpload = false;
		}
	}
	CU(cudaMemcpyAsync(d, host_ptr, bytes, cudaMemcpyHostToDevice, st));
	if(!s->upEvent) CU(cudaEventCreateWithFlags(&s->upEvent, cudaEventDisableTiming));
	CU(cudaEventRecord(s->upEvent, st)); // kernels issued later that read this shadow wait for it
	CU(cudaEventRecord(ctx->evUpload, st));
	s->upSeq = ++ctx->uploadSeq;
	ctx->stats.h2dBytes += bytes;
	return SWCU_OK;
}

extern "C" int swcu_mem_download(swcu_ctx *ctx, void *host_ptr, size_t bytes)
{
	if(!ctx) return SWCU_E_INVALID;
	unsigned char *d = dev_ptr(ctx, host_ptr, bytes);
	if(!d) return fail(ctx, SWCU_E_INVALID, "swcu_mem_download: [%p,+%zu) is not inside a registered range", host_ptr, bytes);
	if(d == (unsigned char *)host_ptr) return fail(ctx, SWCU_E_INVALID, "swcu_mem_download: %p is caller-owned device memory", host_ptr);
	CU(cudaSetDevice(ctx->device));
	ctx->stats.d2hBytes += bytes;
	Shadow *s = find_shadow(ctx, host_ptr, bytes);
	if(!ctx->optCopyStreams)
	{
		int rc = see_upload(ctx, s, ctx->stream, ctx->mainSawUpload);
		if(rc) return rc;
		CU(cudaMemcpyAsync(host_ptr, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
		return SWCU_OK;
	}
	const cudaStream_t st = ctx->d2hStream;
	CU(cudaEventRecord(ctx->evMark, ctx->handover())); // everything issued so far (main stream; hand-over stream: what it waits for) has produced its pixels
	CU(cudaStreamWaitEvent(st, ctx->evMark, 0));
	int rc = see_upload(ctx, s, st, ctx->d2hSawUpload);
	if(rc) return rc;
	CU(cudaMemcpyAsync(host_ptr, d, bytes, cudaMemcpyDeviceToHost, st));
	if(!s->dlEvent) CU(cudaEventCreateWithFlags(&s->dlEvent, cudaEventDisableTiming));
	CU(cudaEventRecord(s->dlEvent, st));
	s->dlPendingMain = s->dlPendingUpload = true;
	CU(cudaEventRecord(ctx->evDownload, st));
	ctx->downloadSeq++;
	return SWCU_OK;
}

// Fences: the completion signal a submission hands back (the reference: the CountedEvent of sw::Renderer::draw, Renderer.cpp:184,
// :499-510, that vk::Fence waits on).  swcu_fence_signal marks "everything issued so far, copies included"; swcu_fence_wait blocks the host on it.
extern "C" int swcu_fence_signal(swcu_ctx *ctx, uint32_t slot)
{
	if(!ctx || slot >= SWCU_MAX_FENCES) return fail(ctx, SWCU_E_INVALID, "swcu_fence_signal: bad slot");
	CU(cudaSetDevice(ctx->device));
	const cudaStream_t st = ctx->d2hStream; // downloads are the tail of a frame: join the other streams here
	CU(cudaEventRecord(ctx->evMark, ctx->stream));
	CU(cudaStreamWaitEvent(st, ctx->evMark, 0));
	if(ctx->sideStream)
	{
		CU(cudaEventRecord(ctx->evMark, ctx->sideStream));
		CU(cudaStreamWaitEvent(st, ctx->evMark, 0));
	}
	int rc = see_uploads(ctx, st, ctx->d2hSawUpload);
	if(rc) return rc;
	CU(cudaEventRecord(ctx->fence[slot], st));
	ctx->fenceValid[slot] = true;
	return SWCU_OK;
}

extern "C" int swcu_fence_wait(swcu_ctx *ctx, uint32_t slot)
{
	if(!ctx || slot >= SWCU_MAX_FENCES) return fail(ctx, SWCU_E_INVALID, "swcu_fence_wait: bad slot");
	if(!ctx->fenceValid[slot]) return SWCU_OK;
	CU(cudaSetDevice(ctx->device));
	CU(cudaEventSynchronize(ctx->fence[slot]));
	return SWCU_OK;
}

extern "C" int swcu_mem_acquire(swcu_ctx *ctx, const void *host_ptr)
{
	if(!ctx || !host_ptr) return fail(ctx, SWCU_E_INVALID, "swcu_mem_acquire: null argument");
	if(!find_shadow(ctx, host_ptr, 1)) return fail(ctx, SWCU_E_INVALID, "swcu_mem_acquire: %p is not inside a registered range", host_ptr);
	CU(cudaSetDevice(ctx->device));
	return main_touches(ctx, host_ptr, true);
}

extern "C" int swcu_mem_release(swcu_ctx *ctx, const void *host_ptr)
{
	if(!ctx || !host_ptr) return fail(ctx, SWCU_E_INVALID, "swcu_mem_release: null argument");
	Shadow *s = find_shadow(ctx, host_ptr, 1);
	if(!s) return fail(ctx, SWCU_E_INVALID, "swcu_mem_release: %p is not inside a registered range", host_ptr);
	CU(cudaSetDevice(ctx->device));
	// The caller's work counts as the newest "upload" of the shadow.  Readers assume that having waited for upload number n they
	// have seen all earlier ones, so the stream first catches up with every upload issued so far.
	int rc = see_uploads(ctx, ctx->stream, ctx->mainSawUpload);
	if(rc) return rc;
	if(!s->upEvent) CU(cudaEventCreateWithFlags(&s->upEvent, cudaEventDisableTiming));
	CU(cudaEventRecord(s->upEvent, ctx->stream));
	s->upSeq = ++ctx->uploadSeq;
	ctx->mainSawUpload = s->upSeq;
	return SWCU_OK;
}

extern "C" int swcu_mem_acquire_on(swcu_ctx *ctx, const void *host_ptr, void *cuda_stream)
{
	if(!ctx || !host_ptr || !cuda_stream) return fail(ctx, SWCU_E_INVALID, "swcu_mem_acquire_on: null argument");
	Shadow *s = find_shadow(ctx, host_ptr, 1);
	if(!s) return fail(ctx, SWCU_E_INVALID, "swcu_mem_acquire_on: %p is not inside a registered range", host_ptr);
	CU(cudaSetDevice(ctx->device));
	// the caller's stream sees the last upload into the shadow (the upload itself has waited for the draws that read the shadow
	// before it), and a download of it still in flight
	if(s->upEvent && s->upSeq) CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, s->upEvent, 0));
	if(s->dlPendingMain) CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, s->dlEvent, 0));
	return SWCU_OK;
}

extern "C" int swcu_mem_release_on(swcu_ctx *ctx, const void *host_ptr, void *cuda_stream)
{
	if(!ctx || !host_ptr || !cuda_stream) return fail(ctx, SWCU_E_INVALID, "swcu_mem_release_on: null argument");
	Shadow *s = find_shadow(ctx, host_ptr, 1);
	if(!s) return fail(ctx, SWCU_E_INVALID, "swcu_mem_release_on: %p is not inside a registered range", host_ptr);
	CU(cudaSetDevice(ctx->device));
	// The caller's work counts as the newest "upload" of the shadow.  Readers assume that having waited for upload number n they have
	// seen all earlier ones, so the caller's stream first catches up with every upload issued so far.
	if(ctx->uploadSeq) CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, ctx->evUpload, 0));
	if(!s->upEvent) CU(cudaEventCreateWithFlags(&s->upEvent, cudaEventDisableTiming));
	CU(cudaEventRecord(s->upEvent, (cudaStream_t)cuda_stream));
	// "the last upload" now ends on the caller's stream; the copy stream queues behind it, so that waiting for any later upload
	// still means having seen this one
	CU(cudaEventRecord(ctx->evUpload, (cudaStream_t)cuda_stream));
	if(ctx->optCopyStreams) CU(cudaStreamWaitEvent(ctx->h2dStream, s->upEvent, 0));
	s->upSeq = ++ctx->uploadSeq;
	return SWCU_OK;
}

extern "C" void *swcu_mem_device_ptr(swcu_ctx *ctx, const void *host_ptr) { return ctx ? dev_ptr(ctx, host_ptr) : nullptr; }

// ------------------------------------------------------------------------------------------------------------------
// plumbing
// ------------------------------------------------------------------------------------------------------------------
extern "C" int swcu_sync(swcu_ctx *ctx)
{
	if(!ctx) return SWCU_E_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->h2dStream));
	for(cudaStream_t ss : ctx->setupStream) CU(cudaStreamSynchronize(ss));
	CU(cudaStreamSynchronize(ctx->stream));
	if(ctx->sideStream) CU(cudaStreamSynchronize(ctx->sideStream));
	CU(cudaStreamSynchronize(ctx->d2hStream));
	// the device is idle: the counters of every finished draw are on the host
	for(auto &S : ctx->set)
	{
		if(!S.countersPending) continue;
		S.countersPending = false;
		ctx->stats.pairs += S.hostCounters->pairTotal;
		if(S.hostCounters->overflow)
			return fail(ctx, SWCU_E_NOMEM, "a draw had big triangles whose bounding boxes cover more than %zu regions in total (%llu): they were NOT rendered; raise swcu_set_option(\"big_pair_budget\")",
			            ctx->optBigPairBudget, (unsigned long long)S.hostCounters->bigReserved);
	}
	return SWCU_OK;
}

extern "C" int swcu_set_stream(swcu_ctx *ctx, void *cuda_stream)
{
	if(!ctx) return SWCU_E_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->ownStream;
	return SWCU_OK;
}

extern "C" int swcu_timer_begin(swcu_ctx *ctx)
{
	if(!ctx) return SWCU_E_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaEventRecord(ctx->t0, ctx->stream));
	return SWCU_OK;
}

extern "C" int swcu_timer_end(swcu_ctx *ctx, float *elapsed_ms)
{
	if(!ctx || !elapsed_ms) return SWCU_E_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaEventRecord(ctx->t1, ctx->stream));
	CU(cudaEventSynchronize(ctx->t1));
	CU(cudaEventElapsedTime(elapsed_ms, ctx->t0, ctx->t1));
	return SWCU_OK;
}

extern "C" int swcu_get_stats(swcu_ctx *ctx, swcu_stats *out)
{
	if(!ctx || !out) return SWCU_E_INVALID;
	*out = ctx->stats;
	return SWCU_OK;
}

extern "C" int swcu_reset_stats(swcu_ctx *ctx)
{
	if(!ctx) return SWCU_E_INVALID;
	ctx->stats = swcu_stats{};
	return SWCU_OK;
}

extern "C" int swcu_set_profiling(swcu_ctx *ctx, int enable)
{
	if(!ctx) return SWCU_E_INVALID;
	ctx->profiling = enable;
	return SWCU_OK;
}

extern "C" int swcu_last_draw_kernels(swcu_ctx *ctx, const char **names, float *ms, int n)
{
	if(!ctx) return SWCU_E_INVALID;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if(ctx->sideStream) cudaStreamSynchronize(ctx->sideStream);
	int k = 0;
	for(const KernelTime &t : ctx->lastKernels)
	{
		if(k >= n) break;
		float v = 0;
		cudaEventElapsedTime(&v, t.e0, t.e1);
		names[k] = t.name;
		ms[k] = v;
		k++;
	}
	return k;
}

// Diagnostics (no counterpart in the reference): with swcu_set_profiling(ctx, 2) the draws stay pipelined over their streams and every
// kernel of the library is bracketed by events on the stream it really runs on; this returns the begin / end of each one in
// milliseconds since the first of them — the timeline of the frames issued since the mode was switched on — and starts a new one.
extern "C" int swcu_timeline(swcu_ctx *ctx, const char **names, float *begin_ms, float *end_ms, int n)
{
	if(!ctx) return SWCU_E_INVALID;
	cudaSetDevice(ctx->device);
	int rc = swcu_sync(ctx);
	if(rc) return rc;
	int k = 0;
	for(const KernelTime &t : ctx->lastKernels)
	{
		if(k >= n) break;
		float a = 0, b = 0;
		cudaEventElapsedTime(&a, ctx->lastKernels[0].e0, t.e0);
		cudaEventElapsedTime(&b, ctx->lastKernels[0].e0, t.e1);
		names[k] = t.name;
		begin_ms[k] = a;
		end_ms[k] = b;
		k++;
	}
	ctx->lastKernels.clear();
	ctx->eventsUsed = 0;
	return k;
}

extern "C" int swcu_set_option(swcu_ctx *ctx, const char *name, int value)
{
	if(!ctx || !name) return SWCU_E_INVALID;
	if(!strcmp(name, "force_binned")) ctx->optForceBinned = value;
	else if(!strcmp(name, "direct_max")) ctx->optDirectMax = value;
	else if(!strcmp(name, "pin_host")) ctx->optPinHost = value;
	else if(!strcmp(name, "tma")) ctx->optTma = value;
	else if(!strcmp(name, "fast_state")) ctx->optFastState = value;
	else if(!strcmp(name, "write_only")) ctx->optWriteOnly = value;
	else if(!strcmp(name, "setup_wide")) ctx->optSetupWide = value;
	else if(!strcmp(name, "pipeline")) ctx->optPipeline = value;
	else if(!strcmp(name, "big_pair_budget")) ctx->optBigPairBudget = (size_t)std::max(value, 0);
	else if(!strcmp(name, "copy_streams"))
	{
		// 0: copies on the main stream (needed when something the library cannot see - e.g. an NCCL collective on the caller's
		// stream - writes a shadow that is also downloaded)
		int rc = swcu_sync(ctx);
		if(rc) return rc;
		ctx->optCopyStreams = value;
	}
	else return fail(ctx, SWCU_E_INVALID, "unknown option '%s'", name);
	return SWCU_OK;
}

// kernel launch bookkeeping: counts OUR kernels and, when profiling, brackets each with events
struct LaunchScope
{
	swcu_ctx *ctx;
	KernelTime kt{};
	bool timed = false;
	cudaStream_t st;
	LaunchScope(swcu_ctx *c, const char *name, cudaStream_t stream = nullptr) : ctx(c), st(stream ? stream : c->stream)
	{
		ctx->stats.kernelLaunches++;
		if(ctx->profiling)
		{
			auto get = [&]() {
				if(ctx->eventsUsed == ctx->eventPool.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->eventPool.push_back(e); }
				return ctx->eventPool[ctx->eventsUsed++];
			};
			kt.name = name; kt.e0 = get(); kt.e1 = get();
			cudaEventRecord(kt.e0, st);
			timed = true;
		}
	}
	~LaunchScope()
	{
		if(timed) { cudaEventRecord(kt.e1, st); ctx->lastKernels.push_back(kt); }
	}
};

// ------------------------------------------------------------------------------------------------------------------
// state gathering (Renderer::draw) and the kernel sequence (DrawCall::run)
// ------------------------------------------------------------------------------------------------------------------
static uint64_t fnv1a(const uint32_t *w, uint32_t n)
{
	uint64_t h = 1469598103934665603ull;
	for(uint32_t i = 0; i < n; i++) { h ^= w[i]; h *= 1099511628211ull; }
	return h ^ n;
}

static int get_shader(swcu_ctx *ctx, const uint32_t *code, uint32_t words, uint32_t stage, swcu_shader_info *out)
{
	if(!code || !words) return fail(ctx, SWCU_E_INVALID, "missing %s shader", stage ? "fragment" : "vertex");
	const uint64_t key = fnv1a(code, words);
	auto range = ctx->shaderCache.equal_range(key);
	auto it = range.first;
	for(; it != range.second; ++it)
		if(it->second.words.size() == words && !memcmp(it->second.words.data(), code, (size_t)words * 4)) break;
	if(it == range.second)
	{
		swcu_ctx::CachedShader entry;
		char err[256];
		int rc = swcu_shader_translate(code, words, &entry.info, err, sizeof(err));
		if(rc != SWCU_OK) return fail(ctx, rc, "%s shader rejected by the SPIR-V subset translator: %s", stage ? "fragment" : "vertex", err);
		entry.words.assign(code, code + words);
		it = ctx->shaderCache.emplace(key, std::move(entry));
	}
	*out = it->second.info;
	if(out->stage != stage) return fail(ctx, SWCU_E_INVALID, "shader stage mismatch (expected %u, module is %u)", stage, out->stage);
	return SWCU_OK;
}

// Context.cpp:1165-1270 (operation folding) and :1272-1300 (factor folding); two SUBTRACT cases fold only for UNORM targets
// (a negative result is clamped to zero there, kept for floating-point targets)
static int fold_blend_op(int op, int sf, int df, bool unorm)
{
	switch(op)
	{
	case BOP_ADD:
		if(sf == BF_ZERO) { if(df == BF_ZERO) return BOP_ZERO_EXT; if(df == BF_ONE) return BOP_DST_EXT; }
		else if(sf == BF_ONE) { if(df == BF_ZERO) return BOP_SRC_EXT; }
		break;
	case BOP_SUBTRACT:
		if(sf == BF_ZERO) { if(df == BF_ZERO || unorm) return BOP_ZERO_EXT; }
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
static int kop(int op)
{
	switch(op)
	{
	case BOP_ADD: return KOP_ADD;
	case BOP_SUBTRACT: return KOP_SUB;
	case BOP_REVERSE_SUBTRACT: return KOP_RSUB;
	case BOP_MIN: return KOP_MIN;
	case BOP_MAX: return KOP_MAX;
	case BOP_SRC_EXT: return KOP_SRC;
	case BOP_DST_EXT: return KOP_DST;
	case BOP_ZERO_EXT: return KOP_ZERO;
	}
	return -1;
}

static int clampi_h(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); } // NaN -> 1 like min(max()) chain is irrelevant here

static int build_const(swcu_ctx *ctx, const swcu_draw_desc *desc, DrawConst &d)
{
	memset(&d, 0, sizeof(d));
	if(desc->structSize != sizeof(swcu_draw_desc)) return fail(ctx, SWCU_E_INVALID, "swcu_draw_desc.structSize %u != %zu", desc->structSize, sizeof(swcu_draw_desc));
	if(desc->topology > TOPO_TRIANGLE_FAN)
		return fail(ctx, SWCU_E_UNSUPPORTED, "topology %u outside the subset (point list, line list/strip, triangle list/strip/fan)", desc->topology);
	if(desc->sampleCount != 1 && desc->sampleCount != 4) return fail(ctx, SWCU_E_UNSUPPORTED, "sample count %u unsupported (1 or 4)", desc->sampleCount);
	if(desc->indexType != 0 && desc->indexType != 2 && desc->indexType != 4) return fail(ctx, SWCU_E_UNSUPPORTED, "index type %u unsupported", desc->indexType);
	if(desc->provokingVertexMode > 1) return fail(ctx, SWCU_E_INVALID, "bad provoking vertex mode");
	const bool srgbTarget = desc->color.buffer && (desc->color.format == VKF_R8G8B8A8_SRGB || desc->color.format == VKF_B8G8R8A8_SRGB);
	const bool floatTarget = desc->color.buffer && (desc->color.format == VKF_R32G32B32A32_SFLOAT || desc->color.format == VKF_R16G16B16A16_SFLOAT);
	if(desc->color.buffer && !srgbTarget && !floatTarget && desc->color.format != VKF_R8G8B8A8_UNORM && desc->color.format != VKF_B8G8R8A8_UNORM)
		return fail(ctx, SWCU_E_UNSUPPORTED, "colour format %u unsupported (R8G8B8A8 / B8G8R8A8 UNORM or SRGB, R16G16B16A16_SFLOAT, R32G32B32A32_SFLOAT)", desc->color.format);
	d.colorEpp = desc->color.format == VKF_R32G32B32A32_SFLOAT ? 4u : (desc->color.format == VKF_R16G16B16A16_SFLOAT ? 2u : 1u);
	if(desc->depth.buffer && desc->depth.format != VKF_D32_SFLOAT && desc->depth.format != VKF_D16_UNORM)
		return fail(ctx, SWCU_E_UNSUPPORTED, "depth format %u unsupported (D32_SFLOAT, D16_UNORM)", desc->depth.format);
	if(desc->depth.buffer && desc->depth.format == VKF_D16_UNORM && desc->stencil.buffer)
		return fail(ctx, SWCU_E_UNSUPPORTED, "D16_UNORM with a stencil aspect is outside the subset");
	if(desc->stencil.buffer && desc->stencil.format != VKF_S8_UINT) return fail(ctx, SWCU_E_UNSUPPORTED, "stencil format %u unsupported (S8_UINT)", desc->stencil.format);
	if(!desc->color.buffer && !desc->depth.buffer && !desc->stencil.buffer) return fail(ctx, SWCU_E_INVALID, "draw without attachments");

	swcu_shader_info vs, fs;
	int rc = get_shader(ctx, desc->vertexShader, desc->vertexShaderWords, 0, &vs);
	if(rc) return rc;
	rc = get_shader(ctx, desc->fragmentShader, desc->fragmentShaderWords, 4, &fs);
	if(rc) return rc;

	{ // Renderer.cpp:300-331
		const float W = 0.5f * desc->viewportWidth, H = 0.5f * desc->viewportHeight;
		const float X0 = desc->viewportX + W, Y0 = desc->viewportY + H;
		d.WxF = W * 256.0f; d.HxF = H * 256.0f;
		volatile float x0f = X0 * 256.0f, y0f = Y0 * 256.0f; // no contraction into an FMA
		d.X0xF = x0f - 128.0f; d.Y0xF = y0f - 128.0f;
		d.depthRange = desc->viewportMaxDepth - desc->viewportMinDepth;
		d.depthNear = desc->viewportMinDepth;
	}
	{ // Renderer.cpp:333-345
		const int x0 = desc->renderArea.x, y0 = desc->renderArea.y;
		const int x1 = x0 + (int)desc->renderArea.width, y1 = y0 + (int)desc->renderArea.height;
		d.scX0 = clampi_h(desc->scissor.x, x0, x1);
		d.scX1 = clampi_h(desc->scissor.x + (int)desc->scissor.width, x0, x1);
		d.scY0 = clampi_h(desc->scissor.y, y0, y1);
		d.scY1 = clampi_h(desc->scissor.y + (int)desc->scissor.height, y0, y1);
		d.suY0 = d.scY0; d.suY1 = d.scY1; // (a group draw widens these to the scissor rows of the whole frame, see swcu_draw)
	}
	d.ms = (int)desc->sampleCount;
	d.sampleMask = desc->sampleMask & (d.ms > 1 ? 0xFu : 0x1u); // Context.cpp:527: bit 0 counts at one sample per pixel too

	// ---- input assembly ----
	d.indexType = desc->indexType;
	d.topology = desc->topology;
	d.provokingFirst = desc->provokingVertexMode == 0;
	d.primCount = desc->primitiveCount;
	d.baseVertex = desc->baseVertex;
	if(desc->indexType)
	{
		// every index the draw fetches must lie inside the registered range: the kernels do no bounds checks on the index stream
		// (the reference reads whatever follows an index buffer that is too short; a device library must not)
		const size_t pc = desc->primitiveCount;
		const size_t nIdx = desc->topology == TOPO_TRIANGLE_LIST ? pc * 3 : desc->topology == TOPO_LINE_LIST ? pc * 2 : desc->topology == TOPO_LINE_STRIP ? pc + 1 :
		                    desc->topology == TOPO_POINT_LIST ? pc : pc + 2;
		d.indexBuffer = desc->primitiveCount ? dev_ptr(ctx, desc->indexBuffer, nIdx * desc->indexType) : dev_ptr(ctx, desc->indexBuffer);
		if(!d.indexBuffer) return fail(ctx, SWCU_E_INVALID, "index buffer [%p, +%zu) is not inside a registered range", desc->indexBuffer, nIdx * desc->indexType);
		if(find_shadow(ctx, desc->indexBuffer, 1)->external) d.inputsExternal = 1;
	}
	// vertex-stage scalars: component c of stream l, or a constant (VertexRoutine::readStream, VertexRoutine.cpp:173-245)
	int vsrcErr = SWCU_OK;
	// a 32-bit word of the draw's push-constant block (DrawData::pushConstants, Renderer.cpp:484-486); words the application never
	// pushed read 0
	if(desc->pushConstantBytes > 4 * SWCU_MAX_PUSH_WORDS || (desc->pushConstantBytes && !desc->pushConstants))
		return fail(ctx, SWCU_E_INVALID, "bad push-constant block (%u bytes)", desc->pushConstantBytes);
	auto push_word = [&](uint32_t w) -> uint32_t {
		uint32_t v = 0;
		if(4 * w + 4 <= desc->pushConstantBytes) memcpy(&v, (const unsigned char *)desc->pushConstants + 4 * w, 4);
		return v;
	};
	// a 32-bit word of uniform block `slot` of the vertex shader: read from the host memory behind the descriptor the draw binds at
	// the block's (set, binding) — BufferDescriptor::ptr / sizeInBytes; words past the end read 0 (robust buffer access)
	if(desc->uniformBufferCount > SWCU_MAX_UNIFORM_BUFFERS) return fail(ctx, SWCU_E_INVALID, "more than %d uniform buffers", SWCU_MAX_UNIFORM_BUFFERS);
	int uboErr = SWCU_OK;
	auto uniform_word = [&](uint32_t value) -> uint32_t {
		const uint32_t slot = value >> 16, w = value & 0xFFFFu;
		if(slot >= vs.uniformCount) { uboErr = fail(ctx, SWCU_E_INVALID, "vertex program reads uniform block %u of %u", slot, vs.uniformCount); return 0; }
		for(uint32_t i = 0; i < desc->uniformBufferCount; i++)
		{
			const swcu_uniform_buffer &u = desc->uniformBuffer[i];
			if(u.set != vs.uniformSet[slot] || u.binding != vs.uniformBinding[slot]) continue;
			uint32_t v = 0;
			if(u.data && (size_t)4 * w + 4 <= u.bytes) memcpy(&v, (const unsigned char *)u.data + 4 * w, 4);
			return v;
		}
		uboErr = fail(ctx, SWCU_E_INVALID, "no uniform buffer bound at set %u, binding %u", vs.uniformSet[slot], vs.uniformBinding[slot]);
		return 0;
	};
	auto vsrc = [&](const swcu_shader_operand &o) -> KVSrc {
		KVSrc k;
		k.ptr = nullptr; k.stride = 0; k.limit = 0xFFFFFFFFu; k.constant = 0.0f; k.pad = 0;
		if(o.kind == SWCU_SRC_CONST) { memcpy(&k.constant, &o.value, 4); return k; }
		if(o.kind == SWCU_SRC_PUSH) { const uint32_t v = push_word(o.value); memcpy(&k.constant, &v, 4); return k; }
		if(o.kind == SWCU_SRC_UNIFORM) { const uint32_t v = uniform_word(o.value); memcpy(&k.constant, &v, 4); return k; }
		if(o.kind == SWCU_SRC_TEMP) return k; // (the caller routes the step's result: posTemp / slotTemp)
		const uint32_t l = o.value >> 2, c = o.value & 3;
		const swcu_vertex_input &in = desc->input[l];
		uint32_t ncomp;
		switch(in.format)
		{
		case VKF_R32_SFLOAT: ncomp = 1; break;
		case VKF_R32G32_SFLOAT: ncomp = 2; break;
		case VKF_R32G32B32_SFLOAT: ncomp = 3; break;
		case VKF_R32G32B32A32_SFLOAT: ncomp = 4; break;
		case VKF_UNDEFINED: ncomp = 0; break;
		default: vsrcErr = fail(ctx, SWCU_E_UNSUPPORTED, "vertex input %u: format %u unsupported (R32..R32G32B32A32_SFLOAT)", l, in.format); return k;
		}
		if(c >= ncomp) { k.constant = c == 3 ? 1.0f : 0.0f; return k; } // missing components read (0,0,0,1)
		if(in.robustnessSize && in.robustnessSize < ncomp * 4) return k;    // every fetch is out of bounds: 0
		unsigned char *base = dev_ptr(ctx, in.buffer);
		if(!base) { vsrcErr = fail(ctx, SWCU_E_INVALID, "vertex input %u: buffer %p is not inside a registered range", l, in.buffer); return k; }
		if(find_shadow(ctx, in.buffer, 1)->external) d.inputsExternal = 1;
		k.ptr = base + 4 * c;
		k.stride = in.vertexStride;
		k.limit = in.robustnessSize ? in.robustnessSize - ncomp * 4 : 0xFFFFFFFFu;
		return k;
	};

	// ---- vertex-stage arithmetic: the translator's steps with their operands resolved for this draw ----
	d.vsProgLen = vs.programLength;
	if(vs.programLength > SWCU_MAX_PROGRAM) return fail(ctx, SWCU_E_INVALID, "vertex program too long");
	{
		uint32_t inKey[SWCU_VS_INPUTS], nIn = 0;
		for(uint32_t i = 0; i < vs.programLength; i++)
		{
			const swcu_shader_operand *src[3] = { &vs.program[i].a, &vs.program[i].b, &vs.program[i].c };
			KVsOperand *dst[3] = { &d.vsProg[i].a, &d.vsProg[i].b, &d.vsProg[i].c };
			d.vsProg[i].op = vs.program[i].op;
			for(int k = 0; k < 3; k++)
			{
				const swcu_shader_operand &o = *src[k];
				if(o.kind == SWCU_SRC_CONST) *dst[k] = { VK_CONST, o.value };
				else if(o.kind == SWCU_SRC_PUSH) *dst[k] = { VK_CONST, push_word(o.value) };
				else if(o.kind == SWCU_SRC_UNIFORM) *dst[k] = { VK_CONST, uniform_word(o.value) };
				else if(o.kind == SWCU_SRC_TEMP)
				{
					if(o.value >= i) return fail(ctx, SWCU_E_INVALID, "vertex program step %u reads a later step", i);
					*dst[k] = { VK_TEMP, o.value };
				}
				else
				{
					uint32_t j = 0;
					while(j < nIn && inKey[j] != o.value) j++;
					if(j == nIn)
					{
						if(nIn == SWCU_VS_INPUTS) return fail(ctx, SWCU_E_UNSUPPORTED, "vertex program reads more than %d attribute components", SWCU_VS_INPUTS);
						inKey[nIn] = o.value;
						d.vsIn[nIn++] = vsrc(o);
					}
					*dst[k] = { VK_INPUT, j };
				}
			}
		}
	}
	auto temp_of = [](const swcu_shader_operand &o) -> int32_t { return o.kind == SWCU_SRC_TEMP ? (int32_t)o.value : -1; };

	// ---- shader routing ----
	for(int k = 0; k < 4; k++) { d.vsPos[k] = vsrc(vs.position[k]); d.posTemp[k] = temp_of(vs.position[k]); }
	for(int k = 0; k < SWCU_MAXSLOTS; k++) d.slotTemp[k] = -1;
	// ---- lines and points ----
	d.primKind = desc->topology == TOPO_POINT_LIST ? PRIM_POINT : ((desc->topology == TOPO_LINE_LIST || desc->topology == TOPO_LINE_STRIP) ? PRIM_LINE : PRIM_TRIANGLE);
	d.lineWidth = desc->lineWidth == 0.0f ? 1.0f : desc->lineWidth;
	d.halfPixelX = 0.5f / (0.5f * desc->viewportWidth); d.halfPixelY = 0.5f / (0.5f * desc->viewportHeight);
	{
		const float one = 1.0f;
		swcu_shader_operand opOne = { SWCU_SRC_CONST, 0 };
		memcpy(&opOne.value, &one, 4);
		d.pointSizeSrc = vsrc(vs.writesPointSize ? vs.pointSize : opOne);
		d.pointSizeTemp = vs.writesPointSize ? temp_of(vs.pointSize) : -1;
	}
	const swcu_shader_operand opZero = { SWCU_SRC_CONST, 0 };
	// a fragment-stage operand that reads an interpolated input becomes a plane slot fed by the vertex stage
	auto slot_from = [&](const swcu_shader_operand &o, int slot) {
		d.slotTemp[slot] = -1;
		if(o.kind == SWCU_SRC_CONST) { d.slotSrc[slot] = vsrc(o); d.slotMode[slot] = IM_FLAT; return; }
		const uint32_t c = o.value; // location*4 + component of the fragment input
		const swcu_shader_operand &vo = ((vs.outputMask >> c) & 1) ? vs.output[c] : opZero; // never written by the vertex stage: 0
		d.slotSrc[slot] = vsrc(vo);
		d.slotTemp[slot] = temp_of(vo);
		d.slotMode[slot] = ((fs.flatMask >> c) & 1) ? IM_FLAT : (((fs.noPerspectiveMask >> c) & 1) ? IM_NOPERSP : IM_PERSP);
	};
	bool anySlot = false;
	for(int ch = 0; ch < 4; ch++)
	{
		d.slotSrc[ch] = vsrc(opZero); d.slotMode[ch] = IM_FLAT;
		const swcu_shader_operand &o = fs.output[ch];
		if(!((fs.outputMask >> ch) & 1)) { d.chanKind[ch] = CK_CONST; d.chanValue[ch] = 0; }
		else if(o.kind == SWCU_SRC_CONST) { d.chanKind[ch] = CK_CONST; d.chanValue[ch] = o.value; }
		else if(o.kind == SWCU_SRC_TEXEL) { d.chanKind[ch] = CK_TEXEL; d.chanValue[ch] = o.value & 3; }
		else { d.chanKind[ch] = CK_SLOT; d.chanValue[ch] = (uint32_t)ch; slot_from(o, ch); anySlot = true; }
	}
	d.shaderClass = fs.usesTexture ? (anySlot ? SH_GENERIC : SH_TEX) : (anySlot ? SH_VARY : SH_CONST);
	d.nslots = d.shaderClass == SH_CONST ? 0 : d.shaderClass == SH_VARY ? 4 : d.shaderClass == SH_TEX ? 2 : 6;
	d.uvSlot = d.shaderClass == SH_TEX ? 0 : 4;
	d.usesTexture = fs.usesTexture;
	if(fs.usesTexture)
	{
		if(d.shaderClass == SH_TEX) // slots [0,4) do not exist in this class: move the (unused) colour slots out of the way
			for(int k = 0; k < 2; k++) { d.slotSrc[k] = vsrc(opZero); d.slotMode[k] = IM_FLAT; }
		slot_from(fs.texCoord[0], d.uvSlot);
		slot_from(fs.texCoord[1], d.uvSlot + 1);
		const swcu_sampled_image *t = nullptr;
		for(uint32_t s = 0; s < desc->sampledImageCount && s < SWCU_MAX_SAMPLED_IMAGES; s++)
			if(desc->sampledImage[s].set == fs.textureSet && desc->sampledImage[s].binding == fs.textureBinding) t = &desc->sampledImage[s];
		if(!t) return fail(ctx, SWCU_E_INVALID, "no sampled image bound at set %u binding %u", fs.textureSet, fs.textureBinding);
		if(t->format != VKF_R8G8B8A8_UNORM && t->format != VKF_R8G8B8A8_SRGB) return fail(ctx, SWCU_E_UNSUPPORTED, "sampled image format %u unsupported (R8G8B8A8_UNORM, R8G8B8A8_SRGB)", t->format);
		d.texSrgb = t->format == VKF_R8G8B8A8_SRGB;
		if(t->anisotropyEnable || t->compareEnable || t->unnormalizedCoordinates) return fail(ctx, SWCU_E_UNSUPPORTED, "sampler state outside the subset (anisotropy / compare / unnormalized)");
		if(t->addressModeU > ADDR_CLAMP_TO_EDGE || t->addressModeV > ADDR_CLAMP_TO_EDGE) return fail(ctx, SWCU_E_UNSUPPORTED, "sampler address mode outside the subset (REPEAT, MIRRORED_REPEAT, CLAMP_TO_EDGE)");
		if(t->magFilter > FILTER_LINEAR || t->minFilter > FILTER_LINEAR || t->mipmapMode > MIPMAP_MODE_LINEAR) return fail(ctx, SWCU_E_UNSUPPORTED, "sampler filter outside the subset");
		if(t->levelCount < 1 || t->levelCount > SWCU_MIPMAP_LEVELS) return fail(ctx, SWCU_E_INVALID, "bad mip level count %u", t->levelCount);
		d.texLevels = t->levelCount;
		for(int l = 0; l < SWCU_MIPMAP_LEVELS; l++)
		{
			const swcu_mip_level &m = t->level[std::min<int>(l, (int)t->levelCount - 1)];
			if(m.width == 0 || m.height == 0 || m.width > 32768 || m.height > 32768) return fail(ctx, SWCU_E_INVALID, "bad mip level %d extent", l);
			d.mip[l].buffer = dev_ptr(ctx, m.buffer, (size_t)m.pitchP * (m.height - 1) * 4 + (size_t)m.width * 4);
			if(!d.mip[l].buffer) return fail(ctx, SWCU_E_INVALID, "mip level %d buffer %p is not inside a registered range", l, m.buffer);
			d.mip[l].width = m.width; d.mip[l].height = m.height; d.mip[l].pitchP = m.pitchP;
			d.mip[l].half = ((0x8000u / m.width) & 0xFFFFu) | (((0x8000u / m.height) & 0xFFFFu) << 16);
		}
		d.magFilter = t->magFilter; d.minFilter = t->minFilter; d.mipmapMode = t->mipmapMode;
		d.addressU = t->addressModeU; d.addressV = t->addressModeV;
		d.mipLodBias = t->mipLodBias; d.minLod = t->minLod; d.maxLod = t->maxLod;
		d.texFast = !d.texSrgb && t->magFilter == FILTER_LINEAR && t->minFilter == FILTER_LINEAR && t->mipmapMode == MIPMAP_MODE_LINEAR &&
		            t->addressModeU == ADDR_REPEAT && t->addressModeV == ADDR_REPEAT;
	}

	// ---- setup state ----
	d.cullMode = desc->cullMode; d.frontFace = desc->frontFace; d.depthClipEnable = desc->depthClipEnable;
	d.minDepthClamp = 0.0f; d.maxDepthClamp = 1.0f; // PixelProcessor.cpp:121-136 (no VK_EXT_depth_range_unrestricted: always clamped)
	if(desc->depthClampEnable)
	{
		d.minDepthClamp = std::min(desc->viewportMinDepth, desc->viewportMaxDepth);
		d.maxDepthClamp = std::max(desc->viewportMinDepth, desc->viewportMaxDepth);
	}
	d.depthBiasConstant = desc->depthBiasConstant; d.depthBiasSlope = desc->depthBiasSlope; d.depthBiasClamp = desc->depthBiasClamp;
	d.depthBiasEnable = desc->depthBiasConstant != 0.0f || desc->depthBiasSlope != 0.0f;
	if(d.primKind != PRIM_TRIANGLE) // SetupProcessor.cpp:75-77: depth bias applies to triangles only
	{
		d.depthBiasConstant = d.depthBiasSlope = d.depthBiasClamp = 0.0f;
		d.depthBiasEnable = 0;
	}

	// ---- pixel state ----
	d.depthTestActive = desc->depthTestEnable && desc->depth.buffer;
	d.depthWriteEnable = d.depthTestActive && desc->depthWriteEnable;
	if(desc->depthCompareOp > CMP_ALWAYS) return fail(ctx, SWCU_E_INVALID, "bad depth compare op");
	d.depthCompareOp = desc->depthCompareOp;
	d.alphaToCoverage = desc->alphaToCoverageEnable != 0;
	d.depthBounds = 0;
	if(desc->depthBoundsTestEnable && desc->depth.buffer) // FragmentState::depthBoundsTestActive, Context.cpp:946-949
	{
		d.depthBounds = d.depthTestActive ? 1 : 2;
		d.minDepthBounds = desc->minDepthBounds; d.maxDepthBounds = desc->maxDepthBounds;
		if(!d.depthTestActive) { d.depthTestActive = 1; d.depthCompareOp = CMP_ALWAYS; d.depthWriteEnable = 0; } // stage the depth plane, test nothing else
	}
	d.stencilActive = desc->stencilTestEnable && desc->stencil.buffer;
	auto face = [](const swcu_stencil_face &s) { KStencilFace k; k.failOp = s.failOp; k.passOp = s.passOp; k.depthFailOp = s.depthFailOp; k.compareOp = s.compareOp; k.compareMask = s.compareMask & 0xFF; k.writeMask = s.writeMask & 0xFF; k.reference = s.reference & 0xFF; k.pad = 0; return k; };
	d.front = face(desc->front); d.back = face(desc->back);
	if(d.stencilActive)
	{
		const swcu_stencil_face *f2[2] = { &desc->front, &desc->back };
		for(const swcu_stencil_face *s : f2)
			if(s->failOp > SOP_DEC_WRAP || s->passOp > SOP_DEC_WRAP || s->depthFailOp > SOP_DEC_WRAP || s->compareOp > CMP_ALWAYS) return fail(ctx, SWCU_E_INVALID, "bad stencil state");
		const bool allKeep = desc->front.passOp == SOP_KEEP && desc->front.depthFailOp == SOP_KEEP && desc->front.failOp == SOP_KEEP &&
		                     desc->back.passOp == SOP_KEEP && desc->back.depthFailOp == SOP_KEEP && desc->back.failOp == SOP_KEEP;
		const bool writeEnabled = (desc->front.writeMask & 0xFF) != 0 || (desc->back.writeMask & 0xFF) != 0;
		d.stencilWrite = !allKeep && writeEnabled;
	}
	{ // Context.cpp:1090-1147
		const int cop = fold_blend_op((int)desc->colorBlendOp, (int)desc->srcColorBlendFactor, (int)desc->dstColorBlendFactor, !floatTarget);
		const int aop = fold_blend_op((int)desc->alphaBlendOp, (int)desc->srcAlphaBlendFactor, (int)desc->dstAlphaBlendFactor, !floatTarget);
		d.colorWriteMask = desc->color.buffer ? (desc->colorWriteMask & 0xF) : 0;
		if(cop == BOP_DST_EXT && aop == BOP_DST_EXT) d.colorWriteMask = 0; // colorWriteActive (Context.cpp:1304-1308) tests the stored factors whether or not blending is enabled
		d.blendEnable = desc->blendEnable && d.colorWriteMask && (cop != BOP_SRC_EXT || aop != BOP_SRC_EXT);
		if(d.blendEnable)
		{
			if(kop(cop) < 0 || kop(aop) < 0) return fail(ctx, SWCU_E_UNSUPPORTED, "blend op %d/%d outside the subset (ADD, SUBTRACT, REVERSE_SUBTRACT, MIN, MAX)", (int)desc->colorBlendOp, (int)desc->alphaBlendOp);
			if(desc->srcColorBlendFactor > BF_SRC_ALPHA_SATURATE || desc->dstColorBlendFactor > BF_SRC_ALPHA_SATURATE ||
			   desc->srcAlphaBlendFactor > BF_SRC_ALPHA_SATURATE || desc->dstAlphaBlendFactor > BF_SRC_ALPHA_SATURATE)
				return fail(ctx, SWCU_E_UNSUPPORTED, "blend factor outside the subset (dual-source factors)");
			d.op = (uint32_t)kop(cop); d.opA = (uint32_t)kop(aop);
			d.srcF = (uint32_t)fold_blend_factor((int)desc->colorBlendOp, (int)desc->srcColorBlendFactor);
			d.dstF = (uint32_t)fold_blend_factor((int)desc->colorBlendOp, (int)desc->dstColorBlendFactor);
			d.srcFA = (uint32_t)fold_blend_factor((int)desc->alphaBlendOp, (int)desc->srcAlphaBlendFactor);
			d.dstFA = (uint32_t)fold_blend_factor((int)desc->alphaBlendOp, (int)desc->dstAlphaBlendFactor);
		}
		// blendConstantU (clamped) for UNORM targets, blendConstantF (as given) for floating-point ones (PixelRoutine.cpp:1203-1223)
		for(int k = 0; k < 4; k++) d.blendConstant[k] = floatTarget ? desc->blendConstants[k] : clamp01(desc->blendConstants[k]);
		d.bgr = desc->color.format == VKF_B8G8R8A8_UNORM || desc->color.format == VKF_B8G8R8A8_SRGB;
		d.srgb = srgbTarget ? 1u : 0u;
		d.blendClass = !d.blendEnable ? BL_OFF : ((d.srcF == BF_SRC_ALPHA && d.dstF == BF_ONE_MINUS_SRC_ALPHA && d.op == KOP_ADD && d.opA == KOP_SRC && d.colorWriteMask == 0xF) ? BL_SRC_ALPHA : BL_GENERIC);
	}

	// ---- attachments ----
	const swcu_attachment *any = desc->color.buffer ? &desc->color : (desc->depth.buffer ? &desc->depth : &desc->stencil);
	d.fbWidth = (int)any->width; d.fbHeight = (int)any->height;
	if(d.fbWidth <= 0 || d.fbHeight <= 0 || d.fbWidth > 8192 || d.fbHeight > 8192) return fail(ctx, SWCU_E_UNSUPPORTED, "framebuffer extent %dx%d outside 1..8192 (OUTLINE_RESOLUTION, Config.hpp:20)", d.fbWidth, d.fbHeight);
	auto att = [&](const swcu_attachment &a, int bpp, unsigned char *&buf, int &pitch, int &slice, const char *what) -> int {
		if(!a.buffer) return SWCU_OK;
		if((int)a.width != d.fbWidth || (int)a.height != d.fbHeight) return fail(ctx, SWCU_E_INVALID, "%s attachment extent differs from the framebuffer", what);
		const size_t need = (size_t)(d.ms - 1) * a.sliceB + (size_t)(a.height - 1) * a.pitchB + (size_t)a.width * bpp;
		buf = dev_ptr(ctx, a.buffer, need);
		if(!buf) return fail(ctx, SWCU_E_INVALID, "%s attachment %p (+%zu) is not inside a registered range", what, a.buffer, need);
		pitch = a.pitchB; slice = a.sliceB;
		return SWCU_OK;
	};
	if((rc = att(desc->color, 4 * (int)d.colorEpp, d.colorBuf, d.colorPitchB, d.colorSliceB, "colour"))) return rc;
	d.depth16 = desc->depth.buffer && desc->depth.format == VKF_D16_UNORM;
	if((rc = att(desc->depth, d.depth16 ? 2 : 4, d.depthBuf, d.depthPitchB, d.depthSliceB, "depth"))) return rc;
	if((rc = att(desc->stencil, 1, d.stencilBuf, d.stencilPitchB, d.stencilSliceB, "stencil"))) return rc;
	d.scX0 = clampi_h(d.scX0, 0, d.fbWidth); d.scX1 = clampi_h(d.scX1, 0, d.fbWidth);
	d.scY0 = clampi_h(d.scY0, 0, d.fbHeight); d.scY1 = clampi_h(d.scY1, 0, d.fbHeight);
	d.suY0 = d.scY0; d.suY1 = d.scY1;

	d.tilesX = (d.fbWidth + SWCU_TILE_W - 1) / SWCU_TILE_W;
	d.tilesY = (d.fbHeight + SWCU_TILE_H - 1) / SWCU_TILE_H;
	d.tileX0 = d.scX0 / SWCU_TILE_W; d.tileY0 = d.scY0 / SWCU_TILE_H;
	d.tileX1 = (d.scX1 + SWCU_TILE_W - 1) / SWCU_TILE_W; d.tileY1 = (d.scY1 + SWCU_TILE_H - 1) / SWCU_TILE_H;
	d.numBins = (uint32_t)(d.tilesX * d.tilesY * 4);
	d.planeOffset = swcu_plane_offset(d.ms);
	d.triStride = swcu_tri_stride(d.nslots, d.ms, d.depthTestActive != 0);
	if(vsrcErr) return vsrcErr;
	if(uboErr) return uboErr;
	return SWCU_OK;
}

// ---- TMA descriptors of the attachments: a 3-D tensor (x, y, sample plane) with a (region width, region height, samples) box ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool tma_eligible(const unsigned char *base, int pitchB, int sliceB, int bpp)
{
	return base && ((uintptr_t)base % 16 == 0) && pitchB > 0 && sliceB > 0 && (pitchB % 16 == 0) && (sliceB % 16 == 0) && (SWCU_REGION_W * bpp) % 16 == 0;
}

// epp: elements per pixel along x (floating-point colour targets are mapped as 2 or 4 32-bit words per pixel)
static bool get_tensor_map(swcu_ctx *ctx, CUtensorMap *out, unsigned char *base, int pitchB, int sliceB, int w, int h, int ms, int bpp, int epp = 1)
{
	if(!ctx->encodeTiled)
	{
		cudaDriverEntryPointQueryResult q;
		void *fn = nullptr;
		if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
		{
			cudaGetLastError();
			return false;
		}
		ctx->encodeTiled = fn;
	}
	const std::vector<uint64_t> key = { (uint64_t)(uintptr_t)base, (uint64_t)pitchB, (uint64_t)sliceB, (uint64_t)w, (uint64_t)h, (uint64_t)ms, (uint64_t)bpp, (uint64_t)epp };
	auto it = ctx->mapCache.find(key);
	if(it == ctx->mapCache.end())
	{
		CUtensorMap m;
		const cuuint64_t dims[3] = { (cuuint64_t)w * epp, (cuuint64_t)h, (cuuint64_t)ms };
		const cuuint64_t strides[2] = { (cuuint64_t)pitchB, (cuuint64_t)sliceB };
		const cuuint32_t box[3] = { (cuuint32_t)(SWCU_REGION_W * epp), SWCU_REGION_H, (cuuint32_t)ms }; // one region: what a warp of the tile kernel stages
		const cuuint32_t estr[3] = { 1, 1, 1 };
		CUresult r = ((EncodeTiledFn)ctx->encodeTiled)(&m, bpp == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : (bpp == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8), 3, base, dims, strides, box, estr,
		                                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if(r != CUDA_SUCCESS) return false;
		if(ctx->mapCache.size() > 256) ctx->mapCache.clear();
		it = ctx->mapCache.emplace(key, m).first;
	}
	*out = it->second;
	return true;
}

template<int MS, int SH, int BL, bool FS>
static void launch_tile4(swcu_ctx *ctx, const DrawConst &d, const TileMaps &maps, dim3 grid)
{
	LaunchScope ls(ctx, MS == 4 ? "k_tile<4>" : "k_tile<1>");
	const int smem = TileLayout<MS, SH>::total(d.depthTestActive != 0, d.stencilActive != 0, FS ? 1 : (int)d.colorEpp);
	static int configured = -1; // per instantiation: the largest dynamic shared-memory size asked for so far
	if(smem > configured)
	{
		cudaFuncSetAttribute(k_tile<MS, SH, BL, FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		cudaFuncSetAttribute(k_tile<MS, SH, BL, FS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		configured = smem;
	}
	k_tile<MS, SH, BL, FS><<<grid, TILE_THREADS, smem, ctx->stream>>>(d, maps);
}

// The common fixed-function state the FS instantiations of k_tile assume (see kernels.cuh); anything else runs the
// instantiation that decodes the state at run time.
static bool fast_state(const swcu_ctx *ctx, const DrawConst &d)
{
	if(!ctx->optFastState) return false;
	if(d.alphaToCoverage || d.depthBounds || d.minDepthClamp != 0.0f || d.maxDepthClamp != 1.0f) return false;
	if(d.stencilActive || d.stencilWrite || !d.colorBuf || d.colorWriteMask != 0xFu || d.bgr || d.srgb || d.colorEpp != 1 || d.depthBiasEnable || d.depth16) return false;
	if(d.depthTestActive && d.depthCompareOp != CMP_LESS && d.depthCompareOp != CMP_LESS_OR_EQUAL) return false;
	if(d.ms == 4 && (d.sampleMask & 0xFu) != 0xFu) return false;
	if(d.blendClass == BL_GENERIC || d.shaderClass == SH_GENERIC) return false;
	for(int ch = 0; ch < 4; ch++)
	{
		const uint32_t kind = d.chanKind[ch];
		if(d.shaderClass == SH_CONST && kind != CK_CONST) return false;
		if(d.shaderClass == SH_VARY && !(kind == CK_CONST || (kind == CK_SLOT && d.slotMode[ch] == IM_PERSP))) return false;
		if(d.shaderClass == SH_TEX && !(kind == CK_TEXEL && d.chanValue[ch] == (uint32_t)ch)) return false;
	}
	if(d.shaderClass == SH_TEX && (d.slotMode[d.uvSlot] != IM_PERSP || d.slotMode[d.uvSlot + 1] != IM_PERSP)) return false;
	return true;
}

template<int MS, int SH, int BL>
static void launch_tile3(swcu_ctx *ctx, const DrawConst &d, const TileMaps &maps, dim3 grid)
{
	if(SH != SH_GENERIC && BL != BL_GENERIC && fast_state(ctx, d)) launch_tile4<MS, SH, BL, (SH != SH_GENERIC && BL != BL_GENERIC)>(ctx, d, maps, grid);
	else launch_tile4<MS, SH, BL, false>(ctx, d, maps, grid);
}
template<int MS, int SH>
static void launch_tile2(swcu_ctx *ctx, const DrawConst &d, const TileMaps &maps, dim3 grid)
{
	switch(d.blendClass)
	{
	case BL_OFF: launch_tile3<MS, SH, BL_OFF>(ctx, d, maps, grid); break;
	case BL_SRC_ALPHA: launch_tile3<MS, SH, BL_SRC_ALPHA>(ctx, d, maps, grid); break;
	default: launch_tile3<MS, SH, BL_GENERIC>(ctx, d, maps, grid); break;
	}
}
template<int MS>
static void launch_tile(swcu_ctx *ctx, const DrawConst &d, const TileMaps &maps, dim3 grid)
{
	switch(d.shaderClass)
	{
	case SH_CONST: launch_tile2<MS, SH_CONST>(ctx, d, maps, grid); break;
	case SH_VARY: launch_tile2<MS, SH_VARY>(ctx, d, maps, grid); break;
	case SH_TEX: launch_tile2<MS, SH_TEX>(ctx, d, maps, grid); break;
	default: launch_tile2<MS, SH_GENERIC>(ctx, d, maps, grid); break;
	}
}

extern "C" int swcu_draw(swcu_ctx *ctx, const swcu_draw_desc *desc)
{
	if(!ctx || !desc) return fail(ctx, SWCU_E_INVALID, "swcu_draw: null argument");
	CU(cudaSetDevice(ctx->device));
	DrawConst d;
	int rc = build_const(ctx, desc, d);
	if(rc) return rc;
	ctx->stats.draws++;
	ctx->stats.primitives += d.primCount;
	if(ctx->profiling == 1) { ctx->lastKernels.clear(); ctx->eventsUsed = 0; } // (2: the timeline keeps growing until swcu_timeline reads it)
	if(d.primCount == 0 || d.sampleMask == 0) return SWCU_OK; // no sample enabled: PixelRoutine.cpp:104-111
	// (an empty render area ends the draw here too — unless this rank is a member of a group: its share of the setup and the barrier of the draw are still due)
	if((d.scX0 >= d.scX1 || d.scY0 >= d.scY1) && !(ctx->group.attached && (ctx->optForceBinned || (int)d.primCount > ctx->optDirectMax))) return SWCU_OK;
	if((unsigned long long)d.primCount * d.triStride > (1ull << 36)) return fail(ctx, SWCU_E_NOMEM, "draw too large");

	const uint32_t n = d.primCount;
	d.direct = (!ctx->optForceBinned && (int)n <= ctx->optDirectMax) ? 1u : 0u;
	swcu_ctx::Group &G = ctx->group;
	const bool grouped = G.attached && !d.direct;
	// The setup phase of this draw uses the set the draw before the previous one used.  A binned draw runs it on the setup
	// stream: it waits only for the inputs (last upload) and for the last reader of this set, not for the tile kernel of the
	// previous draw.  Direct draws, profiling mode (per-kernel events) and inputs in caller-owned device memory (whose producers the
	// library cannot see) stay on the main stream.  No step of a draw waits for the host.
	const int setIndex = ctx->cur;
	swcu_ctx::SetupSet &S = ctx->set[setIndex];
	ctx->cur = (ctx->cur + 1) % SWCU_SETS;
	const bool pipelined = !d.direct && ctx->optPipeline && ctx->profiling != 1 && !d.inputsExternal;
	const int ssIndex = grouped ? setIndex : 0; // a group member: one setup stream per set (see swcu_ctx::setupStream)
	const cudaStream_t ss = pipelined ? ctx->setupStream[ssIndex] : ctx->stream;
	if(pipelined)
	{
		if(S.tileDoneValid) CU(cudaStreamWaitEvent(ss, S.tileDone, 0));
		// The peers' buffers of this draw's set may be written once the barrier of the previous draw has been passed: every rank checks
		// in there only when the last reader of ITS copy of the set has finished (see below).  On one stream the order was implicit.
		if(grouped && ctx->evBarrierValid) CU(cudaStreamWaitEvent(ss, ctx->evBarrier, 0));
		ctx->setupReadsInputs[ssIndex] = true;
	}
	else ctx->mainReadsInputs = true; // the setup phase reads the vertex / index streams on the main stream, asynchronously
	{
		// the vertex / index streams the setup phase reads: wait for THEIR uploads only
		uint64_t &seen = pipelined ? ctx->setupSawUpload[ssIndex] : ctx->mainSawUpload;
		for(int i = 0; i < SWCU_MAX_INPUTS; i++)
			if(desc->input[i].buffer && (rc = see_upload(ctx, find_shadow(ctx, desc->input[i].buffer, 1), ss, seen))) return rc;
		if(desc->indexBuffer && (rc = see_upload(ctx, find_shadow(ctx, desc->indexBuffer, 1), ss, seen))) return rc;
	}
	if((rc = main_touches(ctx, desc->color.buffer, true)) || (rc = main_touches(ctx, desc->depth.buffer, true)) || (rc = main_touches(ctx, desc->stencil.buffer, true))) return rc;
	for(uint32_t t = 0; t < desc->sampledImageCount && t < SWCU_MAX_SAMPLED_IMAGES; t++)
		for(uint32_t l = 0; l < desc->sampledImage[t].levelCount && l < SWCU_MIPMAP_LEVELS; l++)
			if((rc = main_touches(ctx, desc->sampledImage[t].level[l].buffer, false))) return rc;

	// ---- work buffers: every capacity follows from the triangle count, nothing is sized by a device-side result ----
	const size_t scanBlocks = ((size_t)d.numBins + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS);
	const size_t countersBytes = sizeof(DrawCounters) + 4 * (scanBlocks + 1);
	const size_t pairCap = d.direct ? 0 : std::min<size_t>((size_t)4 * n + ctx->optBigPairBudget, 0x3FFFFFF0u);
	if(!d.direct && pairCap < (size_t)4 * n) return fail(ctx, SWCU_E_NOMEM, "%u triangles exceed the pair index space", n);
	d.world = 1; d.rank = 0; d.triLo = 0; d.triHi = n; d.bandRows = d.fbHeight;
	if(grouped)
	{
		// the buffers the peers write lie in the exported arena; this rank sets up its share of the triangles for the whole frame
		if(n > G.maxPrims || d.triStride > G.stride || (uint32_t)d.fbWidth != G.fbW || (uint32_t)d.fbHeight != G.fbH || countersBytes > G.countersBytes)
			return fail(ctx, SWCU_E_INVALID, "group draw outside what swcu_group_reserve sized (%u triangles of %u bytes, %dx%d)", n, d.triStride, d.fbWidth, d.fbHeight);
		if(d.fbHeight % (int)G.world) return fail(ctx, SWCU_E_INVALID, "framebuffer height %d does not split into %u bands", d.fbHeight, G.world);
		d.world = G.world; d.rank = G.rank; d.bandRows = d.fbHeight / (int)G.world;
		if(d.bandRows < 2 * SWCU_REGION_H) return fail(ctx, SWCU_E_UNSUPPORTED, "bands of %d rows are too small for a group", d.bandRows);
		d.triLo = (uint32_t)((unsigned long long)n * G.rank / G.world);
		d.triHi = (uint32_t)((unsigned long long)n * (G.rank + 1) / G.world);
		d.suY0 = clampi_h(desc->scissor.y, 0, d.fbHeight);
		d.suY1 = clampi_h(desc->scissor.y + (int)desc->scissor.height, 0, d.fbHeight);
		// my rows: my band of the scissor (NOT the render area the caller passed, which must be that band)
		d.scY0 = clampi_h(d.suY0, (int)G.rank * d.bandRows, (int)(G.rank + 1) * d.bandRows);
		d.scY1 = clampi_h(d.suY1, (int)G.rank * d.bandRows, (int)(G.rank + 1) * d.bandRows);
		d.tileY0 = d.scY0 / SWCU_TILE_H; d.tileY1 = (d.scY1 + SWCU_TILE_H - 1) / SWCU_TILE_H;
		for(uint32_t p = 0; p < G.world; p++)
		{
			d.peerRecords[p] = G.peer[p] + G.offRecords[setIndex];
			d.peerRect[p] = (uint32_t *)(G.peer[p] + G.offRect[setIndex]);
			d.peerBig[p] = (BigTri *)(G.peer[p] + G.offBig[setIndex]);
			d.peerBinCount[p] = (uint32_t *)(G.peer[p] + G.offBinCount[setIndex]);
			d.peerCounters[p] = (DrawCounters *)(G.peer[p] + G.offCounters[setIndex]);
		}
		d.triRecords = d.peerRecords[G.rank];
		d.triRect = d.peerRect[G.rank];
		d.bigList = d.peerBig[G.rank];
		d.bigCapacity = G.maxPrims;
		d.binCount = d.peerBinCount[G.rank];
		d.counters = d.peerCounters[G.rank];
	}
	else
	{
		if((rc = ensure(ctx, S.triRecords, (size_t)n * d.triStride))) return rc;
		if((rc = ensure(ctx, S.triRect, (size_t)n * 4))) return rc;
		if((rc = ensure(ctx, S.bigList, (size_t)n * sizeof(BigTri)))) return rc; // every triangle may be a big one
		if((rc = ensure(ctx, S.counters, countersBytes))) return rc;
		if(!d.direct)
		{
			const size_t before = S.binCount.cap;
			if((rc = ensure(ctx, S.binCount, (size_t)d.numBins * 4))) return rc;
			if(S.binCount.cap != before) CU(cudaMemsetAsync(S.binCount.p, 0, S.binCount.cap, ss)); // k_fill counts every bin back down to zero
		}
		d.triRecords = (unsigned char *)S.triRecords.p;
		d.triRect = (uint32_t *)S.triRect.p;
		d.counters = (DrawCounters *)S.counters.p;
		d.bigList = (BigTri *)S.bigList.p;
		d.bigCapacity = (uint32_t)std::min<size_t>(S.bigList.cap / sizeof(BigTri), 0x7FFFFFFFu);
		d.binCount = (uint32_t *)S.binCount.p;
		d.peerRecords[0] = d.triRecords; d.peerRect[0] = d.triRect; d.peerBig[0] = d.bigList; d.peerBinCount[0] = d.binCount; d.peerCounters[0] = d.counters;
	}
	if(!d.direct)
	{
		if((rc = ensure(ctx, S.binStart, ((size_t)d.numBins + 1) * 4))) return rc;
		if((rc = ensure(ctx, S.pairs, pairCap * 4))) return rc;
	}
	if(!ctx->zeroPage.p)
	{
		if((rc = ensure(ctx, ctx->zeroPage, 256))) return rc;
		CU(cudaMemsetAsync(ctx->zeroPage.p, 0, 256, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
	}
	d.zeroPage = ctx->zeroPage.p;
	d.hostCounters = d.direct ? nullptr : S.hostCounters;
	d.binStart = (uint32_t *)S.binStart.p;
	d.pairs = (uint32_t *)S.pairs.p;
	d.bigBudget = d.direct ? 0 : pairCap - (size_t)4 * n;
	const dim3 tileGrid((unsigned)std::max(d.tileX1 - d.tileX0, 0), (unsigned)std::max(d.tileY1 - d.tileY0, 0));

	// (a group member's counters and rectangles were reset behind the last draw that used this set: peers may be writing already)
	if(!grouped) CU(cudaMemsetAsync(d.counters, 0, countersBytes, ss));
	// band mode without a group: the scissor / render area keeps less than 3/4 of the framebuffer rows
	d.cullFlags = nullptr;
	if(!d.direct && !grouped && !d.vsProgLen && d.primKind == PRIM_TRIANGLE && n >= 4096 && (long long)(d.scY1 - d.scY0) * 4 < (long long)d.fbHeight * 3)
	{
		if((rc = ensure(ctx, S.cullFlags, n))) return rc;
		LaunchScope ls(ctx, "k_cull", ss);
		k_cull<<<(n + 255) / 256, 256, 0, ss>>>(d, (unsigned char *)S.cullFlags.p);
		d.cullFlags = (const unsigned char *)S.cullFlags.p;
	}
	if(d.triHi > d.triLo)
	{
		LaunchScope ls(ctx, "k_setup", ss);
		const uint32_t share = d.triHi - d.triLo;
		const size_t scratch = (size_t)SWCU_SMALL_ROWS * d.ms * SETUP_THREADS * 4;
		if(d.vsProgLen || d.primKind != PRIM_TRIANGLE) k_setup_prog<<<(share + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, scratch, ss>>>(d); // the vertex stage has arithmetic, or the primitives are lines / points
		else if(d.ms != 1)
		{
			// wave quantisation (see k_setup_wide): the variant with more resident CTAs when it saves enough waves to pay for its slower ones
			const unsigned blocks = (share + SETUP_THREADS - 1) / SETUP_THREADS, sms = (unsigned)std::max(ctx->smCount, 1);
			const unsigned w6 = (blocks + sms * SETUP_BLOCKS_4X - 1) / (sms * SETUP_BLOCKS_4X), w7 = (blocks + sms * SETUP_BLOCKS_WIDE - 1) / (sms * SETUP_BLOCKS_WIDE);
			if(ctx->optSetupWide == 2 /* forced: parity tests */ || (ctx->optSetupWide && w7 * 1.24f < w6 * 0.9f)) k_setup_wide<<<blocks, SETUP_THREADS, scratch, ss>>>(d);
			else k_setup<<<blocks, SETUP_THREADS, scratch, ss>>>(d);
		}
		else k_setup_1x<<<(share + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, scratch, ss>>>(d);
	}
	if(grouped)
	{
		// Every rank's share has to be in my buffers before I bin: the one collective step of a group draw.  Passing this barrier also
		// lets the peers start on the NEXT draw, i.e. write into my next buffer set — so its last reader (the tile kernel of the draw
		// before the previous one) and the reset behind it must be done before I check in.
		swcu_ctx::SetupSet &next = ctx->set[ctx->cur];
		if(pipelined && next.tileDoneValid) CU(cudaStreamWaitEvent(ss, next.tileDone, 0));
		GroupFlags gf;
		memset(&gf, 0, sizeof(gf));
		for(uint32_t p = 0; p < G.world; p++) gf.flags[p] = (uint32_t *)(G.peer[p] + G.offFlags[setIndex]);
		LaunchScope ls(ctx, "k_xbarrier", ss);
		k_xbarrier<<<1, 32, 0, ss>>>(gf, G.world, G.rank, ++G.epoch[setIndex]);
		if(pipelined)
		{
			if(!ctx->evBarrier) CU(cudaEventCreateWithFlags(&ctx->evBarrier, cudaEventDisableTiming));
			CU(cudaEventRecord(ctx->evBarrier, ss));
			ctx->evBarrierValid = true;
		}
	}
	const bool nothingToDraw = d.scX0 >= d.scX1 || d.scY0 >= d.scY1; // (only a group member gets here with an empty band)
	if(!d.direct)
	{
		// ---- binning: count (k_setup + the first blocks of k_binscan), scan, fill; all sized on the host, all asynchronous ----
		const unsigned bigBlocks = 148; // grid-stride over the big list, whose length only the device knows
		{
			LaunchScope ls(ctx, "k_binscan", ss);
			k_binscan<<<bigBlocks + (unsigned)scanBlocks, SCAN_THREADS, 0, ss>>>(d, bigBlocks);
		}
		{
			LaunchScope ls(ctx, "k_fill", ss);
			const unsigned smallBlocks = (n + 256 * FILL_TRIS - 1) / (256 * FILL_TRIS);
			k_fill<<<smallBlocks + bigBlocks, 256, 0, ss>>>(d, smallBlocks);
		}
		S.countersPending = true;
		if(pipelined)
		{
			CU(cudaEventRecord(S.setupDone, ss));
			CU(cudaStreamWaitEvent(ctx->stream, S.setupDone, 0));
		}
	}
	// ---- tensor maps of the attachments the tile kernel stages ----
	TileMaps maps;
	memset(&maps, 0, sizeof(maps));
	{
		const bool colorOn = d.colorWriteMask != 0 && d.colorBuf;
		bool ok = ctx->optTma != 0;
		if(ok && colorOn) ok = tma_eligible(d.colorBuf, d.colorPitchB, d.colorSliceB, 4 * (int)d.colorEpp);
		if(ok && d.depthTestActive) ok = tma_eligible(d.depthBuf, d.depthPitchB, d.depthSliceB, d.depth16 ? 2 : 4);
		if(ok && d.stencilActive) ok = tma_eligible(d.stencilBuf, d.stencilPitchB, d.stencilSliceB, 1);
		if(ok && colorOn) ok = get_tensor_map(ctx, &maps.color, d.colorBuf, d.colorPitchB, d.colorSliceB, d.fbWidth, d.fbHeight, d.ms, 4, (int)d.colorEpp);
		if(ok && d.depthTestActive) ok = get_tensor_map(ctx, &maps.depth, d.depthBuf, d.depthPitchB, d.depthSliceB, d.fbWidth, d.fbHeight, d.ms, d.depth16 ? 2 : 4);
		if(ok && d.stencilActive) ok = get_tensor_map(ctx, &maps.stencil, d.stencilBuf, d.stencilPitchB, d.stencilSliceB, d.fbWidth, d.fbHeight, d.ms, 1);
		d.useTma = ok ? 1u : 0u;
	}
	// No attachment is read: the region warps store the colour of a fragment straight into the framebuffer (no staging, no write-back)
	d.writeOnly = (ctx->optWriteOnly && d.blendClass == BL_OFF && !d.depthTestActive && d.shaderClass != SH_GENERIC && fast_state(ctx, d)) ? 1u : 0u;
	if(!nothingToDraw) { if(d.ms == 4) launch_tile<4>(ctx, d, maps, tileGrid); else launch_tile<1>(ctx, d, maps, tileGrid); }
	CU(cudaGetLastError());
	if(grouped)
	{
		// reset what the peers will write into two draws from now: their next share arrives behind the barrier of the draw in between
		CU(cudaMemsetAsync(d.counters, 0, G.countersBytes, ctx->stream));
		// (the rectangles are put back to "none" by k_fill as it reads them)
	}
	CU(cudaEventRecord(S.tileDone, ctx->stream)); // last reader of this set's records / bins
	S.tileDoneValid = true;
	return SWCU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// groups: work buffers in one IPC-exported allocation per rank, mapped by every peer
// ------------------------------------------------------------------------------------------------------------------
extern "C" int swcu_group_reserve(swcu_ctx *ctx, const swcu_group_desc *g, void *handle64)
{
	if(!ctx || !g || !handle64 || g->structSize != sizeof(swcu_group_desc)) return fail(ctx, SWCU_E_INVALID, "swcu_group_reserve: bad argument");
	if(g->world < 1 || g->world > SWCU_MAX_GROUP || g->rank >= g->world) return fail(ctx, SWCU_E_INVALID, "swcu_group_reserve: rank %u of %u", g->rank, g->world);
	if(g->maxSlots > SWCU_MAXSLOTS || (g->maxSamples != 1 && g->maxSamples != 4) || !g->maxPrimitives || !g->fbWidth || !g->fbHeight || g->fbWidth > 8192 || g->fbHeight > 8192)
		return fail(ctx, SWCU_E_INVALID, "swcu_group_reserve: bad sizes");
	swcu_ctx::Group &G = ctx->group;
	if(G.reserved) return fail(ctx, SWCU_E_INVALID, "swcu_group_reserve: a group is already reserved on this context");
	CU(cudaSetDevice(ctx->device));
	G.rank = g->rank; G.world = g->world; G.maxPrims = g->maxPrimitives; G.fbW = g->fbWidth; G.fbH = g->fbHeight;
	G.stride = swcu_tri_stride((int)g->maxSlots, (int)g->maxSamples, 1);
	const uint32_t tilesX = (G.fbW + SWCU_TILE_W - 1) / SWCU_TILE_W, tilesY = (G.fbH + SWCU_TILE_H - 1) / SWCU_TILE_H;
	G.numBins = tilesX * tilesY * 4;
	const size_t scanBlocks = ((size_t)G.numBins + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS);
	G.countersBytes = (sizeof(DrawCounters) + 4 * (scanBlocks + 1) + 255) & ~(size_t)255;
	size_t off = 0;
	auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
	for(int k = 0; k < SWCU_SETS; k++)
	{
		G.offRecords[k] = take((size_t)G.maxPrims * G.stride);
		G.offRect[k] = take((size_t)G.maxPrims * 4);
		G.offBig[k] = take((size_t)G.maxPrims * sizeof(BigTri));
		G.offBinCount[k] = take((size_t)G.numBins * 4);
		G.offCounters[k] = take(G.countersBytes);
		G.offFlags[k] = take(SWCU_MAX_GROUP * 4);
	}
	G.arenaBytes = off;
	cudaError_t e = cudaMalloc((void **)&G.arena, G.arenaBytes);
	if(e != cudaSuccess) { cudaGetLastError(); return fail(ctx, SWCU_E_NOMEM, "cudaMalloc(%zu) for the group arena failed: %s", G.arenaBytes, cudaGetErrorString(e)); }
	CU(cudaMemsetAsync(G.arena, 0, G.arenaBytes, ctx->stream));
	for(int k = 0; k < SWCU_SETS; k++) CU(cudaMemsetAsync(G.arena + G.offRect[k], 0xFF, (size_t)G.maxPrims * 4, ctx->stream)); // TRI_RECT_NONE
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, G.arena));
	G.reserved = true;
	return SWCU_OK;
}

extern "C" int swcu_group_attach(swcu_ctx *ctx, const void *handles)
{
	if(!ctx || !handles) return fail(ctx, SWCU_E_INVALID, "swcu_group_attach: null argument");
	swcu_ctx::Group &G = ctx->group;
	if(!G.reserved || G.attached) return fail(ctx, SWCU_E_INVALID, "swcu_group_attach: reserve first, attach once");
	CU(cudaSetDevice(ctx->device));
	int rc = swcu_sync(ctx);
	if(rc) return rc;
	for(uint32_t p = 0; p < G.world; p++)
	{
		if(p == G.rank) { G.peer[p] = G.arena; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, (const unsigned char *)handles + 64 * (size_t)p, sizeof(h));
		void *base = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
		if(e != cudaSuccess)
		{
			cudaGetLastError();
			for(uint32_t q = 0; q < p; q++)
				if(q != G.rank && G.peer[q]) { cudaIpcCloseMemHandle(G.peer[q]); G.peer[q] = nullptr; }
			return fail(ctx, SWCU_E_CUDA, "cudaIpcOpenMemHandle of rank %u's work buffers failed: %s", p, cudaGetErrorString(e));
		}
		G.peer[p] = (unsigned char *)base;
	}
	for(auto &e : G.epoch) e = 0;
	ctx->cur = 0; // every rank starts the group with the same buffer set
	G.attached = true;
	CU(make_setup_stream(ctx, true));
	return SWCU_OK;
}

extern "C" int swcu_group_detach(swcu_ctx *ctx)
{
	if(!ctx) return SWCU_E_INVALID;
	swcu_ctx::Group &G = ctx->group;
	if(!G.reserved) return SWCU_OK;
	CU(cudaSetDevice(ctx->device));
	int rc = swcu_sync(ctx);
	for(uint32_t p = 0; p < G.world; p++)
		if(G.attached && p != G.rank && G.peer[p]) cudaIpcCloseMemHandle(G.peer[p]);
	cudaFree(G.arena);
	cudaGetLastError();
	G = swcu_ctx::Group();
	if(ctx->setupStream[0]) make_setup_stream(ctx, false); // (not while the context is being torn down)
	return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// multi-GPU plumbing: CUDA IPC handles of shadows, band copies into peer memory, flags
// ------------------------------------------------------------------------------------------------------------------
extern "C" int swcu_ipc_export(swcu_ctx *ctx, const void *ptr, void *handle64, uint64_t *offset)
{
	if(!ctx || !ptr || !handle64 || !offset) return fail(ctx, SWCU_E_INVALID, "swcu_ipc_export: null argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
	Shadow *s = find_shadow(ctx, ptr, 1);
	if(!s || s->external) return fail(ctx, SWCU_E_INVALID, "swcu_ipc_export: %p is not inside a shadow owned by this context", ptr);
	CU(cudaSetDevice(ctx->device));
	CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, s->dev));
	*offset = (uint64_t)((uintptr_t)ptr - s->host);
	return SWCU_OK;
}

extern "C" int swcu_ipc_open(swcu_ctx *ctx, const void *handle64, void **device_base)
{
	if(!ctx || !handle64 || !device_base) return fail(ctx, SWCU_E_INVALID, "swcu_ipc_open: null argument");
	CU(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	CU(cudaIpcOpenMemHandle(device_base, h, cudaIpcMemLazyEnablePeerAccess));
	return SWCU_OK;
}

extern "C" int swcu_ipc_close(swcu_ctx *ctx, void *device_base)
{
	if(!ctx || !device_base) return SWCU_E_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaIpcCloseMemHandle(device_base));
	return SWCU_OK;
}

extern "C" int swcu_copy_image(swcu_ctx *ctx, const swcu_attachment *src, const swcu_attachment *dst)
{
	if(!ctx || !src || !dst || !src->buffer || !dst->buffer) return fail(ctx, SWCU_E_INVALID, "swcu_copy_image: null argument");
	if(src->format != dst->format || (src->format != VKF_R8G8B8A8_UNORM && src->format != VKF_B8G8R8A8_UNORM)) return fail(ctx, SWCU_E_UNSUPPORTED, "swcu_copy_image: format unsupported");
	if(src->width != dst->width || src->height != dst->height) return fail(ctx, SWCU_E_INVALID, "swcu_copy_image: extent mismatch");
	CU(cudaSetDevice(ctx->device));
	const size_t rowB = (size_t)src->width * 4;
	unsigned char *s = dev_ptr(ctx, src->buffer, (size_t)(src->height - 1) * src->pitchB + rowB);
	unsigned char *t = dev_ptr(ctx, dst->buffer, (size_t)(dst->height - 1) * dst->pitchB + rowB);
	if(!s || !t) return fail(ctx, SWCU_E_INVALID, "swcu_copy_image: image is not inside a registered range");
	const int vec = ((uintptr_t)s % 16 == 0) && ((uintptr_t)t % 16 == 0) && (src->pitchB % 16 == 0) && (dst->pitchB % 16 == 0) && (rowB % 16 == 0);
	const int per = vec ? 16 : 4;
	int rc;
	const cudaStream_t st = ctx->handover();
	if(ctx->sideActive)
	{
		// on the hand-over stream: it has waited for the main stream at swcu_side_begin; what is left are copies in flight on the two
		// images (the bookkeeping of the main stream is not touched: it has not seen these waits)
		const swcu_attachment *both[2] = { src, dst };
		for(int i = 0; i < 2; i++)
		{
			Shadow *sh = find_shadow(ctx, both[i]->buffer, 1);
			if(sh && sh->upEvent && sh->upSeq) CU(cudaStreamWaitEvent(st, sh->upEvent, 0));
			if(sh && i == 1 && sh->dlEvent && (sh->dlPendingMain || sh->dlPendingUpload)) CU(cudaStreamWaitEvent(st, sh->dlEvent, 0));
		}
	}
	else if((rc = main_touches(ctx, src->buffer, false)) || (rc = main_touches(ctx, dst->buffer, true))) return rc;
	LaunchScope ls(ctx, "k_copy_rows", st);
	k_copy_rows<<<dim3((unsigned)((rowB / per + 255) / 256), src->height), 256, 0, st>>>(s, src->pitchB, t, dst->pitchB, (int)rowB, (int)src->height, vec);
	CU(cudaGetLastError());
	return SWCU_OK;
}

extern "C" int swcu_signal(swcu_ctx *ctx, void *flag, uint32_t value)
{
	if(!ctx || !flag) return fail(ctx, SWCU_E_INVALID, "swcu_signal: null argument");
	CU(cudaSetDevice(ctx->device));
	unsigned char *f = dev_ptr(ctx, flag, 4);
	if(!f) return fail(ctx, SWCU_E_INVALID, "swcu_signal: flag is not inside a registered range");
	// A flag tells a peer that this rank's frame may be overwritten, so downloads still in flight come first.  The flag is then
	// written from the download stream, behind them (and behind everything issued on the main stream so far): the main stream
	// itself is not held back and goes on with the next frame's draw.
	cudaStream_t st = ctx->handover();
	if(ctx->optCopyStreams && ctx->mainSawDownload != ctx->downloadSeq)
	{
		CU(cudaEventRecord(ctx->evMark, st));
		st = ctx->d2hStream;
		CU(cudaStreamWaitEvent(st, ctx->evMark, 0));
		ctx->mainSawDownload = ctx->downloadSeq;
	}
	LaunchScope ls(ctx, "k_signal", st);
	k_signal<<<1, 1, 0, st>>>((uint32_t *)f, value);
	CU(cudaGetLastError());
	return SWCU_OK;
}

extern "C" int swcu_wait_flags(swcu_ctx *ctx, const void *flags, uint32_t first, uint32_t count, uint32_t value)
{
	if(!ctx || !flags || count > 64) return fail(ctx, SWCU_E_INVALID, "swcu_wait_flags: bad argument");
	if(count == 0) return SWCU_OK;
	CU(cudaSetDevice(ctx->device));
	unsigned char *f = dev_ptr(ctx, flags, (size_t)(first + count) * 4);
	if(!f) return fail(ctx, SWCU_E_INVALID, "swcu_wait_flags: flags are not inside a registered range");
	LaunchScope ls(ctx, "k_wait_flags", ctx->handover());
	k_wait_flags<<<1, 64, 0, ctx->handover()>>>((const uint32_t *)f, (int)first, (int)count, value);
	CU(cudaGetLastError());
	return SWCU_OK;
}

// ---- the hand-over stream ----
extern "C" int swcu_side_begin(swcu_ctx *ctx)
{
	if(!ctx) return SWCU_E_INVALID;
	if(ctx->sideActive) return fail(ctx, SWCU_E_INVALID, "swcu_side_begin: already begun");
	CU(cudaSetDevice(ctx->device));
	if(!ctx->sideStream)
	{
		CU(cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&ctx->evSideBegin, cudaEventDisableTiming));
	}
	CU(cudaEventRecord(ctx->evSideBegin, ctx->stream));
	CU(cudaStreamWaitEvent(ctx->sideStream, ctx->evSideBegin, 0));
	ctx->sideActive = true;
	return SWCU_OK;
}

extern "C" int swcu_side_end(swcu_ctx *ctx, uint32_t slot)
{
	if(!ctx || slot >= SWCU_SIDE_SLOTS) return fail(ctx, SWCU_E_INVALID, "swcu_side_end: bad slot");
	if(!ctx->sideActive) return fail(ctx, SWCU_E_INVALID, "swcu_side_end without swcu_side_begin");
	CU(cudaSetDevice(ctx->device));
	if(!ctx->sideDone[slot]) CU(cudaEventCreateWithFlags(&ctx->sideDone[slot], cudaEventDisableTiming));
	CU(cudaEventRecord(ctx->sideDone[slot], ctx->sideStream));
	ctx->sideDoneValid[slot] = true;
	ctx->sideActive = false;
	return SWCU_OK;
}

extern "C" int swcu_side_wait(swcu_ctx *ctx, uint32_t slot)
{
	if(!ctx || slot >= SWCU_SIDE_SLOTS) return fail(ctx, SWCU_E_INVALID, "swcu_side_wait: bad slot");
	if(!ctx->sideDoneValid[slot]) return SWCU_OK;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamWaitEvent(ctx->stream, ctx->sideDone[slot], 0));
	return SWCU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// clear / resolve on the resident shadows
// ------------------------------------------------------------------------------------------------------------------
extern "C" int swcu_clear(swcu_ctx *ctx, const swcu_attachment *att, uint32_t samples, const swcu_rect *area, const void *value)
{
	if(!ctx || !att || !area || !value || !att->buffer) return fail(ctx, SWCU_E_INVALID, "swcu_clear: null argument");
	int bpp;
	switch(att->format)
	{
	case VKF_R8G8B8A8_UNORM: case VKF_B8G8R8A8_UNORM: case VKF_R8G8B8A8_SRGB: case VKF_B8G8R8A8_SRGB: case VKF_D32_SFLOAT: bpp = 4; break;
	case VKF_D16_UNORM: bpp = 2; break;
	case VKF_R16G16B16A16_SFLOAT: bpp = 8; break;
	case VKF_R32G32B32A32_SFLOAT: bpp = 16; break;
	case VKF_S8_UINT: bpp = 1; break;
	default: return fail(ctx, SWCU_E_UNSUPPORTED, "swcu_clear: format %u unsupported", att->format);
	}
	if(samples < 1 || area->x < 0 || area->y < 0 || area->x + area->width > att->width || area->y + area->height > att->height)
		return fail(ctx, SWCU_E_INVALID, "swcu_clear: area outside the attachment");
	if(area->width == 0 || area->height == 0) return SWCU_OK;
	CU(cudaSetDevice(ctx->device));
	const size_t need = (size_t)(samples - 1) * att->sliceB + (size_t)(att->height - 1) * att->pitchB + (size_t)att->width * bpp;
	unsigned char *base = dev_ptr(ctx, att->buffer, need);
	if(!base) return fail(ctx, SWCU_E_INVALID, "swcu_clear: attachment is not inside a registered range");
	uint4 v = make_uint4(0, 0, 0, 0);
	memcpy(&v, value, (size_t)bpp);
	int rc;
	if((rc = main_touches(ctx, att->buffer, true))) return rc;
	LaunchScope ls(ctx, "k_clear");
	k_clear<<<dim3((area->width + 255) / 256, area->height), 256, 0, ctx->stream>>>(base, att->pitchB, att->sliceB, bpp, area->x, area->y, (int)area->width, (int)area->height, (int)samples, v);
	CU(cudaGetLastError());
	return SWCU_OK;
}

extern "C" int swcu_resolve(swcu_ctx *ctx, const swcu_attachment *src, uint32_t samples, const swcu_attachment *dst)
{
	if(!ctx || !src || !dst || !src->buffer || !dst->buffer) return fail(ctx, SWCU_E_INVALID, "swcu_resolve: null argument");
	if(samples != 4) return fail(ctx, SWCU_E_UNSUPPORTED, "swcu_resolve: only 4x -> 1x");
	const bool fast = src->format == VKF_R8G8B8A8_UNORM || src->format == VKF_B8G8R8A8_UNORM; // Blitter::fastResolve; everything else: the generic blit
	const int epp = src->format == VKF_R32G32B32A32_SFLOAT ? 4 : (src->format == VKF_R16G16B16A16_SFLOAT ? 2 : 1);
	if(src->format != dst->format || (!fast && epp == 1 && src->format != VKF_R8G8B8A8_SRGB && src->format != VKF_B8G8R8A8_SRGB))
		return fail(ctx, SWCU_E_UNSUPPORTED, "swcu_resolve: format unsupported (RGBA8 / BGRA8 UNORM or SRGB, R16G16B16A16_SFLOAT, R32G32B32A32_SFLOAT)");
	if(src->width != dst->width || src->height != dst->height) return fail(ctx, SWCU_E_INVALID, "swcu_resolve: extent mismatch");
	CU(cudaSetDevice(ctx->device));
	unsigned char *s = dev_ptr(ctx, src->buffer, (size_t)3 * src->sliceB + (size_t)(src->height - 1) * src->pitchB + (size_t)src->width * 4 * epp);
	unsigned char *t = dev_ptr(ctx, dst->buffer, (size_t)(dst->height - 1) * dst->pitchB + (size_t)dst->width * 4 * epp);
	if(!s || !t) return fail(ctx, SWCU_E_INVALID, "swcu_resolve: attachment is not inside a registered range");
	int rc;
	if((rc = main_touches(ctx, src->buffer, false)) || (rc = main_touches(ctx, dst->buffer, true))) return rc;
	LaunchScope ls(ctx, fast ? "k_resolve4" : "k_resolve4_generic");
	if(fast) k_resolve4<<<dim3((src->width + 255) / 256, src->height), 256, 0, ctx->stream>>>(s, src->pitchB, src->sliceB, t, dst->pitchB, (int)src->width, (int)src->height);
	else k_resolve4_generic<<<dim3((src->width + 127) / 128, src->height), 128, 0, ctx->stream>>>(s, src->pitchB, src->sliceB, t, dst->pitchB, (int)src->width, (int)src->height, epp);
	CU(cudaGetLastError());
	return SWCU_OK;
}
