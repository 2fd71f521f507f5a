// api.cu -- the b200ic_* C-ABI (include/b200ic.h): device selection, per-codec dispatch, and the
// host-buffer path (block-row chunks pipelined H2D -> kernel -> D2H over a small ring of streams).
// No CPU fallback exists: without a usable CUDA device every encode call fails loudly.
#include "kernels.h"
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace b200ic {

static thread_local std::string t_error;
static std::atomic<uint64_t> g_launches{0};

static int fail(const char *what, cudaError_t e = cudaSuccess) {
	char buf[512];
	if (e != cudaSuccess) snprintf(buf, sizeof(buf), "b200ic: %s: %s", what, cudaGetErrorString(e));
	else snprintf(buf, sizeof(buf), "b200ic: %s", what);
	t_error = buf;
	return -1;
}

#define B200IC_CUDA(call, what)                      \
	do {                                               \
		cudaError_t e__ = (call);                        \
		if (e__ != cudaSuccess) return fail(what, e__);  \
	} while (0)

static uint32_t block_bytes(int codec) {
	switch (codec) {
	case B200IC_BC1:
	case B200IC_BC23_COLOUR_HALF:
	case B200IC_BC2_ALPHA_HALF:
	case B200IC_BC4: return 8;
	case B200IC_BC2:
	case B200IC_BC3:
	case B200IC_BC5:
	case B200IC_BC6H:
	case B200IC_BC7_AMD:
	case B200IC_BC7_RG: return 16;
	default: return 0;
	}
}

static uint32_t texel_bytes(int fmt) {
	switch (fmt) {
	case B200IC_FMT_R8: return 1;
	case B200IC_FMT_RG8: return 2;
	case B200IC_FMT_RGB8:
	case B200IC_FMT_RGB8_SRGB: return 3;
	case B200IC_FMT_RGBA8:
	case B200IC_FMT_RGBA8_SRGB: return 4;
	case B200IC_FMT_RGBA16F:
	case B200IC_FMT_RGBA16UF: return 8;
	case B200IC_FMT_RGBA32F: return 16;
	default: return 0;
	}
}

static uint32_t blocks_format_bytes(int fmt) { // bytes per pre-gathered block
	switch (fmt) {
	case B200IC_FMT_BLOCKS_F32X1: return 64;
	case B200IC_FMT_BLOCKS_F32X3: return 192;
	case B200IC_FMT_BLOCKS_F32X4: return 256;
	case B200IC_FMT_BLOCKS_RGBA8: return 64;
	default: return 0;
	}
}

// ---- per-device host-path context (process-wide pool): stream ring, grow-only device scratch and pinned staging ---
// One host-path encode runs per device at a time (`mu`); the contexts outlive the calling threads, so a thread that
// encodes and exits leaks nothing, and b200ic_shutdown() releases every device's context.
struct DevCtx {
	static constexpr int kStreams = 3;
	std::mutex mu;
	int device = -1;
	cudaStream_t streams[kStreams] = {};
	void *d_in[kStreams] = {};
	void *d_out[kStreams] = {};
	size_t in_cap[kStreams] = {};
	size_t out_cap[kStreams] = {};
	void *p_in[kStreams] = {};  // pinned staging for pageable caller buffers
	void *p_out[kStreams] = {};
	size_t pin_cap[kStreams] = {};
	size_t pout_cap[kStreams] = {};
	bool ready = false;
	cudaEvent_t fork = nullptr, join[kStreams] = {}, kdone[kStreams] = {};

	int ensure(int dev) { // call with `mu` held and `dev` current
		if (ready) return 0;
		device = dev;
		for (int i = 0; i < kStreams; i++) {
			if (!streams[i]) B200IC_CUDA(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking), "stream create");
			if (!join[i]) B200IC_CUDA(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming), "event create");
			if (!kdone[i]) B200IC_CUDA(cudaEventCreateWithFlags(&kdone[i], cudaEventDisableTiming), "event create");
		}
		if (!fork) B200IC_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "event create");
		ready = true;
		return 0;
	}
	int reserve(int i, size_t in_bytes, size_t out_bytes) {
		if (in_bytes > in_cap[i]) {
			if (d_in[i]) cudaFree(d_in[i]);
			d_in[i] = nullptr;
			in_cap[i] = 0;
			B200IC_CUDA(cudaMalloc(&d_in[i], in_bytes), "cudaMalloc(input chunk)");
			in_cap[i] = in_bytes;
		}
		if (out_bytes > out_cap[i]) {
			if (d_out[i]) cudaFree(d_out[i]);
			d_out[i] = nullptr;
			out_cap[i] = 0;
			B200IC_CUDA(cudaMalloc(&d_out[i], out_bytes), "cudaMalloc(output chunk)");
			out_cap[i] = out_bytes;
		}
		return 0;
	}
	int reserve_pinned(int i, size_t in_bytes, size_t out_bytes) {
		if (in_bytes > pin_cap[i]) {
			if (p_in[i]) cudaFreeHost(p_in[i]);
			p_in[i] = nullptr;
			pin_cap[i] = 0;
			B200IC_CUDA(cudaHostAlloc(&p_in[i], in_bytes, cudaHostAllocDefault), "cudaHostAlloc(input staging)");
			pin_cap[i] = in_bytes;
		}
		if (out_bytes > pout_cap[i]) {
			if (p_out[i]) cudaFreeHost(p_out[i]);
			p_out[i] = nullptr;
			pout_cap[i] = 0;
			B200IC_CUDA(cudaHostAlloc(&p_out[i], out_bytes, cudaHostAllocDefault), "cudaHostAlloc(output staging)");
			pout_cap[i] = out_bytes;
		}
		return 0;
	}
	void release() { // call with `mu` held
		if (device < 0) return;
		int prev = -1;
		cudaGetDevice(&prev);
		cudaSetDevice(device);
		for (int i = 0; i < kStreams; i++) {
			if (streams[i]) cudaStreamDestroy(streams[i]);
			if (d_in[i]) cudaFree(d_in[i]);
			if (d_out[i]) cudaFree(d_out[i]);
			if (p_in[i]) cudaFreeHost(p_in[i]);
			if (p_out[i]) cudaFreeHost(p_out[i]);
			if (join[i]) cudaEventDestroy(join[i]);
			if (kdone[i]) cudaEventDestroy(kdone[i]);
			join[i] = kdone[i] = nullptr;
			streams[i] = nullptr;
			d_in[i] = d_out[i] = p_in[i] = p_out[i] = nullptr;
			in_cap[i] = out_cap[i] = pin_cap[i] = pout_cap[i] = 0;
		}
		if (fork) cudaEventDestroy(fork);
		fork = nullptr;
		ready = false;
		device = -1;
		if (prev >= 0) cudaSetDevice(prev);
	}
};
constexpr int kMaxDevices = 16;
static DevCtx g_ctx[kMaxDevices];
static std::mutex g_init_mutex;
static int g_tables_device_mask_lo = 0; // bit per device whose __constant__/global tables were uploaded

static int ensure_device() {
	int dev = -1;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return fail("no usable CUDA device (this library has no CPU fallback)", e);
	{
		std::lock_guard<std::mutex> lock(g_init_mutex);
		if (!(g_tables_device_mask_lo & (1 << dev))) {
			cudaDeviceProp prop;
			B200IC_CUDA(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
			if (prop.major < 10) return fail("device is not sm_100-class; kernels are built for sm_100a only");
			{ // the per-encode scratch of the AMD BC7 pipeline comes from the stream-ordered pool: keep it cached between encodes
				cudaMemPool_t pool;
				if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
					uint64_t keep = 2ull << 30;
					cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
				}
			}
			B200IC_CUDA(init_bc7rg_tables(), "bc7enc16 table upload");
			B200IC_CUDA(init_bc7amd_tables(), "BC7 table upload");
			B200IC_CUDA(init_bc6h_tables(), "BC6H table upload");
			g_tables_device_mask_lo |= (1 << dev);
		}
	}
	return 0;
}

void count_launches(int extra) { g_launches.fetch_add((uint64_t) extra, std::memory_order_relaxed); }

static int dispatch(int codec, const SrcImage &img, const b200ic_opts &o, void *d_dst, cudaStream_t stream) {
	cudaError_t e;
	switch (codec) {
	case B200IC_BC4: {
		int ch = o.bc4_channel;
		if (img.format == B200IC_FMT_BLOCKS_F32X1) ch = 0;
		e = launch_bc45(img, 1, ch, d_dst, stream);
		break;
	}
	case B200IC_BC5: e = launch_bc45(img, 2, 0, d_dst, stream); break;
	case B200IC_BC1:
		if (o.amd_adaptive_weights) return fail("BC1: AdaptiveColourWeights is not supported (it reads uninitialised memory in the reference)");
		e = launch_bc1(img, o, d_dst, stream);
		break;
	case B200IC_BC2:
	case B200IC_BC3:
	case B200IC_BC23_COLOUR_HALF:
	case B200IC_BC2_ALPHA_HALF:
		if (o.amd_adaptive_weights) return fail("BC2/BC3: AdaptiveColourWeights is not supported (it reads uninitialised memory in the reference)");
		e = launch_bc23(img, o, codec == B200IC_BC3 ? kBc3Colour : (codec == B200IC_BC2 ? kBc2Both : (codec == B200IC_BC23_COLOUR_HALF ? kColourOnly : kAlphaOnly)),
										d_dst, stream);
		break;
	case B200IC_BC7_RG: e = launch_bc7rg(img, o, d_dst, stream); break;
	case B200IC_BC7_AMD: e = launch_bc7amd(img, o, d_dst, stream); break;
	case B200IC_BC6H: e = launch_bc6h(img, o, d_dst, stream); break;
	default: return fail("unsupported codec");
	}
	if (e != cudaSuccess) return fail("kernel launch", e);
	g_launches.fetch_add(1, std::memory_order_relaxed);
	return 0;
}

} // namespace b200ic

using namespace b200ic;

extern "C" {

void b200ic_default_opts(b200ic_opts *o) {
	if (!o) return;
	memset(o, 0, sizeof(*o));
	o->bc1_alpha_threshold = 128 / 255.0f; // src/amd_bc1_compressor.cpp:21-27,57
	o->amd_refinement_steps = 1;           // src/amd_bcx_helpers.cpp:23-31
	o->amd_mode_mask = 0xFF;
	o->src_has_alpha = 1;
	o->rg_perceptual = 1;                  // src/richgel999_bc7enc16.cpp:13-19
	o->rg_fast = 0;
	o->bc4_channel = 1;                    // src/amd_bc4_compressor.cpp:35
}

int b200ic_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int b200ic_codec_available(int codec) {
	switch (codec) {
	case B200IC_BC4:
	case B200IC_BC5: return 1;
	case B200IC_BC1:
	case B200IC_BC2:
	case B200IC_BC3:
	case B200IC_BC23_COLOUR_HALF:
	case B200IC_BC2_ALPHA_HALF: return 1;
	case B200IC_BC7_RG: return 1;
	case B200IC_BC7_AMD: return 1;
	case B200IC_BC6H: return 1;
	default: return 0;
	}
}

int b200ic_init(int device) {
	t_error.clear();
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) return fail("no CUDA device found (this library has no CPU fallback)", e);
	if (device < 0 || device >= n) return fail("device index out of range");
	B200IC_CUDA(cudaSetDevice(device), "cudaSetDevice");
	return ensure_device();
}

void b200ic_shutdown(void) {
	for (int d = 0; d < kMaxDevices; d++) {
		std::lock_guard<std::mutex> lock(g_ctx[d].mu);
		g_ctx[d].release();
	}
}

const char *b200ic_last_error(void) { return t_error.c_str(); }

uint32_t b200ic_block_bytes(int codec) { return block_bytes(codec); }
uint32_t b200ic_texel_bytes(int format) { return texel_bytes(format); }
uint64_t b200ic_output_bytes(int codec, uint32_t w, uint32_t h, uint32_t slices) {
	return (uint64_t) ((w + 3) / 4) * ((h + 3) / 4) * slices * block_bytes(codec);
}
uint64_t b200ic_launch_count(void) { return g_launches.load(); }

int b200ic_encode_device(int codec, const void *d_src, int format, uint32_t width, uint32_t height,
												 uint64_t row_pitch_bytes, uint64_t slice_pitch_bytes, uint32_t slices, const b200ic_opts *opts,
												 void *d_dst, void *stream) {
	t_error.clear();
	if (!d_src || !d_dst) return fail("null buffer");
	if (width == 0 || height == 0 || slices == 0) return fail("empty image");
	if (block_bytes(codec) == 0) return fail("unsupported codec");
	if (ensure_device()) return -1;
	b200ic_opts o;
	if (opts) o = *opts;
	else b200ic_default_opts(&o);
	SrcImage img;
	img.base = static_cast<const uint8_t *>(d_src);
	img.format = format;
	img.width = width;
	img.height = height;
	img.slices = slices;
	if (format >= 100) { // pre-gathered blocks: `width` = number of blocks
		if (blocks_format_bytes(format) == 0) return fail("unsupported block format");
		img.blocks_x = width;
		img.blocks_y = 1;
		img.width = width * 4;
		img.height = 4;
		img.slices = 1;
		img.row_pitch = img.slice_pitch = 0;
	} else {
		const uint32_t tb = texel_bytes(format);
		if (tb == 0) return fail("unsupported source format");
		img.blocks_x = (width + 3) / 4;
		img.blocks_y = (height + 3) / 4;
		img.row_pitch = row_pitch_bytes ? row_pitch_bytes : (uint64_t) width * tb;
		img.slice_pitch = slice_pitch_bytes ? slice_pitch_bytes : img.row_pitch * height;
		if (img.row_pitch < (uint64_t) width * tb) return fail("row pitch smaller than a row");
		// the gather uses 32- / 64- / 128-bit loads for 4- / 8- / 16-byte texels: a misaligned pointer would raise a sticky
		// misaligned-address fault that poisons the CUDA context instead of an error return
		if ((tb == 4 || tb == 8 || tb == 16) && ((img.row_pitch % tb) || (img.slice_pitch % tb) || ((uintptr_t) d_src % tb)))
			return fail("source rows and slices must be aligned to the texel size");
	}
	if (format == B200IC_FMT_BLOCKS_RGBA8 && ((uintptr_t) d_src % 16)) return fail("RGBA8 block sources must be 16-byte aligned");
	if (format >= 100 && format != B200IC_FMT_BLOCKS_RGBA8 && ((uintptr_t) d_src % 4)) return fail("float block sources must be 4-byte aligned");
	if ((uintptr_t) d_dst % block_bytes(codec)) return fail("destination must be aligned to the block size");
	if (codec == B200IC_BC6H) o.bc6h_signed = (format == B200IC_FMT_RGBA16F || format == B200IC_FMT_RGBA32F) ? 1 : o.bc6h_signed;
	return dispatch(codec, img, o, d_dst, static_cast<cudaStream_t>(stream));
}

// ---- host-buffer path --------------------------------------------------------------------------------------------
// Progress is reported per finished block-row with the reference's own expression (e.g. src/amd_bc7_compressor.cpp:
// 71-75: 100 * (y * blocksX) / (blocksX * blocksY) after row y of every slice), in order, from the thread that called
// the encode when one device is used and serialised by a mutex when several are.
struct HostJob {
	int codec, format;
	const uint8_t *src;
	uint8_t *dst;
	uint32_t width, height, slices, blocks_x, blocks_y, tb, bb;
	uint64_t pitch;
	const b200ic_opts *opts;
	b200ic_progress_fn progress;
	void *user;
	bool src_pinned, dst_pinned;
	std::atomic<int> cancelled{0};
	std::mutex progress_mu;
	std::string error; // first error of a worker thread
	std::mutex error_mu;
};

// Staging copies between pageable caller memory and the pinned ring: one core moves ~10 GB/s, less than PCIe 5 and far
// less than the fast codecs consume, so copies of a megabyte or more are cut over a few short-lived threads.
static void staging_copy(void *dst, const void *src, size_t bytes) {
	constexpr size_t kMin = 2u << 20; // (a thread costs ~50 us to start and join)
	unsigned hw = std::thread::hardware_concurrency();
	unsigned parts = (unsigned) std::min<size_t>(bytes / kMin, std::min(4u, hw ? hw : 1u));
	if (parts <= 1) {
		memcpy(dst, src, bytes);
		return;
	}
	const size_t each = ((bytes / parts) + 4095) & ~(size_t) 4095;
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < parts; t++) {
		const size_t off = (size_t) t * each;
		if (off >= bytes) break;
		pool.emplace_back([=]() { memcpy(static_cast<uint8_t *>(dst) + off, static_cast<const uint8_t *>(src) + off, std::min(each, bytes - off)); });
	}
	memcpy(dst, src, std::min(each, bytes));
	for (auto &t : pool) t.join();
}

static bool is_pinned(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static void report_rows(HostJob &job, uint64_t g0, uint64_t g1) { // global rows [g0, g1): slice * blocks_y + block-row
	if (!job.progress) return;
	std::lock_guard<std::mutex> lock(job.progress_mu);
	for (uint64_t g = g0; g < g1 && !job.cancelled.load(); g++) {
		const size_t y = (size_t) (g % job.blocks_y);
		const float pct = 100.f * (y * (size_t) job.blocks_x) / ((size_t) job.blocks_x * job.blocks_y);
		if (job.progress(job.user, pct)) job.cancelled.store(1);
	}
}

// Global rows [g0, g1) of the job on the CURRENT device through its context: chunks of whole block-rows (never across a
// slice) pipelined over the stream ring -- stage (pageable sources: CPU copy into pinned memory), H2D, kernels, D2H,
// un-stage.  The CPU copies of chunk c run while the GPU works on chunks c-1, c-2.
static int encode_rows_on_device(DevCtx &cx, HostJob &job, uint64_t g0, uint64_t g1) {
	if (g0 >= g1) return 0;
	// ~8 MiB of input per chunk for the HBM-speed codecs; AMD BC7 spends ~100 ms on 32 MiB, so bigger chunks cost nothing
	// and keep its per-mode launches long
	const uint64_t chunk_bytes = job.codec == B200IC_BC7_AMD ? (32ull << 20) : (8ull << 20);
	uint32_t rows_per_chunk = (uint32_t) std::max<uint64_t>(1, chunk_bytes / (job.pitch * 4));
	rows_per_chunk = std::min(rows_per_chunk, job.blocks_y);
	const size_t in_cap = (size_t) rows_per_chunk * 4 * job.pitch;
	const size_t out_cap = (size_t) rows_per_chunk * job.blocks_x * job.bb;
	constexpr int ring = DevCtx::kStreams;
	struct Chunk {
		uint64_t g0, g1;
		uint8_t *hd;
		size_t out_bytes;
	};
	std::vector<Chunk> chunks;
	for (uint64_t g = g0; g < g1;) {
		const uint64_t slice_end = (g / job.blocks_y + 1) * job.blocks_y;
		const uint64_t e = std::min<uint64_t>(std::min<uint64_t>(g + rows_per_chunk, slice_end), g1);
		chunks.push_back({g, e, nullptr, 0});
		g = e;
	}
	for (int i = 0; i < ring && (size_t) i < chunks.size(); i++) {
		if (cx.reserve(i, in_cap, out_cap)) return -1;
		if (cx.reserve_pinned(i, job.src_pinned ? 0 : in_cap, job.dst_pinned ? 0 : out_cap)) return -1;
	}
	int rc = 0;
	const size_t total = chunks.size();
	for (size_t c = 0; c < total + ring && rc == 0; c++) {
		if (c >= (size_t) ring) { // retire chunk c - ring before its slot is reused
			Chunk &d = chunks[c - ring];
			const int s = (int) ((c - ring) % ring);
			cudaError_t e = cudaStreamSynchronize(cx.streams[s]);
			if (e != cudaSuccess) { rc = fail("encode chunk", e); break; }
			if (d.out_bytes) {
				if (!job.dst_pinned) staging_copy(d.hd, cx.p_out[s], d.out_bytes);
				report_rows(job, d.g0, d.g1);
			}
		}
		if (c >= total || job.cancelled.load()) continue;
		Chunk &k = chunks[c];
		const int s = (int) (c % ring);
		const uint32_t slice = (uint32_t) (k.g0 / job.blocks_y);
		const uint32_t by0 = (uint32_t) (k.g0 % job.blocks_y), by1 = by0 + (uint32_t) (k.g1 - k.g0);
		const uint32_t y0 = by0 * 4, y1 = std::min(by1 * 4, job.height);
		cudaStream_t st = cx.streams[s];
		const uint8_t *hs = job.src + ((uint64_t) slice * job.height + y0) * job.pitch;
		const size_t in_bytes = (size_t) (y1 - y0) * job.pitch;
		if (!job.src_pinned) {
			staging_copy(cx.p_in[s], hs, in_bytes);
			hs = static_cast<const uint8_t *>(cx.p_in[s]);
		}
		cudaError_t e = cudaMemcpyAsync(cx.d_in[s], hs, in_bytes, cudaMemcpyHostToDevice, st);
		if (e != cudaSuccess) { rc = fail("H2D copy", e); break; }
		if (job.codec == B200IC_BC7_AMD && c > 0) {
			// one mode's kernels at a time on the SMs: chunk c's kernels start after chunk c-1's (the copies still overlap)
			e = cudaStreamWaitEvent(st, cx.kdone[(c - 1) % ring], 0);
			if (e != cudaSuccess) { rc = fail("stream wait", e); break; }
		}
		if (b200ic_encode_device(job.codec, cx.d_in[s], job.format, job.width, y1 - y0, job.pitch, 0, 1, job.opts, cx.d_out[s], st)) { rc = -1; break; }
		if (job.codec == B200IC_BC7_AMD) cudaEventRecord(cx.kdone[s], st);
		k.hd = job.dst + ((uint64_t) slice * job.blocks_y + by0) * job.blocks_x * job.bb;
		k.out_bytes = (size_t) (by1 - by0) * job.blocks_x * job.bb;
		e = cudaMemcpyAsync(job.dst_pinned ? (void *) k.hd : cx.p_out[s], cx.d_out[s], k.out_bytes, cudaMemcpyDeviceToHost, st);
		if (e != cudaSuccess) { rc = fail("D2H copy", e); break; }
	}
	if (rc != 0 || job.cancelled.load())
		for (int i = 0; i < ring; i++) cudaStreamSynchronize(cx.streams[i]);
	return rc;
}

static int encode_host_impl(int codec, const void *h_src, int format, uint32_t width, uint32_t height, uint64_t row_pitch_bytes,
														uint32_t slices, const b200ic_opts *opts, void *h_dst, b200ic_progress_fn progress, void *user, int n_devices) {
	t_error.clear();
	if (!h_src || !h_dst) return fail("null buffer");
	if (width == 0 || height == 0 || slices == 0) return fail("empty image");
	HostJob job;
	job.codec = codec;
	job.format = format;
	job.tb = texel_bytes(format);
	job.bb = block_bytes(codec);
	if (job.tb == 0) return fail("unsupported source format");
	if (job.bb == 0) return fail("unsupported codec");
	if (ensure_device()) return -1;
	int dev = 0;
	cudaGetDevice(&dev);
	job.src = static_cast<const uint8_t *>(h_src);
	job.dst = static_cast<uint8_t *>(h_dst);
	job.width = width;
	job.height = height;
	job.slices = slices;
	job.pitch = row_pitch_bytes ? row_pitch_bytes : (uint64_t) width * job.tb;
	job.blocks_x = (width + 3) / 4;
	job.blocks_y = (height + 3) / 4;
	job.opts = opts;
	job.progress = progress;
	job.user = user;
	job.src_pinned = is_pinned(h_src);
	job.dst_pinned = is_pinned(h_dst);
	const uint64_t rows = (uint64_t) job.blocks_y * slices;
	int nd = n_devices;
	const int visible = b200ic_device_count();
	if (nd > visible) nd = visible;
	if (nd > kMaxDevices) nd = kMaxDevices;
	if ((uint64_t) nd > rows) nd = (int) rows;
	if (nd <= 1) {
		if (dev < 0 || dev >= kMaxDevices) return fail("device index out of range");
		DevCtx &cx = g_ctx[dev];
		std::lock_guard<std::mutex> lock(cx.mu);
		if (cx.ensure(dev)) return -1;
		const int rc = encode_rows_on_device(cx, job, 0, rows);
		if (rc != 0) return rc;
		return job.cancelled.load() ? 1 : 0;
	}
	// block-row shards over devices 0 .. nd-1, one host thread per device (src/amd_bc7_compressor.cpp:48-77 is the loop
	// being split); no collective: every shard lands in its own range of h_dst
	std::atomic<int> failed{0};
	auto work = [&](int d) {
		if (cudaSetDevice(d) != cudaSuccess || ensure_device()) {
			std::lock_guard<std::mutex> lock(job.error_mu);
			if (job.error.empty()) job.error = t_error.empty() ? "b200ic: cudaSetDevice failed" : t_error;
			failed.store(1);
			return;
		}
		DevCtx &cx = g_ctx[d];
		std::lock_guard<std::mutex> lock(cx.mu);
		const uint64_t a = rows * (uint64_t) d / nd, b = rows * (uint64_t) (d + 1) / nd;
		if (cx.ensure(d) || encode_rows_on_device(cx, job, a, b)) {
			std::lock_guard<std::mutex> lock2(job.error_mu);
			if (job.error.empty()) job.error = t_error;
			failed.store(1);
		}
	};
	std::vector<std::thread> pool;
	for (int d = 0; d < nd; d++) pool.emplace_back(work, d);
	for (auto &t : pool) t.join();
	cudaSetDevice(dev);
	if (failed.load()) {
		t_error = job.error;
		return -1;
	}
	return job.cancelled.load() ? 1 : 0;
}

int b200ic_encode_host(int codec, const void *h_src, int format, uint32_t width, uint32_t height, uint64_t row_pitch_bytes,
											 uint32_t slices, const b200ic_opts *opts, void *h_dst, b200ic_progress_fn progress, void *user) {
	return encode_host_impl(codec, h_src, format, width, height, row_pitch_bytes, slices, opts, h_dst, progress, user, 1);
}

int b200ic_encode_host_sharded(int codec, const void *h_src, int format, uint32_t width, uint32_t height, uint64_t row_pitch_bytes,
															 uint32_t slices, const b200ic_opts *opts, void *h_dst, b200ic_progress_fn progress, void *user, int n_devices) {
	return encode_host_impl(codec, h_src, format, width, height, row_pitch_bytes, slices, opts, h_dst, progress, user,
													n_devices <= 0 ? b200ic_device_count() : n_devices);
}

static std::atomic<int> g_shim_devices{-1}; // -1: not set (environment / all visible)
void b200ic_set_devices(int n) { g_shim_devices.store(n < 0 ? 0 : n); }
int b200ic_get_devices(void) {
	int n = g_shim_devices.load();
	if (n < 0) {
		const char *e = getenv("B200IC_DEVICES"); // read once
		n = e ? atoi(e) : 0;
		if (n < 0) n = 0;
		g_shim_devices.store(n);
	}
	const int visible = b200ic_device_count();
	if (n == 0 || n > visible) n = visible;
	return n < 1 ? 1 : n;
}

int b200ic_encode_blocks(int codec, const void *h_blocks, int format, uint64_t nblocks, const b200ic_opts *opts, void *h_dst) {
	t_error.clear();
	if (!h_blocks || !h_dst) return fail("null buffer");
	const uint32_t fb = blocks_format_bytes(format), bb = block_bytes(codec);
	if (fb == 0) return fail("unsupported block format");
	if (bb == 0) return fail("unsupported codec");
	if (nblocks == 0) return 0;
	if (nblocks > 0x3fffffffu) return fail("too many blocks in one call");
	if (ensure_device()) return -1;
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= kMaxDevices) return fail("device index out of range");
	DevCtx &cx = g_ctx[dev];
	std::lock_guard<std::mutex> lock(cx.mu);
	if (cx.ensure(dev)) return -1;
	if (cx.reserve(0, (size_t) nblocks * fb, (size_t) nblocks * bb)) return -1;
	cudaStream_t st = cx.streams[0];
	B200IC_CUDA(cudaMemcpyAsync(cx.d_in[0], h_blocks, (size_t) nblocks * fb, cudaMemcpyHostToDevice, st), "H2D copy");
	if (b200ic_encode_device(codec, cx.d_in[0], format, (uint32_t) nblocks, 1, 0, 0, 1, opts, cx.d_out[0], st)) return -1;
	B200IC_CUDA(cudaMemcpyAsync(h_dst, cx.d_out[0], (size_t) nblocks * bb, cudaMemcpyDeviceToHost, st), "D2H copy");
	B200IC_CUDA(cudaStreamSynchronize(st), "encode blocks");
	return 0;
}

int b200ic_box_mip_rgba8_device(const void *d_src, uint32_t width, uint32_t height, uint64_t src_pitch_bytes, void *d_dst, uint64_t dst_pitch_bytes,
																void *stream) {
	t_error.clear();
	if (!d_src || !d_dst) return fail("null buffer");
	if (width == 0 || height == 0) return fail("empty image");
	if (((uintptr_t) d_src | (uintptr_t) d_dst | src_pitch_bytes | dst_pitch_bytes) & 3u) return fail("RGBA8 levels must be 4-byte aligned");
	if (ensure_device()) return -1;
	B200IC_CUDA(launch_box_mip_rgba8(d_src, width, height, src_pitch_bytes, d_dst, dst_pitch_bytes, static_cast<cudaStream_t>(stream)), "mip kernel launch");
	g_launches.fetch_add(1, std::memory_order_relaxed);
	return 0;
}

int b200ic_decode_device(int codec, const void *d_blocks, uint32_t width, uint32_t height, uint32_t slices, int is_signed, void *d_dst,
												 uint64_t dst_row_pitch_bytes, void *stream) {
	t_error.clear();
	if (!d_blocks || !d_dst) return fail("null buffer");
	if (width == 0 || height == 0 || slices == 0) return fail("empty image");
	const uint32_t bb = block_bytes(codec);
	if (bb == 0 || codec == B200IC_BC23_COLOUR_HALF || codec == B200IC_BC2_ALPHA_HALF) return fail("unsupported codec");
	const uint32_t tb = codec == B200IC_BC4 ? 1 : (codec == B200IC_BC5 ? 2 : (codec == B200IC_BC6H ? 8 : 4));
	if (((uintptr_t) d_blocks % bb) || (tb >= 4 && (((uintptr_t) d_dst | dst_row_pitch_bytes) % tb))) return fail("misaligned buffer");
	if (dst_row_pitch_bytes && dst_row_pitch_bytes < (uint64_t) width * tb) return fail("row pitch smaller than a row");
	if (ensure_device()) return -1;
	B200IC_CUDA(launch_decode(codec, d_blocks, width, height, slices, is_signed, d_dst, dst_row_pitch_bytes, static_cast<cudaStream_t>(stream)), "decode kernel launch");
	g_launches.fetch_add(1, std::memory_order_relaxed);
	return 0;
}

int b200ic_write_dds(const char *path, int codec, int srgb, int is_signed, uint32_t width, uint32_t height, uint32_t levels,
										 const void *const *level_blocks) {
	t_error.clear();
	const int rc = write_dds(path, codec, srgb, is_signed, width, height, levels, level_blocks);
	if (rc == -1) return fail("b200ic_write_dds: bad arguments");
	if (rc == -2) return fail("b200ic_write_dds: cannot open the file");
	if (rc != 0) return fail("b200ic_write_dds: short write");
	return 0;
}

uint64_t b200ic_plan_shards(const uint32_t *widths, const uint32_t *heights, uint64_t n_images, uint32_t chunk_rows, uint32_t world,
														uint32_t rank, b200ic_shard *out, uint64_t cap) {
	if (!widths || !heights || world == 0 || rank >= world) return 0;
	if (chunk_rows == 0) chunk_rows = 64;
	struct Chunk {
		uint64_t blocks;
		uint32_t image, row0, row1;
	};
	std::vector<Chunk> chunks;
	for (uint64_t i = 0; i < n_images; i++) {
		const uint32_t bx = (widths[i] + 3) / 4, by = (heights[i] + 3) / 4;
		for (uint32_t r = 0; r < by; r += chunk_rows) {
			const uint32_t r1 = r + chunk_rows < by ? r + chunk_rows : by;
			chunks.push_back({(uint64_t) bx * (r1 - r), (uint32_t) i, r, r1});
		}
	}
	// Contiguous partition of the chunk sequence (image order, then row order): a chunk goes to the rank in whose share
	// [r T / world, (r + 1) T / world) of the T blocks its midpoint falls.  Every rank gets one contiguous run (adjacent
	// chunks of an image merge into one shard below: fewer, larger launches than a round-robin deal) and the loads differ by
	// at most one chunk.  A pure function of the dimensions.
	uint64_t total = 0;
	for (const Chunk &c : chunks) total += c.blocks;
	std::vector<Chunk> mine;
	uint64_t before = 0;
	for (const Chunk &c : chunks) {
		// rank = floor((before + blocks / 2) * world / total), in 128-bit-safe form (total < 2^40, world < 2^16)
		const uint64_t mid2 = 2 * before + c.blocks; // twice the midpoint
		uint64_t r = total ? (uint64_t) (((unsigned __int128) mid2 * world) / (2 * (unsigned __int128) total)) : 0;
		if (r >= world) r = world - 1;
		if (r == rank) mine.push_back(c);
		before += c.blocks;
	}
	uint64_t n = 0;
	b200ic_shard cur = {0, 0, 0, 0};
	bool have = false;
	auto flush = [&]() {
		if (!have) return;
		if (out && n < cap) out[n] = cur;
		n++;
	};
	for (const Chunk &c : mine) {
		if (have && cur.image == c.image && cur.row1 == c.row0) {
			cur.row1 = c.row1;
			continue;
		}
		flush();
		cur = {c.image, c.row0, c.row1, 0};
		have = true;
	}
	flush();
	return n;
}

int b200ic_encode_batch_device(int codec, const b200ic_image_desc *images, uint64_t n_images, const b200ic_shard *shards,
															 uint64_t n_shards, const b200ic_opts *opts, void *stream) {
	t_error.clear();
	if (!images && n_images) return fail("null image table");
	const uint32_t bb = block_bytes(codec);
	if (bb == 0) return fail("unsupported codec");
	if (ensure_device()) return -1;
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= kMaxDevices) return fail("device index out of range");
	DevCtx &cx = g_ctx[dev];
	std::lock_guard<std::mutex> lock(cx.mu);
	if (cx.ensure(dev)) return -1;
	const uint64_t count = shards ? n_shards : n_images;
	if (count == 0) return 0;
	cudaStream_t user = static_cast<cudaStream_t>(stream);
	const int ring = DevCtx::kStreams;
	B200IC_CUDA(cudaEventRecord(cx.fork, user), "fork");
	for (int i = 0; i < ring; i++) B200IC_CUDA(cudaStreamWaitEvent(cx.streams[i], cx.fork, 0), "fork");
	int rc = 0;
	for (uint64_t k = 0; k < count && rc == 0; k++) {
		b200ic_shard sh;
		if (shards) sh = shards[k];
		else sh = {(uint32_t) k, 0, 0xffffffffu, 0};
		if (sh.image >= n_images) { rc = fail("shard names an image outside the table"); break; }
		const b200ic_image_desc &im = images[sh.image];
		const uint32_t tb = texel_bytes(im.format);
		if (tb == 0) { rc = fail("unsupported source format"); break; }
		if (im.width == 0 || im.height == 0) { rc = fail("empty image"); break; }
		const uint32_t blocks_x = (im.width + 3) / 4, blocks_y = (im.height + 3) / 4;
		const uint32_t r0 = sh.row0, r1 = sh.row1 < blocks_y ? sh.row1 : blocks_y;
		if (r0 >= r1) continue;
		const uint64_t pitch = im.row_pitch_bytes ? im.row_pitch_bytes : (uint64_t) im.width * tb;
		const uint32_t y0 = r0 * 4, y1 = r1 * 4 < im.height ? r1 * 4 : im.height;
		const uint8_t *src = static_cast<const uint8_t *>(im.src) + (uint64_t) y0 * pitch;
		uint8_t *dst = static_cast<uint8_t *>(im.dst) + (uint64_t) r0 * blocks_x * bb;
		// small shards (low mips) overlap on the ring; a shard that fills the GPU on its own keeps to one stream for AMD BC7,
		// whose per-mode kernels slow each other down when launches of different modes share the SMs
		const bool big = (uint64_t) blocks_x * (r1 - r0) >= 4096;
		const int si = (codec == B200IC_BC7_AMD && big) ? 0 : (int) (k % ring);
		rc = b200ic_encode_device(codec, src, im.format, im.width, y1 - y0, pitch, 0, 1, opts, dst, cx.streams[si]);
	}
	for (int i = 0; i < ring; i++) {
		cudaEventRecord(cx.join[i], cx.streams[i]);
		cudaStreamWaitEvent(user, cx.join[i], 0);
	}
	return rc;
}

} // extern "C"
