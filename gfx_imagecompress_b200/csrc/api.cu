// api.cu -- the b200ic_* C-ABI (include/b200ic.h): device selection, per-codec dispatch, and the
// host-buffer path (block-row chunks pipelined H2D -> kernel -> D2H over a small ring of streams).
// No CPU fallback exists: without a usable CUDA device every encode call fails loudly.
#include "kernels.h"
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
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

// ---- per-thread device context: stream ring + grow-only device scratch for the host path -------------
struct HostCtx {
	static constexpr int kStreams = 3;
	int device = -1;
	cudaStream_t streams[kStreams] = {};
	void *d_in[kStreams] = {};
	void *d_out[kStreams] = {};
	size_t in_cap[kStreams] = {};
	size_t out_cap[kStreams] = {};
	bool ready = false;
	cudaEvent_t fork = nullptr, join[kStreams] = {};

	int ensure(int dev) {
		if (ready && dev == device) return 0;
		release();
		device = dev;
		for (int i = 0; i < kStreams; i++) B200IC_CUDA(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking), "stream create");
		B200IC_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "event create");
		for (int i = 0; i < kStreams; i++) B200IC_CUDA(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming), "event create");
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
	void release() {
		if (!ready) return;
		for (int i = 0; i < kStreams; i++) {
			if (streams[i]) cudaStreamDestroy(streams[i]);
			if (d_in[i]) cudaFree(d_in[i]);
			if (d_out[i]) cudaFree(d_out[i]);
			if (join[i]) cudaEventDestroy(join[i]);
			join[i] = nullptr;
			streams[i] = nullptr;
			d_in[i] = d_out[i] = nullptr;
			in_cap[i] = out_cap[i] = 0;
		}
		if (fork) cudaEventDestroy(fork);
		fork = nullptr;
		ready = false;
	}
	~HostCtx() { /* process teardown: the driver reclaims everything */ }
};
static thread_local HostCtx t_ctx;
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
		if (o.amd_3d_refinement || o.amd_adaptive_weights) return fail("BC1: b3DRefinement / AdaptiveColourWeights are not supported");
		e = launch_bc1(img, o, d_dst, stream);
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
#ifdef B200IC_HAVE_BC1
	case B200IC_BC1: return 1;
#endif
#ifdef B200IC_HAVE_BC7RG
	case B200IC_BC7_RG: return 1;
#endif
#ifdef B200IC_HAVE_BC7AMD
	case B200IC_BC7_AMD: return 1;
#endif
#ifdef B200IC_HAVE_BC6H
	case B200IC_BC6H: return 1;
#endif
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

void b200ic_shutdown(void) { t_ctx.release(); }

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
		if ((tb == 4 || tb == 8 || tb == 16) && ((img.row_pitch % tb) || ((uintptr_t) d_src % tb)))
			return fail("source rows must be aligned to the texel size");
	}
	if (codec == B200IC_BC6H) o.bc6h_signed = (format == B200IC_FMT_RGBA16F || format == B200IC_FMT_RGBA32F) ? 1 : o.bc6h_signed;
	return dispatch(codec, img, o, d_dst, static_cast<cudaStream_t>(stream));
}

int b200ic_encode_host(int codec, const void *h_src, int format, uint32_t width, uint32_t height, uint64_t row_pitch_bytes,
											 uint32_t slices, const b200ic_opts *opts, void *h_dst, b200ic_progress_fn progress, void *user) {
	t_error.clear();
	if (!h_src || !h_dst) return fail("null buffer");
	if (width == 0 || height == 0 || slices == 0) return fail("empty image");
	const uint32_t tb = texel_bytes(format), bb = block_bytes(codec);
	if (tb == 0) return fail("unsupported source format");
	if (bb == 0) return fail("unsupported codec");
	if (ensure_device()) return -1;
	int dev = 0;
	cudaGetDevice(&dev);
	if (t_ctx.ensure(dev)) return -1;

	const uint64_t pitch = row_pitch_bytes ? row_pitch_bytes : (uint64_t) width * tb;
	const uint32_t blocks_x = (width + 3) / 4, blocks_y = (height + 3) / 4;
	// chunk = whole block-rows, ~16 MiB of input each so copies overlap the kernels of neighbouring chunks.  AMD BC7
	// spends seconds per 256 MiB (the copies are noise) and runs one launch per mode: big chunks keep launches of
	// DIFFERENT modes from sharing the SMs (they evict each other's code, see launch_bc7amd)
	const uint64_t chunk_bytes = codec == B200IC_BC7_AMD ? (256ull << 20) : (16ull << 20);
	uint32_t rows_per_chunk = (uint32_t) (chunk_bytes / (pitch * 4));
	if (rows_per_chunk < 1) rows_per_chunk = 1;
	if (rows_per_chunk > blocks_y) rows_per_chunk = blocks_y;
	const uint32_t chunks_per_slice = (blocks_y + rows_per_chunk - 1) / rows_per_chunk;
	const uint64_t total_chunks = (uint64_t) chunks_per_slice * slices;
	const size_t in_cap = (size_t) rows_per_chunk * 4 * pitch;
	const size_t out_cap = (size_t) rows_per_chunk * blocks_x * bb;
	const int ring = HostCtx::kStreams;
	for (int i = 0; i < ring && (uint64_t) i < total_chunks; i++)
		if (t_ctx.reserve(i, in_cap, out_cap)) return -1;

	const uint8_t *src = static_cast<const uint8_t *>(h_src);
	uint8_t *dst = static_cast<uint8_t *>(h_dst);
	int rc = 0;
	bool cancelled = false;
	for (uint64_t c = 0; c < total_chunks + ring && rc == 0; c++) {
		// retire chunk c-ring before its slot is reused (also drives the progress callback in order)
		if (c >= (uint64_t) ring) {
			const uint64_t done = c - ring;
			cudaError_t e = cudaStreamSynchronize(t_ctx.streams[done % ring]);
			if (e != cudaSuccess) { rc = fail("encode chunk", e); break; }
			if (progress && !cancelled) {
				const uint32_t cy = (uint32_t) (done % chunks_per_slice);
				const float pct = 100.f * ((float) (done / chunks_per_slice) * blocks_y + (float) cy * rows_per_chunk) /
													((float) blocks_y * slices);
				if (progress(user, pct)) cancelled = true;
			}
		}
		if (c >= total_chunks || cancelled) continue;
		const int s = (int) (c % ring);
		const uint32_t slice = (uint32_t) (c / chunks_per_slice);
		const uint32_t by0 = (uint32_t) (c % chunks_per_slice) * rows_per_chunk;
		const uint32_t by1 = by0 + rows_per_chunk < blocks_y ? by0 + rows_per_chunk : blocks_y;
		const uint32_t y0 = by0 * 4, y1 = by1 * 4 < height ? by1 * 4 : height;
		cudaStream_t st = t_ctx.streams[s];
		const uint8_t *hs = src + ((uint64_t) slice * height + y0) * pitch;
		cudaError_t e = cudaMemcpyAsync(t_ctx.d_in[s], hs, (size_t) (y1 - y0) * pitch, cudaMemcpyHostToDevice, st);
		if (e != cudaSuccess) { rc = fail("H2D copy", e); break; }
		if (b200ic_encode_device(codec, t_ctx.d_in[s], format, width, y1 - y0, pitch, 0, 1, opts, t_ctx.d_out[s], st)) { rc = -1; break; }
		uint8_t *hd = dst + ((uint64_t) slice * blocks_y + by0) * blocks_x * bb;
		e = cudaMemcpyAsync(hd, t_ctx.d_out[s], (size_t) (by1 - by0) * blocks_x * bb, cudaMemcpyDeviceToHost, st);
		if (e != cudaSuccess) { rc = fail("D2H copy", e); break; }
	}
	if (rc != 0 || cancelled) {
		for (int i = 0; i < ring; i++) cudaStreamSynchronize(t_ctx.streams[i]);
		return rc != 0 ? rc : 1;
	}
	return 0;
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
	if (t_ctx.ensure(dev)) return -1;
	if (t_ctx.reserve(0, (size_t) nblocks * fb, (size_t) nblocks * bb)) return -1;
	cudaStream_t st = t_ctx.streams[0];
	B200IC_CUDA(cudaMemcpyAsync(t_ctx.d_in[0], h_blocks, (size_t) nblocks * fb, cudaMemcpyHostToDevice, st), "H2D copy");
	if (b200ic_encode_device(codec, t_ctx.d_in[0], format, (uint32_t) nblocks, 1, 0, 0, 1, opts, t_ctx.d_out[0], st)) return -1;
	B200IC_CUDA(cudaMemcpyAsync(h_dst, t_ctx.d_out[0], (size_t) nblocks * bb, cudaMemcpyDeviceToHost, st), "D2H copy");
	B200IC_CUDA(cudaStreamSynchronize(st), "encode blocks");
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
	// largest first; equal sizes keep (image, row) order so that the plan is a pure function of the dimensions
	std::stable_sort(chunks.begin(), chunks.end(), [](const Chunk &a, const Chunk &b) { return a.blocks > b.blocks; });
	std::vector<uint64_t> load(world, 0);
	std::vector<Chunk> mine;
	for (const Chunk &c : chunks) {
		uint32_t best = 0;
		for (uint32_t r = 1; r < world; r++)
			if (load[r] < load[best]) best = r;
		load[best] += c.blocks;
		if (best == rank) mine.push_back(c);
	}
	std::sort(mine.begin(), mine.end(), [](const Chunk &a, const Chunk &b) { return a.image != b.image ? a.image < b.image : a.row0 < b.row0; });
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
	if (t_ctx.ensure(dev)) return -1;
	const uint64_t count = shards ? n_shards : n_images;
	if (count == 0) return 0;
	cudaStream_t user = static_cast<cudaStream_t>(stream);
	const int ring = HostCtx::kStreams;
	B200IC_CUDA(cudaEventRecord(t_ctx.fork, user), "fork");
	for (int i = 0; i < ring; i++) B200IC_CUDA(cudaStreamWaitEvent(t_ctx.streams[i], t_ctx.fork, 0), "fork");
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
		// big shards fill the GPU on their own and go to the caller's order on stream 0; small ones (low mips) overlap
		rc = b200ic_encode_device(codec, src, im.format, im.width, y1 - y0, pitch, 0, 1, opts, dst, t_ctx.streams[k % ring]);
	}
	for (int i = 0; i < ring; i++) {
		cudaEventRecord(t_ctx.join[i], t_ctx.streams[i]);
		cudaStreamWaitEvent(user, t_ctx.join[i], 0);
	}
	return rc;
}

} // extern "C"
