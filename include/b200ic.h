/* b200ic.h -- thin C-ABI of the B200 (sm_100a) BCn block-compression engine.
 *
 * This is the device-level layer underneath the drop-in `Image_Compress*` API of
 * include/gfx_imagecompress/imagecompress.h.  Plain pointers and sizes only: a host language binds
 * it with cgo / JNI / ctypes without seeing a CUDA or torch type.
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *   b200ic_encode_device / b200ic_encode_host
 *       the per-image triple loop `slices x blocksY x blocksX` of
 *       src/amd_bc1_compressor.cpp:44-71, src/amd_bc4_compressor.cpp:27-50, src/amd_bc5_compressor.cpp:27-54,
 *       src/amd_bc6h_compressor.cpp:34-57, src/amd_bc7_compressor.cpp:48-77, src/richgel999_bc7enc16.cpp:41-70
 *       together with the texel gather / block store of src/block_utils.cpp:7-41,116-160.
 *   b200ic_encode_blocks
 *       the "lowest level block API" of include/gfx_imagecompress/imagecompress.h:111-142
 *       (Image_CompressAMDBC1Block, ...AlphaSingleModeBlock, ...MultiModeLDRBlock, Image_CompressRichGel999BC7enc16)
 *       and BC6HBlockEncoder::CompressBlock (src/amd_bc6h_body.hpp:314), batched.
 *   b200ic_init / b200ic_shutdown
 *       Image_CompressInit / Image_CompressDeinit (src/imagecompress.cpp:11-18).
 *
 * There is NO CPU fallback: every encode entry point fails (non-zero, see b200ic_last_error) when no
 * CUDA device is usable.
 */
#ifndef B200IC_H_
#define B200IC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200IC_API __attribute__((visibility("default")))

typedef enum b200ic_codec {
	B200IC_BC1 = 1,     /* AMD BC1, 8 B/block  (Image_CompressAMDBC1) */
	B200IC_BC2 = 2,     /* AMD BC2, 16 B/block (Image_CompressAMDBC2: 4-bit explicit alpha + colour) */
	B200IC_BC3 = 3,     /* AMD BC3, 16 B/block (Image_CompressAMDBC3: BC4 search on alpha + colour) */
	B200IC_BC4 = 4,     /* AMD BC4, 8 B/block  (Image_CompressAMDBC4; encodes channel 1 like the reference) */
	B200IC_BC5 = 5,     /* AMD BC5, 16 B/block (Image_CompressAMDBC5; channels 0,1) */
	B200IC_BC6H = 6,    /* AMD BC6H, 16 B/block */
	B200IC_BC7_AMD = 7, /* AMD BC7, 16 B/block (Image_CompressAMDBC7) */
	B200IC_BC7_RG = 8,  /* bc7enc16, 16 B/block (Image_CompressRichGel999BC7) */
	/* block API only: the two 8-byte halves of a BC2 / BC3 block on their own */
	B200IC_BC23_COLOUR_HALF = 9, /* Image_CompressAMDRGBSingleModeBlock (F32X3 / F32X4 blocks) */
	B200IC_BC2_ALPHA_HALF = 10   /* Image_CompressAMDExplictAlphaSingleModeBlock (F32X1 blocks) */
} b200ic_codec;

/* Source texel formats.  Values equal the TinyImageFormat tags of compat/tiny_imageformat. */
typedef enum b200ic_format {
	B200IC_FMT_R8 = 1,
	B200IC_FMT_RG8 = 3,
	B200IC_FMT_RGB8 = 5,
	B200IC_FMT_RGB8_SRGB = 6,
	B200IC_FMT_RGBA8 = 7,
	B200IC_FMT_RGBA8_SRGB = 8,
	B200IC_FMT_RGBA16F = 9,   /* signed half  -> BC6H signed path  */
	B200IC_FMT_RGBA32F = 10,
	B200IC_FMT_RGBA16UF = 11, /* half >= 0    -> BC6H unsigned path */
	/* pre-gathered 4x4 blocks (block API): element i of block b at base[(b*16+i)*C] */
	B200IC_FMT_BLOCKS_F32X1 = 101,
	B200IC_FMT_BLOCKS_F32X3 = 103,
	B200IC_FMT_BLOCKS_F32X4 = 104,
	B200IC_FMT_BLOCKS_RGBA8 = 107
} b200ic_format;

/* Options crossing the boundary; zero-initialise then call b200ic_default_opts().  Mirrors
 * Image_CompressBC1Options / Image_CompressAMDBackendOptions / Image_CompressRichGel999BackendOptions
 * (include/gfx_imagecompress/imagecompress.h:35-50) plus the per-block arguments of :111-142. */
typedef struct b200ic_opts {
	float bc1_alpha_threshold;   /* 0..1; <=0 disables punch-through. default 128/255 (reference quirk: active by default) */
	int32_t amd_refinement_steps;/* default 1 */
	int32_t amd_3d_refinement;   /* default 0; non-zero: Refine3D for BC1 / BC2 / BC3 (src/amd_bcx_body.cpp:808-932) */
	int32_t amd_adaptive_weights;/* default 0 (non-zero unsupported: reads uninitialised memory in the reference) */
	int32_t amd_mode_mask;       /* default 0xFF (BC7 / BC6H) */
	int32_t src_has_alpha;       /* BC7: source has 4 channels (filled by the image API) */
	int32_t rg_perceptual;       /* default 1 */
	int32_t rg_fast;             /* default 0 */
	int32_t bc6h_signed;         /* filled from the source format */
	int32_t bc4_channel;         /* default 1 (the reference encodes green, src/amd_bc4_compressor.cpp:35) */
	int32_t reserved[6];
} b200ic_opts;

B200IC_API void b200ic_default_opts(b200ic_opts *opts);

/* Selects `device` (>=0) for the calling thread and uploads the constant tables. Returns 0 on success. */
B200IC_API int b200ic_init(int device);
B200IC_API void b200ic_shutdown(void);
/* Last error message of the calling thread ("" if none). */
B200IC_API const char *b200ic_last_error(void);
B200IC_API int b200ic_device_count(void);
/* 1 if kernels for `codec` are built into this library, else 0 (encode calls then fail with an error). */
B200IC_API int b200ic_codec_available(int codec);

B200IC_API uint32_t b200ic_block_bytes(int codec);
B200IC_API uint64_t b200ic_output_bytes(int codec, uint32_t width, uint32_t height, uint32_t slices);
B200IC_API uint32_t b200ic_texel_bytes(int format);

/* Device-resident encode.  d_src: slices x height rows of `row_pitch_bytes` (slice stride
 * `slice_pitch_bytes`, 0 = height*row_pitch); d_dst: slices x blocksY x blocksX blocks, row-major.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Asynchronous. */
B200IC_API int b200ic_encode_device(int codec, const void *d_src, int format, uint32_t width, uint32_t height,
																		uint64_t row_pitch_bytes, uint64_t slice_pitch_bytes, uint32_t slices,
																		const b200ic_opts *opts, void *d_dst, void *stream);

/* Host-buffer encode: H2D of block-row chunks, kernels and D2H pipelined over internal streams.
 * Pageable caller buffers are staged through pinned memory by the calling thread while the GPU works on the previous
 * chunks; pinned / registered buffers are copied directly.  Synchronous.  `progress` (may be NULL) is called once per
 * finished block-row, in order, with the reference's percentage (100 * y * blocksX / (blocksX * blocksY)); returning
 * non-zero cancels (the call then returns 1).  Mirrors Image_CompressProgressFunc semantics. */
typedef int (*b200ic_progress_fn)(void *user, float percent);
B200IC_API int b200ic_encode_host(int codec, const void *h_src, int format, uint32_t width, uint32_t height,
																	uint64_t row_pitch_bytes, uint32_t slices, const b200ic_opts *opts, void *h_dst,
																	b200ic_progress_fn progress, void *user);

/* The same encode with the block-rows of the image sharded over `n_devices` GPUs of this process (devices
 * 0 .. n_devices-1; <= 0 = every visible device), one host thread and one stream ring per device, every shard written
 * straight into its range of h_dst -- no collective.  This is the loop of reference src/amd_bc7_compressor.cpp:48-77
 * (and its siblings) split by block-row.  `progress` calls are serialised.  Synchronous. */
B200IC_API int b200ic_encode_host_sharded(int codec, const void *h_src, int format, uint32_t width, uint32_t height,
																					uint64_t row_pitch_bytes, uint32_t slices, const b200ic_opts *opts, void *h_dst,
																					b200ic_progress_fn progress, void *user, int n_devices);

/* How many GPUs the Image_Compress* entry points (include/gfx_imagecompress/imagecompress.h) shard a large image over:
 * n >= 1 devices, 0 = every visible device (the default; the environment variable B200IC_DEVICES overrides the default). */
B200IC_API void b200ic_set_devices(int n);
B200IC_API int b200ic_get_devices(void);

/* Batched block API on host buffers: `format` is one of B200IC_FMT_BLOCKS_*; nblocks pre-gathered blocks. */
B200IC_API int b200ic_encode_blocks(int codec, const void *h_blocks, int format, uint64_t nblocks,
																		const b200ic_opts *opts, void *h_dst);

/* ---- batches and multi-GPU sharding (BASELINE config[4]: texture batches with mip chains, block-row sharded) -------
 * The reference compresses one image per call and never walks mip chains: callers loop over
 * textures and levels (reference src/amd_bc7_compressor.cpp:25-80 once per level).  Here the loop is a batch. */
typedef struct b200ic_image_desc {
	const void *src;             /* device pointer, `height` rows of `row_pitch_bytes` (0 = tightly packed) */
	void *dst;                   /* device pointer, blocksY x blocksX blocks, row-major */
	int32_t format;              /* b200ic_format (texel formats only) */
	uint32_t width, height;
	uint32_t reserved;
	uint64_t row_pitch_bytes;
} b200ic_image_desc;

/* A shard = block-rows [row0, row1) of image `image`. */
typedef struct b200ic_shard {
	uint32_t image, row0, row1, reserved;
} b200ic_shard;

/* Deterministic partition of the block-rows of `n_images` images over `world` ranks (every rank computes the same
 * plan, no communication): images are cut into chunks of at most `chunk_rows` block-rows (0 = 64); rank r takes the
 * chunks (in image, row order) whose midpoint falls into its share [r T / world, (r + 1) T / world) of the T blocks -- one
 * contiguous run per rank, loads within one chunk of each other -- returned sorted by (image, row0) with adjacent ranges
 * merged.  Blocks are independent, so the encode needs no halo and no
 * collective.  Writes at most `cap` shards of `rank` to `out` and returns how many it has.  Host-only. */
B200IC_API uint64_t b200ic_plan_shards(const uint32_t *widths, const uint32_t *heights, uint64_t n_images, uint32_t chunk_rows,
																			 uint32_t world, uint32_t rank, b200ic_shard *out, uint64_t cap);

/* Encodes `n_shards` shards (or, with shards == NULL, every image whole) of device-resident images.  The launches are
 * spread over internal streams forked from / joined to `stream`, so small mip levels overlap.  Asynchronous. */
B200IC_API int b200ic_encode_batch_device(int codec, const b200ic_image_desc *images, uint64_t n_images,
																					const b200ic_shard *shards, uint64_t n_shards, const b200ic_opts *opts, void *stream);

/* ---- either side of the encode ------------------------------------------------------------------------
 * Next mip level of a device-resident RGBA8 image: 2x2 box filter, floor(x + 0.5) rounding, dimensions floor(size / 2)
 * (a dimension that is already 1 stays 1).  d_dst holds max(1, width/2) x max(1, height/2) texels.  Asynchronous. */
B200IC_API int b200ic_box_mip_rgba8_device(const void *d_src, uint32_t width, uint32_t height, uint64_t src_pitch_bytes,
																					 void *d_dst, uint64_t dst_pitch_bytes, void *stream);
/* Decodes blocks (slices x blocksY x blocksX, row-major) back to texels: RGBA8 for BC1 / BC2 / BC3 / BC7, R8 for BC4,
 * RG8 for BC5, RGBA16F (A = 1) for BC6H (`is_signed` selects the signed format).  The reference has no decoder; this
 * follows the S3TC / RGTC / BPTC specifications.  d_dst: slices x height rows of `dst_row_pitch_bytes` (0 = tight).
 * Asynchronous. */
B200IC_API int b200ic_decode_device(int codec, const void *d_blocks, uint32_t width, uint32_t height, uint32_t slices, int is_signed,
																		void *d_dst, uint64_t dst_row_pitch_bytes, void *stream);
/* Writes a .dds file (DX10 header) with `levels` mip levels of host-resident blocks, level l being
 * max(1, width >> l) x max(1, height >> l) texels.  The reference leaves file output to the external gfx_imageio. */
B200IC_API int b200ic_write_dds(const char *path, int codec, int srgb, int is_signed, uint32_t width, uint32_t height,
																uint32_t levels, const void *const *level_blocks);

/* Per-kernel timing of the AMD BC7 pipeline (bench.py's roofline): while enabled, every kernel launch is bracketed by
 * CUDA events on the launching stream.  b200ic_profile_read waits for the recorded events, adds their durations (ms)
 * and counts into ms[mode * 4 + kind] / launches[mode * 4 + kind] (32 entries each; kind 0 quantise, 1 cube, 2 window,
 * 3 thread-per-block kernel), clears the record and returns the number of launches read. */
B200IC_API void b200ic_profile(int enable);
B200IC_API int b200ic_profile_read(double *ms, uint64_t *launches);

/* Number of kernel launches issued by this library since b200ic_init (for bench.py's gpu_launches). */
B200IC_API uint64_t b200ic_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200IC_H_ */
