/*
 * sloth_b200.h -- C ABI of the B200 raster path (libsloth_b200.so).
 *
 * The reference (ecumene/rust-sloth) has no plugin/FFI interface: its raster
 * path is the group of Rust calls made once per frame from src/main.rs:76-89
 *
 *     let rot = Rotation3::from_euler_angles(..).to_homogeneous();   main.rs:76-77
 *     context.update(size, &mesh_queue)?;                            main.rs:78   (context.rs:93-141)
 *     context.clear();                                               main.rs:79   (context.rs:35-45)
 *     for mesh in &mesh_queue { draw_mesh(&mut context, &mesh, rot, default_shader); }   main.rs:80-83
 *     context.flush(..)   <- reads context.frame_buffer                main.rs:89   (context.rs:50-92)
 *
 * A Rust host replaces exactly that group with the calls below (extern "C"
 * block + build.rs shown in INTEGRATION.md).  Conventions: every function
 * returns 0 on success and a negative SLOTH_E_* code on failure, with a
 * human-readable message available from sloth_last_error() (thread-local);
 * nothing is thrown across the boundary; the caller owns every host buffer;
 * device memory belongs to the context; a context is bound to one GPU and is
 * not thread-safe (use one context per host thread / per GPU).
 *
 * There is NO CPU fallback: if no CUDA device is usable, sloth_ctx_create
 * fails with SLOTH_E_CUDA.
 *
 * Frame-buffer cell format (replaces `(char, (u8,u8,u8))`, context.rs:16):
 *     uint32 cell = glyph | r << 8 | g << 16 | b << 24        (glyph is ASCII)
 * blank = (' ',0,0,0), image-mode row marker = ('\n',0,0,0)  (rasterizer.rs:89-91).
 * Matrices are column-major float[16], like nalgebra's Matrix4<f32> storage.
 */
#ifndef SLOTH_B200_H
#define SLOTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLOTH_API __attribute__((visibility("default")))

enum {
    SLOTH_OK = 0,
    SLOTH_E_ARG = -1,      /* bad argument (null pointer, zero size, out of range) */
    SLOTH_E_CUDA = -2,     /* CUDA runtime error / no usable device */
    SLOTH_E_STATE = -3,    /* call order (render before scene_set / resize) */
    SLOTH_E_TOO_LARGE = -4, /* more than 2^27-1 triangles, or W*H+H >= 2^31, or W/H > 65535 */
    SLOTH_E_IO = -5,        /* a model or material file could not be read */
    SLOTH_E_PARSE = -6,     /* malformed model file (message in the reference's wording) */
    SLOTH_E_UNSUPPORTED = -7 /* well-formed input the device parser does not decide (inf/nan, >19 digits on a rounding boundary ...) */
};

typedef struct sloth_ctx sloth_ctx;

/* Context::blank(image)  -- src/context.rs:22-34.  `device` = CUDA ordinal. */
SLOTH_API int sloth_ctx_create(int device, int image_mode, sloth_ctx **out);
SLOTH_API int sloth_ctx_destroy(sloth_ctx *ctx);

/*
 * The mesh queue (main.rs:31, inputs.rs:95-129) as one triangle soup in draw
 * order (mesh order, then triangle order: main.rs:80-83, rasterizer.rs:43-45).
 *   xyz  n_tri*9 floats: v1.xyz v2.xyz v3.xyz, w = 1 implied (geometry.rs:17-23)
 *   rgb  n_tri*3 bytes:  Triangle.color
 *   scene_max  fold(max) over meshes of bounding_box.max.{x,y,z} starting at
 *              0.0 -- the `scale` loop of Context::update, context.rs:106-113.
 * Uploaded once; the soup stays resident in HBM (40 B/triangle).
 */
SLOTH_API int sloth_scene_set(sloth_ctx *ctx, const float *xyz, const uint8_t *rgb, size_t n_tri,
                              float scene_max);

/*
 * The same mesh queue before to_simple_mesh de-indexes it (geometry.rs:99-107 reads
 * positions[indices[..]]; tobj's Mesh and stl_io's IndexedMesh are both indexed):
 *   positions  n_vert*3 floats, indices  n_tri*3 vertex ids in draw order (meshes concatenated, ids offset by
 *   the caller), rgb / scene_max as above.  SLOTH_E_ARG when an id is >= n_vert.
 * Both entry points end in the same resident scene: corners are deduplicated by exact bit pattern, and when
 * triangles share vertices the frame path runs Triangle::mul (geometry.rs:43-48) once per unique vertex
 * (k_xform) instead of once per triangle corner -- bit-identical, since the product is a pure function of the
 * vertex.  sloth_ctx_set_path (before the scene is set) pins the choice: AUTO picks the indexed path at <= 1.5
 * unique vertices per triangle, SOUP / INDEXED force one (tests, profiling).
 */
SLOTH_API int sloth_scene_set_indexed(sloth_ctx *ctx, const float *positions, size_t n_vert, const uint32_t *indices,
                                      const uint8_t *rgb, size_t n_tri, float scene_max);
enum { SLOTH_PATH_AUTO = 0, SLOTH_PATH_SOUP = 1, SLOTH_PATH_INDEXED = 2 };
SLOTH_API int sloth_ctx_set_path(sloth_ctx *ctx, int path);

/*
 * Model loading on the device: match_meshes (inputs.rs:95-129: tobj 3.2.2 / stl_io 0.4.2) and
 * to_meshes (geometry.rs:83-189: de-indexed soup, fan triangulation, colour rules, bounding boxes) for files of
 * any size.  The file's bytes go to the GPU once; line splitting, decimal -> f32 conversion (correctly rounded,
 * like Rust's f32::from_str), index resolution, triangulation, colours and the max-coordinate fold of
 * Context::update (context.rs:106-113) all run there and the soup never visits the host.
 *
 *   sloth_scene_load   the whole of match_meshes for one CLI value ("a.obj b.stl", split on ' ', extension
 *                      dispatch, .mtl files resolved next to the .obj); replaces the scene like sloth_scene_set.
 *   sloth_loader_*     the same for bytes the caller already holds: begin, add files in draw order, commit.
 *                      `mtl_dir` is the directory (with trailing '/', or "" / NULL for the cwd) where mtllib
 *                      statements are looked up.  Texts must be shorter than 4 GiB each.
 *   sloth_scene_size / sloth_scene_get   read the resident soup back (tests, tools).
 * Errors carry the reference's wording (SLOTH_E_PARSE / SLOTH_E_IO).  SLOTH_E_UNSUPPORTED marks input that is
 * legal but outside what the device parser decides exactly; nothing is loaded then and the caller may parse on
 * the host and use sloth_scene_set (the `sloth` CLI does, with a note on stderr).
 */
SLOTH_API int sloth_scene_load(sloth_ctx *ctx, const char *models_arg, size_t *n_tri_out, float *scene_max_out);
SLOTH_API int sloth_loader_begin(sloth_ctx *ctx);
SLOTH_API int sloth_loader_add_obj(sloth_ctx *ctx, const char *text, size_t len, const char *mtl_dir);
SLOTH_API int sloth_loader_add_stl(sloth_ctx *ctx, const void *bytes, size_t len);
SLOTH_API int sloth_loader_commit(sloth_ctx *ctx, size_t *n_tri_out, float *scene_max_out);
SLOTH_API size_t sloth_scene_size(const sloth_ctx *ctx);
SLOTH_API int sloth_scene_get(sloth_ctx *ctx, float *xyz, uint8_t *rgb, float *scene_max_out);

/* match_dimensions (inputs.rs:159-169) / the size adoption in Context::update
 * (context.rs:134-137).  (Re)allocates the frame state for W x H cells. */
SLOTH_API int sloth_ctx_resize(sloth_ctx *ctx, uint32_t width, uint32_t height);

/*
 * One frame = update + clear + draw_mesh over the whole queue (main.rs:78-83).
 *   rot        the `transform` passed to draw_mesh (column-major 4x4)
 *   cells_out  host buffer, W*H cells (+H in image mode, context.rs:36-41)
 *   z_out      optional host buffer, W*H floats = Context.z_buffer after the
 *              frame (may be NULL).  A winning -0.0 is reported as +0.0.
 */
SLOTH_API int sloth_render(sloth_ctx *ctx, const float rot[16], uint32_t *cells_out, float *z_out);

/* The -j / turntable loop (main.rs:60-111): n_frames frames, frame k uses
 * rots[16*k ..], written to cells_out + k*cells_per_frame.  Device work and
 * device->host copies are pipelined across frames. */
SLOTH_API int sloth_render_batch(sloth_ctx *ctx, const float *rots, size_t n_frames, uint32_t *cells_out);

/*
 * What crosses PCIe when a frame's destination is host memory (Context.frame_buffer, context.rs:16-17).
 *   SLOTH_WIRE_CELLS  the plain 4-byte cells (default): 33 MB per 3840x2160 frame, i.e. the PCIe rate of the box
 *                     is the frame rate of sloth_render / sloth_render_batch.
 *   SLOTH_WIRE_SPANS  run-length: the device sends one (start, cell) pair per run of equal cells -- the blank
 *                     background and the doubled cells of rasterizer.rs:83-85 make frames mostly runs -- and a small
 *                     pool of host threads inside the library rebuilds the 4-byte cells into cells_out while the
 *                     next frames render (streaming stores).  Frames whose run list would exceed half the plain
 *                     bytes are sent as plain cells.  cells_out receives the same bytes either way; z_out requests
 *                     and the device-resident entry points are unaffected.  Threads: SLOTH_WIRE_THREADS, default =
 *                     hardware threads / LOCAL_WORLD_SIZE (2..32).
 * sloth_wire_stats: out[0] = frames sent since sloth_ctx_set_wire, out[1] = of those as plain cells, out[2] =
 * device->host bytes they took, out[3] = pool threads.
 * sloth_expand_spans: the host half on its own (no GPU involved): runs = n_runs pairs (start, cell), ascending
 * starts, the first at 0; run i covers [start_i, start_{i+1}), the last one ends at n_cells.
 */
enum { SLOTH_WIRE_CELLS = 0, SLOTH_WIRE_SPANS = 1 };
SLOTH_API int sloth_ctx_set_wire(sloth_ctx *ctx, int wire);
SLOTH_API int sloth_wire_stats(const sloth_ctx *ctx, uint64_t out[4]);
SLOTH_API int sloth_expand_spans(const uint32_t *runs, size_t n_runs, uint32_t *cells_out, size_t n_cells);

/* Same frame, but the result stays on the device: d_cells is a 16-byte aligned device pointer
 * (same GPU) to W*H(+H) cells -- or band_rows*W cells when a band is set.
 * Runs on the context's stream; sloth_ctx_sync() waits for it. */
SLOTH_API int sloth_render_device(sloth_ctx *ctx, const float rot[16], void *d_cells);
/* n_frames frames with device-resident results: frame k goes to d_cells + k*frame_stride_cells

 * (a multiple of 4 cells; stride 0 = every frame overwrites the same buffer).  Inside the call the geometry of frame k+1
 * overlaps the resolve of frame k on a second internal stream; when the context stream (or
 * sloth_ctx_sync) completes, all frames are complete. */
SLOTH_API int sloth_render_device_batch(sloth_ctx *ctx, const float *rots, size_t n_frames, void *d_cells,
                                        size_t frame_stride_cells);
SLOTH_API int sloth_ctx_sync(sloth_ctx *ctx);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on, so
 * a caller can record its own events around sloth_render_device calls. */
SLOTH_API void *sloth_ctx_stream(sloth_ctx *ctx);

/*
 * Row-band mode for one huge frame split across GPUs: this context produces
 * only cells of rows [row0, row1) (destination rows: a fragment that wraps
 * past the end of row y belongs to row y+1, rasterizer.rs:80).  The output
 * of sloth_render_device is then (row1-row0)*W cells; the image-mode tail
 * (H blank cells) is produced by whoever assembles the frame.
 * row0 = row1 = 0 restores whole-frame mode.
 */
SLOTH_API int sloth_ctx_set_band(sloth_ctx *ctx, uint32_t row0, uint32_t row1);

/*
 * Device buffers shared between the processes of one box (one process per GPU), for band mode without a gather
 * step: the rank that assembles the frame allocates it and exports a handle; every other rank opens the handle
 * and passes `mapped + 4*row0*W` to sloth_render_device, so its resolve kernel stores the band straight into
 * the owner's frame over NVLink (peer access is enabled by the open).  sloth_device_read / _write copy between
 * device memory (own or mapped) and the host on the context's stream and wait.
 */
#define SLOTH_IPC_HANDLE_BYTES 64
SLOTH_API int sloth_device_alloc(int device, size_t bytes, void **d_ptr_out);
SLOTH_API int sloth_device_free(int device, void *d_ptr);
SLOTH_API int sloth_ipc_export(int device, const void *d_ptr, unsigned char handle_out[SLOTH_IPC_HANDLE_BYTES]);
SLOTH_API int sloth_ipc_open(int device, const unsigned char handle[SLOTH_IPC_HANDLE_BYTES], void **d_ptr_out);
SLOTH_API int sloth_ipc_close(int device, void *d_ptr);
SLOTH_API int sloth_device_read(sloth_ctx *ctx, const void *d_ptr, void *host_out, size_t bytes);
SLOTH_API int sloth_device_write(sloth_ctx *ctx, void *d_ptr, const void *host_in, size_t bytes);

/*
 * Context::flush on the device (src/context.rs:50-92).  The exact bytes the
 * reference prints for the cell buffer, produced on the GPU:
 *   mode 0  plain glyphs                                  flush(color = false)   (without println's '\n')
 *   mode 1  ESC[48;2;25;25;25m ESC[38;2;R;G;Bm c ESC[0m   flush(true, false)     (crossterm 0.18)
 *   mode 2  <span style="color:rgb(R,G,B)">c              flush(true, true)      (the -j export)
 * Frame prefixes/suffixes (cursor move, "`\n", "`,\n" ...) stay with the host.
 * sloth_text_capacity: worst-case bytes of one frame.  sloth_render_text(_batch): render + flush,
 * text of frame k at text_out + k*frame_stride (>= capacity), its length in lens_out[k]; rendering,
 * serialisation and device->host copies of consecutive frames are pipelined.
 */
SLOTH_API size_t sloth_text_capacity(const sloth_ctx *ctx, int mode);
SLOTH_API int sloth_render_text(sloth_ctx *ctx, const float rot[16], int mode, char *text_out, size_t cap,
                                size_t *len_out);
SLOTH_API int sloth_render_text_batch(sloth_ctx *ctx, const float *rots, size_t n_frames, int mode, char *text_out,
                                      size_t frame_stride, size_t *lens_out);
/* Serialise a cell buffer that is already on the device (same GPU); returns the text length. */
SLOTH_API int sloth_flush_device(sloth_ctx *ctx, int mode, const void *d_cells, size_t n_cells, void *d_text,
                                 size_t cap, size_t *len_out);

/* The `shader: F` closure of draw_mesh (rasterizer.rs:39-41), restricted to
 * what the reference ever passes: 9 ascending `<=` thresholds and 10 glyphs
 * (the last one for "above all thresholds or NaN").  Default = default_shader,
 * rasterizer.rs:5-27. */
SLOTH_API int sloth_shader_set(sloth_ctx *ctx, const float thr[9], const char glyph[10]);

typedef struct sloth_stats {
    uint64_t frames;           /* frames rendered by this context */
    uint64_t kernel_launches;  /* CUDA kernels launched by this library */
    uint64_t fragments;        /* covered fragments (depth tests) of the last frame; needs count_fragments */
    uint32_t n_tri;
    uint32_t walk_tris;        /* last frame: triangles sent to the warp-per-row-band kernel */
    uint32_t walk_items;       /* last frame: row-band work items */
    uint32_t irregular_tris;   /* last frame: triangles sent to the brute-force kernel */
    uint32_t stamp_fixups;     /* last frame: newline-vs-wrapped-fragment order fix-ups */
    float last_frame_ms;       /* device time of the last sloth_render / per frame of the last batch */
    float geom_ms, walk_ms, resolve_ms; /* per-kernel device times of the last sloth_render when timing is on */
    uint32_t chunks_processed; /* last frame: chunks of 32 triangles the geometry kernel did not band-cull; needs count_fragments */
    float load_read_ms, load_parse_ms, load_commit_ms; /* last sloth_scene_load: file -> pinned memory, copy + device
                                                          parse, soup -> resident scene (host wall clock) */
    uint32_t n_vert;           /* unique vertices of the resident scene (0 when it renders through the soup path) */
    uint32_t geom_path;        /* SLOTH_PATH_SOUP or SLOTH_PATH_INDEXED: what the resident scene uses */
    float xform_ms;            /* per-vertex transform kernel of the last sloth_render when timing is on (indexed path) */
    uint64_t l2_window_bytes;  /* bytes of transformed vertices kept L2-resident by an access-policy window (0 = none) */
    uint64_t l2_persist_max;   /* the device's persisting-L2 set-aside limit */
    uint32_t tile_tris;        /* last frame: queued triangles rasterised by the binned tile path (the rest of walk_tris walked) */
    uint32_t tile_pairs;       /* last frame: (tile, triangle) pairs */
    uint32_t tiles_used;       /* last frame: non-empty screen tiles */
} sloth_stats;

SLOTH_API int sloth_stats_get(sloth_ctx *ctx, sloth_stats *out);
/* flags: bit 0 = count fragments (adds atomics; off by default), bit 1 = per-kernel event timing */
SLOTH_API int sloth_stats_enable(sloth_ctx *ctx, uint32_t flags);

SLOTH_API const char *sloth_last_error(void);

/* ---- host-side helpers (no GPU work): the scalar arithmetic around the path ---- */

/* Rotation3::from_euler_angles(roll,pitch,yaw).to_homogeneous(), main.rs:76-77
 * (libm sinf/cosf on the host, like Rust's f32::sin_cos). */
SLOTH_API void sloth_rotation_from_euler(float roll, float pitch, float yaw, float out[16]);
/* Context::update's matrix, context.rs:93-133 (sizes taken `as u16`). */
SLOTH_API void sloth_utransform(uint32_t width, uint32_t height, float scene_max, float out[16]);
/* pitch sequence of `image -j N` starting from the -y argument (inputs.rs:131-149,
 * main.rs:55-58,92-106); returns the number of frames the reference renders. */
SLOTH_API size_t sloth_turntable_pitches(float y_arg, uint32_t n_frames, float *out, size_t cap);
/* Page-locked host memory for cells_out: device->host copies of a batch only
 * overlap with rendering when the destination is pinned. */
SLOTH_API int sloth_pinned_alloc(size_t bytes, void **out);
SLOTH_API int sloth_pinned_free(void *ptr);
/* Page-lock memory the caller already owns (cudaHostRegister) -- e.g. a POSIX shared-memory frame that the per-GPU
 * processes of a box all map: in band mode every rank then passes `frame + row0*W` as cells_out of sloth_render and
 * its band travels straight over its own PCIe link into the final host frame (no root GPU, no gather). */
SLOTH_API int sloth_host_register(void *ptr, size_t bytes);
SLOTH_API int sloth_host_unregister(void *ptr);
/* number of cells per frame for this context: W*H (+H in image mode), or band size */
SLOTH_API size_t sloth_cells_per_frame(const sloth_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SLOTH_B200_H */
