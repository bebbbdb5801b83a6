/*
 * nrays_b200.h — C-ABI drop-in boundary for the nrays per-pixel render hot path.
 *
 * The reference (sebcrozet/nrays, Rust) has no FFI/plugin layer; the seam this
 * header replaces is the pair
 *     Scene::new(nodes, lights, background)            src/scene.rs:119-133
 *     scene::render(&scene, &resolution, ray_per_pixel,
 *                   window_width, camera_eye, projection) -> Image
 *                                                      src/scene.rs:29-116
 * whose only caller is examples/loader3d.rs:61,86-93.  The Rust host keeps its
 * Scene/SceneNode/Material/Light API; a `Scene::flatten()` shim
 * fills the plain-old-data tables below and hands them across this boundary.
 *
 * Conventions
 *   - plain pointers + sizes only; no C++/torch types; every struct is POD.
 *   - geometry parameters are f64 exactly as the reference holds them
 *     (src/lib.rs:33 `Scalar = f64`); the device converts to f32 once at upload.
 *   - colours / energies are f32 as in the reference (src/ray_with_energy.rs:7).
 *   - all descriptor memory is caller-owned and only read during the call.
 *   - no unwinding across the boundary: every entry point returns an int status
 *     (NRB_OK == 0); nrb_last_error() gives the message for the calling thread.
 */
#ifndef NRAYS_B200_H
#define NRAYS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRB_ABI_VERSION 1u

/* ---- status codes --------------------------------------------------------- */
enum {
  NRB_OK = 0,
  NRB_ERR_INVALID_ARG = 1,    /* mirrors assert!(ray_per_pixel > 0) src/scene.rs:37 and malformed tables */
  NRB_ERR_CUDA = 2,           /* any CUDA runtime failure (message has the CUDA string) */
  NRB_ERR_NO_DEVICE = 3,      /* no sm_100 device visible: the product path never falls back to CPU */
  NRB_ERR_QUEUE_OVERFLOW = 4, /* ray queues would exceed the configured memory ceiling */
  NRB_ERR_UNSUPPORTED = 5,    /* feature outside the flattened format (e.g. user-defined Material impl) */
  NRB_ERR_INTERNAL = 6        /* a self-check of the library failed (BVH invariant): a bug in this library, not in the input */
};

/* ---- shape kinds: ncollide3d shapes built at examples/loader3d.rs:593-659,695 */
enum {
  NRB_SHAPE_BALL = 0,     /* param[0] = radius                          loader3d.rs:601 */
  NRB_SHAPE_CUBOID = 1,   /* param[0..2] = half extents                 loader3d.rs:612 */
  NRB_SHAPE_CYLINDER = 2, /* param[0] = half_height, param[1] = radius  loader3d.rs:623 (axis = local Y) */
  NRB_SHAPE_CAPSULE = 3,  /* param[0] = half_height, param[1] = radius  loader3d.rs:634 */
  NRB_SHAPE_CONE = 4,     /* param[0] = half_height, param[1] = radius  loader3d.rs:645 (apex at +Y) */
  NRB_SHAPE_PLANE = 5,    /* param[0..2] = unit normal                  loader3d.rs:656 */
  NRB_SHAPE_TRIMESH = 6   /* first_index / tri_count / vertex_base      loader3d.rs:695 */
};

/* ---- material kinds: the three Material impls of the reference -------------- */
enum {
  NRB_MAT_PHONG = 0,  /* src/phong_material.rs:9-151  */
  NRB_MAT_NORMAL = 1, /* src/normal_material.rs:7-15  */
  NRB_MAT_UV = 2      /* src/uv_material.rs:8-21      */
};

/* ---- texture sampling modes: src/texture2d.rs:50-58 ------------------------- */
enum { NRB_INTERP_BILINEAR = 0, NRB_INTERP_NEAREST = 1 };
enum { NRB_OVERFLOW_CLAMP = 0, NRB_OVERFLOW_WRAP = 1 };

/* One SceneNode (src/scene_node.rs:8-19). */
typedef struct NrbNodeDesc {
  int32_t shape;            /* NRB_SHAPE_* */
  int32_t material;         /* index into NrbSceneDesc.materials */
  double param[3];          /* see shape kinds */
  double rot[9];            /* rotation of `transform: Isometry3<f64>`, row-major 3x3 */
  double trans[3];          /* translation of `transform` */
  double refr_coeff;        /* SceneNode.refr_coeff (f64) */
  float refl_mix;           /* SceneNode.refl_mix */
  float refl_atenuation;    /* SceneNode.refl_atenuation (sic) */
  float alpha;              /* SceneNode.alpha */
  int32_t solid;            /* SceneNode.solid (bool) */
  int32_t nmap_texture;     /* SceneNode.nmap (depth shift, src/scene_node.rs:60-70): texture index or -1; <= 32 such nodes */
  int32_t _pad;
  /* TRIMESH only: triangles [first_index/3, first_index/3 + tri_count) of `indices`,
   * each index relative to `vertex_base` in positions/uvs (TriMesh::new(coords, faces, Some(uvs))). */
  uint64_t first_index;
  uint64_t tri_count;
  uint64_t vertex_base;
} NrbNodeDesc;

/* Light (src/light.rs:8-23).  racsample = floor(sqrt(nsample)) is computed by the host exactly as Light::new does. */
typedef struct NrbLightDesc {
  double pos[3];
  double radius;
  uint32_t racsample;
  float color[3];
} NrbLightDesc;

/* Material table row.  For NRB_MAT_PHONG the fields are PhongMaterial's (src/phong_material.rs:9-16). */
typedef struct NrbMaterialDesc {
  int32_t kind;          /* NRB_MAT_* */
  float ambient[3];
  float diffuse[3];
  float specular[3];
  float shininess;
  int32_t texture;       /* diffuse texture index or -1 */
  int32_t alpha_texture; /* opacity-map texture index or -1 */
} NrbMaterialDesc;

/* Texture2d (src/texture2d.rs:60-64) over a shared RGBA32F texel pool
 * (ImageData.pixels: Vec<Point4<f32>>, row-major y*W+x, y already flipped at load: texture2d.rs:99-107). */
typedef struct NrbTextureDesc {
  uint32_t width, height;
  int32_t interpolation; /* NRB_INTERP_* */
  int32_t overflow;      /* NRB_OVERFLOW_* */
  uint64_t texel_offset; /* in texels (4 floats each) into NrbSceneDesc.texels */
} NrbTextureDesc;

/* The flattened scene: what Scene::new receives, as tables. */
typedef struct NrbSceneDesc {
  uint32_t struct_size; /* sizeof(NrbSceneDesc), for forward compatibility */
  uint32_t abi_version; /* NRB_ABI_VERSION */
  uint32_t n_nodes;
  uint32_t n_lights;
  uint32_t n_materials;
  uint32_t n_textures;
  const NrbNodeDesc *nodes;
  const NrbLightDesc *lights;
  const NrbMaterialDesc *materials;
  const NrbTextureDesc *textures;
  uint64_t n_texels;
  const float *texels;      /* 4 floats per texel */
  uint64_t n_vertices;
  const float *positions;   /* 3 floats per vertex, mesh-local frame (already /4: loader3d.rs:669) */
  const float *uvs;         /* 2 floats per vertex, or NULL (then zero, as src/obj.rs:383 zero-fills) */
  uint64_t n_indices;
  const uint32_t *indices;  /* 3 per triangle */
  float background[3];      /* Scene.background (src/scene.rs:22); loader passes (1,1,1): loader3d.rs:61 */
  uint32_t _pad;
} NrbSceneDesc;

/* Arguments of scene::render (src/scene.rs:29-36) plus the two things the reference
 * leaves implicit: the RNG seed (rand::random is OS-seeded there) and a recursion cap
 * (the reference has none and overflows its stack on mirror boxes). */
typedef struct NrbCamera {
  uint32_t width, height;   /* resolution (Vector2<f64> holding integers in the reference) */
  uint32_t ray_per_pixel;   /* > 0, else NRB_ERR_INVALID_ARG (assert at src/scene.rs:37) */
  uint32_t max_depth;       /* recursion cap; 0 -> default 64.  Paths cut by it are counted in NrbStats */
  double window_width;      /* AA jitter width in pixels */
  double eye[3];            /* camera_eye */
  double projection[16];    /* inverse view-projection, column-major (nalgebra Matrix4 storage order) */
  uint64_t seed;            /* Philox4x32-10 key; same seed => same image on oracle and device */
} NrbCamera;

/* Optional sharding of the image into 16x16-pixel tiles (multi-GPU).
 * The rank renders tiles t = first, first+stride, ... (< n_tiles_total), in that order, into a
 * packed buffer [n_local_tiles][16][16][3]; pixels outside the image are written as 0. */
#define NRB_TILE 16u
typedef struct NrbTileSet {
  uint32_t first;
  uint32_t stride;
} NrbTileSet;

/* Counters returned by every render (each counted ray is one BVH query). */
typedef struct NrbStats {
  uint64_t rays_primary;
  uint64_t rays_reflect;
  uint64_t rays_refract;
  uint64_t rays_shadow;     /* light samples of the reference semantics = shadow queries cast + rays_shadow_culled */
  uint64_t paths_truncated; /* children not spawned because of max_depth */
  uint32_t waves;           /* wavefront iterations */
  uint32_t kernel_launches; /* this library's kernels launched by the call */
  float ms_device;          /* CUDA-event span raygen -> resolve on the render stream */
  float ms_trace;           /* CUDA-event time of the trace_kernel launches ONLY (closest hit + shadow rays; not the tail kernel) */
  float ms_shade;           /* ms_device - ms_trace - ms_tail: shade + resolve + inter-kernel gaps */
  uint32_t _pad;
  uint64_t bvh_nodes;       /* device BVH size, for the roofline's scene_bytes */
  uint64_t triangles;
  uint64_t scene_bytes;
  uint32_t launches_trace;  /* launches of the persistent trace kernel (ms_trace is their CUDA-event time) */
  uint32_t launches_shade;  /* launches of the shade kernel */
  uint64_t rays_shadow_culled; /* of rays_shadow: light samples whose contribution is exactly zero (weight 0, e.g. hits on
                                * fully transparent texels) — the reference casts them, this library does not */
  float ms_tail;            /* CUDA-event time of the tail kernel launches (ray chains followed per lane) */
  float ms_shade_kernel;    /* CUDA-event time of the shade_kernel launches alone */
  uint32_t launches_tail;   /* launches of the tail kernel */
  uint32_t _pad2;
  uint64_t rays_tail;       /* closest-hit queries answered inside the tail kernel (the rest went through trace_kernel);
                             * every shadow query goes through trace_kernel */
} NrbStats;

typedef struct NrbScene NrbScene; /* opaque handle == Arc<Scene> of the reference */

/* ---- entry points ----------------------------------------------------------- */

/* Number of usable CUDA devices (0 if none). */
int nrb_device_count(void);

/* Replaces Scene::new (src/scene.rs:119-133): validates the tables, builds the BVH
 * (replacing BVT::new_balanced, src/scene.rs:126, and TriMesh::new's inner BVT,
 * loader3d.rs:695), uploads everything to `device`.  The descriptor is not retained. */
int nrb_scene_create(const NrbSceneDesc *desc, int device, NrbScene **out);

/* Host-only half of nrb_scene_create: validates the tables and builds the BVH WITHOUT touching a
 * device (same status codes for malformed input), then checks the tree's structural invariants.
 * Lets a host validate a flattened scene, and the CPU test-suite exercise the builder. */
typedef struct NrbBuildInfo {
  uint64_t bvh_nodes, triangles, shapes, planes, transparent_candidates;
  uint32_t max_depth;  /* deepest leaf, <= 60 (traversal stack) */
  float build_ms;      /* host wall time of flatten + BVH build (+ upload for nrb_scene_create*) */
  float gpu_build_ms;  /* CUDA-event time of the device build kernels (NRB_BUILDER_LBVH), else 0 */
  uint32_t builder;    /* NRB_BUILDER_* actually used */
  uint32_t node_format; /* device node record the scene was uploaded in: 0 fp32 (64 B), 2 bf16 half extents (48 B),
                           3 / 4 16-bit grid (32 B), plain / speculative loop; 0 from nrb_scene_validate (nothing uploaded) */
  uint32_t _pad;
} NrbBuildInfo;
int nrb_scene_validate(const NrbSceneDesc *desc, NrbBuildInfo *info);

/* BVH builders.  The tree shape never changes render results, only build and traversal cost. */
enum {
  NRB_BUILDER_SAH = 0, /* host binned-SAH build (threaded): best traversal, default */
  NRB_BUILDER_LBVH = 1, /* device build (Morton sort + Karras radix tree + bottom-up fit): fastest to build, slowest to
                           traverse (the topology follows Morton prefixes, not the SAH) */
  NRB_BUILDER_PLOC = 2  /* device build (Morton sort + parallel locally-ordered clustering, Meister & Bittner 2018): every merge
                           minimises the merged surface area, so the tree traverses close to the SAH tree while still building
                           in milliseconds — the device builder of choice for one-shot renders of large meshes */
};
typedef struct NrbBuildOptions {
  uint32_t builder; /* NRB_BUILDER_* */
  uint32_t _reserved[3];
} NrbBuildOptions;

/* nrb_scene_create with options.  opts == NULL: defaults, and only then the environment variable NRB_BUILDER=sah|lbvh|ploc
 * is consulted (experiments); an explicit NrbBuildOptions always wins. */
int nrb_scene_create_opts(const NrbSceneDesc *desc, int device, const NrbBuildOptions *opts, NrbScene **out);

/* How the scene behind a handle was built. */
int nrb_scene_build_info(const NrbScene *scene, NrbBuildInfo *info);

/* Drops the handle and all device memory. */
void nrb_scene_destroy(NrbScene *scene);

/* Replaces Scene::set_background (src/scene.rs:136-139). */
int nrb_scene_set_background(NrbScene *scene, const float rgb[3]);

/* Runs all of this scene's device work on `cuda_stream` (a cudaStream_t owned by the caller, e.g. the
 * harness's current stream, so one pair of events brackets render + collective); NULL restores the
 * scene's own non-blocking stream.  The reference's render is synchronous; so is ours on any stream. */
int nrb_scene_set_stream(NrbScene *scene, void *cuda_stream);

/* Replaces scene::render (src/scene.rs:29-116).  `out_rgb` is a HOST buffer of
 * width*height*3 floats, row-major, pixel (x,y) at 3*(x + y*width) — the layout of
 * Image.pixels (src/image.rs:12-24).  Synchronous: the image is complete on return.
 * `stats` may be NULL.  If `out_rgb` is pinned (nrb_host_alloc / cudaHostAlloc), the device->host copy of the
 * image overlaps the last phase of the frame instead of following it. */
int nrb_render(NrbScene *scene, const NrbCamera *camera, float *out_rgb, NrbStats *stats);

/* Same render, result left in DEVICE memory (`d_out_rgb`, same layout) on the scene's
 * device; returns after the work is complete. */
int nrb_render_device(NrbScene *scene, const NrbCamera *camera, float *d_out_rgb, NrbStats *stats);

/* Tile-sharded render for multi-GPU use: renders only the tiles of `tiles` into the packed
 * DEVICE buffer `d_out_tiles` ([n_local][16][16][3] floats).  *n_local_tiles receives the count. */
int nrb_render_tiles_device(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles,
                            float *d_out_tiles, uint32_t *n_local_tiles, NrbStats *stats);

/* Tile-sharded render that resolves this rank's tiles straight into the ROW-MAJOR image `d_image_rgb`
 * (width*height*3 floats, layout of nrb_render_device) and writes only the pixels it owns.  The image may live on
 * ANOTHER GPU (memory opened with nrb_ipc_open): the finished pixels then cross NVLink once, in the resolve
 * kernel itself, and no gather / un-tile pass follows (the one exchange of SURVEY 8e, fused into K5).  The
 * caller orders the ranks (e.g. one tiny all-reduce) before the owner reads the image. */
int nrb_render_tiles_to_image(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, float *d_image_rgb,
                              NrbStats *stats);

/* End-to-end form of the same exchange: the destination of scene::render is a HOST image.  Renders this rank's tiles and
 * drops them into `host_image_rgb` — pinned / registered host memory (nrb_host_register), e.g. one shared-memory segment
 * mapped by every rank — with ONE strided 2-D DMA over this GPU's own PCIe link; N ranks fill the image through N links in
 * parallel.  Needs width % 16 == 0, tiles_x % stride == 0 and first < stride (each rank then owns whole tile columns);
 * otherwise NRB_ERR_UNSUPPORTED and the caller uses nrb_render_tiles_to_image. */
int nrb_render_tiles_to_host(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, float *host_image_rgb,
                             NrbStats *stats);

/* RGB8 forms of the two exchanges above ("next" row 8f-2 fused into the exchange): the finished pixels leave the GPU quantised
 * as Image::to_png encodes them (src/image.rs:64-77: clamp(c*255, 0, 255) truncated), 3 bytes per pixel instead of 12 over
 * NVLink / PCIe.  `d_image_rgb8` / `host_image_rgb8`: width*height*3 bytes, row-major; same geometry rules as the float forms. */
int nrb_render_tiles_to_image_rgb8(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, uint8_t *d_image_rgb8,
                                   NrbStats *stats);
int nrb_render_tiles_to_host_rgb8(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, uint8_t *host_image_rgb8,
                                  NrbStats *stats);

/* Device memory that other processes on the node can map (one process per GPU): the owner allocates and
 * passes the 64-byte handle around (any transport), peers open it and get a pointer valid in their kernels. */
typedef struct NrbIpcHandle {
  unsigned char bytes[64];
} NrbIpcHandle;
int nrb_ipc_alloc(int device, uint64_t bytes, void **d_ptr, NrbIpcHandle *handle); /* zero-filled */
int nrb_ipc_open(int device, const NrbIpcHandle *handle, void **d_ptr);            /* in ANOTHER process than the owner */
int nrb_ipc_close(int device, void *d_ptr);
int nrb_ipc_free(int device, void *d_ptr);

/* Pins and maps caller-owned host memory (a shared-memory segment for nrb_render_tiles_to_host; any buffer nrb_render
 * should overlap its copy into) and returns the address kernels use for it.  (Kernel stores into such memory work —
 * nrb_render_tiles_to_image accepts the pointer — but reach only ~10 GB/s over PCIe; DMA is the way in.) */
int nrb_host_register(int device, void *host_ptr, uint64_t bytes, void **d_ptr);
int nrb_host_unregister(int device, void *host_ptr);

/* Number of 16x16 tiles covering width x height, and how many of them a tile set owns. */
uint32_t nrb_tile_count(uint32_t width, uint32_t height);
uint32_t nrb_tile_count_local(uint32_t width, uint32_t height, const NrbTileSet *tiles);

/* After the gather: scatter `n_ranks` packed tile buffers laid end to end in DEVICE memory
 * (rank r owning tiles r, r+n_ranks, ...; each rank's buffer padded to `tiles_per_rank` tiles)
 * into a row-major width*height*3 DEVICE image.  Runs on `cuda_stream` (NULL = default stream) and
 * returns after it completed. */
int nrb_untile_device(int device, void *cuda_stream, const float *d_gathered, uint32_t n_ranks,
                      uint32_t tiles_per_rank, uint32_t width, uint32_t height, float *d_out_rgb);

/* "Next" row 8f-2: Image::to_png's quantisation (src/image.rs:64-77): clamp(c*255, 0, 255) truncated
 * to u8, RGB8 row-major, on device; `out_rgb8` is a HOST buffer of width*height*3 bytes. */
int nrb_render_rgb8(NrbScene *scene, const NrbCamera *camera, uint8_t *out_rgb8, NrbStats *stats);

/* Pinned host memory helpers so the device->host read of the image is a single async copy. */
void *nrb_host_alloc(uint64_t bytes);
void nrb_host_free(void *p);

/* Message for the last non-zero status on this thread ("" if none). */
const char *nrb_last_error(void);

/* Library/build identification: "nrays_b200 <abi> sm_100a". */
const char *nrb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NRAYS_B200_H */
