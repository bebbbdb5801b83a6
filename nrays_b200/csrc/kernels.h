// kernels.h — launch wrappers of kernels.cu (host side sees only these).
#pragma once
#include "device_types.cuh"

namespace nrb {

constexpr int kTraceBlock = 128;
#ifndef NRB_FETCH_PACKETS
#define NRB_FETCH_PACKETS 2
#endif
#ifndef NRB_TRACE_MIN_BLOCKS
#define NRB_TRACE_MIN_BLOCKS 10
#endif
#ifndef NRB_SMALL_QUEUE
#define NRB_SMALL_QUEUE (3u << 20)
#endif
constexpr unsigned kSmallQueue = NRB_SMALL_QUEUE;  // default TraceOpts.small_queue
constexpr int kFetchPackets = NRB_FETCH_PACKETS;  // 32-ray packets a warp takes per cursor atomic
constexpr int kTraceMinBlocks = NRB_TRACE_MIN_BLOCKS;  // resident CTAs / SM the trace kernel is compiled for: 10 -> 48 registers, 8 bytes of
                                                      // spill outside the visit loop; measured 7 / 8 / 9 / 10 / 12 on C3 2.22 / 2.22 / 2.23 /
                                                      // 2.21 / 2.34 ms, on C4 7.62 / 7.32 / 7.13 / 6.95 / 7.32 ms (profiles/r2_experiments.txt)
#ifndef NRB_TAIL_MIN_BLOCKS
#define NRB_TAIL_MIN_BLOCKS 6
#endif
constexpr int kTailMinBlocks = NRB_TAIL_MIN_BLOCKS;  // resident CTAs / SM of the tail kernel (mesh-only scenes): 6 -> 80 registers with ~200 B of
                                                   // spills beats 4 (114 registers, no spills) by 2.5 % of the C3 frame: the chains are latency-bound
#ifndef NRB_SHADE_BLOCK
#define NRB_SHADE_BLOCK 128
#endif
constexpr int kShadeBlock = NRB_SHADE_BLOCK;  // >= 64 (two reservation leaders)
#ifndef NRB_SHADE_MIN_BLOCKS
#define NRB_SHADE_MIN_BLOCKS 6
#endif
constexpr int kShadeMinBlocks = NRB_SHADE_MIN_BLOCKS;  // 6: shade capped at 80 registers -> 768 resident threads / SM

// Run-time knobs of one trace launch.
struct TraceOpts {
  int min_active_closest;  // dynamic fetch: a warp refills its idle lanes once fewer than this many are traversing (0: never early)
  int min_active_shadow;
  int reverse_shadow;      // opaque any-hit phase of shadow rays walked from the light end
  uint32_t small_queue;    // queues shorter than this are fetched one 32-ray packet at a time (else kFetchPackets packets)
};
void launch_trace(const SceneView &sc, bool has_shapes, const FrameParams &fp, bool primary, RayQueue q, float4 *hits,
                  WaveCounters *wc_closest, uint32_t slot_lo, uint32_t n_slots, ShadowQueue sq, float4 *accum,
                  WaveCounters *wc_shadow, TraceOpts opts, int grid, cudaStream_t st);
// scenes with depth-shift (nmap) nodes: general packet kernel, every lane runs the whole query
void launch_trace_general(const SceneView &sc, const FrameParams &fp, bool primary, RayQueue q, float4 *hits,
                          WaveCounters *wc_closest, uint32_t slot_lo, uint32_t n_slots, ShadowQueue sq, float4 *accum,
                          WaveCounters *wc_shadow, int grid, cudaStream_t st);
void launch_shade(const SceneView &sc, bool has_shapes, const FrameParams &fp, bool primary, RayQueue qin,
                  const float4 *hits, WaveCounters *wc, uint32_t slot_lo, uint32_t n_slots, uint32_t lo, uint32_t hi,
                  RayQueue qout, ShadowQueue sq, Counters *ctr, float4 *accum, int grid, cudaStream_t st);
void launch_tail(const SceneView &sc, bool has_shapes, const FrameParams &fp, RayQueue qin, WaveCounters *wc,
                 RayQueue qspill, ShadowQueue sq, Counters *ctr, float4 *accum, WaveCounters *wc_sh, int grid,
                 cudaStream_t st);
int shade_blocks_per_sm(bool has_shapes);
void launch_resolve(const float4 *accum, uint32_t n, uint32_t spp, float *out_rgb, cudaStream_t st);
void launch_resolve_rgb8(const float4 *accum, uint32_t n, uint32_t spp, uint8_t *out, cudaStream_t st);
void launch_patch_host_image(const float4 *accum, const float *early_rgb, uint32_t n, uint32_t spp, float *host_rgb, cudaStream_t st);
void launch_resolve_tiles_to_image(const float4 *accum, const FrameParams &fp, float *out_rgb, cudaStream_t st);
void launch_resolve_tiles_to_segments(const float4 *accum, const FrameParams &fp, uint32_t cols_per_rank, float *stage, cudaStream_t st);
void launch_resolve_tiles_to_image_rgb8(const float4 *accum, const FrameParams &fp, uint8_t *out_rgb8, cudaStream_t st);
void launch_resolve_tiles_to_segments_rgb8(const float4 *accum, const FrameParams &fp, uint32_t cols_per_rank, uint8_t *stage, cudaStream_t st);
void launch_untile(const float *gathered, uint32_t n_ranks, uint32_t tiles_per_rank, uint32_t width, uint32_t height,
                   float *out_rgb, cudaStream_t st);
int trace_blocks_per_sm(bool has_shapes);
void debug_visit_counters(unsigned long long out[4], bool reset);

}  // namespace nrb
