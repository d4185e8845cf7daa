/* include/skr_mgpu.h -- single-process multi-GPU frame split, on top of include/skr.h.
 *
 * The reference has no multi-GPU path (its only parallelism is the OpenMP row loop, src/main.cpp:33); this is the
 * additive part (3) of the north star: interleaved image tiles across the GPUs of one box.  Frame assembly:
 *   - row bands (default where every GPU pair has peer access, i.e. NVLink boxes): every GPU renders its interleaved tiles
 *     but stores each finished pixel block, over NVLink and while it is still tracing, into the memory of the GPU that OWNS
 *     the block's band of rows (skr_render_bands_device); then every GPU copies its contiguous band to the host itself:
 *     N copies over N PCIe links (SKR_MGPU_NO_BANDS=1 skips this path);
 *   - direct path (no peer access): the caller's host frame is page-locked and mapped (once; cudaHostRegister, or used as it is
 *     when already page-locked) and every GPU's render kernel stores its finished pixel blocks straight into it over its
 *     OWN PCIe link while it is still tracing -- no device frame, no gather, no D2H copy through one GPU;
 *   - peer path (SKR_MGPU_NO_DIRECT=1, when every GPU can map GPU 0's memory): skr_render_peers_device() on every GPU with GPU 0's
 *     frame as the target -- the render kernels store each finished pixel straight into it over NVLink; no collective,
 *     no de-interleave pass, one D2H copy once all kernels are done;
 *   - NCCL path (SKR_MGPU_NO_DIRECT=1 and no peer access, or SKR_MGPU_NO_P2P=1): skr_render_tiles_device(), ONE ncclAllGather of the quantised RGB8
 *     tiles, a de-interleave kernel, one D2H copy.
 * The scene is replicated; the RNG is keyed by pixel, so the frame is byte-identical for any GPU count and either path.
 *
 * One persistent host thread per GPU drives its context (the --gillum wavefront schedules with host read-backs).  For the
 * one-process-per-GPU model (torchrun, MPI) use skr.h directly: skr_render_tiles_device + your all-gather +
 * skr_deinterleave_device (skele_raytracer_b200/distributed.py does that over torch.distributed).
 */
#ifndef SKR_MGPU_H
#define SKR_MGPU_H

#include "skr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skr_mgpu skr_mgpu;

/* n_gpus <= 0: all visible devices.  Devices 0 .. n_gpus-1; peer access to GPU 0 is enabled where possible, NCCL
 * communicators (ncclCommInitAll) are only created when it is not. */
int skr_mgpu_init(int n_gpus, skr_mgpu **out);
void skr_mgpu_destroy(skr_mgpu *m);
const char *skr_mgpu_last_error(const skr_mgpu *m); /* m may be NULL: last error of skr_mgpu_init on this thread */
int skr_mgpu_world(const skr_mgpu *m);

/* Scene replicated on every GPU (each builds its own LBVH: cheaper than broadcasting one). */
int skr_mgpu_scene_upload(skr_mgpu *m, const skr_scene_desc *scene);

/* Renders the frame on all GPUs and returns it in HOST memory (rgb8: H*W*3 bytes, as skr_render).
 * opt->rank / opt->world are ignored (set per GPU).  stats (optional): counters and queue figures summed over the
 * GPUs, ms_total = slowest GPU's render, ms_d2h = copy on GPU 0 (NCCL path: gather + de-interleave + copy). */
int skr_mgpu_render(skr_mgpu *m, const skr_options *opt, uint8_t *rgb8, skr_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
