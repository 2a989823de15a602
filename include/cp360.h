/*
 * cp360.h — C ABI of libcp360.so: the B200 (sm_100a) spherical-projection hot path of
 * hsientzucheng/CP-360-Weakly-Supervised-Saliency.
 *
 * The reference is pure Python and has no FFI layer; its boundary for this path is three
 * classes:
 *     CubePad(lrtd_pad).forward(x)            model/cube_pad.py:23-42
 *     Equi2Cube(w, img).to_cube(img)          utils/equi_to_cube.py:11-129
 *     Cube2Equi(w).to_equi_nn(cube)           utils/cube_to_equi.py:11-66
 * Each entry point below replaces the body of one of those methods (file:line cited per
 * function). A maintainer binds them with ctypes; the host mirror in
 * cp-360-weakly-supervised-saliency_b200/ does exactly that.
 *
 * Conventions
 *   - plain pointers + sizes; no torch / CUDA types in signatures. `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).
 *   - pointers named *_dev are device pointers, *_host are host pointers; the library never
 *     allocates or frees caller memory and keeps no reference after the call returns.
 *   - every call returns a cp360_status (0 = ok). Kernels are launched asynchronously on
 *     `stream`; launch errors are reported, execution errors surface at the caller's next sync.
 *   - thread-safe for distinct streams. Global state, all mutex- or atomic-guarded: a thread-local
 *     last-error string, the launch counter, the per-device pool of work counters (common.cu), the
 *     CubePad tiling table (cubepad.cu) and the registry of cp360_host_alloc buffers.
 *   - there is NO CPU fallback: without a CUDA device the *_fwd/_bwd calls return
 *     CP360_ERR_CUDA. The *_build_* calls are host-only (map construction, once per
 *     resolution) and work anywhere.
 *
 * Face order everywhere: 0=Back 1=Down 2=Front 3=Left 4=Right 5=Top (cube_pad.py:49).
 */
#ifndef CP360_H_
#define CP360_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CP360_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define CP360_API __attribute__((visibility("default")))
#else
#define CP360_API
#endif

typedef enum cp360_status {
  CP360_OK = 0,
  CP360_ERR_BAD_ARG = 1,      /* null pointer, negative size, unsupported elem size ...        */
  CP360_ERR_GROUP = 2,        /* batch is not a multiple of 6 ("CubePad size mismatch!",       */
                              /*   cube_pad.py:33-35 prints and exit()s; here: error code)     */
  CP360_ERR_SHAPE = 3,        /* H != W, pad > H, Hin*2 != Win (equi_to_cube.py:15 assert) ... */
  CP360_ERR_RANGE = 4,        /* map value outside the representable / interpolation range     */
  CP360_ERR_CUDA = 5,         /* CUDA runtime / launch failure; see cp360_last_error()         */
  CP360_ERR_ALIGN = 6         /* pointer not aligned to the element size                       */
} cp360_status;

CP360_API int cp360_version(void);
CP360_API const char* cp360_status_string(int status);
/* Thread-local, human readable detail of the last non-OK status on this thread ("" if none). */
CP360_API const char* cp360_last_error(void);
/* Number of kernels launched by this library since load (all threads); for bench accounting. */
CP360_API uint64_t cp360_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * CubePad — model/cube_pad.py:23-216
 * ---------------------------------------------------------------------------------------- */

/* Host: output extent. Replaces the implicit shape arithmetic of cube_pad.py:179-215. */
CP360_API int cp360_cubepad_out_shape(int H, int W, int pl, int pr, int pt, int pd, int* Ho, int* Wo);

/* Host: the integer index map the kernels implement, as a table.
 * map_host[6*Ho*Wo] (int32): for output pixel (face f, oy, ox) the flat index
 * (src_face*H*W + src_row*W + src_col) into one channel's [6,H,W] cube.
 * This is the closed form of cube_pad.py:106-215 (plates :114-162, corners :165-176). */
CP360_API int cp360_cubepad_build_map(int H, int W, int pl, int pr, int pt, int pd, int32_t* map_host);

/* Device: y[6N,C,Ho,Wo] = CubePad(x[6N,C,H,W]); contiguous NCHW, any element size in
 * {1,2,4,8,16} bytes (pure data movement). n_faces = 6N must be a multiple of 6.
 * Replaces CubePad.forward, cube_pad.py:28-42 (+ CubePadding.forward :95-216).
 * Pad order is the reference's: l, r, t, d. */
CP360_API int cp360_cubepad_fwd(const void* x_dev, void* y_dev, int64_t n_faces, int64_t C, int H, int W,
                      int pl, int pr, int pt, int pd, int elem_bytes, void* stream);

/* Same, with the kernel forced (tests / benchmarks): 0 auto, 1 generic gather, 2 band kernel with
 * streaming stores, 3 band kernel with TMA bulk stores, 4 cube-tile kernel (all six faces of a
 * channel group staged in shared memory), 5 row kernel (TMA bulk loads, warp-specialised). Non-applicable choices return CP360_ERR_SHAPE/ALIGN. */
CP360_API int cp360_cubepad_fwd_algo(const void* x_dev, void* y_dev, int64_t n_faces, int64_t C, int H, int W,
                           int pl, int pr, int pt, int pd, int elem_bytes, int algo, void* stream);

/* Host: the kernel cp360_cubepad_fwd (algo 0) runs for this problem: 1 generic, 3 band, 4
 * cube-tile (v1), 5 row, 6 cube-tile; aligned16 = both tensor pointers are 16 B aligned. Negative: -status. */
CP360_API int cp360_cubepad_pick_algo(int64_t n_faces, int64_t C, int H, int W, int pl, int pr, int pt,
                            int pd, int elem_bytes, int aligned16);

/* Device, fp32: y = CubePad(act(x * scale[c] + shift[c])) in one pass — the eval-mode BatchNorm affine + ReLU that
 * precede CubePad at model/resnet_cubic.py:89-92 (separate multiply and add; act = ReLU if relu != 0). scale_dev /
 * shift_dev: device float[C] or NULL (1 / 0). out_C > 0: y has out_C >= C channels and the padded planes land in
 * channels [out_c_off, out_c_off + C) — CubePad of a channel concatenation (model/clstm.py:57-58) written one
 * source at a time, without materialising the cat. out_C == 0: y has C channels. Same status codes as cp360_cubepad_fwd. */
CP360_API int cp360_cubepad_fused_fwd(const float* x_dev, float* y_dev, int64_t n_faces, int64_t C, int H, int W,
                            int pl, int pr, int pt, int pd, const float* scale_dev, const float* shift_dev,
                            int relu, int64_t out_C, int64_t out_c_off, void* stream);

/* Device, fp32, EXPLICIT tuning — the one CubePad call that allocates and synchronises: times candidate tilings
 * of this problem on the caller's tensors (y is written with the correct result by every candidate; a 160 MB
 * flush buffer and two events are created and destroyed inside the call, which waits for its own launches)
 * and remembers the winner for later cp360_cubepad_fwd calls with the same (device, geometry, C, n_faces).
 * effort 1..8 scales the repetitions. Not allowed during stream capture. cp360_cubepad_fwd itself never tunes:
 * it uses, in this order, a result of this call, the built-in table measured on B200 for the cubic-ResNet-50 /
 * ConvLSTM shapes (csrc/cubepad_tuned.h), and shape heuristics. (CP360_AUTOTUNE=1 restores implicit
 * first-call tuning for experiments.) */
CP360_API int cp360_cubepad_autotune(const void* x_dev, void* y_dev, int64_t n_faces, int64_t C, int H, int W,
                           int pl, int pr, int pt, int pd, int effort, void* stream);

/* Host: register a tiling for one problem on the current device (what cp360_cubepad_autotune does with its winner;
 * used by tools/tune_chain.py, which times candidates inside a whole chain of kernels instead of in isolation).
 * algo 5 = row kernel (row_rb rows per band or row_tile_kb KB of whole planes per tile, dealing order 0 / 2, ring
 * depth), 6 = cube-tile kernel (stage KB, stages, consumer warps); 0 = forget the registration. A tiling that does
 * not apply to the problem makes the launch fall back to the other kernel's heuristics. No allocation, no sync. */
CP360_API int cp360_cubepad_set_tiling(int64_t n_faces, int64_t C, int H, int W, int pl, int pr, int pt, int pd,
                             int algo, int row_rb, int row_order, int row_slots, int row_tile_kb,
                             int cube_stage_kb, int cube_stages, int cube_warps);

/* Host: human-readable tiling cp360_cubepad_fwd uses for this problem on the current device and where it
 * came from ("" = shape heuristics: no autotune result and no table row for this site). */
CP360_API int cp360_cubepad_tune_info(int64_t n_faces, int64_t C, int H, int W, int pl, int pr, int pt, int pd,
                            char* buf, int buf_len);

/* Host: the transpose of cp360_cubepad_build_map in CSR form. For source pixel s = (face*H + y)*W + x of
 * one channel's cube, entries_host[offsets_host[s] .. offsets_host[s+1]) are the flat output indices
 * (face'*Ho*Wo + oy*Wo + ox) that copy it: the interior copy first, then the halo copies in the order
 * cp360_cubepad_bwd_f32 sums them. offsets_host[6*H*W + 1]; entries_host[6*Ho*Wo] (every output pixel
 * appears exactly once) or NULL to only count. */
CP360_API int cp360_cubepad_build_inverse_map(int H, int W, int pl, int pr, int pt, int pd, int32_t* offsets_host,
                                    int32_t* entries_host);

/* Device, fp32: gx[6N,C,H,W] = dCubePad^T(gy[6N,C,Ho,Wo]) — every input pixel receives the sum
 * of the gradients of all output pixels that copied it (what autograd derives from the
 * cat/index_select/repeat chain; needed by temporal_model/train_temporal.py:167-170). No atomics
 * (faces up to 32 px: one pass over the staged padded gradient of a whole cube; larger faces: a copy
 * kernel for the face interiors and a gather kernel for the pixels that halo positions copy): the sum
 * runs in the fixed order of cp360_cubepad_build_inverse_map, so gradients are reproducible bit for bit.
 * The small-face kernel keeps its position tables (a few KB, a pure function of H, W and the pads) in a
 * per-device cache: the FIRST call for a geometry allocates them, builds them with a one-CTA launch on
 * `stream` and synchronises that stream once; later calls do neither. A geometry first seen while
 * `stream` is being captured builds the tables inside every CTA instead (no allocation, no
 * synchronisation during capture). CP360_BWD_TABLE_CACHE=0 disables the cache. */
CP360_API int cp360_cubepad_bwd_f32(const float* gy_dev, float* gx_dev, int64_t n_faces, int64_t C, int H,
                          int W, int pl, int pr, int pt, int pd, void* stream);

/* ------------------------------------------------------------------------------------------
 * Equi2Cube — utils/equi_to_cube.py:11-129
 * ---------------------------------------------------------------------------------------- */

/* Host: sampling map for Hin x Win (Win == 2*Hin) -> 6 faces of w x w, vertical fov in degrees.
 * Replaces Equi2Cube.__init__, equi_to_cube.py:12-110 (float64, table-lookup inverse trig) and
 * the float32 cast + cv2 fixed-point conversion of :122-125 / cv2.remap.
 * Outputs (any may be NULL):
 *   packed_host[cp360_e2c_map_words(w,Hin,Win)] uint32 — the device map. Frames up to 2047 x 1023 (the
 *     reference's 1920 x 960): one word per pixel, x0<<20 | y0<<10 | fx<<5 | fy (x0 = sx>>5, fx = sx&31, ...).
 *     Larger frames (4K / 8K equirects): two words per pixel, x0<<16 | y0 then fx<<5 | fy; the device
 *     copy must then be 8 B aligned. The *_fwd calls pick the form from Hin / Win the same way.
 *   sx_host, sy_host[6*w*w] int32   cvRound(float32(inX)*32), cvRound(float32(inY)*32)
 *   inx_host, iny_host[6*w*w] double  the reference's self.inXs / self.inYs (1-based coords)
 * CP360_ERR_RANGE beyond 65535 x 32767. */
CP360_API int64_t cp360_e2c_map_words(int w, int Hin, int Win);
CP360_API int cp360_e2c_build_map(int w, int Hin, int Win, double vfov_deg, uint32_t* packed_host,
                        int32_t* sx_host, int32_t* sy_host, double* inx_host, double* iny_host);

#define CP360_LAYOUT_NCHW 0 /* faces[(b*6+f), c, y, x]   — what the cubic ResNet consumes        */
#define CP360_LAYOUT_NHWC 1 /* faces[(b*6+f), y, x, c]   — the reference's dict of w x w x C     */

/* Device, fp32: bilinear resampling with cv2.remap(INTER_LINEAR) fixed-point semantics
 * (1/32-pixel weights, fp32, no FMA, BORDER_CONSTANT 0).
 * frames_dev [B,Hin,Win,C] channel-contiguous; faces_dev [6B,...] in `out_layout`.
 * If mean_host/std_host are non-NULL (C floats each) the per-channel normalisation
 * (v - mean[c]) / std[c] of utils/utils.py:28-33 (im_norm, dataset_feat_extractor.py:148-151)
 * is fused into the store.
 * Replaces Equi2Cube.to_cube, equi_to_cube.py:112-129. */
CP360_API int cp360_e2c_fwd(const float* frames_dev, const uint32_t* packed_dev, float* faces_dev,
                  int64_t B, int Hin, int Win, int C, int w, int out_layout,
                  const float* mean_host, const float* std_host, void* stream);

/* Device: the same resampling for uint8 frames [B,Hin,Win,C] (a decoded / resized video frame,
 * dataset_feat_extractor.py:126-142, before its "/255.0"): pixel = float32(u8) / denom, which for
 * denom = 255 equals the reference's float32(u8 / 255.0) for every code; then identical arithmetic,
 * so faces are bit-identical to cp360_e2c_fwd on the converted frame. A quarter of the host->device
 * and DRAM read traffic. Fused normalisation needs C == 3. */
CP360_API int cp360_e2c_fwd_u8(const uint8_t* frames_dev, const uint32_t* packed_dev, float* faces_dev,
                     int64_t B, int Hin, int Win, int C, int w, int out_layout, float denom,
                     const float* mean_host, const float* std_host, void* stream);

/* Device: padded[6B,C,w+pt+pd,w+pl+pr] = CubePad(im_norm(to_cube(frame))) in ONE kernel — the chain
 * dataset_feat_extractor.py:145-157 (to_cube, im_norm, stack, NHWC->NCHW) + CubePad(3) in front of conv1
 * (model/resnet_cubic.py:116-117): the faces tensor is never materialised, every padded pixel is
 * resampled through the map entry of the face pixel CubePad would have copied (same geometry table
 * as cp360_cubepad_fwd), so the result is bit-identical to cp360_e2c_fwd followed by cp360_cubepad_fwd.
 * frames_u8 != 0: frames_dev is uint8 [B,Hin,Win,C] converted as float32(u8)/denom (see cp360_e2c_fwd_u8),
 * else float [B,Hin,Win,C] and denom is ignored. mean/std as in cp360_e2c_fwd. Output is NCHW.
 * Pad order l, r, t, d (cube_pad.py:12-20); CP360_ERR_SHAPE if a pad exceeds w. */
CP360_API int cp360_e2c_cubepad_fwd(const void* frames_dev, int frames_u8, const uint32_t* packed_dev,
                          float* padded_dev, int64_t B, int Hin, int Win, int C, int w, int pl, int pr,
                          int pt, int pd, float denom, const float* mean_host, const float* std_host,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Cube2Equi — utils/cube_to_equi.py:11-66
 * ---------------------------------------------------------------------------------------- */

/* Host: replaces Cube2Equi.__init__, cube_to_equi.py:12-35 (+ sph_utils.py:53-153).
 *   face_host[2w*4w] int8 (0..5), coord_host[2w*4w*2] double (x,y in face pixels). */
CP360_API int cp360_c2e_build_map(int w, int8_t* face_host, double* coord_host);

/* Host: the fp32 sampling plan to_equi_nn implies (cube_to_equi.py:58-64 + grid_sample):
 *   M = max(float32(coord)); gn = (g - M/2)/(M/2); unnormalise (align_corners 0/1), floor,
 *   four bilinear weights in torch's order nw, ne, sw, se — all in fp32.
 *   tap_host[2w*4w] uint32: face<<28 | (y0+1)<<14 | (x0+1)   (x0,y0 in [-1, w-1])
 *   wts_host[2w*4w*4] float.  M_out: the data-dependent normaliser (may be NULL). */
CP360_API int cp360_c2e_build_plan(int w, int align_corners, uint32_t* tap_host, float* wts_host,
                         float* M_out);

/* Device, fp32: equi[B,C,2w,4w] from cube[6B,C,w,w]; out-of-face taps contribute 0
 * (padding_mode='zeros'). Replaces Cube2Equi.to_equi_nn, cube_to_equi.py:37-66. */
CP360_API int cp360_c2e_fwd(const float* cube_dev, const uint32_t* tap_dev, const float* wts_dev,
                  float* equi_dev, int64_t B, int64_t C, int w, void* stream);

/* Device, fp32: sal[B,2w,4w] = max over channels of the above, without materialising it
 * (test_temporal.py:82-84, train_temporal.py:105-106, dataset_feat_extractor.py:174-175).
 * NaN semantics are torch.max's: a NaN in any channel of a pixel makes that pixel NaN.
 * w <= 16: one kernel, no atomics — a thread-block cluster per frame splits the channels, the partial maps meet in
 * distributed shared memory and are combined with warp shuffles; bit-reproducible. Larger faces: a -inf fill
 * plus an order-preserving atomic max. */
CP360_API int cp360_c2e_max_fwd(const float* cube_dev, const uint32_t* tap_dev, const float* wts_dev,
                      float* sal_dev, int64_t B, int64_t C, int w, void* stream);

/* Device, fp32: the same channel max together with the channel it came from — the differentiable
 * form the training path needs (train_temporal.py:105-107 backprops through to_equi_nn + torch.max):
 * sal[B,2w,4w], argmax[B,2w,4w] int32 (lowest channel among equal maxima, the first NaN channel if any, as
 * torch.max(dim)). scratch: caller-owned uint64[B*2w*4w] work buffer, needed (and overwritten) only for
 * w > 16; may be NULL for smaller faces. */
CP360_API int cp360_c2e_max_arg_fwd(const float* cube_dev, const uint32_t* tap_dev, const float* wts_dev,
                          float* sal_dev, int32_t* argmax_dev, uint64_t* scratch_dev, int64_t B, int64_t C,
                          int w, void* stream);

/* Host: the TRANSPOSE of the sampling plan, for the backward passes: for cube pixel s = (face*w + y)*w + x,
 * entries [offsets_host[s], offsets_host[s+1]) list the equirect pixels that sample it (pix_host, increasing) and
 * the bilinear weight each applies (wts_host). offsets_host[6*w*w + 1]; pix_host / wts_host hold
 * offsets_host[6*w*w] <= 32*w*w entries (pass NULL for both to only count). */
CP360_API int cp360_c2e_build_bwd_plan(int w, int align_corners, int32_t* offsets_host, int32_t* pix_host,
                             float* wts_host);

/* Device, fp32: backward of the fused channel max: gcube[6B,C,w,w] (zero-filled here) receives
 * gsal[b,pix] * weight at the four taps of channel argmax[b,pix] — what autograd yields for
 * torch.max(to_equi_nn(x), 1)[0] without materialising the [B,C,2w,4w] map or its gradient. Gather over the
 * transposed plan (cp360_c2e_build_bwd_plan, device copies): no atomics, fixed summation order. */
CP360_API int cp360_c2e_max_bwd(const float* gsal_dev, const int32_t* argmax_dev, const int32_t* offsets_dev,
                      const int32_t* pix_dev, const float* bwts_dev, float* gcube_dev, int64_t B, int64_t C, int w,
                      void* stream);

/* Host: sampling plan of Cube2Equi.to_equi_cv2, cube_to_equi.py:68-91 — cv2.remap(INTER_CUBIC) fed
 * float32(out_coord) as it is (face pixels, no normalisation): s = cvRound(float32(coord)*32),
 *   tap_host[2w*4w] uint32: face<<28 | (sy&31)<<23 | (sx&31)<<18 | (sy>>5)<<9 | (sx>>5)
 * ((s>>5)-1 is the origin of the 4x4 window, s&31 the row of OpenCV's bicubic table). w <= 512. */
CP360_API int cp360_c2e_build_cubic_plan(int w, uint32_t* tap_host);

/* Device, fp32: equi[B,C,2w,4w] from cube[6B,C,w,w] with cv2.remap(INTER_CUBIC, BORDER_CONSTANT 0)
 * arithmetic (A = -0.75 table, 4x4 fp32 weight products, OpenCV's summation order, no FMA), each
 * output pixel sampled from the face face_map assigns it. Replaces Cube2Equi.to_equi_cv2,
 * cube_to_equi.py:68-91 (6 x 250 cv2.remap calls + boolean-mask scatters); any C (the reference
 * hard-codes 1000 channels, :88). */
CP360_API int cp360_c2e_cubic_fwd(const float* cube_dev, const uint32_t* tap_dev, float* equi_dev, int64_t B,
                        int64_t C, int w, void* stream);

/* Device, fp32: gcube[6B,C,w,w] = d(c2e)^T(gequi[B,C,2w,4w]) — the gradient of cp360_c2e_fwd
 * (train_temporal.py:167-170 back-propagates through to_equi_nn). A gather over the transposed plan
 * (cp360_c2e_build_bwd_plan, device copies): every cube pixel sums its contributors in a fixed order, no
 * atomics — bit-reproducible, and gcube needs no zero fill. n_entries = offsets_host[6*w*w], the length of
 * pix_dev / bwts_dev (the launch sizes its shared-memory copy of the plan by it). */
CP360_API int cp360_c2e_bwd(const float* gequi_dev, const int32_t* offsets_dev, const int32_t* pix_dev,
                  const float* bwts_dev, int n_entries, float* gcube_dev, int64_t B, int64_t C, int w,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * .npy files either side of the path: cube score files `cube_feat/%06d.npy` [6,1000,7,7]
 * (static_model/dataset_feat_extractor.py:187-189 writes, temporal_model/test_temporal.py:64,70 and
 * data/dataset.py:65 read) and equirect result maps `%05d.npy` [14,28] (test_temporal.py:86-88).
 * Host-only calls.
 * ---------------------------------------------------------------------------------------- */

/* Parse the header of a .npy file (format 1.0 / 2.0 / 3.0). Any output may be NULL.
 * descr: dtype string such as "<f4"; shape[max_dims]; data_offset: byte offset of the array data. */
CP360_API int cp360_npy_read_header(const char* path, char* descr, int descr_len, int* ndim, int64_t* shape,
                          int max_dims, int64_t* data_offset, int* fortran_order);

/* Read a C-order little-endian f4 / f8 / f2 / u1 / i4 / i8 array into dst_host[n_elems] as float32
 * (what torch.FloatTensor(np.load(path)) yields, test_temporal.py:70-78). n_elems must equal the
 * file's element count (CP360_ERR_SHAPE otherwise). dst_host may be pinned memory. */
CP360_API int cp360_npy_read_f32(const char* path, float* dst_host, int64_t n_elems);

/* Write src_host as a float32 C-order .npy, byte-identical to numpy.save (format 1.0). Written to
 * a temporary file and renamed, so readers never see a partial file. */
CP360_API int cp360_npy_write_f32(const char* path, const float* src_host, int ndim, const int64_t* shape);

/* ------------------------------------------------------------------------------------------
 * Host staging memory for the end-to-end path: the reference reads decoded video frames into host
 * arrays (static_model/dataset_feat_extractor.py:119-142) and copies per frame (class_activation_model.py:58).
 * Page-locked buffers let the copy engines stream them at PCIe speed.
 * mode 0: cudaHostAlloc(portable); 1: + write-combined; 2: 2 MB-aligned anonymous mapping with
 * MADV_HUGEPAGE, touched and cudaHostRegister'ed (fewer IOMMU translations per byte on multi-GPU hosts).
 * These two calls are the only ones in the library that allocate; the memory belongs to the caller
 * until cp360_host_free.
 * ---------------------------------------------------------------------------------------- */
CP360_API int cp360_host_alloc(uint64_t bytes, int mode, void** out_host);
CP360_API int cp360_host_free(void* host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* CP360_H_ */
