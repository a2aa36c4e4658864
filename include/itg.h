/*
 * itg.h -- C ABI of libitg_b200.so: B200 (sm_100a) kernels for the patch-by-patch Generator inference path
 * of Infinite_Texture_GANs (local padding).
 *
 * The reference has no native code and no FFI: its "operator interface" for this path is the Python
 * surface of models/layers.py, models/generators.py and utils.py.  Each entry point below names the
 * reference code it replaces (file:line relative to the reference repo).  The Python mirror of that
 * surface (infinite_texture_gans_b200/{layers,generators,utils}.py) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every function returns 0 on success and a negative code on error; itg_last_error() returns a
 *     thread-local message.  No exceptions, no allocation, no implicit device synchronisation.
 *   - the caller owns every buffer (device pointers unless stated otherwise) and passes the CUDA stream
 *     (a cudaStream_t cast to void*).  Launches are stream-ordered and can be captured in a CUDA graph.
 *   - one device per call: the current CUDA device of the calling thread.
 *
 * Data layout ("grid tensors")
 *   The patch grid of the reference (B = nph*npw patches of r x r pixels, NCHW) is stored merged and
 *   channels-last:  a tensor of H x W pixels (H = nph*r, W = npw*r) and C storage channels is a buffer
 *   of (H+2) x (W+2) x C elements: the interior plus a 1-pixel FRAME.  The frame holds what LocalPadder
 *   (models/layers.py:78-101) would concatenate around the merged image: the outer padding (replicate
 *   or zeros), the stored halo row/column of the sequential protocol (models/layers.py:103-143), or a
 *   neighbouring GPU's border row.  A conv therefore never materialises padded patches: tap (dy,dx) of
 *   output pixel (y,x) reads buffer pixel (y+1+dy, x+1+dx).  `in`/`out` pointers below always address
 *   the buffer origin (frame pixel (-1,-1)).  C is a multiple of 8; padded channels hold zeros.
 */
#ifndef ITG_H_
#define ITG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ITG_ABI_VERSION 3

enum itg_status {
  ITG_OK = 0,
  ITG_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  ITG_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed */
  ITG_ERR_UNSUPPORTED = -3   /* valid request this build cannot serve */
};

enum itg_dtype { ITG_F32 = 0, ITG_F16 = 1, ITG_BF16 = 2 };

/* What the conv computes on its M-grid (the pixels it is evaluated on):
 *   ITG_CONV3X3 : 9 taps dy,dx in {-1,0,1}; weight tap t = (dy+1)*3+(dx+1).     conv2d_lp, layers.py:29-36
 *   ITG_CONV1X1 : 1 tap.                                                          conv1x1 shortcut, layers.py:294-299
 *   ITG_UPCONV  : conv3x3(pad(nearest_up2(x))) evaluated as 4 output phases (a,b) of 2x2 taps on the
 *                 low-res framed tensor with pre-summed weights; weight tap = phase*4 + i*2 + j,
 *                 dy = a-1+i, dx = b-1+j, output pixel (2y+a, 2x+b).             generators.py:95-111 + layers.py:29-36
 */
enum itg_conv_mode { ITG_CONV3X3 = 0, ITG_CONV1X1 = 1, ITG_UPCONV = 2 };

/* How the producer fills the frame of the tensors it writes (outer padding, layers.py:82-99):
 * NONE leaves the frame untouched (the caller fills it: sequential halos, neighbour-GPU rows). */
enum itg_border { ITG_BORDER_NONE = 0, ITG_BORDER_REPLICATE = 1, ITG_BORDER_CONSTANT = 2 };

enum itg_residual { ITG_RES_NONE = 0, ITG_RES_GRID = 1 /* framed grid tensor, operand dtype */,
                    ITG_RES_F32 = 2 /* unframed fp32 NHWC */ };

enum itg_impl { ITG_IMPL_AUTO = 0 /* tcgen05 kernels: pair / halo-tile / streaming for 16-bit operands, split-precision for fp32 */,
                ITG_IMPL_DIRECT = 1 /* CUDA-core direct conv (any dtype; the on-device cross-check) */,
                ITG_IMPL_UMMA = 2 /* tcgen05 implicit GEMM, operands streamed per tap (16-bit only) */,
                ITG_IMPL_TILE = 3 /* tcgen05 persistent halo-tile kernel: k_pad <= 64, n_pad <= 64 (16-bit only) */,
                ITG_IMPL_PAIR = 4 /* tcgen05 cta_group::2 halo-tile kernel, weights resident: 3x3 | 1x1, k_pad <= 128, n_pad <= 256 (16-bit only) */,
                ITG_IMPL_SPLIT = 5 /* fp32 tensors on tensor cores: operands split into two fp16 terms, three tcgen05.mma per step (fp32 only) */ };

enum itg_img_layout { ITG_IMG_MERGED = 0 /* (C, H, W) planar */, ITG_IMG_PATCHES = 1 /* (B, C, P, P) */ };

/* One fused convolution launch.  Pointers are device pointers; unused ones are NULL.
 *
 * acc[y,x,n] = sum_taps sum_k W[tap][n][k] * in[y+dy, x+dx, in_c_off+k]            (fp32 accumulate)
 * v          = acc + bias[n] (+ residual)
 * SSM mode (mod_x != NULL, layers.py:228-234): columns are interleaved (gamma_c, beta_c) and
 *   v_c = (1+gamma_c) * (x_c - mod_mean_c) * mod_rstd_c + beta_c,  x read from mod_x at (oy>>mod_shift, ox>>mod_shift)
 * outputs (any subset):
 *   out_raw : v                                   (grid tensor)
 *   out_act : act(scale[n]*v + shift[n])          (grid tensor; BN eval + (Leaky)ReLU, frame per `border`)
 *   out_f32 : v                                   (unframed fp32 NHWC, out_c channels)
 *   out_img : tanh(v[0..img_c))                   (fp32 planar image, generators.py:119-121)
 */
typedef struct itg_conv_desc {
  int32_t dtype;        /* itg_dtype of in / weights / grid outputs */
  int32_t mode;         /* itg_conv_mode */
  int32_t impl;         /* itg_impl */
  int32_t border;       /* itg_border applied to out_act's frame */

  const void* in;       /* framed grid tensor: pixel (-1,-1) of an (in_h+2) x (in_w+2) x in_c window */
  int32_t in_h, in_w;   /* interior size of `in` == M-grid size */
  int32_t in_pitch;     /* pixels between buffer rows of `in`; 0 = in_w + 2 (a window into a wider buffer otherwise) */
  int32_t in_c;         /* storage channels of `in` (multiple of 8) */
  int32_t in_c_off;     /* first channel of the slice this conv reads (multiple of 8) */
  int32_t k;            /* channels contracted per tap (<= k_pad) */

  const void* w;        /* [taps][n_pad][k_pad], operand dtype; taps = 9 | 1 | 16 */
  int32_t n_pad;        /* GEMM N: multiple of 16, zero-padded rows */
  int32_t k_pad;        /* multiple of 16 (64 when k > 64), zero-padded */
  const float* bias;    /* [n_pad] or NULL */

  int32_t out_h, out_w; /* interior size of the outputs (2*in_h, 2*in_w for ITG_UPCONV) */
  int32_t out_c;        /* storage channels of out_raw / out_act / out_f32 (multiple of 8, <= n_pad; n_pad/2 in SSM mode) */

  int32_t res_kind;     /* itg_residual */
  int32_t res_shift;    /* residual read at (oy>>res_shift, ox>>res_shift) */
  const void* res;
  int32_t res_c;        /* storage channels of the residual tensor */
  int32_t res_h, res_w; /* interior size of the residual tensor (ITG_RES_GRID) / its size (ITG_RES_F32) */

  const void* mod_x;    /* SSM: grid tensor to modulate, operand dtype */
  int32_t mod_c, mod_shift, mod_h, mod_w;
  const float* mod_mean;  /* [out_c] running_mean   (layers.py:218) */
  const float* mod_rstd;  /* [out_c] 1/sqrt(running_var+eps) */

  void* out_raw;
  void* out_act;
  const float* scale;   /* [n_pad] BN eval scale  w/sqrt(var+eps); NULL = 1 */
  const float* shift;   /* [n_pad] BN eval shift  b-mean*scale;    NULL = 0 */
  float leak;           /* LeakyReLU slope (0 = ReLU) */
  int32_t act_linear;   /* 1: out_act = scale*v+shift without activation (SSM shortcut bn3) */
  float* out_f32;
  float* out_img;
  int32_t img_c;        /* image channels (generators.py:83) */
  int32_t img_layout;   /* itg_img_layout */
  int32_t patch;        /* P, for ITG_IMG_PATCHES */

  /* Optional second input (ABI v3): the 1x1 shortcut of a residual block folded into its conv2 (layers.py:294-299, 319-320) --
   * result = conv(in, w) + conv1x1(in2, w2) + bias, one accumulator, no intermediate tensor.  ITG_CONV3X3 on the CTA-pair kernel
   * only (ITG_IMPL_PAIR, or ITG_IMPL_AUTO when the layer is eligible); NULL = off. */
  const void* in2;      /* framed grid tensor of the same interior size as `in` */
  int32_t in2_c;        /* its storage channels (multiple of 8) */
  int32_t in2_c_off;    /* first channel of the slice read (multiple of 8) */
  int32_t k2;           /* channels contracted (multiple of 8, <= k2_pad <= 128) */
  const void* w2;       /* [1][n_pad][k2_pad], operand dtype */
  int32_t k2_pad;
} itg_conv_desc;

/* Library / ABI version (ITG_ABI_VERSION). */
int itg_version(void);
/* Message of the last failing call on this thread. */
const char* itg_last_error(void);
/* sizeof(itg_conv_desc) as compiled, so a binding can check its struct mirror. */
int itg_conv_desc_size(void);

/* conv2d_lp.forward + the elementwise ops around it (layers.py:29-36,301-322; generators.py:91-121). */
int itg_conv_fwd(const itg_conv_desc* desc, void* stream);

/* Attention.forward (layers.py:246-258) evaluated per patch of `patch` x `patch` pixels on a grid tensor
 * of th x tw patches; C = channels (C/8 query/key, C/2 value channels).
 * x: framed grid tensor (storage channels xc).  Weights are fp32: w_theta/w_phi [C/8][C], w_g [C/2][C],
 * w_o [C][C/2], biases likewise; gamma is read from device memory (attention.gamma, layers.py:244).
 * out_raw = gamma*o + x;  out_act = act(scale*out_raw + shift) with frame per `border`.  Either may be NULL. */
int itg_attention_fwd(int32_t dtype, const void* x, int32_t th, int32_t tw, int32_t patch, int32_t C, int32_t xc,
                      const float* w_theta, const float* b_theta, const float* w_phi, const float* b_phi,
                      const float* w_g, const float* b_g, const float* w_o, const float* b_o, const float* gamma,
                      void* out_raw, void* out_act, const float* scale, const float* shift, float leak,
                      int32_t border, void* stream);

/* StochasticSpatialModulation.forward (models/layers.py:228-234) as one launch:
 *   m1 = relu(conv3x3_valid(map) + b1)            mlp_shared (layers.py:220-222, 229), 1 -> 128 channels
 *   [gamma|beta] = conv3x3_valid(m1) + b2         embed (layers.py:224, 230), 128 -> 2C channels
 *   out = [act]((1 + gamma) * (x - mean) * rstd + beta)       (layers.py:228, 231-233; the activation of ResBlockGenerator.forward,
 *                                                               layers.py:303-312, unless `linear`, which is the shortcut's bn3)
 * The 128-channel hidden map stays on chip (16-bit operand types only; the fp32 exact mode runs the two convs as itg_conv_fwd launches).
 *   map     : fp32 noise map of this level, (h+4) x (w+4) values, `map_pitch` floats per row (utils.py:246, map_dim = 1 as in test_sample.py:56)
 *   w_mlp   : [128][16] operand dtype, K-major: columns 0..8 = mlp_shared.0.weight[n][0][ky][kx] (tap ky*3+kx), columns 9 and 10 = the bias
 *             mlp_shared.0.bias[n] split into a 16-bit hi + lo pair (the kernel feeds constant ones in those two K slots), rest zero
 *   w_embed : [9][n_pad][128] operand dtype, rows interleaved (gamma_c, beta_c) per stored channel c (2*c <= n_pad, n_pad % 16 == 0)
 *   b_embed : [n_pad] fp32, same interleaving
 *   x       : grid tensor to modulate, x_c storage channels, interior x_h x x_w, read at (y >> x_shift, x >> x_shift)
 *             (x_shift = 1: the nearest-2x up-sampling of generators.py:95-111 folded into the addressing)
 *   mean / rstd : [c] running_mean and 1/sqrt(running_var + eps) of the affine-free BatchNorm (layers.py:218)
 *   out     : grid tensor h x w x c, frame written per `border`
 *   zero_ring : 0 for the local-padding Generator (valid convs on over-sized noise maps, utils.py:237-256).  1 for the non-local Generator
 *             (--padding_mode zeros: mlp_shared and embed are conv3x3(..., p=1), layers.py:213-224): the caller zero-extends the r x r map by
 *             2 px and the kernel forces the hidden map to zero outside the image, which is what zero-padding embed's input means */
typedef struct itg_ssm_desc {
  int32_t dtype;        /* ITG_F16 | ITG_BF16 */
  int32_t border;       /* itg_border applied to out's frame */
  int32_t h, w;         /* interior size of out */
  int32_t c;            /* storage channels of out (multiple of 8) */
  int32_t n_pad;        /* GEMM columns of the embed conv */
  const float* map;
  int32_t map_pitch;
  int32_t x_shift;
  const void* w_mlp;
  const void* w_embed;
  const float* b_embed;
  const void* x;
  int32_t x_c, x_h, x_w;
  int32_t linear;       /* 1: no activation */
  const float* mean;
  const float* rstd;
  void* out;
  float leak;
  int32_t zero_ring;
} itg_ssm_desc;
int itg_ssm_desc_size(void);
int itg_ssm_fwd(const itg_ssm_desc* desc, void* stream);

/* Host-supplied noise -> grid tensor.  src: fp32 planar (C, H, W) (the z grid with its random 1-px ring,
 * utils.py:228, or an SSM map, utils.py:246); dst: (H x W x dst_c) channels-last in `dtype`, channels
 * >= C zero-filled.  The whole H x W extent is copied: the caller interprets the outermost ring as the frame. */
int itg_pack_nchw(int32_t dtype, const float* src, int32_t C, int32_t H, int32_t W, void* dst, int32_t dst_c,
                  void* stream);

/* SSM noise map (utils.py:246, one channel, fp32, Hm x Wm = interior + 4) -> stack of its nine 3x3 taps as a
 * grid tensor with interior (Hm-2) x (Wm-2) and dst_c >= 16 channels (channel t = map shifted by tap t,
 * channels >= 9 zero).  The first SSM conv `mlp_shared` (1 -> 128 channels, layers.py:220,229) is then an
 * ITG_CONV1X1 launch with k = 16. */
int itg_pack_map_taps(int32_t dtype, const float* src, int32_t Hm, int32_t Wm, void* dst, int32_t dst_c,
                      void* stream);

/* Copy a rectangle of pixels between grid tensors (all channels): the halo moves of the sequential
 * protocol (layers.py:103-143) and of the row-band multi-GPU split.  Coordinates are buffer pixels
 * (frame included); pitches are in pixels. */
int itg_copy_rect(int32_t dtype, const void* src, int32_t src_pitch, int32_t sy, int32_t sx,
                  void* dst, int32_t dst_pitch, int32_t dy, int32_t dx, int32_t h, int32_t w, int32_t c,
                  void* stream);

/* Row-band multi-GPU split: halo rows of one conv2d_lp input over peer-mapped memory (NVLink P2P), one launch, no host
 * involvement.  Pushes this rank's first / last interior pixel row into the up / down neighbour's inbox row and
 * publishes *step in the neighbour's flag; waits for the neighbours' flags to reach *step and copies the local inbox
 * rows into the top / bottom frame row of `grid`.  NULL inbox = no neighbour on that side.  `roles` selects what this
 * launch does (bit 0 push up, 1 push down, 2 pull top, 3 pull bottom; 15 = everything) so that the pushes can be
 * issued right after the producer and the pulls right before the consumer, with independent work in between.  `up_*` / `down_*` are
 * pointers into the NEIGHBOURS' memory (CUDA IPC mappings), `top_*` / `bot_*` are local.  Replaces the `.cpu()` / `.to(device)`
 * halo hand-off of LocalPadder.update_padding_variables (layers.py:117-139) for the multi-GPU case. */
int itg_halo_exchange(int32_t dtype, void* grid, int32_t h, int32_t w, int32_t c, void* up_inbox, void* down_inbox,
                      int32_t* up_flag, int32_t* down_flag, const void* top_inbox, const void* bot_inbox,
                      int32_t* top_flag, int32_t* bot_flag, const int32_t* step, int32_t roles, void* stream);
/* Advance the device-resident step counter that itg_halo_exchange publishes / waits for (once per Generator pass). */
int itg_step_advance(int32_t* step, void* stream);

/* Exchange buffers for itg_halo_exchange: device memory of THIS process that neighbour ranks map through CUDA IPC
 * (one process per GPU).  alloc zero-fills `bytes` on `device` and returns the pointer plus a 64-byte handle to send to
 * the neighbours; open maps a neighbour's handle into this process (peer access is enabled lazily by the runtime). */
int itg_ipc_alloc(int32_t device, uint64_t bytes, void** ptr, void* handle64);
int itg_ipc_open(int32_t device, const void* handle64, void** ptr);
int itg_ipc_close(void* ptr);
int itg_ipc_free(void* ptr);

/* Fill the frame of a grid tensor from its interior (replicate) or with zeros (constant): F.pad of
 * layers.py:82.  sides: bit0 top, bit1 bottom, bit2 left, bit3 right. */
int itg_fill_frame(int32_t dtype, void* t, int32_t h, int32_t w, int32_t c, int32_t border, int32_t sides,
                   void* stream);

/* Counter-based replacement of the host-side noise draw (utils.py:228, 246: torch.randn of the whole z grid / noise map, then .to(device)):
 * fills dst (C, h, w) fp32, contiguous, with the window [y0, y0+h) x [x0, x0+w) of a standard-normal field of C x Hf x Wf values that is a
 * pure function of (seed, field, element index): Philox4x32-10 keyed by seed, counter = (element / 4, field), Box-Muller.  Ranks / sub-images
 * generate exactly the part they consume; windows of one field agree bit for bit wherever they overlap.  field: 0 = z, 1 + i = map of level i.
 * Same distribution as torch.randn, not the same stream. */
int itg_noise_normal(float* dst, int32_t C, int32_t h, int32_t w, int32_t y0, int32_t x0, int32_t Hf, int32_t Wf, uint64_t seed, uint32_t field,
                     void* stream);

/* Output stage of test_sample.py:75-79: `save_image(img * 0.5 + 0.5, path)` quantises the fp32 image with
 * torchvision's `mul(255).add_(0.5).clamp_(0, 255).to(uint8)` after the `* 0.5 + 0.5`, on the host.  This does the
 * same arithmetic (same fp32 operations in the same order, no FMA contraction: bit-identical bytes) on the device, so
 * that 1 byte instead of 4 crosses PCIe per sample.  img: planar fp32 (c, h, w) with `row_pitch` floats between rows
 * and `plane_pitch` floats between channels (a cropped view of the Generator's output buffer); out: interleaved
 * (h, w, c) uint8, contiguous -- the layout PIL / image encoders take.  c <= 8 (the Generator's img_ch range). */
int itg_image_to_u8(const float* img, int32_t c, int32_t h, int32_t w, int64_t row_pitch, int64_t plane_pitch,
                    uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ITG_H_ */
