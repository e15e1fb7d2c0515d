/* uav_b200.h — C ABI of libuav_b200.so: the sm_100a kernels behind the Upscale-A-Video
 * diffusion sampling hot path.
 *
 * Conventions
 *  - Plain C: raw device pointers + explicit shapes. The library never allocates, frees or
 *    retains caller memory; every call is stream-ordered on `stream` and never synchronises.
 *  - Activations are channels-last: a "b c t h w" tensor of the reference is stored as
 *    [b][t][h][w][c] (c contiguous). `ld_*` arguments are the element distance between two
 *    consecutive pixels (>= channel count), so a tensor may be a channel slice of a wider
 *    buffer (this is how the skip-connection torch.cat of unet_blocks.py:573,645 disappears).
 *  - Weights are K-major fp16: conv weight (Cout,Cin,kh,kw) of the reference is passed as
 *    [Cout][kh][kw][Cin]; Conv3d (Cout,Cin,kt,kh,kw) as [Cout][kt][kh][kw][Cin]; Linear as
 *    [Cout][Cin] (unchanged). Bias is fp32.
 *  - Every function returns uav_status_t; on failure uav_last_error_string() describes it.
 *    No exceptions, no exit(), no CPU fallback.
 *
 * Each entry point cites the reference code (relative to /root/reference/models_video/)
 * whose device work it replaces.
 */
#ifndef UAV_B200_H_
#define UAV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  UAV_OK = 0,
  UAV_ERR_INVALID = 1,     /* bad argument (shape / alignment / null) */
  UAV_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed */
  UAV_ERR_UNSUPPORTED = 3  /* valid but not implemented for this configuration */
} uav_status_t;

typedef enum { UAV_F16 = 0, UAV_F32 = 1 } uav_dtype_t;
typedef enum { UAV_ACT_NONE = 0, UAV_ACT_SILU = 1, UAV_ACT_GEGLU = 2 } uav_act_t;

typedef void* uav_stream_t; /* cudaStream_t */

const char* uav_version(void);
const char* uav_last_error_string(void);
/* number of kernels this library has launched in the calling process (bench: gpu_launches) */
uint64_t uav_launch_count(void);

/* Fused epilogue of every implicit-GEMM entry point:
 *   v = acc + bias[n] + rowvec[row / rows_per_vec][n]
 *   v = act(v)            (GEGLU: out[n] = v[n] * gelu(v[n + N/2]), output has N/2 columns)
 *   v = v + residual[row][n]
 *   out[row][n] = (out_dtype) v
 * Replaces the separate ATen kernels for "+ temb[:, :, None, None, None]" (resnet.py:272-276),
 * "(input_tensor + hidden_states) / output_scale_factor" (resnet.py:292, scale == 1 in every
 * shipped config), "attn(...) + hidden_states" (attention.py:531,537,549,559,563),
 * "hidden_states + residual" (attention.py:405), "input_tensor + hidden_states * w"
 * (temporal_module.py:192, w == 1) and GEGLU (diffusers_attention.py:802-823). */
typedef struct {
  const float* bias;     /* [N] or NULL */
  const void* rowvec;    /* fp16 [num_vec][ld_rowvec] or NULL */
  int64_t rows_per_vec;  /* output rows sharing one rowvec row (t*h*w of one batch item) */
  int64_t ld_rowvec;
  const void* residual;  /* fp16 [M][ld_res] or NULL */
  int64_t ld_res;
  int act;               /* uav_act_t */
  int out_dtype;         /* uav_dtype_t */
  int64_t ld_out;        /* element distance between output rows */
} uav_epilogue_t;

/* out[M][N] = epilogue(a[M][K] @ w[N][K]^T).  a: fp16, row stride lda (K % 8 == 0).
 * Replaces nn.Linear call sites: attention.py:97-106,156,177-178,202,327,355,382,400;
 * diffusers FeedForward (attention.py:493); time_emb_proj (resnet.py:243,273);
 * TimestepEmbedding (unet_video.py:176,478); AttentionBlock q/k/v/proj
 * (diffusers_attention.py:296-301). */
uav_status_t uav_linear(const void* a, int64_t M, int64_t K, int64_t lda, const void* w,
                        int64_t N, void* out, const uav_epilogue_t* epi, uav_stream_t stream);

/* 2-D convolution over NB = b*t images, channels-last, ksize 1 or 3, stride 1 or 2.
 * pad_mode 0: symmetric zero padding ksize/2 (InflatedConv3d resnet.py:94-101,
 *             Downsample3D padding=1 resnet.py:172);
 * pad_mode 1: stride 2, F.pad (0,1,0,1) then no padding (Downsample3D padding=0,
 *             resnet.py:188-192, VAE encoder).
 * x: fp16 [NB][H][W][ld_in] (first Cin channels used, Cin % 8 == 0); w: fp16
 * [Cout][ksize][ksize][Cin]; out: [NB][Ho][Wo][ld_out].  Stride 2 needs even H and W. */
uav_status_t uav_conv2d(const void* x, int64_t NB, int64_t H, int64_t W, int64_t Cin,
                        int64_t ld_in, const void* w, int64_t Cout, int ksize, int stride,
                        int pad_mode, void* out, const uav_epilogue_t* epi,
                        uav_stream_t stream);

/* Temporal (k,1,1) convolution with zero padding (k-1)/2 in t (nn.Conv3d in
 * ResnetBlock3DCNN, resnet.py:332,348,361).  x: fp16 [B][T][HW][ld_in];
 * w: fp16 [Cout][k][Cin]; out: [B][T][HW][ld_out]. */
uav_status_t uav_conv_temporal(const void* x, int64_t B, int64_t T, int64_t HW, int64_t Cin,
                               int64_t ld_in, const void* w, int64_t Cout, int k, void* out,
                               const uav_epilogue_t* epi, uav_stream_t stream);

/* 3x3x3 convolution, zero padding 1 in t,h,w (ResnetBlock3D_plus.conv_3d, resnet.py:461).
 * x: fp16 [B][T][H][W][ld_in]; w: fp16 [Cout][3][3][3][Cin]. */
uav_status_t uav_conv3d(const void* x, int64_t B, int64_t T, int64_t H, int64_t W, int64_t Cin,
                        int64_t ld_in, const void* w, int64_t Cout, void* out,
                        const uav_epilogue_t* epi, uav_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UAV_B200_H_ */
