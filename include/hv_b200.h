/*
 * hv_b200.h — C ABI of the B200-native two-stage pseudo-healthy synthesis path.
 *
 * The reference (zhibaishouheilab/HealthiVert-GAN) has no FFI of its own: the path sits
 * behind Python nn.Module classes.  This header is the boundary a
 * maintainer binds (ctypes stub in INTEGRATION.md); every entry point names the
 * reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless a
 *     parameter is documented as host memory;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, and returns 0 on success or a negative hv_status; the message of
 *     the last failure on the calling thread is returned by hv_last_error();
 *   - image tensors are NCHW fp32, contiguous, exactly like the reference's tensors;
 *   - no CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef HV_B200_H
#define HV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hv_stream_t;

enum hv_status {
  HV_OK = 0,
  HV_ERR_INVALID = -1,      /* bad argument / unsupported shape  (reference: assert / assert 0) */
  HV_ERR_CUDA = -2,         /* CUDA runtime error                                         */
  HV_ERR_UNSUPPORTED = -3,  /* valid request, not built for this configuration            */
  HV_ERR_STATE = -4         /* call order violation (e.g. forward before prepare)         */
};

enum hv_act { HV_ACT_NONE = 0, HV_ACT_ELU = 1, HV_ACT_RELU = 2, HV_ACT_SIGMOID = 3,
              HV_ACT_LRELU02 = 4, HV_ACT_CLAMP1 = 5 /* identity then clamp[-1,1] */,
              HV_ACT_HEADS = 6 /* Cout==2: ch0 clamp[-1,1] -> y, ch1 sigmoid -> y2 */ };

enum hv_src_mode { HV_SRC_DIRECT = 0, /* [N,ch,H,W]                                       */
                   HV_SRC_UP2 = 1,    /* [N,ch,H/2,W/2] read through nearest x2 upsample  */
                   HV_SRC_SUB2 = 2,   /* [N,ch,2H,2W]  read through nearest x0.5 (::2)    */
                   HV_SRC_SCALAR = 3  /* [N] one value per sample broadcast to a plane    */ };

enum hv_precision { HV_PREC_FP32 = 0, /* SIMT FFMA, parity mode (max-abs <= 1e-3)         */
                    HV_PREC_BF16 = 1  /* bf16 operands on tcgen05, fp32 accumulate        */ };

const char* hv_last_error(void);
int hv_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t hv_launch_count(void);

/* ---- A1: Conv2dBlock = spectral_norm(Conv2d)+bias+act ---------------------------------
 * replaces torch.nn.utils.spectral_norm's forward pre-hook (torch/nn/utils/
 * spectral_norm.py:92-114) as installed by models/inpaint_networks.py:491-492.
 * training!=0: v <- normalize(W^T u), u <- normalize(W v) in place (eps 1e-12), then
 * sigma = u^T W v; w_eff = w_orig / sigma (fp32, same [cout, kdim] layout).           */
int hv_sn_prepare(const float* w_orig, float* u, float* v, int cout, int kdim, int training,
                  float* w_eff, float* sigma_out, hv_stream_t stream);

typedef struct hv_conv_src { const float* ptr; int channels; int mode; } hv_conv_src;

typedef struct hv_conv_desc {
  int n, cin, cout;          /* cin = sum of src[i].channels                              */
  int hin, win;              /* virtual input extent (after the per-source up/sub-sample) */
  int k, stride, pad, dil;
  int act;                   /* hv_act                                                    */
  int nsrc;                  /* 1..4 sources concatenated along channels (torch.cat)      */
  hv_conv_src src[4];
} hv_conv_desc;

/* replaces Conv2dBlock.forward (models/inpaint_networks.py:494-503) and the
 * torch.cat / F.interpolate that feed it (:77,:97-99,:105-106,:179,:207,:219,:222,:225);
 * also nn.Conv2d(+LeakyReLU) of NLayerDiscriminator (models/networks.py:575-598).
 * w: [cout,cin,k,k] fp32 effective weights; bias may be NULL; y2 only for HV_ACT_HEADS.  */
int hv_conv2d_fwd(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2,
                  hv_stream_t stream);

/* bf16 tensor-core variant of hv_conv2d_fwd (tcgen05.mma, fp32 accumulate in TMEM): same
 * descriptor and fp32 NCHW tensors; inputs and weights are rounded to bf16, the output is rounded to
 * bf16 (heads: fp32).  k in {3,5}, 'same' padding, stride 1 or (k=3) 2, cout <= 64.
 * up2_out != 0 additionally applies the nearest x2 upsample of the NEXT layer to the stored
 * result (y is [n,cout,2*hout,2*wout]).                                                        */
int hv_conv2d_bf16(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2,
                   int up2_out, hv_stream_t stream);

/* replaces global_pool + fc_height + sigmoid (models/inpaint_networks.py:90-93,:211-214) */
int hv_gap_fc_sigmoid(const float* x, const float* fc_w, const float* fc_b, float* out,
                      int n, int c, int hw, hv_stream_t stream);

/* ---- A3: ContextualAttention.forward(f, f, mask) -------------------------------------
 * replaces models/inpaint_networks.py:247-410 for ksize=3, stride=1, rate=2, fuse_k=3.
 * f: [n,c,h,w]; mask: [n,1,4h,4w]; y: [n,c,h,w]; offsets (may be NULL): [n,2,h/2,w/2]
 * int32 (argmax row/col minus own position); flow (may be NULL): [n,3,4h,4w] fp32
 * colour-wheel image in [0,1].  per_sample_mask == 0 reproduces the reference, which
 * reads the mask of sample 0 for the whole batch (:314); != 0 uses each sample's own.
 * workspace: hv_ctx_attn_workspace_bytes(n,c,h,w) bytes of device memory.               */
size_t hv_ctx_attn_workspace_bytes(int n, int c, int h, int w);
int hv_ctx_attn_fwd(const float* f, const float* mask, float* y, int32_t* offsets, float* flow,
                    int n, int c, int h, int w, float softmax_scale, int fuse,
                    int per_sample_mask, void* workspace, hv_stream_t stream);

/* bf16 tensor-core variant (tcgen05 similarity and paste contractions, fused fuse+softmax): same arguments, c = h = w = 64;
 * the workspace is taken from the stream-ordered allocator.                                                   */
int hv_ctx_attn_fwd_bf16(const float* f, const float* mask, float* y, int32_t* offsets, float* flow,
                         int n, int c, int h, int w, float softmax_scale, int fuse,
                         int per_sample_mask, hv_stream_t stream);

/* ---- A4: threshold + height-adaptive stitch -------------------------------------------
 * replaces models/pix2pix_model.py:201-252 and eval_3d_sagittal_twostage.py:103-118.
 * For sample i: pred = ceil(pred_h[i]*maxheight); hgt = max(pred, height[i]);
 * d = hgt-height[i]; xu = x1[i]-d/2; xb = xu+hgt;  out rows [xu,xb) <- gen,
 * rows [0,xu) <- real[d/2 : x1), rows [xb,H) <- real[x2 : x2+H-xb).  No host sync.
 * rows_out (may be NULL): [n,4] int32 = (hgt, d, xu, xb).                              */
int hv_stitch(const float* gen, const float* real, const float* pred_h, const int32_t* x1,
              const int32_t* x2, const int32_t* height, int maxheight, float* out,
              int32_t* rows_out, int n, int h, int w, hv_stream_t stream);
/* out = p > 0.5 ? 1 : 0 (pix2pix_model.py:201-202, eval:105) scaled by `value`; u8 and f32 */
int hv_threshold(const float* p, float* out_f32, uint8_t* out_u8, float value, size_t count,
                 hv_stream_t stream);

/* ---- A5: Sobel + edge loss ------------------------------------------------------------
 * replaces Sobel.forward (models/edge_operator.py:41-48).                               */
int hv_sobel(const float* img, float* edges, int n, int h, int w, hv_stream_t stream);
/* 800*mse(sobel(fake), sobel(real)) as an integer XOR count on {0,1} masks
 * (pix2pix_model.py:109,:263-264,:349). xor_count: device uint64; loss: device f32 */
int hv_edge_xor_loss(const float* fake_mask, const float* real_mask, unsigned long long* xor_count,
                     float* loss, int n, int h, int w, hv_stream_t stream);

/* ---- A10: per-column vertebral heights ------------------------------------------------
 * replaces the integer part of calculate_heights (evaluation/RHLV_quantification.py:41-73;
 * coronal twin slices axis 1).  vol_fake / vol_label: u8 {0,1} volumes [d0,d1,d2]
 * C-contiguous; slices are taken along `axis` (2 = sagittal, 1 = coronal) for
 * z in [z0,z1).  Per slice s=z-z0 writes counts[s][8][ncols] int32 in the order
 * all/pre/mid/post fake, all/pre/mid/post label, and meta[s][8] int32 =
 * (valid, t1, t2, center_fake, center_label, ymin, ymax, 0).                            */
int hv_column_heights(const uint8_t* vol_fake, const uint8_t* vol_label, int d0, int d1, int d2,
                      int axis, int z0, int z1, int32_t* counts, int32_t* meta, hv_stream_t stream);

/* ---- A2: the two-stage generator as one native plan -----------------------------------
 * replaces Generator.forward (models/inpaint_networks.py:28-32) in eval()/no_grad mode:
 * CoarseGenerator.forward (:68-117) + FineGenerator.forward (:169-232).                 */
typedef struct hv_generator hv_generator;

int hv_generator_num_layers(void); /* 47 conv blocks, reference state_dict order */
/* name_out: >= 64 bytes, "coarse_generator.conv1" ... ; geometry of the layer */
int hv_generator_layer_info(int idx, char* name_out, int* cin, int* cout, int* k, int* stride,
                            int* pad, int* dil, int* act);
int hv_generator_create(hv_generator** out, int max_batch, int precision);
int hv_generator_destroy(hv_generator* g);
/* device pointers to the reference parameters of layer idx; copied/packed by prepare()  */
int hv_generator_set_layer(hv_generator* g, int idx, const float* w_orig, float* u, float* v,
                           const float* bias);
/* which: 0 = coarse_generator.fc_height, 1 = fine_generator.fc_height; w [64], b [1]     */
int hv_generator_set_fc(hv_generator* g, int which, const float* w, const float* b);
/* sigma, W/sigma and operand packing for all layers; training!=0 runs one power iteration */
int hv_generator_prepare(hv_generator* g, int training, hv_stream_t stream);
/* x, mask, cam: [n,1,256,256]; ratio: [n]; outputs [n,1,256,256] x4, flow [n,3,256,256]
 * (NULL to skip), pred1_h/pred2_h [n]; offsets (NULL to skip) [n,2,32,32] int32.
 * per_sample_mask != 0 uses every sample's own mask in the attention (faithful for the
 * batch-1 eval driver); 0 reproduces the reference's sample-0 quirk.                     */
int hv_generator_forward(hv_generator* g, const float* x, const float* mask, const float* cam,
                         const float* ratio, int n, float* coarse_seg, float* fine_seg,
                         float* x_stage1, float* x_stage2, float* flow, float* pred1_h,
                         float* pred2_h, int32_t* offsets, int per_sample_mask,
                         hv_stream_t stream);
/* measurement hook (bf16 plan): launch the tensor-core conv kernel of layer idx alone on the
 * activations of the last forward (bench.py times the dominant kernel with it)            */
int hv_generator_run_layer(hv_generator* g, int idx, int n, hv_stream_t stream);
/* debug/parity tap: copy the fp32 NCHW activation of layer idx (or idx==47: attention
 * output) of the LAST forward into out; returns element count or negative status        */
long long hv_generator_read_tap(hv_generator* g, int idx, float* out, hv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HV_B200_H */
