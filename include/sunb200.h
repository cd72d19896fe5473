/* sunb200 -- C ABI of the B200-native SUN / Visformer episodic hot path.
 *
 * Drop-in boundary.  The reference (DongSky/few-shot-vit) is pure PyTorch and has no FFI of its own; every
 * FLOP of its hot path is executed by torch.nn modules.  Each entry point below replaces the device work of
 * the reference function cited beside it (paths relative to the reference checkout).  The Python shims in
 * few-shot-vit_b200/models/ keep the reference's module API (models.make / forward signatures / state_dict
 * names) and call these functions through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a CUDA stream passed as void* (cudaStream_t).
 *   - every function is asynchronous and stream-ordered; it returns 0 on success or a negative SunbStatus.
 *     sunb_last_error() returns a thread-local message for the last failure.  No exceptions cross the ABI.
 *   - the caller owns all memory (PyTorch's caching allocator in the shims); the library allocates nothing.
 *   - "bf16" buffers are raw uint16 bfloat16 bit patterns; activations are NHWC.
 *   - sm_100a only.  There is no CPU fallback.
 */
#ifndef SUNB200_H
#define SUNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUNB_ABI_VERSION 7

int sunb_abi_version(void);
const char* sunb_last_error(void);

/* ---------------------------------------------------------------------------------------------------
 * Generic GEMM / implicit GEMM with fused epilogue (tcgen05 + TMEM + TMA).
 * Replaces nn.Conv2d 1x1 / 3x3 (+ folded BatchNorm, bias, GELU / LeakyReLU, residual add) as used in
 * test_phase/models/visformer.py:144-163 (Mlp), :175-191 (Attention.qkv/proj), :209-214 (stem conv2/conv3),
 * :276-287 (PatchEmbed) and nn.Linear in sun_meta_training/models/classifier.py:27-35.
 *   C[m, g*c_goff + n] = act( rs[m / rows_per_img] * sum_{tap,k} A_tap[m, g*a_goff + k] * W[(g*taps + tap)*N + n, k]
 *                             + resid[m, g*c_goff + n] + bias[(m % bias_mod)*bias_ld + g*c_goff + n] )
 * a_mode 0: plain rows (taps = 1).  a_mode 1: 3x3 / pad 1 / stride 1 convolution over NHWC [B,H,W,lda]
 * (taps = 9, tap = (dy+1)*3 + (dx+1)); tiles are built from bw x bh pixel boxes (bw*bh divides 128).
 * ------------------------------------------------------------------------------------------------- */
typedef struct SunbGemmDesc {
    int32_t M, N, K;
    int32_t taps, groups;
    int32_t a_goff, c_goff;
    int32_t a_mode, H, W, bw, bh;
    const void* A;          /* bf16 */
    int32_t lda;
    const void* Wt;         /* bf16, K-major rows of ldw elements */
    int32_t ldw;
    const float* bias;
    int32_t bias_mod, bias_ld;
    int32_t act;            /* 0 none, 1 LeakyReLU(0.1), 2 GELU(erf) */
    const void* resid;      /* bf16, nullable */
    int32_t ldr;
    const float* row_scale; /* per-image scale of the accumulator (DropPath mask / keep), nullable */
    int32_t rows_per_img;
    void* out;              /* bf16, nullable */
    int32_t ldc;
    float* out_f32;         /* nullable */
    int32_t ldc_f32;
    int32_t out_map, oH, oW; /* 0 identity; 1 = 2x2 space-to-depth of an oH x oW raster (feeds PatchEmbed) */
    void* out2;             /* bf16 copy of the value before `act` (saved for the backward pass), nullable */
    int32_t ldc2;
    const void* dact_aux;   /* backward: result *= act'(dact_aux[m, col]) with activation kind `dact`, nullable */
    int32_t ld_aux;
    int32_t dact;
} SunbGemmDesc;

/* impl: 0 = tcgen05 kernel (product path), 1 = SIMT cross-check kernel (tests only) */
int sunb_gemm(const SunbGemmDesc* desc, int impl, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Encoder: Visformer.forward of 'visformer_micro_80' in eval mode, BatchNorm folded
 * (test_phase/models/visformer.py:424-462; sun_meta_training/models/visformer.py:464 for the dense output).
 * All weight buffers are produced by few-shot-vit_b200/sunb200/packing.py (layouts documented there).
 * ------------------------------------------------------------------------------------------------- */
typedef struct SunbConvMlpW {        /* stage-1 Block: norm2 folded into conv1 */
    const void* w1; const float* b1; /* bf16 [256][128], fp32 [256] */
    const void* w2;                  /* bf16 grouped 3x3 weights [8 groups][9 taps][32 n][32 k] (stand-alone sunb_gconv3x3) */
    const void* w3;                  /* bf16 [128][256] (stand-alone conv3 GEMM) */
    const void* w23;                 /* operand blob of the fused block tail (sunb_convmlp_tail): per group g 26,624 bytes =
                                      * grouped taps [9][4 k-chunks][32 n][8 k] then the conv3 slice [4 k-chunks][128 n][8 k] */
} SunbConvMlpW;

typedef struct SunbAttnBlockW {      /* stage-2/3 Block: norm1 folded into qkv, norm2 into mlp.conv1 */
    /* heads padded to dp = 48 (stage 2, d = 42) / 96 (stage 3, d = 85) channels with zero rows / columns, so every
     * (token, head) segment of the qkv buffer is 16-byte aligned: row (x*6 + y)*dp + z of wqkv is qkv.weight row
     * x*6*d + y*d + z for z < d and zero for d <= z < dp; the same for bqkv; column y*dp + z of wproj likewise. */
    const void* wqkv; const float* bqkv;   /* bf16 [3*6*dp][C], fp32 [3*6*dp] */
    const void* wproj;                     /* bf16 [C][6*dp] */
    const void* w1; const float* b1;       /* bf16 [4C][C], fp32 [4C] */
    const void* w3;                        /* bf16 [C][4C] */
} SunbAttnBlockW;

typedef struct SunbEncoderWeights {
    const float* stem_w1; const float* stem_b1;    /* fp32 [64][27] (BN1 folded), [64] */
    const float* stem_wd; const float* stem_bd;    /* fp32 [128][27] (downsample BN folded), [128] */
    const void* stem_w2; const float* stem_b2;     /* bf16 [9][128][64], fp32 [128] */
    const void* stem_w3; const float* stem_b3;     /* bf16 [9][128][128], fp32 [128] */
    const float* pos1;                             /* fp32 [400][128] (NHWC) */
    SunbConvMlpW s1[4];
    const void* pe2_w; const float* pe2_bias;      /* bf16 [256][4*128] k=(dy,dx,c); fp32 [100][256] = bias*bn + pos2 */
    SunbAttnBlockW s2[2];
    const void* pe3_w; const float* pe3_bias;      /* bf16 [512][4*256]; fp32 [25][512] */
    SunbAttnBlockW s3[3];
    const float* final_scale; const float* final_shift;   /* fp32 [512] each */
} SunbEncoderWeights;

/* optional copies of the residual stream after each layer boundary (bf16 NHWC), for the per-layer tests */
typedef struct SunbEncoderTaps {
    void* stem;          /* [B,20,20,128] after pos_embed1 */
    void* stage1[4];     /* [B,20,20,128] */
    void* patch_embed2;  /* [B,10,10,256] */
    void* stage2[2];
    void* patch_embed3;  /* [B,5,5,512] */
    void* stage3[3];
} SunbEncoderTaps;

int sunb_encoder_workspace_bytes(int B, size_t* bytes);

/* x: fp32 NCHW [B,3,80,80].  pooled: fp32 [B,512].  dense (nullable): fp32 NHWC [B,5,5,512] after the final BN.
 * dense_bf16 / pooled_bf16 (nullable): bf16 copies that feed the SUN classifier GEMMs. */
int sunb_encoder_forward(const SunbEncoderWeights* w, const float* x, int B, void* workspace, size_t workspace_bytes,
                         float* pooled, float* dense, void* dense_bf16, void* pooled_bf16, const SunbEncoderTaps* taps,
                         void* stream);

/* Attention core (visformer.py:183-190) on tcgen05.  qkv bf16 [B*S, ld_qkv], channel (x*heads + y)*d_stride + z for x in
 * {q,k,v}, head y, z < d; out bf16 [B*S, ld_out], channel y*d_stride + z.  Heads are PADDED: d_stride = 48 for S = 100
 * (d <= 48), d_stride = 96 for S = 25 (48 < d <= 96) -- the two attention stages of visformer_micro_80; pad channels must
 * be zero on input and are written as zero.  Any other layout (the reference's packed d_stride == d included) returns
 * SUNB_ERR_ARG: there is one implementation and no fallback. */
int sunb_attention(const void* qkv, void* out, int B, int S, int d, int d_stride, int heads, int ld_qkv, int ld_out,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Episode head: MetaBaseline (test_phase/models/meta_baseline.py:36-46) and utils.compute_logits
 * (test_phase/utils/__init__.py:78-101).  metric: 0 dot, 1 cos, 2 sqr.
 *   feat_shot fp32 [E, way, shot, D]; feat_query fp32 [E, Q, D]; logits fp32 [E, Q, way].
 *   temp: device scalar if temp_dev != NULL (the learnable nn.Parameter), else temp_host.
 * sunb_logits_ce_acc: out[0] = F.cross_entropy mean, out[1] = utils.compute_acc (test_few_shot.py:89-90).
 * ------------------------------------------------------------------------------------------------- */
int sunb_episode_logits(const float* feat_shot, const float* feat_query, float* logits, int E, int way, int shot, int Q,
                        int D, int metric, const float* temp_dev, float temp_host, void* stream);
int sunb_logits_ce_acc(const float* logits, const int64_t* label, int R, int W, float* out2, void* stream);
int sunb_hard_ce_backward(const float* logits, const int64_t* label, int R, int W, const float* gout, float gscale,
                          float* dlogits, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SUN local head.
 *   sunb_softlabel: generate_softlabel (sun_meta_training/offline.py:57-76), bug-compatible (background column 1).
 *     logits element (b, c, p) at logits[b*sb + c*sc + p*sp]; out fp32 [B*hw, n_cls+1].
 *   sunb_soft_ce_forward/backward: SoftTargetCrossEntropy (offline.py:34-45).  row_loss: scratch fp32 [R].
 * ------------------------------------------------------------------------------------------------- */
int sunb_softlabel(const float* logits, int64_t sb, int64_t sc, int64_t sp, int B, int n_cls, int hw, int k, int bp,
                   double smoothing, float* out, void* stream);
int sunb_soft_ce_forward(const float* x, int ldx, const float* target, int ldt, int R, int Rt, int C, float* row_loss,
                         float* loss, void* stream);
int sunb_soft_ce_backward(const float* x, int ldx, const float* target, int ldt, int R, int Rt, int C, const float* gout,
                          float gscale, float* dx, int lddx, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Train-mode building blocks (meta-tuning step, reference: meta_tuning_sun_m/train_meta.py:168-174 driving
 * autograd through the modules above with BatchNorm in batch-statistics mode and DropPath).
 * The step schedule lives in few-shot-vit_b200/sunb200/train.py (torch.autograd.Function); every device
 * operation it issues is one of the entry points below or sunb_gemm / sunb_attention.
 * ------------------------------------------------------------------------------------------------- */

/* Weight gradient on tcgen05 (MN-major operands, split-K, fp32 atomic accumulation into `out`):
 *   out[g][tap][m, n] += sum_p dY[p, g*a_goff + m] * X_tap[p, g*b_goff + n]
 * mode 0: plain rows.  mode 1: X shifted per 3x3 tap over NHWC [B,H,W,ldx] (same box geometry as sunb_gemm). */
typedef struct SunbWgradDesc {
    int32_t P, Ma, Nb, Ca, Cb;
    int32_t groups, a_goff, b_goff;
    int32_t taps;
    int32_t mode, H, W, bw, bh;
    const void* dY; int32_t ldy;     /* bf16 */
    const void* X; int32_t ldx;      /* bf16 */
    float* out; int32_t ldo;         /* fp32 [groups][taps][Ma][ldo] */
    int32_t ksplit;                  /* 0 = choose */
} SunbWgradDesc;
int sunb_wgrad(const SunbWgradDesc* desc, void* stream);

/* stem conv1 (3->64, s2) + downsample (3->128, s2) from the fp32 NCHW image (visformer.py:209-210,216);
 * w1 fp32 [64][27], wd fp32 [128][27], biases fp32; lrelu != 0 applies LeakyReLU(0.1) to the conv1 branch. */
int sunb_stem_in(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, void* a1, void* idn,
                 int B, int lrelu, void* stream);
/* scratch: bf16 [B*1600*32] (the im2col matrix of x), 32-byte aligned; dw1 [64][27] / dwd [128][27] are accumulated into. */
int sunb_stem_wgrad(const float* x, const void* da1, const void* didn, float* dw1, float* dwd, int B, void* scratch,
                    void* stream);

/* BatchNorm2d, training mode (visformer.py:118-124): column statistics, finalize (+ running stats, momentum),
 * apply (+ activation, + positional table), and the backward pair. */
int sunb_colstats(const void* x, int ldx, const void* u, int ldu, long M, int C, float* sum, float* sq, void* stream);
int sunb_bn_finalize(const float* sum, const float* sq, float count, const float* gamma, const float* beta, float* rmean,
                     float* rvar, int64_t* nbt, float momentum, float eps, int C, float* scale, float* shift, float* mean,
                     float* rstd, void* stream);
/* one-launch variants: column statistics + finalize by the last block (sum / sq / sdz / sdzx and the 4-byte ticket must be
 * zero on entry).  sunb_bn_stats_forward = sunb_colstats + sunb_bn_finalize; sunb_bn_stats_backward = sunb_colstats(dz, x) +
 * sunb_bn_bwd_finalize. */
int sunb_bn_stats_forward(const void* x, int ldx, long M, int C, float* sum, float* sq, void* ticket, const float* gamma,
                          const float* beta, float* rmean, float* rvar, int64_t* nbt, float momentum, float eps, float* scale,
                          float* shift, float* mean, float* rstd, void* stream);
int sunb_bn_stats_backward(const void* dz, int lddz, const void* x, int ldx, long M, int C, float* sdz, float* sdzx, void* ticket,
                           float count, const float* mean, const float* rstd, const float* gamma, int frozen, float* a, float* c1,
                           float* c2, float* dgamma, float* dbeta, void* stream);
int sunb_bn_apply(const void* x, int ldx, const float* scale, const float* shift, int act, const float* tab, int tab_mod,
                  void* out, int ldo, long M, int C, void* stream);
/* frozen BatchNorm inside a training step (utils.freeze_bn, test_phase/utils/__init__.py:150-153): running statistics */
int sunb_bn_frozen(const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps, int C,
                   float* scale, float* shift, float* mean, float* rstd, void* stream);
int sunb_bn_bwd_finalize(const float* sdz, const float* sdzx, float count, const float* mean, const float* rstd,
                         const float* gamma, int C, int frozen, float* a, float* c1, float* c2, float* dgamma, float* dbeta,
                         void* stream);
int sunb_bn_bwd_apply(const void* dz, int lddz, const void* x, int ldx, const float* a, const float* c1, const float* c2,
                      const float* mean, const void* res, int ldr, void* out, int ldo, long M, int C, void* stream);

/* stem tail (visformer.py:232-237 + pos_embed1 :431): BN3(c3) + BNd(id) -> LeakyReLU -> 2x2 max-pool -> + pos */
int sunb_stem_tail_forward(const void* c3, const void* idn, const float* s3, const float* t3, const float* sd,
                           const float* td, const float* pos, void* out, int B, void* stream);
int sunb_stem_tail_backward(const void* c3, const void* idn, const float* s3, const float* t3, const float* sd,
                            const float* td, const void* g, void* dz, int B, void* stream);

/* final BN (as scale/shift) + global average pool (visformer.py:455-462) and its pooling backward */
int sunb_final_norm_pool(const void* x, const float* scale, const float* shift, float* dense, void* dense_bf16,
                         float* pooled, void* pooled_bf16, int B, int T, int C, void* stream);
int sunb_pool_backward(const float* dpooled, const float* ddense, void* dy, int B, int T, int C, void* stream);

/* helpers: DropPath row scaling (visformer.py:89-97), space-to-depth reorder, positional-embedding gradient,
 * weight preparation (fp32 master -> bf16 operand layouts), grouped-conv pair packing / gradient extraction */
int sunb_scale_rows(const void* in, const float* rs, int rows_per_img, void* out, long M, int C, void* stream);
int sunb_s2d_reorder(const void* in, void* out, int B, int H, int W, int C, int dir, void* stream);
int sunb_batch_sum(const void* g, int B, long n, float* out, void* stream);
/* All weight operand copies of one training step in ONE launch: entry i writes the bf16 tensor
 * dst[a][b][c][d] = src[off + a*strides[0] + b*strides[1] + c*strides[2] + d*strides[3]] for d < dims[3], zero for
 * dims[3] <= d < ldd (and for c >= valid2 when valid2 != 0).  descs: HOST array, passed to the kernel by value (nothing is uploaded; CUDA-graph capturable). */
typedef struct SunbPackDesc {
    const void* src;    /* fp32 master weight */
    void* dst;          /* bf16 operand, contiguous [dims0][dims1][dims2][ldd] */
    int64_t off;        /* element offset added to every source index (mirrored taps start at the last tap) */
    int32_t strides[4]; /* source strides in elements (may be negative or zero) */
    int32_t dims[4];
    int32_t ldd;        /* padded length of the last destination dimension */
    int32_t valid2;     /* 0, or the number of dims[2] entries that come from the source (the rest are zero: padded heads) */
} SunbPackDesc;
int sunb_pack_weights(const SunbPackDesc* descs, int n, void* stream);
int sunb_permute_cast(const float* src, long off, long sa, long sb, long sc, int A, int B, int Cd, int ldd, void* dst,
                      void* stream);
int sunb_grouped_pairs(const float* w, void* dst, int transpose_flip, void* stream);
int sunb_grouped_wgrad_extract(const float* scratch, float* dw, void* stream);

/* channel LayerNorm over NHWC rows, fp32 (LayerNorm wrapper, test_phase/models/visformer.py:109-115; not instantiated by
 * 'visformer_micro_80', provided for API completeness) */
int sunb_layernorm_rows(const float* x, const float* gamma, const float* beta, float* y, long M, int C, float eps,
                        void* stream);

/* Grouped 3x3 convolution (8 groups x 32 channels, pad 1) on 20x20 maps: Mlp.conv2 (visformer.py:146-148,157-159), forward
 * (act = 2: GELU, y2 = pre-activation copy) and data gradient (transposed/mirrored weights from sunb_gconv_pack(.., 1),
 * aux/dact = chain-rule factor act'(aux)).  x, y, y2, aux: bf16 [B*400, ld]; wg: bf16 [8][9][32][32]. */
int sunb_gconv3x3(const void* x, int ldx, const void* wg, void* y, int ldy, void* y2, int ldy2, const void* aux, int ldaux,
                  int B, int act, int dact, void* stream);
int sunb_gconv_pack(const float* w, void* dst, int transpose_flip, void* stream);

/* Fused tail of the stage-1 conv-MLP block, eval mode (Mlp.conv2 + GELU + conv3 and the Block residual, visformer.py:146-163,
 * 259-263):  out = resid + conv3(gelu(gconv3x3(h1))).  h1: bf16 [B*400, 256] = gelu(conv1(bn(x))); wblob: SunbConvMlpW.w23;
 * resid, out: bf16 [B*400, 128] (out != resid); s2d = 1 stores the rows 2x2 space-to-depth (input order of the following
 * PatchEmbed GEMM).  The 256-channel hidden tensor between the grouped conv and conv3 never leaves the SM. */
int sunb_convmlp_tail(const void* h1, const void* wblob, const void* resid, void* out, int B, int s2d, void* stream);

/* Fused MLP of a stage-2 attention block, eval mode (visformer.py:127-163 inside Block :259-263 with BatchNorm folded into conv1):
 *   out = x + conv3(gelu(conv1(x) + b1)),  x / out: bf16 [M, 256] (out may alias x unless s2d), w1: bf16 [1024, 256], b1: fp32
 *   [1024], w3: bf16 [256, 1024].  s2d = 1 stores the rows 2x2 space-to-depth of an oH x oW raster (input order of the next
 *   PatchEmbed GEMM).  The 1024-wide hidden tensor never leaves the SM. */
int sunb_mlp_fused(const void* x, const void* w1, const float* b1, const void* w3, void* out, int M, int s2d, int oH, int oW,
                   void* stream);

/* backward of the attention core (same padded layouts as sunb_attention, tcgen05; dout: gradient of `out`; the pad channels
 * of dqkv are written as zero) and of the episode head */
int sunb_attention_backward(const void* qkv, const void* dout, void* dqkv, int B, int S, int d, int d_stride, int heads,
                            int ld_qkv, int ld_out, void* stream);
int sunb_episode_logits_backward(const float* feat_shot, const float* feat_query, const float* dlogits, float* dshot,
                                 float* dquery, float* dtemp, int E, int way, int shot, int Q, int D, int metric,
                                 const float* temp_dev, float temp_host, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SUN-D head (DeepEMD on Visformer node features), evaluation path: meta_tuning_sun_d/Models/models/Network.py:48-81,
 * 109-128, 143-175 and emd_utils.py:65-76.  proto fp32 [W, n, D], query fp32 [Q, n, D] (node-major rows, n <= 32 nodes);
 * logits fp32 [Q, W] = sum_ij sim_ij * flow_ij * temperature / n with flow = the optimal transport plan of the reference's
 * cv2.EMD call (solved on the device); flows (nullable) fp32 [Q, W, n, n] returns the plans.
 * ------------------------------------------------------------------------------------------------- */
int sunb_emd_head(const float* proto, const float* query, float* logits, float* flows, int W, int Q, int n, int D,
                  float temperature, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * On-device input path (test_phase/datasets/mini_imagenet.py:50-56 default_transform): uint8 HWC images [N, in, in, 3]
 * resident in HBM -> PIL-exact bilinear Resize -> CenterCrop -> ToTensor -> Normalize -> fp32 NCHW [n, 3, out, out].
 * idx (nullable): int64 [n] gather index into the image store (the flat batch of a CategoriesSampler).
 * tab_min int32 [out], tab_k int32 [out][3]: first source index and 22-bit fixed-point weights of every CROPPED output
 * position (PIL precompute_coeffs / normalize_coeffs_8bpc; built by sunb200/input.py); mean_std fp32 [6].
 * ------------------------------------------------------------------------------------------------- */
int sunb_preprocess_u8(const void* data, const int64_t* idx, int n, int in_size, int out_size, const int32_t* tab_min,
                       const int32_t* tab_k, const float* mean_std, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fused multi-tensor optimizers: one launch updates every parameter tensor (all fp32).
 *   sunb_fused_sgd   : torch.optim.SGD(momentum, weight_decay) as utils.make_optimizer builds it
 *                      (meta_tuning_sun_m/utils/__init__.py:128-139; train_meta_warmup.py:140).  m = momentum buffer.
 *   sunb_fused_adamw : AdamW(betas, eps, decoupled weight decay) (sun_meta_training/offline.py:229).  m, v = moments.
 * tensors: HOST array of n_tensors entries (device pointers inside); it is passed to the kernel by value, so nothing is
 * copied at step time and a CUDA-graph capture records it.  g == NULL skips the tensor.  hp_dev: device fp32 [8]:
 *   [0] lr  [1] momentum (SGD) | beta1 (AdamW)  [2] weight_decay  [3] beta2  [4] eps  [5] step counter (AdamW: incremented
 *   on the device by every call, so a captured CUDA graph replays correctly).
 * ------------------------------------------------------------------------------------------------- */
typedef struct SunbOptTensor {
    void* p;            /* fp32 parameter */
    const void* g;      /* fp32 gradient (nullable) */
    void* m;            /* fp32 momentum / first moment */
    void* v;            /* fp32 second moment (AdamW; NULL for SGD) */
    int64_t n;          /* elements */
} SunbOptTensor;
int sunb_fused_sgd(const SunbOptTensor* tensors, int n_tensors, float* hp_dev, void* stream);
int sunb_fused_adamw(const SunbOptTensor* tensors, int n_tensors, float* hp_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUNB200_H */
