// C ABI (include/sunb200.h) and the eval-mode encoder schedule.
#include "common.cuh"
#include "gemm_desc.cuh"
#include "../../include/sunb200.h"

#include <map>
#include <mutex>
#include <utility>

#include <stdarg.h>
#include <string.h>

// launchers defined in the other translation units
int sunb_launch_stem_in(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, bf16* a1,
                        bf16* idn, int B, int lrelu, cudaStream_t stream);
int sunb_launch_attention(const bf16* qkv, bf16* out, int B, int S, int d, int ds, int heads, int ld_qkv, int ld_out,
                          cudaStream_t stream);
int sunb_launch_final_norm_pool(const bf16* x, const float* scale, const float* shift, float* dense, bf16* dense_bf16,
                                float* pooled, bf16* pooled_bf16, int B, int T, int C, cudaStream_t stream);
int sunb_launch_episode_logits(const float* feat_shot, const float* feat_query, float* logits, int E, int way, int shot,
                               int Q, int D, int metric, const float* temp_dev, float temp_host, cudaStream_t stream);
int sunb_launch_logits_ce_acc(const float* logits, const long long* label, int R, int W, float* out, cudaStream_t stream);
int sunb_launch_softlabel(const float* logits, long sb, long sc, long sp, int B, int n_cls, int hw, int k, int bp,
                          double smoothing, float* out, cudaStream_t stream);
int sunb_launch_soft_ce_forward(const float* x, int ldx, const float* t, int ldt, int R, int Rt, int C, float* row_loss,
                                float* loss, cudaStream_t stream);
int sunb_launch_soft_ce_backward(const float* x, int ldx, const float* t, int ldt, int R, int Rt, int C,
                                 const float* gout, float gscale, float* dx, int lddx, cudaStream_t stream);
int sunb_launch_hard_ce_backward(const float* l, const long long* label, int R, int W, const float* gout, float gscale,
                                 float* dl, cudaStream_t stream);

int sunb_launch_episode_logits_bwd(const float* feat_shot, const float* feat_query, const float* dlogits, float* dshot,
                                   float* dquery, float* dtemp, int E, int way, int shot, int Q, int D, int metric,
                                   const float* temp_dev, float temp_host, cudaStream_t stream);
int sunb_launch_wgrad_tc(WgradParams p, cudaStream_t stream);

int sunb_launch_layernorm_rows(const float* x, const float* gamma, const float* beta, float* y, long M, int C, float eps,
                               cudaStream_t stream);

extern "C" int sunb_convmlp_tail(const void* h1, const void* wblob, const void* resid, void* out, int B, int s2d, void* stream);
extern "C" int sunb_gconv3x3(const void* x, int ldx, const void* wg, void* y, int ldy, void* y2, int ldy2, const void* aux,
                             int ldaux, int B, int act, int dact, void* stream);

static thread_local char g_err[512] = "";
static thread_local bool g_pdl = true;
bool sunb_pdl_allowed() { return g_pdl; }
void sunb_pdl_allow(bool on) { g_pdl = on; }
namespace {
struct PdlScope {                       // restores the calling thread's setting on every return path
    bool prev;
    explicit PdlScope(bool on) : prev(g_pdl) { g_pdl = on; }
    ~PdlScope() { g_pdl = prev; }
};
}  // namespace

void sunb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// per (kernel, device) shared-memory opt-in and per-device SM count (see common.cuh)
int sunb_opt_in_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;
    int dev = 0;
    SUNB_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int& have = done[std::make_pair(kernel, dev)];
    if (have < bytes) {
        SUNB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
    return SUNB_OK;
}

int sunb_num_sms() {
    static std::mutex mu;
    static std::map<int, int> sms;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> lock(mu);
    int& n = sms[dev];
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

namespace {

constexpr int HEADS = 6;

struct Workspace {
    bf16 *a1, *idn, *a2, *s1, *h1, *s1b, *s1d, *t2, *qkv2, *ao2, *hid2, *t2d, *t3, *qkv3, *ao3, *hid3;
    size_t bytes;
};

Workspace carve(void* base, int B) {
    Workspace w;
    size_t off = 0;
    auto take = [&](size_t elems) {
        bf16* p = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(base) + off);
        off += (elems * sizeof(bf16) + 255) & ~(size_t)255;
        return p;
    };
    const size_t b = (size_t)B;
    w.a1 = take(b * 1600 * 64);
    w.idn = take(b * 1600 * 128);
    w.a2 = take(b * 1600 * 128);
    w.s1 = take(b * 400 * 128);
    w.h1 = take(b * 400 * 256);
    w.s1b = take(b * 400 * 128);       // stage-1 residual stream ping-pong (the fused block tail cannot run in place)
    w.s1d = take(b * 400 * 128);
    w.t2 = take(b * 100 * 256);
    w.qkv2 = take(b * 100 * 864);      // 3 x 6 heads x 48 (d = 42 padded)
    w.ao2 = take(b * 100 * 288);
    w.hid2 = take(b * 100 * 1024);
    w.t2d = take(b * 100 * 256);
    w.t3 = take(b * 25 * 512);
    w.qkv3 = take(b * 25 * 1728);      // 3 x 6 heads x 96 (d = 85 padded)
    w.ao3 = take(b * 25 * 576);
    w.hid3 = take(b * 25 * 2048);
    w.bytes = off;
    return w;
}

GemmParams base_gemm(int M, int N, int K, const bf16* A, int lda, const void* W, int ldw, bf16* out, int ldc) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K;
    p.taps = 1; p.groups = 1;
    p.A = A; p.lda = lda;
    p.Wt = reinterpret_cast<const bf16*>(W); p.ldw = ldw;
    p.bias_mod = 1; p.bias_ld = 0;
    p.rows_per_img = 1;
    p.out = out; p.ldc = ldc;
    return p;
}

int tap_copy(void* dst, const bf16* src, size_t elems, cudaStream_t s) {
    if (!dst) return SUNB_OK;
    SUNB_CHECK_CUDA(cudaMemcpyAsync(dst, src, elems * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
    return SUNB_OK;
}

constexpr int SUNB_MLP_FUSED_MIN_ROWS = 6000;    // 60 images: measured break-even of the fused stage-2 MLP (tools/mlp_time.py)

// stage-2/3 Block (visformer.py:259-263 with attention enabled); x is updated in place, last block may store
// its output 2x2 space-to-depth for the following PatchEmbed.
int attn_block(const SunbAttnBlockW& w, bf16* x, int B, int S, int C, int d, int dp, bf16* qkv, bf16* ao,
               bf16* hid, bf16* s2d_out, int side, cudaStream_t st) {
    // heads are padded to dp channels (zero weights): qkv rows are [3][6][dp], attention output rows [6][dp]
    const int M = B * S, inner = HEADS * dp;
    GemmParams p = base_gemm(M, 3 * inner, C, x, C, w.wqkv, C, qkv, 3 * inner);
    p.bias = w.bqkv;
    SUNB_TRY(sunb_launch_gemm(p, st));
    SUNB_TRY(sunb_launch_attention(qkv, ao, B, S, d, dp, HEADS, 3 * inner, inner, st));
    p = base_gemm(M, C, inner, ao, inner, w.wproj, inner, x, C);
    p.resid = x; p.ldr = C;
    SUNB_TRY(sunb_launch_gemm(p, st));
    if (C == 256 && M >= SUNB_MLP_FUSED_MIN_ROWS)
        // stage 2: conv1 + GELU + conv3 + residual in one kernel, the 1024-wide hidden tensor stays on the SM (mlp_fused.cu).
        // Bit-identical to the two GEMMs below; those win only on tiny batches (6 tiles of 8 serial chunks vs 4x more CTAs).
        return sunb_mlp_fused(x, w.w1, w.b1, w.w3, s2d_out ? s2d_out : x, M, s2d_out ? 1 : 0, side, side, st);
    p = base_gemm(M, 4 * C, C, x, C, w.w1, C, hid, 4 * C);
    p.bias = w.b1; p.act = ACT_GELU;
    SUNB_TRY(sunb_launch_gemm(p, st));
    p = base_gemm(M, C, 4 * C, hid, 4 * C, w.w3, 4 * C, s2d_out ? s2d_out : x, C);
    p.resid = x; p.ldr = C;
    if (s2d_out) { p.out_map = MAP_S2D; p.oH = side; p.oW = side; }
    SUNB_TRY(sunb_launch_gemm(p, st));
    return SUNB_OK;
}

}  // namespace

extern "C" {

int sunb_abi_version(void) { return SUNB_ABI_VERSION; }
const char* sunb_last_error(void) { return g_err; }

int sunb_gemm(const SunbGemmDesc* d, int impl, void* stream) {
    SUNB_REQUIRE(d != nullptr, "sunb_gemm: null descriptor");
    GemmParams p = sunb_desc_to_params(d);
    SUNB_REQUIRE(p.taps >= 1 && p.groups >= 1, "sunb_gemm: taps/groups must be >= 1");
    SUNB_REQUIRE(p.out || p.out_f32, "sunb_gemm: no output buffer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SUNB_REQUIRE(impl == 0, "sunb_gemm: impl %d is not part of the product library (the SIMT checker lives in tests/native)", impl);
    return sunb_launch_gemm_tc(p, st);
}

int sunb_encoder_workspace_bytes(int B, size_t* bytes) {
    SUNB_REQUIRE(B > 0 && bytes, "workspace_bytes: bad arguments");
    *bytes = carve(nullptr, B).bytes;
    return SUNB_OK;
}

int sunb_encoder_forward(const SunbEncoderWeights* w, const float* x, int B, void* workspace, size_t workspace_bytes,
                         float* pooled, float* dense, void* dense_bf16, void* pooled_bf16, const SunbEncoderTaps* taps,
                         void* stream) {
    SUNB_REQUIRE(w && x && pooled && workspace, "encoder_forward: null argument");
    SUNB_REQUIRE(B > 0, "encoder_forward: B must be positive");
    Workspace ws = carve(workspace, B);
    if (ws.bytes > workspace_bytes) {
        sunb_set_error("encoder_forward: workspace too small (%zu < %zu bytes for B=%d)", workspace_bytes, ws.bytes, B);
        return SUNB_ERR_WORKSPACE;
    }
    SUNB_REQUIRE((((size_t)workspace) & 255) == 0, "encoder_forward: workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    PdlScope pdl(B <= SUNB_PDL_MAX_IMAGES);       // long kernels gain nothing from overlapped set-up (common.cuh)
    SunbEncoderTaps none;
    memset(&none, 0, sizeof(none));
    const SunbEncoderTaps& tp = taps ? *taps : none;

    // ---- stem (visformer.py:220-239) + pos_embed1 (:431)
    SUNB_TRY(sunb_launch_stem_in(x, w->stem_w1, w->stem_b1, w->stem_wd, w->stem_bd, ws.a1, ws.idn, B, 1, st));
    {
        GemmParams p = base_gemm(B * 1600, 128, 64, ws.a1, 64, w->stem_w2, 64, ws.a2, 128);
        p.taps = 9; p.a_mode = 1; p.H = 40; p.W = 40; p.bw = 8; p.bh = 8;
        p.bias = w->stem_b2; p.act = ACT_LRELU;
        SUNB_TRY(sunb_launch_gemm(p, st));
        // conv3 + bn3 + shortcut + LeakyReLU with the 2x2 max-pool and pos_embed1 fused into the epilogue (visformer.py:229-237,
        // 431): the 40x40 map never reaches HBM, only the pooled 20x20 map is written
        p = base_gemm(B * 1600, 128, 128, ws.a2, 128, w->stem_w3, 128, nullptr, 128);
        p.taps = 9; p.a_mode = 1; p.H = 40; p.W = 40; p.bw = 8; p.bh = 8;
        p.bias = w->stem_b3; p.act = ACT_LRELU;
        p.resid = ws.idn; p.ldr = 128;
        p.pool_out = ws.s1; p.pool_pos = w->pos1;
        SUNB_TRY(sunb_launch_gemm(p, st));
    }
    SUNB_TRY(tap_copy(tp.stem, ws.s1, (size_t)B * 400 * 128, st));

    // ---- stage 1: x + conv3(gelu(gconv3x3(gelu(conv1(bn(x))))))  (visformer.py:152-163, 259-263)
    //      conv1 + GELU as a tcgen05 GEMM, then ONE fused kernel for grouped 3x3 + GELU + conv3 + residual (convmlp_tc.cu)
    bf16* cur = ws.s1;
    bf16* nxt = ws.s1b;
    for (int i = 0; i < 4; ++i) {
        const SunbConvMlpW& bw = w->s1[i];
        const bool last = (i == 3);
        GemmParams p = base_gemm(B * 400, 256, 128, cur, 128, bw.w1, 128, ws.h1, 256);
        p.bias = bw.b1; p.act = ACT_GELU;
        SUNB_TRY(sunb_launch_gemm(p, st));
        // the last block stores its output 2x2 space-to-depth for the PatchEmbed GEMM
        SUNB_TRY(sunb_convmlp_tail(ws.h1, bw.w23, cur, last ? ws.s1d : nxt, B, last ? 1 : 0, stream));
        if (tp.stage1[i]) {
            if (last) {   // the raster copy only exists for the test tap
                SUNB_TRY(sunb_convmlp_tail(ws.h1, bw.w23, cur, tp.stage1[i], B, 0, stream));
            } else {
                SUNB_TRY(tap_copy(tp.stage1[i], nxt, (size_t)B * 400 * 128, st));
            }
        }
        bf16* t = cur; cur = nxt; nxt = t;
    }

    // ---- patch_embed2 + pos_embed2 (visformer.py:438-441): GEMM over the space-to-depth view [B*100, 512]
    {
        GemmParams p = base_gemm(B * 100, 256, 512, ws.s1d, 512, w->pe2_w, 512, ws.t2, 256);
        p.bias = w->pe2_bias; p.bias_mod = 100; p.bias_ld = 256;
        SUNB_TRY(sunb_launch_gemm(p, st));
    }
    SUNB_TRY(tap_copy(tp.patch_embed2, ws.t2, (size_t)B * 100 * 256, st));
    for (int i = 0; i < 2; ++i) {
        const bool last = (i == 1);
        SUNB_TRY(attn_block(w->s2[i], ws.t2, B, 100, 256, 42, 48, ws.qkv2, ws.ao2, ws.hid2,
                            last ? ws.t2d : nullptr, 10, st));
        if (tp.stage2[i]) {
            if (last) {   // t2 still holds the pre-MLP stream: redo the last GEMM identity-mapped for the tap
                GemmParams q = base_gemm(B * 100, 256, 1024, ws.hid2, 1024, w->s2[i].w3, 1024,
                                         reinterpret_cast<bf16*>(tp.stage2[i]), 256);
                q.resid = ws.t2; q.ldr = 256;
                SUNB_TRY(sunb_launch_gemm(q, st));
            } else {
                SUNB_TRY(tap_copy(tp.stage2[i], ws.t2, (size_t)B * 100 * 256, st));
            }
        }
    }

    // ---- patch_embed3 + pos_embed3 (visformer.py:447-450)
    {
        GemmParams p = base_gemm(B * 25, 512, 1024, ws.t2d, 1024, w->pe3_w, 1024, ws.t3, 512);
        p.bias = w->pe3_bias; p.bias_mod = 25; p.bias_ld = 512;
        SUNB_TRY(sunb_launch_gemm(p, st));
    }
    SUNB_TRY(tap_copy(tp.patch_embed3, ws.t3, (size_t)B * 25 * 512, st));
    for (int i = 0; i < 3; ++i) {
        SUNB_TRY(attn_block(w->s3[i], ws.t3, B, 25, 512, 85, 96, ws.qkv3, ws.ao3, ws.hid3, nullptr, 5, st));
        SUNB_TRY(tap_copy(tp.stage3[i], ws.t3, (size_t)B * 25 * 512, st));
    }

    // ---- final BN + global average pool (visformer.py:455-462)
    SUNB_TRY(sunb_launch_final_norm_pool(ws.t3, w->final_scale, w->final_shift, dense, reinterpret_cast<bf16*>(dense_bf16),
                                         pooled, reinterpret_cast<bf16*>(pooled_bf16), B, 25, 512, st));
    return SUNB_OK;
}

int sunb_attention(const void* qkv, void* out, int B, int S, int d, int d_stride, int heads, int ld_qkv, int ld_out,
                   void* stream) {
    SUNB_REQUIRE(qkv && out, "attention: null argument");
    return sunb_launch_attention(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, S, d, d_stride, heads, ld_qkv,
                                 ld_out, reinterpret_cast<cudaStream_t>(stream));
}

int sunb_episode_logits(const float* feat_shot, const float* feat_query, float* logits, int E, int way, int shot, int Q,
                        int D, int metric, const float* temp_dev, float temp_host, void* stream) {
    SUNB_REQUIRE(feat_shot && feat_query && logits, "episode_logits: null argument");
    return sunb_launch_episode_logits(feat_shot, feat_query, logits, E, way, shot, Q, D, metric, temp_dev, temp_host,
                                      reinterpret_cast<cudaStream_t>(stream));
}

int sunb_logits_ce_acc(const float* logits, const int64_t* label, int R, int W, float* out2, void* stream) {
    SUNB_REQUIRE(logits && label && out2, "logits_ce_acc: null argument");
    return sunb_launch_logits_ce_acc(logits, reinterpret_cast<const long long*>(label), R, W, out2,
                                     reinterpret_cast<cudaStream_t>(stream));
}

int sunb_hard_ce_backward(const float* logits, const int64_t* label, int R, int W, const float* gout, float gscale,
                          float* dlogits, void* stream) {
    SUNB_REQUIRE(logits && label && dlogits, "hard_ce_backward: null argument");
    return sunb_launch_hard_ce_backward(logits, reinterpret_cast<const long long*>(label), R, W, gout, gscale, dlogits,
                                        reinterpret_cast<cudaStream_t>(stream));
}

int sunb_softlabel(const float* logits, int64_t sb, int64_t sc, int64_t sp, int B, int n_cls, int hw, int k, int bp,
                   double smoothing, float* out, void* stream) {
    SUNB_REQUIRE(logits && out, "softlabel: null argument");
    return sunb_launch_softlabel(logits, (long)sb, (long)sc, (long)sp, B, n_cls, hw, k, bp, smoothing, out,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int sunb_soft_ce_forward(const float* x, int ldx, const float* target, int ldt, int R, int Rt, int C, float* row_loss,
                         float* loss, void* stream) {
    SUNB_REQUIRE(x && target && row_loss && loss, "soft_ce_forward: null argument");
    return sunb_launch_soft_ce_forward(x, ldx, target, ldt, R, Rt, C, row_loss, loss, reinterpret_cast<cudaStream_t>(stream));
}

int sunb_soft_ce_backward(const float* x, int ldx, const float* target, int ldt, int R, int Rt, int C, const float* gout,
                          float gscale, float* dx, int lddx, void* stream) {
    SUNB_REQUIRE(x && target && dx, "soft_ce_backward: null argument");
    return sunb_launch_soft_ce_backward(x, ldx, target, ldt, R, Rt, C, gout, gscale, dx, lddx,
                                        reinterpret_cast<cudaStream_t>(stream));
}

int sunb_wgrad(const SunbWgradDesc* d, void* stream) {
    SUNB_REQUIRE(d && d->dY && d->X && d->out, "sunb_wgrad: null argument");
    WgradParams p;
    p.P = d->P; p.Ma = d->Ma; p.Nb = d->Nb; p.Ca = d->Ca; p.Cb = d->Cb;
    p.groups = d->groups > 0 ? d->groups : 1; p.a_goff = d->a_goff; p.b_goff = d->b_goff;
    p.taps = d->taps > 0 ? d->taps : 1;
    p.mode = d->mode; p.H = d->H; p.W = d->W; p.bw = d->bw; p.bh = d->bh;
    p.dY = reinterpret_cast<const bf16*>(d->dY); p.ldy = d->ldy;
    p.X = reinterpret_cast<const bf16*>(d->X); p.ldx = d->ldx;
    p.out = d->out; p.ldo = d->ldo; p.ksplit = d->ksplit;
    return sunb_launch_wgrad_tc(p, reinterpret_cast<cudaStream_t>(stream));
}

int sunb_stem_in(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, void* a1, void* idn,
                 int B, int lrelu, void* stream) {
    SUNB_REQUIRE(x && w1 && b1 && wd && bd && a1 && idn, "stem_in: null argument");
    return sunb_launch_stem_in(x, w1, b1, wd, bd, reinterpret_cast<bf16*>(a1), reinterpret_cast<bf16*>(idn), B, lrelu,
                               reinterpret_cast<cudaStream_t>(stream));
}

int sunb_final_norm_pool(const void* x, const float* scale, const float* shift, float* dense, void* dense_bf16,
                         float* pooled, void* pooled_bf16, int B, int T, int C, void* stream) {
    SUNB_REQUIRE(x && scale && shift && pooled, "final_norm_pool: null argument");
    return sunb_launch_final_norm_pool(reinterpret_cast<const bf16*>(x), scale, shift, dense,
                                       reinterpret_cast<bf16*>(dense_bf16), pooled, reinterpret_cast<bf16*>(pooled_bf16), B, T,
                                       C, reinterpret_cast<cudaStream_t>(stream));
}

int sunb_episode_logits_backward(const float* feat_shot, const float* feat_query, const float* dlogits, float* dshot,
                                 float* dquery, float* dtemp, int E, int way, int shot, int Q, int D, int metric,
                                 const float* temp_dev, float temp_host, void* stream) {
    SUNB_REQUIRE(feat_shot && feat_query && dlogits && dshot && dquery, "episode_logits_backward: null argument");
    return sunb_launch_episode_logits_bwd(feat_shot, feat_query, dlogits, dshot, dquery, dtemp, E, way, shot, Q, D, metric,
                                          temp_dev, temp_host, reinterpret_cast<cudaStream_t>(stream));
}

int sunb_layernorm_rows(const float* x, const float* gamma, const float* beta, float* y, long M, int C, float eps,
                        void* stream) {
    return sunb_launch_layernorm_rows(x, gamma, beta, y, M, C, eps, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
