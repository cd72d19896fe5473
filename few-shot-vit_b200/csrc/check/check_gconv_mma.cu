// Grouped 3x3 convolution (8 groups x 32 channels, pad 1) of the stage-1 conv-MLP on 20x20 maps
// (reference: Mlp.conv2, test_phase/models/visformer.py:146-148,157-159), forward and data-gradient.
//
// Warp-level tensor-core kernel staged through shared memory (north_star: "grouped 3x3 conv ... warp-level kernels"):
// a CTA owns one (image, group).  The group's 32-channel slice of the image is loaded ONCE into shared memory with a
// zero halo (22 x 22 pixels, cp.async 16-byte chunks), the group's 9 x 32 x 32 weights sit beside it, and every filter tap
// is just a different ldmatrix row address into that halo tile -- the activation is read from L2 once instead of nine
// times (the tcgen05 implicit-GEMM formulation re-fetched a shifted box per tap and wasted half of each 64-wide
// block-diagonal MMA).  5 warps x 5 m16 tiles cover the 400 pixels; a warp keeps the B fragments of a k-step in registers
// across its 5 tiles.  Epilogue: GELU (+ pre-activation copy for training) or the chain-rule factor gelu'(aux) for dgrad.
// TEST-ONLY warp-MMA cross-check of gconv_tc.cu (tests/native/libsunb200_check.so): never linked into libsunb200.so.
#include "../common.cuh"

#include <stdlib.h>
#include <string.h>


namespace {

constexpr int HW = 20, HP = 22, NPIX = 400, GC = 32;      // map side, padded side, pixels, channels per group
constexpr int A_LD = GC + 8;                               // 40 bf16 = 80 B rows (conflict-free ldmatrix)
constexpr int W_LD = GC + 8;
constexpr int A_ELEMS = HP * HP * A_LD;                    // 19360
constexpr int W_ELEMS = 9 * GC * W_LD;                     // 11520
constexpr int WARPS = 5, TILES_PER_WARP = 5, THREADS = WARPS * 32;
constexpr size_t SMEM = (size_t)(A_ELEMS + W_ELEMS) * sizeof(bf16);   // 61,760 B -> 3 CTAs per SM

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* smem) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x, y, y2, aux: bf16 [B*400, ld] NHWC rows; wg: bf16 [8 groups][9 taps][32 n][32 k]
__global__ void __launch_bounds__(THREADS, 3) gconv3x3_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ wg,
                                                              bf16* __restrict__ y, int ldy, bf16* __restrict__ y2, int ldy2,
                                                              const bf16* __restrict__ aux, int ldaux, int act, int dact) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* sA = reinterpret_cast<bf16*>(smem_raw);
    bf16* sW = sA + A_ELEMS;
    const int img = blockIdx.x, grp = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // zero the halo ring (interior is overwritten by cp.async; disjoint addresses)
    for (int i = tid; i < HP * HP; i += THREADS) {
        const int hy = i / HP, hx = i % HP;
        if (hy == 0 || hy == HP - 1 || hx == 0 || hx == HP - 1) {
            uint4* p = reinterpret_cast<uint4*>(sA + i * A_LD);
#pragma unroll
            for (int j = 0; j < A_LD / 8; ++j) p[j] = make_uint4(0, 0, 0, 0);
        }
    }
    // interior: 400 pixels x 4 chunks of 16 B
    const bf16* xin = x + (size_t)img * NPIX * ldx + grp * GC;
    for (int i = tid; i < NPIX * 4; i += THREADS) {
        const int p = i >> 2, c = i & 3;
        const int py = p / HW, px = p % HW;
        cp_async16(sA + ((py + 1) * HP + px + 1) * A_LD + c * 8, xin + (size_t)p * ldx + c * 8);
    }
    const bf16* wsrc = wg + (size_t)grp * 9 * GC * GC;
    for (int i = tid; i < 9 * GC * 4; i += THREADS) {
        const int row = i >> 2, c = i & 3;
        cp_async16(sW + row * W_LD + c * 8, wsrc + row * GC + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ldmatrix row assignment: lane -> (row within the m16 tile, 8-wide k half)
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lk = (lane >> 4) * 8;
    int a_off[TILES_PER_WARP];                 // element offset of this lane's pixel row (halo coords, tap (0,0)) per tile
#pragma unroll
    for (int mt = 0; mt < TILES_PER_WARP; ++mt) {
        const int p = (warp * TILES_PER_WARP + mt) * 16 + lrow;
        a_off[mt] = ((p / HW) * HP + (p % HW)) * A_LD + lk;
    }
    const int g = lane >> 2, t = lane & 3;
    float acc[TILES_PER_WARP][4][4];
#pragma unroll
    for (int mt = 0; mt < TILES_PER_WARP; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;

#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        const int tap_off = ((tap / 3) * HP + (tap % 3)) * A_LD;
        const bf16* wt = sW + tap * GC * W_LD;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t b[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b[j][0] = *reinterpret_cast<const uint32_t*>(wt + (j * 8 + g) * W_LD + ks * 16 + t * 2);
                b[j][1] = *reinterpret_cast<const uint32_t*>(wt + (j * 8 + g) * W_LD + ks * 16 + 8 + t * 2);
            }
#pragma unroll
            for (int mt = 0; mt < TILES_PER_WARP; ++mt) {
                uint32_t a0, a1, a2, a3;
                ldmatrix_x4(a0, a1, a2, a3, sA + a_off[mt] + tap_off + ks * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) mma16816(acc[mt][j], a0, a1, a2, a3, b[j][0], b[j][1]);
            }
        }
    }

    // epilogue: rows g / g+8 of each tile, channels j*8 + t*2 (+1)
#pragma unroll
    for (int mt = 0; mt < TILES_PER_WARP; ++mt) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int p = (warp * TILES_PER_WARP + mt) * 16 + g + hr * 8;
            const size_t row = (size_t)img * NPIX + p;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = grp * GC + j * 8 + t * 2;
                float v0 = acc[mt][j][hr * 2], v1 = acc[mt][j][hr * 2 + 1];
                if (y2) *reinterpret_cast<__nv_bfloat162*>(y2 + row * ldy2 + ch) = __floats2bfloat162_rn(v0, v1);
                if (act == ACT_GELU) { v0 = gelu_fast(v0); v1 = gelu_fast(v1); }
                if (aux) {
                    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(aux + row * ldaux + ch);
                    v0 *= act_grad(__bfloat162float(a.x), dact);
                    v1 *= act_grad(__bfloat162float(a.y), dact);
                }
                *reinterpret_cast<__nv_bfloat162*>(y + row * ldy + ch) = __floats2bfloat162_rn(v0, v1);
            }
        }
    }
}

}  // namespace

extern "C" int sunb_check_gconv3x3(const void* x, int ldx, const void* wg, void* y, int ldy, void* y2, int ldy2, const void* aux,
                                   int ldaux, int B, int act, int dact, void* stream) {
    SUNB_REQUIRE(x && wg && y && B > 0, "check_gconv3x3: bad arguments");
    SUNB_REQUIRE(ldx % 8 == 0 && ldy % 2 == 0 && (((size_t)x) & 15) == 0 && (((size_t)wg) & 15) == 0,
                 "check_gconv3x3: operands must be 16-byte aligned");
    SUNB_CHECK_CUDA(cudaFuncSetAttribute(gconv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    gconv3x3_kernel<<<dim3(B, 8), THREADS, SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(wg), reinterpret_cast<bf16*>(y), ldy,
        reinterpret_cast<bf16*>(y2), ldy2, reinterpret_cast<const bf16*>(aux), ldaux, act, dact);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
