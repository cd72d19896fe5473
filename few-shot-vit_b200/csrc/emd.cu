// SUN-D head: DeepEMD-style patch-to-patch matching on Visformer node features, evaluation path
// (reference: meta_tuning_sun_d/Models/models/Network.py:48-81 emd_forward_1shot, :109-128 get_emd_distance (solver 'opencv'),
//  :143-175 normalize_feature / get_similiarity_map, emd_utils.py:65-76 emd_inference_opencv).
// One CTA per (query, class) pair does the whole head for that pair:
//   1. node weights   w1[i] = relu(<q_i, mean_j p_j>) + 1e-3,  w2[j] = relu(<p_j, mean_i q_i>) + 1e-3     (get_weight_vector)
//   2. centring       every node minus its mean over the channels                                          (normalize_feature)
//   3. similarity     sim[i][j] = cos(q_i, p_j)                                                             (get_similiarity_map)
//   4. EMD            min sum_ij (1 - sim_ij) f_ij  s.t. row sums = w1 * n / sum(w1), column sums = w2 * n / sum(w2), f >= 0,
//                     weights first clamped as relu(w) + 1e-5 (emd_inference_opencv); logit = sum_ij sim_ij f_ij * T / n.
// The reference solves step 4 with cv2.EMD on the CPU, one device->host copy per pair (375 per episode).  Here the
// transportation problem is solved on the device by successive shortest augmenting paths with node potentials (exact for
// the LP; the logit only depends on the optimal objective, which is unique, so any exact solver reproduces cv2.EMD up to
// floating-point tolerance -- pinned against cv2.EMD in the tests).  n <= 32 nodes, fp64 flows / potentials.
#include "common.cuh"
#include "../../include/sunb200.h"

namespace {

constexpr int MAXN = 32;
constexpr int THREADS = 128;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// proto [W, n, D], query [Q, n, D] fp32 (node-major rows); logits [Q, W]; flows (optional) [Q, W, n, n]
__global__ void __launch_bounds__(THREADS) emd_head_kernel(const float* __restrict__ proto, const float* __restrict__ query,
                                                           float* __restrict__ logits, float* __restrict__ flows, int W, int n,
                                                           int D, float temperature) {
    extern __shared__ float sm[];
    float* pbar = sm;                   // [D] mean over the proto nodes
    float* qbar = pbar + D;             // [D]
    __shared__ float pmean[MAXN], qmean[MAXN], pnorm[MAXN], qnorm[MAXN], w1[MAXN], w2[MAXN];
    __shared__ float sim[MAXN][MAXN + 1];
    __shared__ double cost[MAXN][MAXN + 1], flow[MAXN][MAXN + 1];
    __shared__ double supply[MAXN], demand[MAXN], pot[2 * MAXN], dist[2 * MAXN];
    __shared__ int prev[2 * MAXN], done[2 * MAXN];

    const int qi = blockIdx.x / W, cj = blockIdx.x % W;
    const float* P = proto + (size_t)cj * n * D;
    const float* Qn = query + (size_t)qi * n * D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- means over the nodes (per channel), per-node channel means
    for (int c = tid; c < D; c += THREADS) {
        float sp = 0.f, sq = 0.f;
        for (int i = 0; i < n; ++i) { sp += P[i * D + c]; sq += Qn[i * D + c]; }
        pbar[c] = sp / (float)n;
        qbar[c] = sq / (float)n;
    }
    __syncthreads();
    // ---- per node: channel mean, weight against the other side's node mean
    for (int i = warp; i < 2 * n; i += THREADS / 32) {
        const bool isq = i >= n;
        const float* x = isq ? Qn + (i - n) * D : P + i * D;
        const float* other = isq ? pbar : qbar;
        float s = 0.f, d = 0.f;
        for (int c = lane; c < D; c += 32) { const float v = x[c]; s += v; d = fmaf(v, other[c], d); }
        s = warp_sum_f(s);
        d = warp_sum_f(d);
        if (lane == 0) {
            if (isq) { qmean[i - n] = s / (float)D; w1[i - n] = fmaxf(d, 0.f) + 1e-3f; }
            else     { pmean[i] = s / (float)D;     w2[i] = fmaxf(d, 0.f) + 1e-3f; }
        }
    }
    __syncthreads();
    // ---- centred norms and cosine similarities
    for (int i = warp; i < 2 * n; i += THREADS / 32) {
        const bool isq = i >= n;
        const float* x = isq ? Qn + (i - n) * D : P + i * D;
        const float m = isq ? qmean[i - n] : pmean[i];
        float s = 0.f;
        for (int c = lane; c < D; c += 32) { const float v = x[c] - m; s = fmaf(v, v, s); }
        s = warp_sum_f(s);
        if (lane == 0) { if (isq) qnorm[i - n] = sqrtf(s); else pnorm[i] = sqrtf(s); }
    }
    __syncthreads();
    for (int pr = warp; pr < n * n; pr += THREADS / 32) {
        const int i = pr / n, j = pr % n;               // query node i, proto node j
        const float mq = qmean[i], mp = pmean[j];
        float s = 0.f;
        for (int c = lane; c < D; c += 32) s = fmaf(Qn[i * D + c] - mq, P[j * D + c] - mp, s);
        s = warp_sum_f(s);
        if (lane == 0) {
            const float v = s / (fmaxf(qnorm[i], 1e-8f) * fmaxf(pnorm[j], 1e-8f));     // F.cosine_similarity, eps 1e-8
            sim[i][j] = v;
            cost[i][j] = 1.0 - (double)v;
            flow[i][j] = 0.0;
        }
    }
    __syncthreads();
    if (warp != 0) return;

    // ---- transportation problem on one warp: successive shortest augmenting paths (Dijkstra on reduced costs)
    // nodes 0..n-1 = query nodes (sources), n..2n-1 = proto nodes (sinks); lane l owns node l (n <= 32: two passes for sinks)
    if (lane == 0) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < n; ++i) { s1 += (double)(fmaxf(w1[i], 0.f) + 1e-5f); s2 += (double)(fmaxf(w2[i], 0.f) + 1e-5f); }
        for (int i = 0; i < n; ++i) {
            supply[i] = (double)(fmaxf(w1[i], 0.f) + 1e-5f) * ((double)n / s1);
            demand[i] = (double)(fmaxf(w2[i], 0.f) + 1e-5f) * ((double)n / s2);
        }
        for (int i = 0; i < 2 * n; ++i) pot[i] = 0.0;
    }
    __syncwarp();
    const double EPS = 1e-9 * n;
    for (int iter = 0; iter < 16 * n * n; ++iter) {
        // pick the source with the largest remaining supply
        int s = -1;
        double best = EPS;
        for (int i = 0; i < n; ++i) if (supply[i] > best) { best = supply[i]; s = i; }
        if (s < 0) break;
        // Dijkstra from s over the residual graph with reduced costs (dense, 2n nodes): lane-parallel relaxations
        for (int v = lane; v < 2 * n; v += 32) { dist[v] = 1e300; prev[v] = -1; done[v] = 0; }
        __syncwarp();
        if (lane == 0) dist[s] = 0.0;
        __syncwarp();
        int target = -1;
        for (int step = 0; step < 2 * n; ++step) {
            // closest unfinished node (all lanes scan; 2n <= 64 entries)
            int u = -1;
            double du = 1e299;
            for (int v = 0; v < 2 * n; ++v) if (!done[v] && dist[v] < du) { du = dist[v]; u = v; }
            if (u < 0) break;
            if (u >= n && demand[u - n] > EPS) { target = u; break; }      // nearest sink with residual demand
            __syncwarp();
            if (lane == 0) done[u] = 1;
            if (u < n) {                                                   // source u -> every sink j (forward arcs, cap inf)
                for (int j = lane; j < n; j += 32) {
                    const double nd = du + cost[u][j] - pot[u] + pot[n + j];
                    if (!done[n + j] && nd < dist[n + j]) { dist[n + j] = nd; prev[n + j] = u; }
                }
            } else {                                                       // sink u -> source i along backward arcs with flow
                const int j = u - n;
                for (int i = lane; i < n; i += 32) {
                    if (flow[i][j] > EPS) {
                        const double nd = du - cost[i][j] - pot[u] + pot[i];
                        if (!done[i] && nd < dist[i]) { dist[i] = nd; prev[i] = u; }
                    }
                }
            }
            __syncwarp();
        }
        if (target < 0) break;                                             // no sink with demand reachable: done (imbalance ~ eps)
        if (lane == 0) {
            // bottleneck
            double delta = fmin(supply[s], demand[target - n]);
            for (int v = target; v != s; v = prev[v]) {
                const int u = prev[v];
                if (v < n) delta = fmin(delta, flow[v][u - n]);            // backward arc sink u -> source v
            }
            for (int v = target; v != s; v = prev[v]) {
                const int u = prev[v];
                if (v >= n) flow[u][v - n] += delta; else flow[v][u - n] -= delta;
            }
            supply[s] -= delta;
            demand[target - n] -= delta;
            const double dt = dist[target];
            for (int v = 0; v < 2 * n; ++v) pot[v] -= (done[v] ? dist[v] : dt) - dt;   // pot[v] += dt - min(dist[v], dt)
        }
        __syncwarp();
    }
    __syncwarp();
    // ---- logit = sum sim * flow * T / n
    double acc = 0.0;
    for (int pr = lane; pr < n * n; pr += 32) acc += (double)sim[pr / n][pr % n] * flow[pr / n][pr % n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[(size_t)qi * W + cj] = (float)(acc * (double)temperature / (double)n);
    if (flows) {
        float* f = flows + (size_t)blockIdx.x * n * n;
        for (int pr = lane; pr < n * n; pr += 32) f[pr] = (float)flow[pr / n][pr % n];
    }
}

}  // namespace

extern "C" int sunb_emd_head(const float* proto, const float* query, float* logits, float* flows, int W, int Q, int n, int D,
                             float temperature, void* stream) {
    SUNB_REQUIRE(proto && query && logits && W > 0 && Q > 0 && D > 0, "emd_head: bad arguments");
    SUNB_REQUIRE(n >= 1 && n <= MAXN, "emd_head: 1 <= nodes <= %d (got %d)", MAXN, n);
    const size_t smem = 2 * (size_t)D * sizeof(float);
    SUNB_REQUIRE(smem <= 32 * 1024, "emd_head: D too large");
    emd_head_kernel<<<Q * W, THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(proto, query, logits, flows, W, n, D, temperature);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
