// spmm_cluster.cu -- K-spmm-db in its cluster form: banded complex SpMM + |z|^2 + power_to_db with the
// kernel coefficients stationary in shared memory.
//
// Replaces sprs::prod::mul_acc_mat_vec_csr x (4 + 2), norm_sqr and power_to_db (vqt.rs:889-910, :922-954).
//
//   * A thread-block cluster of CS CTAs splits the kernel rows into CS contiguous, work-balanced parts.
//     Each CTA copies its part's coefficients (<= ~100 KB) into shared memory once and then loops over
//     16-frame rounds (two 8-frame tiles): nothing but spectra streams through the SM, and every
//     coefficient load is an LDS with a compile-time offset instead of a dependent L2 round trip.
//   * A warp owns 8 row pairs ("units").  Lane = (half h, tile t, unit u): the two halves of a unit's band
//     go to lanes l and l ^ 16 and are added with shuffles at the end, which halves the longest band a
//     warp walks and balances the warps; the two tiles of the round share the coefficient load.
//   * The staged tile is de-swizzled into four planes [chunk][column] of 16-byte entries, so a lane's four
//     spectrum loads are one pointer plus compile-time offsets; units are placed (host) on distinct column
//     residues modulo 8, which makes every quarter-warp LDS.128 conflict-free for the whole band.
//   * power_to_db needs each frame's max / min over all rows: every CTA reduces its part and writes the pair
//     into the shared memory of all CTAs of the cluster (DSMEM); one cluster barrier per round.
#include <algorithm>

#include <cooperative_groups.h>

#include "device_helpers.cuh"
#include "vqt_device.cuh"

namespace cg = cooperative_groups;

namespace pvqt_dev {
namespace {

constexpr int kPlaneF4 = kClusterPlaneCols;          // float4 entries per plane
constexpr int kTileF4 = 4 * kPlaneF4;                // float4 entries per staged tile

__global__ void __launch_bounds__(kClusterThreads, 2) spmm_db_cluster_kernel(const __grid_constant__ ClusterParams P)
{
    extern __shared__ __align__(128) float4 cl_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned cs = cluster.num_blocks();
    const unsigned cid = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_warps_cta = blockDim.x >> 5;
    const ClusterPart part = P.parts[rank];

    // shared memory: coefficients | tile[2] (4 planes each; reused as ls[16][max_rows] after the band walk) | mm[2][cs][16]
    float4 *coef = cl_smem;
    float4 *tile = cl_smem + P.coef_bytes / 16;
    float *ls = reinterpret_cast<float *>(tile);
    float2 *mm = reinterpret_cast<float2 *>(reinterpret_cast<char *>(cl_smem) + P.mm_offset);

    // ---- prologue (plan data only: overlaps the producer kernel's tail) ----
    {
        const float4 *src = P.coef + (size_t)part.coef_base * 16;
        const int n = part.coef_slots * 16;
        for (int i = tid; i < n; i += blockDim.x) cp_async16(coef + i, src + i);
    }
    const int h = lane >> 4, t = (lane >> 3) & 1, u = lane & 7;
    const bool has_work = warp < part.n_warps;
    ClusterWarp W = {0, 0, 0, 0};
    ClusterLane L = {0, 0, -1, 0};
    if (has_work) {
        W = P.warps[part.desc_base + warp];
        L = P.lanes[(size_t)(part.desc_base + warp) * 16 + h * 8 + u];
    }
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();
    pdl_wait();  // the spectra come from K-fft / K-sdft

    const float4 *kp0 = coef + (size_t)W.slot_base * 16 + h * 8 + u;
    const float4 *np0 = coef + (size_t)W.nslot_base * 16 + h * 8 + u;
    const float4 *xt = tile + t * kTileF4;   // this lane's tile of the round
    const int nb = P.n_buckets;
    unsigned parity = 0;

    for (unsigned round = cid; round * 2 < P.n_tiles; round += n_clusters, parity ^= 1) {
        // ---- stage the round's two tiles, de-swizzled into planes: plane q, column c <- chunk q ^ ((c >> 1) & 3)
        for (int tt = 0; tt < 2; ++tt) {
            const unsigned tl = round * 2 + tt;
            if (tl >= P.n_tiles) break;
            const float4 *src = reinterpret_cast<const float4 *>(P.spec) + ((size_t)tl * P.spec_stride + part.col_lo) * 4;
            float4 *dst = tile + tt * kTileF4;
            const int n = part.n_cols * 4;
            for (int i = tid; i < n; i += blockDim.x) {
                const int c = i >> 2, p = i & 3;               // physical chunk p of column c (col_lo is a multiple of 8)
                const int q = p ^ ((c >> 1) & 3);              // logical chunk
                cp_async16(dst + q * kPlaneF4 + c, src + i);
            }
            // columns the band halves may touch past the staged range carry zero coefficients: keep them finite
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < 4 * (part.cols_touched - part.n_cols); i += blockDim.x)
                dst[(i & 3) * kPlaneF4 + part.n_cols + (i >> 2)] = z;
        }
        cp_async_wait_all();
        __syncthreads();

        float2 re0[4], im0[4], re1[4], im1[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) re0[p] = im0[p] = re1[p] = im1[p] = make_float2(0.f, 0.f);
        if (has_work) {
            {
                const float4 *kp = kp0;
                const float4 *xp = xt + L.col;
#pragma unroll 2
                for (int j = 0; j < W.width; ++j) {
                    const float4 k = kp[0];
                    const float4 xr03 = xp[0], xr47 = xp[kPlaneF4], xi03 = xp[2 * kPlaneF4], xi47 = xp[3 * kPlaneF4];
                    kp += 16;
                    xp += 1;
                    mac8<false>(re0, im0, k.x, k.y, xr03, xr47, xi03, xi47);
                    mac8<false>(re1, im1, k.z, k.w, xr03, xr47, xi03, xi47);
                }
            }
            {
                const float4 *kp = np0;
                const float4 *xp = xt + L.ncol;
                for (int j = 0; j < W.nwidth; ++j) {
                    const float4 k = kp[0];  // conj(Kneg), see the device plan
                    const float4 xr03 = xp[0], xr47 = xp[kPlaneF4], xi03 = xp[2 * kPlaneF4], xi47 = xp[3 * kPlaneF4];
                    kp += 16;
                    xp += 1;
                    mac8<true>(re0, im0, k.x, k.y, xr03, xr47, xi03, xi47);
                    mac8<true>(re1, im1, k.z, k.w, xr03, xr47, xi03, xi47);
                }
            }
            // add the two halves of the band (lanes l and l ^ 16): both lanes end up with the full sums
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                re0[p].x += __shfl_xor_sync(0xffffffffu, re0[p].x, 16); re0[p].y += __shfl_xor_sync(0xffffffffu, re0[p].y, 16);
                im0[p].x += __shfl_xor_sync(0xffffffffu, im0[p].x, 16); im0[p].y += __shfl_xor_sync(0xffffffffu, im0[p].y, 16);
                re1[p].x += __shfl_xor_sync(0xffffffffu, re1[p].x, 16); re1[p].y += __shfl_xor_sync(0xffffffffu, re1[p].y, 16);
                im1[p].x += __shfl_xor_sync(0xffffffffu, im1[p].x, 16); im1[p].y += __shfl_xor_sync(0xffffffffu, im1[p].y, 16);
            }
        }
        __syncthreads();  // every warp is done with the staged tiles: their memory becomes ls[16][max_rows]
        // half 0 takes the pair's first row, half 1 the second: |z|^2 (norm_sqr) and log_spec, vqt.rs:930
        if (has_work && L.n_rows > h) {
            const uint32_t frame0 = (round * 2 + t) * kTileFrames;
            const int row = L.row + h;
#pragma unroll
            for (int f = 0; f < kTileFrames; ++f) {
                const float2 r2 = h == 0 ? re0[f >> 1] : re1[f >> 1], i2 = h == 0 ? im0[f >> 1] : im1[f >> 1];
                const float zr = (f & 1) ? r2.y : r2.x, zi = (f & 1) ? i2.y : i2.x;
                const float p = zr * zr + zi * zi;
                ls[(t * kTileFrames + f) * P.max_rows + row] = log_spec(p, P.ref_db);
                if (P.power != nullptr && frame0 + f < P.n_frames)
                    P.power[(size_t)(frame0 + f) * nb + part.row_lo + row] = p;
            }
        }
        __syncthreads();

        // ---- this part's max / min per frame -> every CTA of the cluster (vqt.rs:933-938) ----
        for (int f = warp; f < kClusterRoundFrames; f += n_warps_cta) {
            const float *l = ls + f * P.max_rows;
            float mx = -CUDART_INF_F, mn = CUDART_INF_F;
            for (int r = lane; r < part.n_rows; r += 32) {
                mx = fmaxf(mx, l[r]);
                mn = fminf(mn, l[r]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            }
            if (lane < (int)cs) {
                float2 *remote = cluster.map_shared_rank(mm, lane);
                remote[(parity * cs + rank) * kClusterRoundFrames + f] = make_float2(mx, mn);
            }
        }
        cluster.sync();

        // ---- power_to_db's frame-wise part (vqt.rs:939-950) for the rows this CTA owns ----
        for (int f = warp; f < kClusterRoundFrames; f += n_warps_cta) {
            const uint32_t frame = round * kClusterRoundFrames + f;
            if (frame >= P.n_frames) continue;
            float mx = -CUDART_INF_F, mn = CUDART_INF_F;
            for (unsigned r = 0; r < cs; ++r) {
                const float2 v = mm[(parity * cs + r) * kClusterRoundFrames + f];
                mx = fmaxf(mx, v.x);
                mn = fminf(mn, v.y);
            }
            const float floor_db = mx - kTopDb, log_spec_min = fmaxf(mn, floor_db);
            const float *l = ls + f * P.max_rows;
            float *out = P.out_db + (size_t)frame * nb + part.row_lo;
            for (int r = lane; r < part.n_rows; r += 32) out[r] = db_out(l[r], floor_db, log_spec_min);
        }
        __syncthreads();  // ls and the tiles are rewritten by the next round
    }
}

}  // namespace

size_t cluster_smem_bytes(int coef_bytes, int max_rows, int cluster_size)
{
    // the log-spectrum buffer ls[16][max_rows] aliases the two staged tiles
    const size_t tiles = std::max(2 * (size_t)kTileF4 * sizeof(float4), (size_t)kClusterRoundFrames * max_rows * sizeof(float));
    return (size_t)coef_bytes + tiles + (size_t)2 * cluster_size * kClusterRoundFrames * sizeof(float2);
}

namespace {
void fill_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int n_clusters, int cluster_size, size_t smem,
                 cudaStream_t stream)
{
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)(n_clusters * cluster_size));
    cfg.blockDim = dim3(kClusterThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
}
}  // namespace

cudaError_t configure_cluster(int coef_bytes, int max_rows, int cluster_size, int *max_clusters)
{
    const size_t smem = cluster_smem_bytes(coef_bytes, max_rows, cluster_size);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(spmm_db_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    fill_config(cfg, attr, 1, cluster_size, smem, nullptr);
    cfg.numAttrs = 1;  // occupancy of the cluster shape only
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, spmm_db_cluster_kernel, &cfg);
    if (e != cudaSuccess) return e;
    *max_clusters = n;
    return cudaSuccess;
}

cudaError_t launch_spmm_db_cluster(const ClusterParams &p, int n_clusters, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    fill_config(cfg, attr, n_clusters, p.cluster_size, cluster_smem_bytes(p.coef_bytes, p.max_rows, p.cluster_size), stream);
    return cudaLaunchKernelEx(&cfg, spmm_db_cluster_kernel, p);
}

}  // namespace pvqt_dev
