// L2 -> SM operand-delivery probe for the tensor-core GEMM (gemm_tc_q_kernel): the same TMA boxes, stage ring and cluster
// shapes as the production kernel, but NO MMA and NO stagers -- a consumer thread releases every stage as soon as it lands.
// Answers (profiles/r2_l2_ingest_probe.json):
//   * how many bytes per clock and SM the chip delivers from L2 into shared memory when every SM streams unicast tiles,
//   * whether that is a per-SM or a chip-wide limit (fewer active SMs),
//   * whether it is latency (stage count) or throughput bound,
//   * whether TMA multicast inside a cluster (A tile shared by PA CTAs, W tile shared by PW CTAs) raises the bytes DELIVERED
//     per clock, i.e. whether a multicast line costs the L2 one read or one per destination.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/_bin/l2_ingest_probe scripts/l2_ingest_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../densefusion_b200/csrc/tc_ptx.cuh"
using namespace df_tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Probe {
    int iters;          // stage fills per CTA
    int stages;         // ring depth
    int PA, PW;         // CTAs sharing one A tile (along N) / one W tile (along M); cluster = 2 * PA * PW
    int mc;             // 1: each CTA loads 1/PA of A and 1/PW of W and multicasts; 0: every CTA loads all of its tiles itself;
                        // 2 (PA == 2, PW == 1): each CTA loads HALF of the A tile from L2 and pushes it to its twin (the CTA of the other
                        // pair that needs the same rows) through distributed shared memory (cp.async.bulk shared::cta -> shared::cluster)
    int a_rows, w_rows; // rows per CTA tile: A 128 rows x 128 B, W w_rows x 128 B
    int a_blocks, w_blocks, kblocks;
};

__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, uint16_t mask)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
                 : "memory");
}

__device__ __forceinline__ void dsmem_push(void* dst_local, const void* src, uint32_t bytes, uint64_t* bar_local, uint32_t cta)
{
    uint32_t rdst, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(dst_local)), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(bar_local)), "r"(cta));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rdst), "r"(smem_u32(src)), "r"(bytes), "r"(rbar) : "memory");
}

__global__ void __launch_bounds__(96, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const Probe p, long long* cycles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = (uint32_t)p.a_rows * 128, w_bytes = (uint32_t)p.w_rows * 128;
    const uint32_t stage_bytes = (a_bytes + w_bytes + 1023) & ~1023u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + 16;
    uint64_t* half_full = full + 32;                            // mc == 2: this CTA's own half of the A tile has landed
    const int CL = 2 * p.PA * p.PW;
    const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int h = rank & 1, ia = (rank >> 1) % p.PA, iw = (rank >> 1) / p.PA;
    const int cid = blockIdx.x / CL;
    const int warp = threadIdx.x >> 5;
    const int writers = p.mc ? p.PA + p.PW - 1 : 1;             // CTAs whose TMA writes into my stages (me included)
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, writers); mbar_init(half_full + i, 1); }
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();
    uint16_t mask_a = 0, mask_w = 0;
    for (int j = 0; j < p.PA; ++j) mask_a |= (uint16_t)(1u << (h + 2 * (j + p.PA * iw)));
    for (int j = 0; j < p.PW; ++j) mask_w |= (uint16_t)(1u << (h + 2 * (ia + p.PA * j)));
    const long long t0 = clock64();
    if (warp == 0) {
        if (threadIdx.x == 0) {
            const int a_part = p.mc ? p.a_rows / p.PA : p.a_rows, w_part = p.mc ? p.w_rows / p.PW : p.w_rows;
            for (int it = 0; it < p.iters; ++it) {
                const int s = it % p.stages;
                mbar_wait(empty + s, ((it / p.stages) & 1) ^ 1);
                uint8_t* dst = smem + (size_t)s * stage_bytes;
                if (p.mc == 2) {
                    // own half of A from L2 (signals half_full), W from L2 and the twin's half by DSMEM push (both signal full)
                    const int kb2 = it % p.kblocks, tile2 = it / p.kblocks;
                    const int ablk2 = (int)(((long long)cid * 2 + h + (long long)tile2 * 37) % p.a_blocks);
                    const int wblk2 = (int)(((long long)(tile2 * p.PA + ia) * 2 + h + cid % 5) % p.w_blocks);
                    mbar_expect_tx(half_full + s, a_bytes / 2);
                    mbar_expect_tx(full + s, a_bytes / 2 + w_bytes);
                    tma_load_2d(&tm_a, dst + ia * (a_bytes / 2), half_full + s, kb2 * 32, ablk2 * p.a_rows + ia * (p.a_rows / 2));
                    tma_load_2d(&tm_w, dst + a_bytes, full + s, kb2 * 32, wblk2 * p.w_rows);
                    continue;
                }
                mbar_expect_tx(full + s, a_bytes + w_bytes);
                const int kb = it % p.kblocks, tile = it / p.kblocks;
                // the A tile of (cluster, iw, h): shared by the PA CTAs along N; the W tile of (tile, ia, h): shared along M
                const int ablk = (int)(((long long)(cid * p.PW + iw) * 2 + h + (long long)tile * 37) % p.a_blocks);
                const int wblk = (int)(((long long)(tile * p.PA + ia) * 2 + h + cid % 5) % p.w_blocks);
                if (p.mc && p.PA > 1)
                    tma_load_2d_mc(&tm_a, dst + ia * a_part * 128, full + s, kb * 32, ablk * p.a_rows + ia * a_part, mask_a);
                else
                    tma_load_2d(&tm_a, dst, full + s, kb * 32, ablk * p.a_rows);
                if (p.mc && p.PW > 1)
                    tma_load_2d_mc(&tm_w, dst + a_bytes + iw * w_part * 128, full + s, kb * 32, wblk * p.w_rows + iw * w_part, mask_w);
                else
                    tma_load_2d(&tm_w, dst + a_bytes, full + s, kb * 32, wblk * p.w_rows);
            }
        }
    } else if (threadIdx.x == 64) {
        if (p.mc == 2) {                                        // pusher: forward my half of A to the twin as soon as it has landed
            const uint32_t twin = (uint32_t)(h + 2 * (1 - ia));
            for (int it = 0; it < p.iters; ++it) {
                const int s = it % p.stages;
                mbar_wait(half_full + s, (it / p.stages) & 1);
                uint8_t* mine = smem + (size_t)s * stage_bytes + ia * (a_bytes / 2);
                dsmem_push(mine, mine, a_bytes / 2, full + s, twin);
            }
        }
    } else if (threadIdx.x == 32) {
        for (int it = 0; it < p.iters; ++it) {
            const int s = it % p.stages;
            mbar_wait(full + s, (it / p.stages) & 1);
            if (p.mc == 2) mbar_wait(half_full + s, (it / p.stages) & 1);
            if (!p.mc) mbar_arrive(empty + s);
            else {
                for (int j = 0; j < p.PA; ++j) mbar_arrive_remote(empty + s, (uint32_t)(h + 2 * (j + p.PA * iw)));
                for (int j = 0; j < p.PW; ++j)
                    if (j != iw) mbar_arrive_remote(empty + s, (uint32_t)(h + 2 * (ia + p.PA * j)));
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    cluster_sync_all();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeTiledFn fn, CUtensorMap* map, const float* base, long long rows, int K, int box_rows)
{
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char** argv)
{
    const char* out_path = argc > 1 ? argv[1] : "gpurun_out/l2_ingest_probe.json";
    const bool quick = argc > 2 && !strcmp(argv[2], "quick");     // one pass per configuration (for ncu)
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ptr);
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const int K = 384, kblocks = K / 32;
    const int a_blocks = 128, w_blocks = 20;                     // A: 128 x 128 rows x 1.5 KB = 25 MB, W: 20 x 144 rows -- L2 resident
    float *A, *W;
    long long* cyc;
    CK(cudaMalloc(&A, (size_t)a_blocks * 128 * K * 4));
    CK(cudaMalloc(&W, (size_t)w_blocks * 144 * K * 4));
    CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
    CK(cudaMemset(A, 0, (size_t)a_blocks * 128 * K * 4));
    CK(cudaMemset(W, 0, (size_t)w_blocks * 144 * K * 4));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    FILE* out = fopen(out_path, "w");
    if (!out) { perror(out_path); return 1; }
    fprintf(out, "{\"sms\": %d, \"sm_khz_max\": %d, \"a_tile_bytes\": 16384, \"w_tile_bytes\": 18432, \"runs\": [\n", sms, khz);
    struct Cfg { int PA, PW, mc, stages, frac_num, frac_den; const char* what; };
    const Cfg cfgs[] = {
        {1, 1, 0, 4, 1, 1, "pairs, unicast, 4 stages (the production kernel's traffic)"},
        {1, 1, 0, 2, 1, 1, "pairs, unicast, 2 stages"},
        {1, 1, 0, 6, 1, 1, "pairs, unicast, 6 stages"},
        {1, 1, 0, 4, 1, 2, "pairs, unicast, half of the SMs"},
        {1, 1, 0, 4, 1, 4, "pairs, unicast, a quarter of the SMs"},
        {2, 1, 0, 4, 1, 1, "cluster 4 (2 pairs along N), every CTA loads its own copy"},
        {2, 1, 1, 4, 1, 1, "cluster 4 (2 pairs along N), A tile multicast to 2"},
        {1, 2, 0, 4, 1, 1, "cluster 4 (2 pairs along M), every CTA loads its own copy"},
        {1, 2, 1, 4, 1, 1, "cluster 4 (2 pairs along M), W tile multicast to 2"},
        {2, 2, 0, 4, 1, 1, "cluster 8 (2x2 pairs), every CTA loads its own copy"},
        {2, 2, 1, 4, 1, 1, "cluster 8 (2x2 pairs), A and W multicast to 2"},
        {4, 1, 1, 4, 1, 1, "cluster 8 (4 pairs along N), A tile multicast to 4"},
        {1, 4, 1, 4, 1, 1, "cluster 8 (4 pairs along M), W tile multicast to 4"},
        {2, 4, 1, 4, 1, 1, "cluster 16 (2 along N x 4 along M), multicast"},
        {2, 1, 2, 4, 1, 1, "cluster 4 (2 pairs along N), half of A from L2 + half pushed by the twin over DSMEM"},
        {2, 1, 2, 6, 1, 1, "the same, 6 stages"},
        {2, 1, 2, 4, 1, 2, "the same, half of the clusters"},
    };
    const int ncfg = (int)(sizeof(cfgs) / sizeof(cfgs[0]));
    bool first = true;
    for (int ci = 0; ci < ncfg; ++ci) {
        const Cfg& c = cfgs[ci];
        Probe p = {};
        p.iters = quick ? 2400 : 12000; p.stages = c.stages; p.PA = c.PA; p.PW = c.PW; p.mc = c.mc;
        p.a_rows = 128; p.w_rows = 144; p.a_blocks = a_blocks; p.w_blocks = w_blocks; p.kblocks = kblocks;
        const int CL = 2 * c.PA * c.PW;
        CUtensorMap ma, mw;
        const int a_box = c.mc ? p.a_rows / c.PA : p.a_rows, w_box = c.mc == 1 ? p.w_rows / c.PW : p.w_rows;
        if (!make_map(fn, &ma, A, (long long)a_blocks * 128, K, a_box) || !make_map(fn, &mw, W, (long long)w_blocks * 144, K, w_box)) {
            fprintf(stderr, "tensor map failed\n"); return 1;
        }
        const size_t smem_bytes = 1024 + (size_t)c.stages * ((p.a_rows * 128 + p.w_rows * 128 + 1023) & ~1023) + 1024;
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(96); cfg.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cfg.gridDim = dim3(sms / CL * CL);
        int occ = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&occ, probe_kernel, &cfg);
        if (e != cudaSuccess || occ <= 0) {
            (void)cudaGetLastError();
            fprintf(stderr, "cfg %d: cluster %d not launchable (%s)\n", ci, CL, cudaGetErrorString(e));
            continue;
        }
        int clusters = occ < sms / CL ? occ : sms / CL;
        clusters = clusters * c.frac_num / c.frac_den;
        if (clusters < 1) clusters = 1;
        cfg.gridDim = dim3(clusters * CL);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best_ms = 1e30f;
        const int reps = quick ? 1 : 4;
        for (int rep = 0; rep < reps; ++rep) {                     // rep 0 warms L2
            CK(cudaEventRecord(e0));
            CK(cudaLaunchKernelEx(&cfg, probe_kernel, ma, mw, p, cyc));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if ((rep > 0 || quick) && ms < best_ms) best_ms = ms;
        }
        long long hc[1024];
        CK(cudaMemcpy(hc, cyc, sizeof(long long) * clusters * CL, cudaMemcpyDeviceToHost));
        long long cmax = 0; double csum = 0;
        for (int i = 0; i < clusters * CL; ++i) { if (hc[i] > cmax) cmax = hc[i]; csum += (double)hc[i]; }
        const double bytes_cta = (double)p.iters * (p.a_rows + p.w_rows) * 128.0;
        const double ctas = (double)clusters * CL;
        const double delivered = bytes_cta * ctas;
        const double requested = c.mc ? (double)p.iters * ctas * (p.a_rows * 128.0 / c.PA + p.w_rows * 128.0 / (c.mc == 2 ? 1 : c.PW)) : delivered;
        const double cyc_avg = csum / ctas;
        fprintf(out, "%s {\"what\": \"%s\", \"cluster\": %d, \"share_a\": %d, \"share_w\": %d, \"multicast\": %d, \"stages\": %d, \"ctas\": %d, "
                     "\"ms\": %.4f, \"delivered_TBps\": %.3f, \"l2_requested_TBps\": %.3f, \"delivered_B_per_clk_per_sm\": %.2f, "
                     "\"cycles_per_stage_fill\": %.1f, \"sm_mhz_effective\": %.0f}",
                first ? "" : ",\n", c.what, CL, c.PA, c.PW, c.mc, c.stages, clusters * CL, best_ms, delivered / best_ms * 1e-9,
                requested / best_ms * 1e-9, bytes_cta / cyc_avg, cyc_avg / p.iters, (double)cmax / best_ms * 1e-3);
        first = false;
        fflush(out);
        printf("%-70s CL=%2d ctas=%3d  %.3f ms  delivered %.2f TB/s (L2 requests %.2f TB/s)  %.1f B/clk/SM  %.0f cyc/fill\n", c.what, CL,
               clusters * CL, best_ms, delivered / best_ms * 1e-9, requested / best_ms * 1e-9, bytes_cta / cyc_avg, cyc_avg / p.iters);
    }
    fprintf(out, "\n]}\n");
    fclose(out);
    return 0;
}
