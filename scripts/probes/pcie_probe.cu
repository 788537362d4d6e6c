// pcie_probe.cu — measures, on the box, the three ways a frame's busy tiles can reach host memory:
//   (a) one pinned D2H DMA of the whole frame (what grb_read_frames_async does),
//   (b) a kernel storing tiles straight into mapped pinned host memory (zero copy, 128-bit stores),
//   (c) one DMA of a device-compacted tile buffer + a host scatter.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// tiles[i] = tile index; each block copies one 32x32 tile (colour uchar4 + depth float) of frame f
__global__ void __launch_bounds__(256) scatter_tiles(const uint4 *__restrict__ srcC, const uint4 *__restrict__ srcZ, uint4 *dstC, uint4 *dstZ,
                              const int *tiles, int nt, int W, int H, int ntx, size_t framePix4) {
    const int t = tiles[blockIdx.x], f = blockIdx.y;
    const int tx = t % ntx, ty = t / ntx;
    const int px = (threadIdx.x & 7), py = threadIdx.x >> 3;
    const int gy = ty * 32 + py;
    if (gy >= H) return;
    const size_t i = (size_t)f * framePix4 + ((size_t)gy * W) / 4 + tx * 8 + px;
    dstC[i] = srcC[i];
    dstZ[i] = srcZ[i];
}
// variant: a block takes a group of 4 horizontally adjacent tiles (all assumed busy) and every warp writes whole 512-byte
// rows of the group, so that the stores of a warp are contiguous over 4 tiles instead of 4 separate 128-byte tile rows
__global__ void __launch_bounds__(256) scatter_groups(const uint4 *__restrict__ srcC, const uint4 *__restrict__ srcZ, uint4 *dstC, uint4 *dstZ,
                               const int *groups, int ng, int W, int H, int ntx, size_t framePix4) {
    const int g = groups[blockIdx.x], f = blockIdx.y;          // group index: tile (tx0 = 4 * (g % (ntx/4)), ty)
    const int gx4 = ntx / 4;
    const int tx0 = (g % gx4) * 4, ty = g / gx4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 32; r += 8) {
        const int gy = ty * 32 + r;
        if (gy >= H) break;
        const size_t i = (size_t)f * framePix4 + ((size_t)gy * W) / 4 + tx0 * 8 + lane;
        dstC[i] = srcC[i];
        dstZ[i] = srcZ[i];
    }
}

__global__ void __launch_bounds__(256) compact_tiles(const uint4 *__restrict__ srcC, const uint4 *__restrict__ srcZ, uint4 *out,
                              const int *tiles, int nt, int W, int H, int ntx, size_t framePix4) {
    const int t = tiles[blockIdx.x], f = blockIdx.y;
    const int tx = t % ntx, ty = t / ntx;
    const int px = (threadIdx.x & 7), py = threadIdx.x >> 3;
    const int gy = ty * 32 + py;
    if (gy >= H) return;
    const size_t i = (size_t)f * framePix4 + ((size_t)gy * W) / 4 + tx * 8 + px;
    uint4 *o = out + ((size_t)f * nt + blockIdx.x) * 512;
    o[threadIdx.x] = srcC[i];
    o[256 + threadIdx.x] = srcZ[i];
}

int main(int argc, char **argv) {
    const int W = 1280, H = 720, F = 64, ntx = W / 32, nty = (H + 31) / 32;
    const size_t pix = (size_t)W * H;
    uint4 *dC, *dZ, *hC, *hZ, *dOut, *hOut;
    CK(cudaMalloc(&dC, pix * 4 * F)); CK(cudaMalloc(&dZ, pix * 4 * F));
    CK(cudaMemset(dC, 1, pix * 4 * F)); CK(cudaMemset(dZ, 2, pix * 4 * F));
    CK(cudaHostAlloc(&hC, pix * 4 * F, cudaHostAllocMapped)); CK(cudaHostAlloc(&hZ, pix * 4 * F, cudaHostAllocMapped));
    memset(hC, 0, pix * 4 * F); memset(hZ, 0, pix * 4 * F);
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char *name, double bytes, auto fn) {
        fn(); CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        const int reps = 5;
        for (int r = 0; r < reps; r++) fn();
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-58s %8.3f ms/rep  %7.2f GB/s\n", name, ms / reps, bytes / (ms / reps * 1e-3) / 1e9);
    };
    // (a) full-frame DMA
    timeit("(a) DMA 64 full frames colour+depth", 2.0 * pix * 4 * F, [&] {
        CK(cudaMemcpyAsync(hC, dC, pix * 4 * F, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(hZ, dZ, pix * 4 * F, cudaMemcpyDeviceToHost, s));
    });
    timeit("(a') DMA 1 full frame colour+depth", 2.0 * pix * 4, [&] {
        CK(cudaMemcpyAsync(hC, dC, pix * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(hZ, dZ, pix * 4, cudaMemcpyDeviceToHost, s));
    });
    // busy tile sets of different sizes: a disc of tiles around the centre (like the sphere)
    for (int nt : {130, 260, 920}) {
        std::vector<int> tiles;
        for (int r2 = 0; (int)tiles.size() < nt && r2 < 4000; r2++)
            for (int ty = 0; ty < nty && (int)tiles.size() < nt; ty++)
                for (int tx = 0; tx < ntx && (int)tiles.size() < nt; tx++) {
                    const int dx = tx - ntx / 2, dy = ty - nty / 2;
                    if (dx * dx + dy * dy == r2) tiles.push_back(ty * ntx + tx);
                }
        int *dT; CK(cudaMalloc(&dT, nt * 4)); CK(cudaMemcpy(dT, tiles.data(), nt * 4, cudaMemcpyHostToDevice));
        char name[128];
        const double bytes = (double)nt * 8192 * F;
        snprintf(name, sizeof name, "(b) zero-copy scatter %d tiles x %d frames", nt, F);
        timeit(name, bytes, [&] { scatter_tiles<<<dim3(nt, F), 256, 0, s>>>(dC, dZ, hC, hZ, dT, nt, W, H, ntx, pix / 4); });
        snprintf(name, sizeof name, "(b1) zero-copy scatter %d tiles x 1 frame", nt);
        timeit(name, bytes / F, [&] { scatter_tiles<<<dim3(nt, 1), 256, 0, s>>>(dC, dZ, hC, hZ, dT, nt, W, H, ntx, pix / 4); });
        {   // (b2) the same tiles as groups of 4 adjacent ones (rounded down to whole groups)
            std::vector<int> groups;
            std::vector<char> isBusy(ntx * nty, 0);
            for (int t : tiles) isBusy[t] = 1;
            for (int ty = 0; ty < nty; ty++)
                for (int gxi = 0; gxi < ntx / 4; gxi++) {
                    bool all = true;
                    for (int k = 0; k < 4; k++) all = all && isBusy[ty * ntx + gxi * 4 + k];
                    if (all) groups.push_back(ty * (ntx / 4) + gxi);
                }
            if (!groups.empty()) {
                int *dG; CK(cudaMalloc(&dG, groups.size() * 4)); CK(cudaMemcpy(dG, groups.data(), groups.size() * 4, cudaMemcpyHostToDevice));
                snprintf(name, sizeof name, "(b2) zero-copy, %zu groups of 4 tiles (512 B rows) x %d frames", groups.size(), F);
                const int ng = (int)groups.size();
                timeit(name, (double)ng * 4 * 8192 * F, [&] { scatter_groups<<<dim3(ng, F), 256, 0, s>>>(dC, dZ, hC, hZ, dG, ng, W, H, ntx, pix / 4); });
                cudaFree(dG);
            }
        }
        CK(cudaMalloc(&dOut, (size_t)nt * 8192 * F)); CK(cudaHostAlloc(&hOut, (size_t)nt * 8192 * F, cudaHostAllocDefault));
        snprintf(name, sizeof name, "(c) compact + DMA %d tiles x %d frames (no host scatter)", nt, F);
        timeit(name, bytes, [&] {
            compact_tiles<<<dim3(nt, F), 256, 0, s>>>(dC, dZ, dOut, dT, nt, W, H, ntx, pix / 4);
            CK(cudaMemcpyAsync(hOut, dOut, (size_t)nt * 8192 * F, cudaMemcpyDeviceToHost, s));
        });
        // host scatter cost of (c), single thread
        {
            auto t0 = std::chrono::steady_clock::now();
            for (int f = 0; f < F; f++)
                for (int k = 0; k < nt; k++) {
                    const int t = tiles[k], tx = t % ntx, ty = t / ntx;
                    const char *src = (const char *)(hOut + ((size_t)f * nt + k) * 512);
                    for (int py = 0; py < 32 && ty * 32 + py < H; py++) {
                        memcpy((char *)hC + ((size_t)f * pix + (size_t)(ty * 32 + py) * W + tx * 32) * 4, src + py * 128, 128);
                        memcpy((char *)hZ + ((size_t)f * pix + (size_t)(ty * 32 + py) * W + tx * 32) * 4, src + 4096 + py * 128, 128);
                    }
                }
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("    host scatter of (c), 1 thread: %.3f ms  %.2f GB/s\n", sec * 1e3, bytes / sec / 1e9);
        }
        cudaFree(dOut); cudaFreeHost(hOut); cudaFree(dT);
    }
    // check the zero-copy result landed
    CK(cudaDeviceSynchronize());
    const unsigned char *b = (const unsigned char *)hC;
    size_t ones = 0;
    for (size_t i = 0; i < pix * 4; i += 4096) ones += b[i] == 1;
    printf("zero-copy landed: %zu of %zu probes\n", ones, pix * 4 / 4096);
    return 0;
}
