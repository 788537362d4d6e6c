// ipc_probe.cu — two processes, two GPUs: can a kernel of process B store straight into a cudaMalloc'ed buffer of
// process A (cudaIpc* + NVLink peer mapping), how fast, and what does a device-side flag hand-off cost?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ipc_probe ipc_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/socket.h>
#include <sys/wait.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("[%d] %s: %s\n", getpid(), #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void fill(uint4 *dst, size_t n, unsigned v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_uint4(v, v, v, v);
}
__global__ void signal_flag(volatile unsigned *flag, unsigned v) { __threadfence_system(); *flag = v; }
__global__ void wait_flag(volatile unsigned *flag, unsigned v) { while (*flag < v) { } __threadfence_system(); }

int main() {
    int sv[2];
    socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
    const size_t bytes = 64u << 20;
    pid_t pid = fork();
    if (pid != 0) {   // A: owner, GPU 0
        CK(cudaSetDevice(0));
        void *buf; unsigned *flag;
        CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&flag, 256)); CK(cudaMemset(flag, 0, 256)); CK(cudaMemset(buf, 0, bytes));
        cudaIpcMemHandle_t h[2];
        CK(cudaIpcGetMemHandle(&h[0], buf)); CK(cudaIpcGetMemHandle(&h[1], flag));
        CK(cudaDeviceSynchronize());
        write(sv[0], h, sizeof h);
        // wait for the 10 rounds through the device flag
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (unsigned r = 1; r <= 10; r++) {
            CK(cudaEventRecord(e0));
            wait_flag<<<1, 1>>>(flag, r);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            unsigned probe[4];
            CK(cudaMemcpy(probe, (char *)buf + bytes - 16, 16, cudaMemcpyDeviceToHost));
            if (probe[0] != r) printf("A: round %u sees %u at the end of the buffer (STALE)\n", r, probe[0]);
        }
        printf("A: all rounds observed through the device flag\n");
        char c; read(sv[0], &c, 1);
        int st; waitpid(pid, &st, 0);
        printf("A: child exit %d\n", WEXITSTATUS(st));
        return 0;
    }
    // B: writer, GPU 1
    CK(cudaSetDevice(1));
    cudaIpcMemHandle_t h[2];
    read(sv[1], h, sizeof h);
    void *buf; unsigned *flag;
    cudaError_t e = cudaIpcOpenMemHandle(&buf, h[0], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { printf("B: cudaIpcOpenMemHandle failed: %s\n", cudaGetErrorString(e)); char c = 0; write(sv[1], &c, 1); return 2; }
    CK(cudaIpcOpenMemHandle((void **)&flag, h[1], cudaIpcMemLazyEnablePeerAccess));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (unsigned r = 1; r <= 10; r++) {
        CK(cudaEventRecord(e0));
        fill<<<148 * 4, 256>>>((uint4 *)buf, bytes / 16, r);
        CK(cudaEventRecord(e1));
        signal_flag<<<1, 1>>>(flag, r);
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("B: round %u peer store of %zu MB over NVLink: %.3f ms = %.1f GB/s\n", r, bytes >> 20, ms, bytes / (ms * 1e-3) / 1e9);
        usleep(20000);
    }
    char c = 1; write(sv[1], &c, 1);
    return 0;
}
